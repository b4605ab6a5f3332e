#!/usr/bin/env python3
"""Summarise ncu reports (gpurun_out/*.ncu-rep) into profiles/: one JSON + text per kernel.
    python tools/ncu_summary.py gpurun_out/ctc_rows_r1.ncu-rep [...]"""
import csv, io, json, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            d[h] = (v, u)
        res.append(d)
    return res


def main():
    os.makedirs("profiles", exist_ok=True)
    traffic_path = os.path.join("profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for path in sys.argv[1:]:
        for d in raw(path):
            name = d.get("Kernel Name", ("?", ""))[0]
            short = name.split("(")[0].replace("void ", "").replace("asr::", "").split("<")[0]
            summ = {"report": os.path.basename(path), "kernel": name}
            for k in KEYS:
                if k in d:
                    v, u = d[k]
                    try:
                        fv = float(v.replace(",", ""))
                    except ValueError:
                        continue
                    summ[k] = {"value": fv, "unit": u}
            rd, wr = summ.get("dram__bytes_read.sum"), summ.get("dram__bytes_write.sum")
            if rd and wr:
                tb = rd["value"] * UNIT.get(rd["unit"], 1) + wr["value"] * UNIT.get(wr["unit"], 1)
                summ["dram_bytes_per_launch"] = tb
                traffic[short + "_bytes_per_launch"] = tb
            base = os.path.join("profiles", os.path.basename(path).replace(".ncu-rep", ""))
            json.dump(summ, open(base + ".json", "w"), indent=1)
            with open(base + ".txt", "w") as f:
                f.write("ncu --set full --clock-control none (one launch)  report: %s\n%s\n" % (os.path.basename(path), name))
                for k, v in summ.items():
                    if isinstance(v, dict):
                        f.write("%-95s %14.4f %s\n" % (k, v["value"], v["unit"]))
                if "dram_bytes_per_launch" in summ:
                    f.write("%-95s %14.0f byte\n" % ("dram bytes read+write per launch", summ["dram_bytes_per_launch"]))
            print(base + ".txt")
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
