#!/usr/bin/env python3
"""Per-kernel SASS opcode evidence for the Blackwell-native claims (B200_PROFILING.md, "What proves a Blackwell-native
kernel"): disassembles csrc/libasr_sm100.so with cuobjdump and counts, per kernel, the tcgen05 / TMEM / TMA opcodes
(UTC*MMA, LDTM / STTM, UTMALDG / UTMASTG / UTMAREDG, UTCBAR), the packed fp32 pairs (FFMA2 / FADD2 / FMUL2), MUFU and
the legacy tensor path (HMMA - must be absent).  Runs without a GPU.

    python tools/sass_histogram.py > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "end-to-end_asr_pytorch_b200", "csrc", "libasr_sm100.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "LDGMC", "SYNCS", "FFMA2", "FADD2",
         "FMUL2", "MUFU", "HMMA", "LDGSTS", "REDUX", "SHFL", "BAR", "ELECT"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1).split(".")[0]] += 1
    names = demangle(list(per))
    print("# SASS opcode counts per kernel of %s (cuobjdump -sass; instructions, not executions)" % os.path.relpath(LIB, ROOT))
    print("# columns: total | " + " ".join(WATCH))
    totals = collections.Counter()
    for fn, cnt in per.items():
        short = re.sub(r"\(.*", "", names.get(fn, fn))
        short = re.sub(r"^void ", "", short)
        row = [cnt.get(k, 0) for k in WATCH]
        for k, v in zip(WATCH, row):
            totals[k] += v
        print("%-64s %6d | %s" % (short[:64], sum(cnt.values()), " ".join("%s=%d" % (k, v) for k, v in zip(WATCH, row) if v)))
    print("# library totals: " + " ".join("%s=%d" % (k, totals[k]) for k in WATCH))
    if totals["HMMA"]:
        print("# WARNING: legacy mma.sync tensor path present", file=sys.stderr)


if __name__ == "__main__":
    main()
