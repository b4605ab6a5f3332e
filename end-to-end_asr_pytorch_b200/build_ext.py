"""Builds csrc/*.cu into csrc/libasr_sm100.so with nvcc for sm_100a (in-tree).

    python end-to-end_asr_pytorch_b200/build_ext.py [--force] [--verbose]

The shared library is a plain C-ABI library (include/asr_sm100.h); it links
only against the CUDA runtime, not against torch or libcuda.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libasr_sm100.so")
STAMP = os.path.join(CSRC, ".build_stamp")
SOURCES = ["common.cu", "cif.cu", "assigner.cu", "specaug.cu", "ctc.cu", "mha.cu", "gemm.cu", "gemm2.cu", "allreduce.cu", "ln.cu"]
HEADERS = ["common.cuh", "tcgen05.cuh", "philox.cuh", os.path.join("..", "..", "include", "asr_sm100.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("ASR_NVCC_EXTRA", "").split()   # e.g. -DASR_MHA_TRACE for the clock64 traces of tools/mha_trace.py


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _fingerprint():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _fingerprint()


def build(force=False, verbose=False):
    """Compile if the sources changed.  Returns the path of the shared library."""
    if not force and is_current():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libasr_sm100.so (set NVCC=/path/to/nvcc)")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for src in srcs:
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("==== %s ====\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(log)[-6000:])
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(_fingerprint())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
