"""Build libmfb200.so (sm_100a) in-tree with nvcc.  `python -m microstructure_fingerprinting_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmfb200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]
# exact.cu reproduces the reference's separately-rounded arithmetic: no FMA contraction.
UNITS = [("exact.cu", ["-fmad=false"]), ("fast.cu", []), ("mc.cu", []), ("api.cu", [])]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "mfb200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False, experiments=False):
    """experiments=True builds libmfb200_exp.so with -DMFB_EXPERIMENTS (timing ablations and
    counters steered by environment variables; never loaded unless MFB_LIB points at it)."""
    lib = LIB.replace(".so", "_exp.so") if experiments else LIB
    if not force and not experiments and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src.replace(".cu", "_exp.o" if experiments else ".o"))
        if experiments:
            extra = extra + ["-DMFB_EXPERIMENTS"]
        cmd = [nvcc] + ARCH + COMMON + extra + ["-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
