"""Build libbtsdsp.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m openbts_ttsou_b200.build          # or: from openbts_ttsou_b200.build import build; build()

nvcc cross-compiles without a GPU.  Flags that matter for parity with the reference's scalar float
arithmetic: -fmad=false (no FMA contraction), IEEE division and square root
(-prec-div/-prec-sqrt true), no flush-to-zero; host code with -ffp-contract=off.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbtsdsp.so")
SOURCES = ["capi.cu", "kernels.cu", "resample.cu"]
# every header of csrc/ (a stale-check that names them one by one misses the next one added) + the public header
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))) + ["../../include/btsdsp.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden",
    "-cudart", "shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile and link libbtsdsp.so if any source is newer.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    # parity fence: packed multiplies are used for code density, but a fused packed multiply-add would break the
    # reference's separately-rounded arithmetic -- the only FFMA2 allowed is cplx.cuh's pmul0 (addend = zero register)
    for obj in objs:
        sass = subprocess.run([os.path.join(os.path.dirname(nvcc), "cuobjdump"), "-sass", obj], capture_output=True, text=True)
        bad = [l for l in sass.stdout.splitlines() if "FFMA2" in l and not re.search(r"RZ\.F32\s*;", l)]
        if sass.returncode == 0 and bad:
            raise RuntimeError("%s contains FFMA2 with a live addend: a packed mul+add was fused, results would not be "
                               "bit-exact\n%s" % (obj, bad[0]))
    cmd = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
