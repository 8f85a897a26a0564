"""Build libvulkansift.so (the C-ABI product) with nvcc for sm_100a, in-tree.

    python -m vulkansift_b200.build [--force] [--verbose]

The library is self-contained (static cudart, driver entry points resolved at
run time) and lands in vulkansift_b200/lib/ so that it travels with a repository
snapshot.  -fmad=false / -ffp-contract=off are part of the arithmetic contract
of include/vksift_arith.h, not tuning flags.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libvulkansift.so")
# same sources with -DVKS_ANALYSIS: adds vksiftx_setDebugSkip (stage ablation, invalid results).  Only tools/ablation.py and
# bench.py's marginal-cost measurement load it; every timed or parity-checked detection runs on LIB, which has no such switch.
LIB_ANALYSIS = os.path.join(HERE, "lib", "libvulkansift_analysis.so")
SOURCES = ["api.cu", "plan.cu", "pyramid.cu", "pyramid_strip.cu", "extrema.cu", "describe.cu", "match.cu", "exchange.cu"]
HEADERS = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
    os.path.join(ROOT, "include", "vksift_arith.h"), os.path.join(ROOT, "include", "vksift_b200_ext.h"),
    os.path.join(ROOT, "include", "vulkansift", "vulkansift.h"), os.path.join(ROOT, "include", "vulkansift", "vulkansift_types.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall,-Wno-unused-function",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    objs = []
    procs = []
    analysis_objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s, __file__] + HEADERS):
            cmd = [cc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        if src == "api.cu":  # the only translation unit that looks at VKS_ANALYSIS
            oa = os.path.join(OBJ, "api_analysis.o")
            analysis_objs.append(oa)
            if force or _stale(oa, [s, __file__] + HEADERS):
                cmd = [cc] + NVCC_FLAGS + ["-DVKS_ANALYSIS", "-c", s, "-o", oa]
                procs.append((src + " (analysis)", subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        else:
            analysis_objs.append(o)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [cc, "-shared", "-o", LIB] + objs + ["-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    if force or procs or _stale(LIB_ANALYSIS, analysis_objs):
        cmd = [cc, "-shared", "-o", LIB_ANALYSIS] + analysis_objs + ["-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed (analysis library)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
