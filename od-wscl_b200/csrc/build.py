"""Build libodwscl_sm100.so (the C-ABI kernel library) with nvcc for sm_100a, in-tree.

    python od-wscl_b200/csrc/build.py [--force] [--verbose]

Each .cu is compiled to an object in csrc/build/ (parallel), then linked into
od-wscl_b200/lib/libodwscl_sm100.so.  No torch dependency: the library is plain CUDA runtime.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(HERE, "build")
SO = os.path.join(OUT_DIR, "libodwscl_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(HERE, "*.cu")))
    hdrs = glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(" ".join(cmd[-3:]) + "\n" + res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError("nvcc failed: " + cmd[-3])
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(SO, objs):
        subprocess.check_call([NVCC, "-shared", "-o", SO] + objs + ["-lcudart"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
