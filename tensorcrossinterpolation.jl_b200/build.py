"""Builds libtci_b200.so (sm_100a only) from csrc/*.cu with nvcc, in-tree.

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the
repository snapshot.  Usage: python tensorcrossinterpolation.jl_b200/build.py [--force] [-v]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtci_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["ctx.cu", "group.cu", "bond.cu", "pi_eval.cu", "rrlu.cu", "rrlu_lazy.cu", "luci.cu", "dgemm.cu", "tt.cu", "mpo.cu", "zgemm.cu", "zpath.cu", "cache.cu", "user_target.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def _deps():
    inc = os.path.join(HERE, "..", "include")
    return [os.path.join(CSRC, "tci_internal.h"), os.path.join(CSRC, "rrlu_common.cuh"), os.path.join(inc, "tci_b200.h"),
            os.path.join(inc, "tci_targets.h"), os.path.join(inc, "tci_zarith.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    if verbose and log:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
