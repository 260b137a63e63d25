"""Build liblavila_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m lavila_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to lavila_b200/_build/, the library to
lavila_b200/liblavila_b200.so (git-ignored, but shipped to the GPU box by gpurun).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "liblavila_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(open(path, "rb").read())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(INCLUDE, "lavila_b200.h"), "rb").read())
    return h.hexdigest()


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    log = p.stdout + p.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs, relink = [], [], force or not os.path.exists(LIB_PATH)
    for src in _sources():
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ_DIR, base + ".o")
        stamp = obj + ".sha1"
        dig = _digest(src)
        objs.append(obj)
        if force or not os.path.exists(obj) or not os.path.exists(stamp) or open(stamp).read() != dig:
            jobs.append((nvcc, src, obj, verbose, stamp, dig))
    if jobs:
        relink = True
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_compile_one, [j[:4] for j in jobs]))
        for j in jobs:
            with open(j[4], "w") as f:
                f.write(j[5])
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
