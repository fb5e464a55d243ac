"""In-tree build of libkge_b200.so (CUDA kernels + C ABI) for sm_100a.

Usage: ``python -m torchkge_b200._build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU; the resulting .so lives in torchkge_b200/lib/ (git-ignored,
but shipped to the GPU box by gpurun).  Objects are rebuilt only when a source or header
is newer than them.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(LIBDIR, "libkge_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libkge_b200.so")


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") or f.endswith(".cpp"))
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdr.append(os.path.join(ROOT, "include", "kge_b200.h"))
    return [os.path.join(CSRC, f) for f in cu], hdr


def build(force=False, verbose=False):
    srcs, hdrs = _sources()
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
        objs.append(obj)
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_hdr))
        if stale:
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        text = out.decode(errors="replace")
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, text))
        elif verbose:
            sys.stderr.write(text)
    if failed:
        raise RuntimeError("libkge_b200 build failed")
    relink = force or procs or not os.path.exists(LIB) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
