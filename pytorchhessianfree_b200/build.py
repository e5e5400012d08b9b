"""Compile the sm_100a CUDA sources into the in-tree C-ABI shared library.

    python -m pytorchhessianfree_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is written next to the sources
(``csrc/libhf_b200.so``) so that it travels with the tree to the GPU box; it is
git-ignored.  cudart is linked statically: the library only shares driver-level
handles (device pointers, CUstream) with PyTorch.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libhf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "hf_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Build ``libhf_b200.so`` if any source is newer than it.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("HF_B200_DEFS", "").split(), "-c", src, "-o", obj]  # e.g. -DHF_TC_ITER_TRACE=1
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB + ".tmp", *objs]
    subprocess.run(link, check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
