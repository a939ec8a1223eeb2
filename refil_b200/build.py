"""Compile refil_b200/csrc/*.cu into refil_b200/librefil_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m refil_b200.build [--force]

The .so is built IN-TREE so it travels to the GPU box with the repo snapshot; it is git-ignored.
"""
import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "librefil_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "refil_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Build the shared library if any source is newer than it.  Returns the path."""
    if not force and not _stale():
        return SO_PATH
    objs = []
    bdir = os.path.join(_HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        newest = max([os.path.getmtime(src), os.path.getmtime(HEADER)] +
                     [os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))])
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            continue
        cmd = [_nvcc()] + flags + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode(errors="replace")))
    cmd = [_nvcc(), "-shared", "--cudart", "shared", "-Wno-deprecated-gpu-targets", "-o", SO_PATH] + objs
    subprocess.check_call(cmd)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
