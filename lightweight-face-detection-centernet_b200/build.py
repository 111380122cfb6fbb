"""In-tree nvcc build of the C-ABI library (sm_100a only; cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcenterface_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-lcuda", "-ldl"]


def sources():
    out = [os.path.join(CSRC, "engine.cu")]
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "centerface_b200.h"))
    return out, deps


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources()[1])


def build(force=False, verbose=False):
    """Compile csrc/engine.cu -> libcenterface_b200.so next to this file.  Returns the path."""
    if not force and not is_stale():
        return LIB
    srcs, _ = sources()
    extra = os.environ.get("CF_NVCC_FLAGS", "").split()  # development A/B builds, e.g. -DCF_MBF_SWISHQ=0
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
