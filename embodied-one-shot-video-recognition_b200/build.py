"""In-tree build of the CUDA library (sm_100a only).

    python embodied-one-shot-video-recognition_b200/build.py   ->  .../libeosvr.so

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU
box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libeosvr.so")
SOURCES = ["eosvr_api.cu", "eosvr_match.cu", "eosvr_episode.cu"]
HEADERS = ["eosvr_internal.h", "eosvr_ptx.cuh", os.path.join("..", "..", "include", "eosvr.h")]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", OUT] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
