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
SOURCES = ["eosvr_api.cu", "eosvr_match.cu", "eosvr_episode.cu", "eosvr_selfcheck.cu"]
HEADERS = ["eosvr_internal.h", "eosvr_ptx.cuh", os.path.join("..", "..", "include", "eosvr.h")]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True builds libeosvr_exp.so with -DEOSVR_EXPERIMENTS: the timing modes that skip work
    (EOSVR_EXP bits 1/2/4/32, wrong results) exist only there, never in the shipped library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    out = OUT.replace("libeosvr.so", "libeosvr_exp.so") if experiments else OUT
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest(deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", out] + srcs
    if experiments:
        cmd.insert(1, "-DEOSVR_EXPERIMENTS")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
