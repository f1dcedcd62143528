"""Builds libm3p2i_b200.so in-tree for sm_100a:  python -m m3p2i_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
OUT = os.path.join(HERE, "libm3p2i_b200.so")
SOURCES = ["kernels.cu", "api.cu"]
HEADERS = ["common.cuh", "point_env.cuh", "panda_env.cuh", "panda_team.cuh", "panda_far.cuh", "rollout_common.cuh", "kernels.cuh", "params_host.h", os.path.join("..", "..", "include", "m3p2i_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
