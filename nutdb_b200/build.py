"""Builds the in-tree native libraries (no JIT cache: the .so files travel with the repo snapshot).

  libnutdb_gpu.so       CUDA kernels + C ABI (include/nutdb_gpu.h), sm_100a only
  libnutdb_workload.so  synthetic batch generator used by bench.py and the tests
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GPU_SO = os.path.join(HERE, "libnutdb_gpu.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]
GPU_SOURCES = ["nutdb_gpu.cu", "hydrate.cpp", "dispatch.cpp"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp"))]
    d.append(os.path.join(os.path.dirname(HERE), "include", "nutdb_gpu.h"))
    return d


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build_gpu(force=False, verbose=False):
    if not force and not _stale(GPU_SO, _deps()):
        return GPU_SO
    # the parser bytecode is generated from the grammar description
    subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_parse_program.py")], stdout=subprocess.DEVNULL)
    srcs = [os.path.join(CSRC, s) for s in GPU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", GPU_SO] + srcs
    subprocess.check_call(cmd)
    return GPU_SO


def build_all(force=False, verbose=False):
    from . import workload
    build_gpu(force, verbose)
    workload.build(force)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    from nutdb_b200 import workload
    build_gpu(force="--force" in sys.argv, verbose="-v" in sys.argv)
    workload.build(force="--force" in sys.argv)
    print("built", GPU_SO)
