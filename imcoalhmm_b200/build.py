"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "imc_lib.cu")
OUT = os.environ.get("IMC_LIB_PATH") or os.path.join(HERE, "libimcoalhmm_b200.so")   # IMC_LIB_PATH: experiment builds
DEPS = [SRC, os.path.join(HERE, "csrc", "forward_kernels.cuh"),
        os.path.join(HERE, "csrc", "zip_kernels.cuh"),
        os.path.join(HERE, "csrc", "tokenizer.inl"),
        os.path.join(HERE, "csrc", "model_host.inl"),
        os.path.join(HERE, "csrc", "zip_host.inl"),
        os.path.join(HERE, "csrc", "ingest_host.inl"),
        os.path.join(HERE, "csrc", "comm_host.inl"),
        os.path.join(HERE, "csrc", "model_kernels.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "imcoalhmm_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl", "--split-compile", "0"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + os.environ.get("IMC_EXTRA_NVCC_FLAGS", "").split() +
           (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC])
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
