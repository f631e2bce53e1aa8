"""Builds dxvae_b200/libdxvae_b200.so from csrc/ with nvcc for sm_100a (in-tree, so the
.so travels with the repo snapshot to the GPU box).  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdxvae_b200.so")
FILES = ["dx_gemm.cu", "dx_tc_gemm.cu", "dx_encoder.cu", "dx_decoder.cu", "dx_data.cu", "dx_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA extension")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".h", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "dxvae_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(SRC, f) for f in FILES] + ["-o", OUT + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
