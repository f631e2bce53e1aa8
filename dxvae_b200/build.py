"""Builds dxvae_b200/libdxvae_b200.so from csrc/ with nvcc for sm_100a (in-tree, so the
.so travels with the repo snapshot to the GPU box).  nvcc cross-compiles without a GPU.
Translation units are compiled in parallel into build/obj (one object per .cu, rebuilt when
the source or any header is newer) and linked into the shared library."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdxvae_b200.so")
OBJ = os.path.join(os.path.dirname(HERE), "build", "obj")
FILES = ["dx_gemm.cu", "dx_tc_gemm.cu", "dx_encoder.cu", "dx_decoder.cu", "dx_data.cu", "dx_api.cu", "dx_small.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "--extended-lambda", "-Xcompiler", "-fPIC"] + os.environ.get("DX_NVCC_EXTRA", "").split()   # (experiments: extra -D switches)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA extension")


def _headers():
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "dxvae_b200.h"))
    return deps


def _sources():
    return [f for f in FILES if os.path.exists(os.path.join(SRC, f))]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    objs = []
    for f in _sources():
        src = os.path.join(SRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for rc in ex.map(lambda c: subprocess.call(c), jobs):
            if rc != 0:
                raise RuntimeError("nvcc failed (exit %d)" % rc)
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", OUT + ".tmp"])
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force="--incremental" not in sys.argv, verbose="-v" in sys.argv))
