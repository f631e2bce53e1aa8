"""Loader of the CUDA library.  There is deliberately no fallback: if the .so is missing
or no GPU is visible, compute entry points raise."""
import ctypes
import os

from . import _abi

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdxvae_b200.so")
_lib = None


def lib_path():
    return _PATH


def lib():
    """The bound ctypes library (host-only calls — parameter table, host batcher — work
    without a GPU; anything that launches kernels needs one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(
                "dxvae_b200: CUDA extension %s is not built (run `python -m dxvae_b200.build`); "
                "there is no CPU fallback" % _PATH)
        _lib = _abi.bind(ctypes.CDLL(_PATH))
        if _lib.dxvae_abi_version() != 4:
            raise RuntimeError("dxvae_b200: ABI version mismatch")
    return _lib


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("dxvae_b200: no CUDA device visible; the hot path runs only on the GPU (no CPU fallback)")
    return lib()


def check(rc, what):
    _abi.check(lib(), rc, what)


def launch_count():
    return int(lib().dxvae_launch_count())
