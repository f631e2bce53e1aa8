"""dxvae_b200 — B200-native DX-VAE hot path (encode / decode / ELBO train step).

Python host code + a C-ABI CUDA library (dxvae_b200/libdxvae_b200.so, built from
csrc/ for sm_100a).  The library is loaded on first use and there is no CPU fallback:
without it (or without a GPU) every compute entry point raises.
"""
__all__ = ["DXVAE", "DXGraph", "DXGraphBatch", "DXDataset", "graph_to_syx"]


def __getattr__(name):  # lazy: importing the package must not need torch.cuda
    if name == "DXVAE":
        from .model import DXVAE
        return DXVAE
    if name in ("DXGraph", "DXGraphBatch", "DXDataset", "graph_to_syx"):
        from . import dxdata
        return getattr(dxdata, name)
    raise AttributeError(name)
