"""Flat parameter blob <-> state_dict (53 tensors, model.py:24-72 registration order).

The native library reads weights (and accumulates gradients) as ONE fp32 blob whose
tensor offsets come from dxvae_param_entry(); every tensor starts 256-byte aligned.
"""
import ctypes as C

import numpy as np

from ._abi import ParamEntry, check


def param_table(lib):
    """[(name, offset, shape)] in state_dict order."""
    out = []
    e = ParamEntry()
    k = 0
    while True:
        if k >= 53:
            break
        check(lib, lib.dxvae_param_entry(k, C.byref(e)), "dxvae_param_entry")
        shape = (e.rows, e.cols) if e.cols else (e.rows,)
        out.append((e.name.decode(), int(e.offset), shape))
        k += 1
    return out


def flatten_state_dict(sd, table, total):
    """state_dict (torch tensors or arrays) -> np.float32 blob of `total` floats."""
    blob = np.zeros(total, np.float32)
    for name, off, shape in table:
        t = sd[name]
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
        n = int(np.prod(shape))
        blob[off:off + n] = a.reshape(-1)
    return blob


def unflatten(blob, table):
    """blob -> {name: view} (numpy or torch, whatever `blob` is)."""
    out = {}
    for name, off, shape in table:
        n = int(np.prod(shape))
        out[name] = blob[off:off + n].reshape(shape)
    return out
