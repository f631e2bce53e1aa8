"""Synthetic DX7 voices for the benchmark configs: every field uniform over
its legal range (dxdata.py:8-74), packed in the 128-byte bulk-dump layout, so that
_make_graph semantics (and hence topology ~ uniform over the 32 DX_ALGO entries) apply."""
import numpy as np


def random_voices(n, seed=0):
    """(n,128) uint8 packed voices, numpy PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    r = lambda hi, size=None: rng.integers(0, hi + 1, size=(n,) if size is None else (n, size), dtype=np.int64)
    v = np.zeros((n, 128), np.int64)
    for op in range(6):                     # bytes 0..101: OP6 .. OP1, 17 bytes each
        o = op * 17
        v[:, o:o + 8] = r(99, 8)            # EG R1-4, L1-4
        v[:, o + 8] = r(99); v[:, o + 9] = r(99); v[:, o + 10] = r(99)      # BP, LD, RD
        v[:, o + 11] = r(3) * 4 + r(3)      # RC | LC
        v[:, o + 12] = r(14) * 8 + r(7)     # DET | RS
        v[:, o + 13] = r(7) * 4 + r(3)      # KVS | AMS
        v[:, o + 14] = r(99)                # OL
        v[:, o + 15] = r(31) * 2 + r(1)     # FC | M
        v[:, o + 16] = r(99)                # FF
    v[:, 102:110] = r(99, 8)                # pitch EG
    v[:, 110] = r(31)                       # ALG
    v[:, 111] = r(1) * 8 + r(7)             # OKS | FB
    v[:, 112] = r(99); v[:, 113] = r(99); v[:, 114] = r(99); v[:, 115] = r(99)   # LFS LFD LPMD LAMD
    v[:, 116] = r(7) * 16 + r(5) * 2 + r(1)  # LPMS | LFW | LKS
    v[:, 117] = r(48)                       # TRNSP
    v[:, 118:128] = np.frombuffer(b"SYNTHETIC ", np.uint8)
    return v.astype(np.uint8)
