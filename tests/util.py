"""Shared helpers for the test-suite (fixtures -> tensors)."""
import os

import numpy as np
import torch

import dxvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def voices():
    if "v" not in _cache:
        _cache["v"] = np.load(os.path.join(GOLDEN, "synprez_voices.npz"))
    return _cache["v"]


def dataset_graphs(indices):
    """(X (n,7,27), params (n,7,21), edge lists, A (n,7,7)) of dataset graphs via the oracle's make_graph."""
    v = voices()["voices"]
    Xs, Ps, Es = [], [], []
    for i in indices:
        X, P, s, d = O.make_graph(v[i])
        Xs.append(X); Ps.append(P); Es.append((s, d))
    return torch.stack(Xs), torch.stack(Ps), Es, adj_dense(Es)


def adj_dense(edge_lists):
    A = torch.zeros(len(edge_lists), 7, 7)
    for b, (s, d) in enumerate(edge_lists):
        for x, y in zip(s, d):
            A[b, x, y] = 1
    return A


def adj_from_masks(masks):
    m = np.asarray(masks, np.uint64)
    A = np.zeros((len(m), 7, 7), np.uint8)
    for s in range(7):
        for d in range(7):
            A[:, s, d] = (m >> np.uint64(s * 7 + d)) & np.uint64(1)
    return A


def random_edge_lists(n, p, seed):
    """Arbitrary directed 7-node graphs (decoded graphs can have any of the 49 edges)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        A = rng.random((7, 7)) < p
        s, d = np.nonzero(A)
        out.append((s.tolist(), d.tolist()))
    return out


def alg_of(indices):
    v = voices()["voices"]
    return [int(v[i][110]) for i in indices]


def pick_by_alg(algs):
    """First dataset index of each requested algorithm."""
    v = voices()["voices"]
    out = []
    for a in algs:
        hits = np.nonzero(v[:, 110] == a)[0]
        if len(hits):
            out.append(int(hits[0]))
    return out
