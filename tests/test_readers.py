"""File readers of dxvae_b200.dxdata that stand in for mido / dgl (dxdata.py:314-338)."""
import os

import numpy as np
import pytest
import torch

from tests import util

REF_BIN = "/root/reference/DX_data/DXDataset.bin"


def test_read_syx_matches_golden_voices(tmp_path):
    from dxvae_b200.dxdata import read_syx
    v = read_syx(os.path.join(util.GOLDEN, "gen_patch.syx"))
    assert v.shape == (32, 128) and v.dtype == np.uint8
    raw = open(os.path.join(util.GOLDEN, "gen_patch.syx"), "rb").read()
    assert v.tobytes() == raw[6:6 + 4096]
    bad = tmp_path / "bad.syx"
    bad.write_bytes(b"\x00" * 100)
    with pytest.raises(ValueError):
        read_syx(str(bad))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference data set only exists in the build container")
def test_read_dgl_bin_reproduces_the_reference_dataset():
    """The DGL save_graphs container is parsed without DGL; content = golden voices through make_graph."""
    import dxvae_oracle as O
    from dxvae_b200.dxdata import read_dgl_bin
    G = read_dgl_bin(REF_BIN)
    assert len(G) == 1024
    v = util.voices()
    for i in (0, 1, 77, 511, 1023):
        X, P, s, d = O.make_graph(v["voices"][i])
        assert torch.equal(G[i].ndata["X"], X) and torch.equal(G[i].ndata["params"], P)
        es, ed = G[i].edges()
        assert (es.tolist(), ed.tolist()) == (s, d)


def test_graph_objects_expose_the_dgl_surface_the_reference_uses():
    from dxvae_b200.dxdata import DXGraph, DXGraphBatch, edges_from_mask, mask_from_edges
    import dxvae_oracle as O
    g = DXGraph(torch.zeros(7, 27), torch.zeros(7, 21), [1, 2, 2, 6], [0, 1, 2, 6])
    assert g.successors(2).tolist() == [1, 2] and g.predecessors(0).tolist() == [1]
    assert g.adj().to_dense()[2, 1] == 1 and g.num_nodes() == 7 and g.num_edges() == 4
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.random((7, 7)) < 0.4
        A[0, 0] = False
        s, d = O.edges_from_adj(A.tolist())
        assert edges_from_mask(mask_from_edges(s, d)) == (s, d)        # reference insertion order
    gb = DXGraphBatch.from_graphs([g, g])
    assert len(gb) == 2 and gb[1].edges()[0].tolist() == [1, 2, 2, 6] and len(gb[0:1]) == 1
