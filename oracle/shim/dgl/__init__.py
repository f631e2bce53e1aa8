"""Minimal pure-Python stand-in for the handful of DGL calls the reference makes.

TEST INFRASTRUCTURE ONLY.  `dgl` is not installable in this image (no network),
so the reference's unmodified model.py / dxdata.py are executed against this
stand-in when (a) validating the oracle restatement and (b) generating the
golden vectors under tests/golden/ (see oracle/make_golden.py).  Nothing in the
product package imports this.

Calls covered (everything the reference relies on):
  model.py:167,174,190   g.predecessors(v), g.successors(v)
  model.py:183,273-277   g.ndata[...]
  model.py:219-222       dgl.graph(([], [])), g.to(device), g.add_nodes(1, {...})
  model.py:239,248-250   g.add_edges(u, v)
  model.py:279           g.adj().to_dense()      (row = src, col = dst)
  main.py:9              g.edges()
  dxdata.py:77,172       dgl.data.DGLDataset
  dxdata.py:332,335      dgl.save_graphs / dgl.load_graphs
"""
import struct

import torch

from . import data  # noqa: F401  (dgl.data.DGLDataset)

__all__ = ["graph", "DGLGraph", "load_graphs", "save_graphs", "data"]


class _Adj:
    def __init__(self, g):
        self._g = g

    def to_dense(self):
        n = self._g.num_nodes()
        a = torch.zeros(n, n)
        for s, d in zip(self._g._src, self._g._dst):
            a[s, d] = 1.0
        return a


class DGLGraph:
    def __init__(self, src=(), dst=()):
        self._src = [int(s) for s in src]
        self._dst = [int(d) for d in dst]
        self._n = (max(self._src + self._dst) + 1) if self._src else 0
        self.ndata = _NData(self)

    # -- structure ---------------------------------------------------------
    def num_nodes(self):
        return self._n

    def num_edges(self):
        return len(self._src)

    def to(self, device):
        return self

    def add_nodes(self, num, data=None):
        self._n += num
        if data:
            for k, t in data.items():
                cur = self.ndata._d.get(k)
                self.ndata._d[k] = t.clone() if cur is None else torch.cat([cur, t], 0)

    def add_edges(self, u, v):
        self._src.append(int(u))
        self._dst.append(int(v))

    def predecessors(self, v):
        return torch.tensor([s for s, d in zip(self._src, self._dst) if d == v], dtype=torch.int64)

    def successors(self, v):
        return torch.tensor([d for s, d in zip(self._src, self._dst) if s == v], dtype=torch.int64)

    def edges(self):
        return (torch.tensor(self._src, dtype=torch.int64), torch.tensor(self._dst, dtype=torch.int64))

    def adj(self):
        return _Adj(self)


class _NData:
    def __init__(self, g):
        self._g = g
        self._d = {}

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        if self._g._n == 0:
            self._g._n = v.shape[0]
        self._d[k] = v

    def __contains__(self, k):
        return k in self._d

    def keys(self):
        return self._d.keys()


def graph(data):
    src, dst = data
    if torch.is_tensor(src):
        src = src.tolist()
    if torch.is_tensor(dst):
        dst = dst.tolist()
    return DGLGraph(src, dst)


# ---------------------------------------------------------------------------
# DGL "save_graphs" v2 container, just enough to read DX_data/DXDataset.bin.
# Layout as found in the file: NDArray blobs are tagged
# with the magic below; every graph owns 11 of them, of which #6/#7 are the
# int64 src/dst lists, #9 is X (7,27) f32 and #10 is params (7,21) f32.
# ---------------------------------------------------------------------------
_ND_MAGIC = struct.pack("<Q", 0xDD5E40F096B4A13F)
_DT = {(0, 64): torch.int64, (2, 32): torch.float32, (0, 32): torch.int32, (2, 64): torch.float64}


def _read_ndarrays(buf):
    out = []
    pos = buf.find(_ND_MAGIC)
    while pos >= 0:
        p = pos + 16  # magic + reserved
        _devtype, _devid, ndim = struct.unpack_from("<iii", buf, p)
        p += 12
        code, bits, _lanes = struct.unpack_from("<BBH", buf, p)
        p += 4
        shape = struct.unpack_from("<%dq" % ndim, buf, p)
        p += 8 * ndim
        (nbytes,) = struct.unpack_from("<q", buf, p)
        p += 8
        t = torch.frombuffer(bytearray(buf[p:p + nbytes]), dtype=_DT[(code, bits)]).reshape(shape) \
            if nbytes else torch.zeros(shape, dtype=_DT[(code, bits)])
        out.append(t)
        pos = buf.find(_ND_MAGIC, p + nbytes)
    return out


def load_graphs(path):
    with open(path, "rb") as f:
        buf = f.read()
    arrs = _read_ndarrays(buf)
    assert len(arrs) % 11 == 0, len(arrs)
    graphs = []
    for i in range(0, len(arrs), 11):
        g = DGLGraph(arrs[i + 6].tolist(), arrs[i + 7].tolist())
        g._n = arrs[i + 9].shape[0]
        g.ndata._d["X"] = arrs[i + 9]
        g.ndata._d["params"] = arrs[i + 10]
        graphs.append(g)
    return graphs, {}


def save_graphs(path, graphs):  # the reference tree is read-only: never write
    raise RuntimeError("shim: save_graphs is disabled (reference tree is read-only)")
