"""Graph format of the reference's dxdata.py, without DGL / mido.

Mirrors the surface the reference exposes (dxdata.py:77-397, main.py:6-9):
  * graph objects with  ndata['X'] (7,27) f32,  ndata['params'] (7,21) f32,  edges() ->
    (src, dst) int64 — `DXGraph` is a duck-typed stand-in for the DGLGraph; real DGL
    graphs are accepted wherever a graph is expected, nothing here imports dgl;
  * `DXDataset(raw_dir, save_dir)`: .syx banks -> graphs (`_make_graph`, dxdata.py:174-312,
    runs as the CUDA kernel dxvae_voices_to_graphs), or the cached DGL `.bin` when present
    (dxdata.py:334-338), including the reference's quirk that a cached load stores the
    `(graphs, labels)` tuple so `dataset[0]` is the whole list (main.py:55);
  * `graph_to_syx(G, file)` (dxdata.py:341-397): voice packing runs as dxvae_pack_syx.
`DXGraphBatch` is the batched (structure-of-arrays) form the kernels consume; it behaves
like a list of DXGraph.
"""
import os
import struct
from pathlib import Path

import operator
import weakref

import numpy as np
import torch

from . import _lib

N_NODES, N_PARAMS, SIZE_X = 7, 21, 27
_SYX_HEAD = bytes([0xF0, 67, 0, 9, 32, 0])      # dxdata.py:343 + sysex start
_SYX_TAIL = bytes([88, 0xF7])                   # dxdata.py:344 (constant "checksum") + sysex end


def edges_from_mask(mask):
    """Edge list of a decoded graph in the reference's insertion order (model.py:237-250):
    for vi=1..6: (vi,vi)?, then for vj=vi-1..0: (vj,vi)?, (vi,vj)?.  Any edge not reachable by
    that walk (a self-loop on node 0) is appended last."""
    m = int(mask)
    src, dst = [], []
    bit = lambda s, d: (m >> (s * 7 + d)) & 1
    for vi in range(1, N_NODES):
        if bit(vi, vi):
            src.append(vi); dst.append(vi)
        for vj in range(vi - 1, -1, -1):
            if bit(vj, vi):
                src.append(vj); dst.append(vi)
            if bit(vi, vj):
                src.append(vi); dst.append(vj)
    if bit(0, 0):
        src.append(0); dst.append(0)
    return src, dst


def mask_from_edges(src, dst):
    m = 0
    for s, d in zip(src, dst):
        m |= 1 << (int(s) * 7 + int(d))
    return m


class _Adj:
    def __init__(self, g):
        self._g = g

    def to_dense(self):
        a = torch.zeros(N_NODES, N_NODES)
        for s, d in zip(*self._g._edges):
            a[s, d] = 1.0
        return a


class _ViewData(dict):
    """ndata of a graph that is a row view of a DXGraphBatch: replacing or removing an entry detaches the graph from its
    batch (the batcher's fast path then treats it as a foreign graph), so the fast path needs no per-graph comparison."""
    __slots__ = ("_g",)

    def _detach(self):
        self._g._oid = None

    def __setitem__(self, k, v):
        self._detach(); dict.__setitem__(self, k, v)

    def __delitem__(self, k):
        self._detach(); dict.__delitem__(self, k)

    def pop(self, *a):
        self._detach(); return dict.pop(self, *a)

    def popitem(self):
        self._detach(); return dict.popitem(self)

    def update(self, *a, **kw):
        self._detach(); dict.update(self, *a, **kw)

    def clear(self):
        self._detach(); dict.clear(self)

    def setdefault(self, *a):
        self._detach(); return dict.setdefault(self, *a)


class DXGraph:
    """One 7-node patch graph: the subset of the DGLGraph API the reference touches."""

    def __init__(self, X, params, src, dst):
        self.ndata = {"X": X, "params": params}
        self._edges = ([int(s) for s in src], [int(d) for d in dst])
        self._oid = None        # (batch token << 32 | row) while this graph is an untouched row view of a DXGraphBatch

    def edges(self):
        return (torch.tensor(self._edges[0], dtype=torch.int64), torch.tensor(self._edges[1], dtype=torch.int64))

    def num_nodes(self):
        return N_NODES

    def num_edges(self):
        return len(self._edges[0])

    def to(self, device):
        return self

    def adj(self):
        return _Adj(self)

    def predecessors(self, v):
        return torch.tensor([s for s, d in zip(*self._edges) if d == v], dtype=torch.int64)

    def successors(self, v):
        return torch.tensor([d for s, d in zip(*self._edges) if s == v], dtype=torch.int64)


def _graph_edges(g):
    if isinstance(g, DXGraph):
        return g._edges
    s, d = g.edges()
    return s.tolist(), d.tolist()


_GET_OID = operator.attrgetter("_oid")
_BATCHES = weakref.WeakValueDictionary()      # batch token -> DXGraphBatch that handed out row views
_next_token = [1]


class _LazyEdgeLists:
    """Edge lists of a gathered sub-batch, materialised on first use (only the host batcher and graph views need them)."""

    def __init__(self, base, idx):
        self._base, self._idx, self._lists = base, idx, None

    def _get(self):
        if self._lists is None:
            b = self._base
            self._lists = [b[int(i)] for i in self._idx]
        return self._lists

    def __len__(self):
        return len(self._idx)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self._get()[i]
        return self._base[int(self._idx[i])]

    def __iter__(self):
        return iter(self._get())


class DXGraphBatch:
    """B graphs as three arrays: X (B,7,27) f32, params (B,7,21) f32, adj (B,) int64 masks
    (bit src*7+dst).  Indexing / iteration yields DXGraph views, so it can be passed anywhere
    the reference takes a list of graphs (print_data, graph_to_syx, DXVAE.encode ...)."""

    def __init__(self, X, params, adj, edge_lists=None):
        self.X, self.params, self.adj = X, params, adj
        self._edge_lists = edge_lists          # original insertion order when known

    @classmethod
    def from_graphs(cls, graphs, staging=False):
        """staging=True (the batcher's own call, DXVAE._prepare): rows drawn from a PINNED host batch come back as an
        IndexedBatch — the parent plus an index list — which the device gathers straight out of the pinned memory
        (dxvae_pack_graphs_indexed); nothing is gathered or staged on the host."""
        if isinstance(graphs, DXGraphBatch):
            return graphs
        graphs = list(graphs)
        if not graphs:
            raise ValueError("empty batch")
        fast = cls._from_views(graphs, staging)
        if fast is not None:
            return fast
        X = torch.stack([g.ndata["X"].detach().to("cpu", torch.float32) for g in graphs])
        P = torch.stack([g.ndata["params"].detach().to("cpu", torch.float32) for g in graphs])
        edges = [_graph_edges(g) for g in graphs]
        adj = torch.tensor([_to_i64(mask_from_edges(*e)) for e in edges], dtype=torch.int64)
        return cls(X, P, adj, edges)

    @classmethod
    def _from_views(cls, graphs, staging=False):
        """Batcher fast path: graphs that are untouched row views of ONE DXGraphBatch (what iterating / indexing a batch or
        a DXDataset hands out) are re-batched by index — one gather (a plain slice when the rows are consecutive) instead
        of stacking thousands of (7,27) tensors.  Returns None when any graph is foreign or was modified (a view whose
        ndata was touched has detached itself, see _ViewData).  With staging=True rows of a pinned host batch are not
        gathered at all: the result is an IndexedBatch (parent + index list) for the device to gather."""
        try:                                                    # ONE pass over the (scattered) graph objects, at C speed: this
            oid = np.fromiter(map(_GET_OID, graphs), np.int64, count=len(graphs))   # runs once per graph of every training batch
        except (TypeError, AttributeError):                     # a foreign graph object or a detached view (_oid None)
            return None
        tok = oid >> 32
        if not bool(np.all(tok == tok[0])):
            return None
        base = _BATCHES.get(int(tok[0]))
        if base is None:
            return None
        idx_np = oid & 0xFFFFFFFF
        owners = graphs
        n = len(owners)
        el = base._edge_lists
        if n == 1 or bool(np.all(np.diff(idx_np) == 1)):
            sl = slice(int(idx_np[0]), int(idx_np[0]) + n)
            return cls(base.X[sl], base.params[sl], base.adj[sl], el[sl] if el is not None else None)
        idx = idx_np
        ii = torch.from_numpy(idx_np)
        if base.X.is_cuda:
            ii = ii.to(base.X.device)
        sub_el = _LazyEdgeLists(el, idx) if el is not None else None
        pinned = (not base.X.is_cuda) and base.X.is_pinned()
        if pinned and staging:
            return IndexedBatch(base, idx_np, sub_el)
        take = (lambda t: t.index_select(0, ii).pin_memory()) if pinned else (lambda t: t.index_select(0, ii))
        return cls(take(base.X), take(base.params), take(base.adj), sub_el)

    def __len__(self):
        return self.X.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            el = self._edge_lists[i] if self._edge_lists is not None else None
            return DXGraphBatch(self.X[i], self.params[i], self.adj[i], el)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._view(i, int(self.adj[i]) if self._edge_lists is None else 0)

    def _view(self, i, mask):
        e = self._edge_lists[i] if self._edge_lists is not None else edges_from_mask(mask & (2 ** 49 - 1))
        X, P = self.X[i], self.params[i]
        if X.is_cuda:
            return DXGraph(X.cpu(), P.cpu(), e[0], e[1])
        g = DXGraph.__new__(DXGraph)
        nd = _ViewData(); dict.__setitem__(nd, "X", X); dict.__setitem__(nd, "params", P); nd._g = g
        g.ndata = nd
        g._edges = (list(e[0]), list(e[1]))
        tok = self.__dict__.get("_token")
        if tok is None:
            tok = self._token = _next_token[0]
            _next_token[0] += 1
            _BATCHES[tok] = self
        g._oid = (tok << 32) | i
        g._ob = self                      # (keeps the batch alive as long as one of its views is)
        return g

    def __iter__(self):
        masks = self.adj.cpu().tolist() if self._edge_lists is None else None
        for i in range(len(self)):
            yield self._view(i, masks[i] if masks is not None else 0)

    def pin_memory(self):
        return DXGraphBatch(self.X.pin_memory(), self.params.pin_memory(), self.adj.pin_memory(), self._edge_lists)

    def edge_lists(self):
        if self._edge_lists is not None:
            return self._edge_lists
        return [edges_from_mask(int(m)) for m in self.adj.cpu().tolist()]

    def cpu(self):
        return DXGraphBatch(self.X.cpu(), self.params.cpu(), self.adj.cpu(), self._edge_lists)


class IndexedBatch(DXGraphBatch):
    """Rows `index` of a pinned host DXGraphBatch, not gathered: what the batcher hands the device when a training loop
    draws a shuffled batch from its dataset.  X / params / adj materialise (host gather) only if somebody asks for them."""

    def __init__(self, base, index, edge_lists=None):
        self.base, self.index = base, index
        self._edge_lists = edge_lists
        self._mat = None

    def _m(self):
        if self._mat is None:
            ii = torch.from_numpy(self.index)
            self._mat = (self.base.X.index_select(0, ii), self.base.params.index_select(0, ii), self.base.adj.index_select(0, ii))
        return self._mat

    X = property(lambda self: self._m()[0])
    params = property(lambda self: self._m()[1])
    adj = property(lambda self: self._m()[2])

    def __len__(self):
        return len(self.index)

    def materialise(self):
        return DXGraphBatch(self.X, self.params, self.adj, self._edge_lists)


def _to_i64(m):
    return m - (1 << 64) if m >= (1 << 63) else m


# ---------------------------------------------------------------------------------------
# file formats
# ---------------------------------------------------------------------------------------
def read_syx(file):
    """dxdata.py:314-318: one 32-voice bulk dump -> (32,128) uint8 packed voices."""
    raw = Path(file).read_bytes()
    if len(raw) < 4104 or raw[0] != 0xF0:
        raise ValueError("%s is not a DX7 32-voice bulk dump" % file)
    end = raw.index(0xF7)
    data = raw[1:end]                       # what mido hands out as msg.data
    return np.frombuffer(data[5:-1], np.uint8).reshape(32, -1).copy()


_ND_MAGIC = struct.pack("<Q", 0xDD5E40F096B4A13F)


def read_dgl_bin(path):
    """Minimal reader of the DGL save_graphs v2 container found at DX_data/DXDataset.bin:
    every graph owns 11 NDArray blobs; #6/#7 are int64 src/dst, #9 is X, #10 is params."""
    buf = Path(path).read_bytes()
    blobs = []
    pos = buf.find(_ND_MAGIC)
    while pos >= 0:
        p = pos + 16
        ndim = struct.unpack_from("<i", buf, p + 8)[0]
        code, bits = struct.unpack_from("<BB", buf, p + 12)
        shape = struct.unpack_from("<%dq" % ndim, buf, p + 16)
        q = p + 16 + 8 * ndim
        nbytes = struct.unpack_from("<q", buf, q)[0]
        q += 8
        dt = {(0, 64): np.int64, (2, 32): np.float32, (0, 32): np.int32}[(code, bits)]
        blobs.append(np.frombuffer(buf, dt, count=nbytes // np.dtype(dt).itemsize, offset=q).reshape(shape))
        pos = buf.find(_ND_MAGIC, q + nbytes)
    if len(blobs) % 11:
        raise ValueError("unexpected DGL .bin layout in %s" % path)
    graphs = []
    for i in range(0, len(blobs), 11):
        graphs.append(DXGraph(torch.from_numpy(blobs[i + 9].copy()), torch.from_numpy(blobs[i + 10].copy()),
                              blobs[i + 6].tolist(), blobs[i + 7].tolist()))
    return graphs


def voices_to_batch(voices, device="cuda"):
    """dxdata.py:174-312 for many voices at once on the GPU: (n,128) uint8 -> DXGraphBatch
    (tensors on `device`)."""
    L = _lib.require_cuda()
    if torch.is_tensor(voices):                      # (a pinned host tensor uploads without a staging copy)
        v = voices.to(device, torch.uint8, non_blocking=True).contiguous()
    else:
        v = torch.as_tensor(np.array(voices, np.uint8, copy=True)).to(device)
    n = v.shape[0]
    Xg = torch.empty(n, N_NODES, SIZE_X, device=device)
    Pg = torch.empty(n, N_NODES, N_PARAMS, device=device)
    adj = torch.empty(n, dtype=torch.int64, device=device)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.dxvae_voices_to_graphs(n, v.data_ptr(), None, None, adj.data_ptr(), Xg.data_ptr(), Pg.data_ptr(), st),
               "dxvae_voices_to_graphs")
    return DXGraphBatch(Xg, Pg, adj)


class DXDataset:
    """Drop-in for dxdata.py:77-338 (`DXDataset(raw_dir='DX_data')`)."""

    name = "DXDataset.bin"

    def __init__(self, raw_dir=None, save_dir=None):
        self._raw_dir = raw_dir
        self._save_dir = save_dir if save_dir is not None else raw_dir
        self.save_path = os.path.join(self._save_dir, self.name)
        if self.has_cache():
            self.load()
        else:
            self.process()

    def has_cache(self):
        return os.path.exists(self.save_path)

    def load(self):
        # dxdata.py:334-335 keeps dgl.load_graphs' (graphs, labels) tuple; main.py:55 relies on it
        self.graphs = (read_dgl_bin(self.save_path), {})

    def process(self):
        files = list(Path(self._raw_dir).rglob("*.syx"))
        voices = np.concatenate([read_syx(f) for f in files])
        self.graphs = list(voices_to_batch(voices).cpu())
        # the packed algorithm byte keys DX_ALGO un-modded (dxdata.py:308): keep original edge order
        from .algo import DX_ALGO
        for g, v in zip(self.graphs, voices):
            g._edges = (list(DX_ALGO[int(v[110])][0]), list(DX_ALGO[int(v[110])][1]))

    def __getitem__(self, idx):
        return self.graphs[idx]

    def __len__(self):
        return len(self.graphs)


def graph_to_syx(G, file="gen_patch.syx"):
    """dxdata.py:341-397: write the graphs' params as a DX7 bulk dump (no 32-voice chunking,
    constant trailer byte 88 — as the reference does)."""
    data = graph_to_syx_bytes(G)
    Path(file).write_bytes(data)


def graph_to_syx_bytes(G):
    L = _lib.require_cuda()
    if isinstance(G, DXGraphBatch):
        P = G.params
    else:
        P = torch.stack([g.ndata["params"] for g in G])
    P = P.to("cuda", torch.float32).contiguous()
    n = P.shape[0]
    out = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.dxvae_pack_syx(n, P.data_ptr(), out.data_ptr(), st), "dxvae_pack_syx")
    return _SYX_HEAD + out.cpu().numpy().tobytes() + _SYX_TAIL
