"""TEST INFRASTRUCTURE -- clean-room stand-in for the slice of the DGL 0.7 API that the
reference's hot path touches, so the *unmodified* reference module
(``/root/reference/src/model/rigid_docking_model.py``) can be imported and run in a container
where DGL (``requirements.txt:5``, not vendored, not installable offline) is absent.

Only ``oracle/`` scripts and ``tests/`` put this directory on ``sys.path``.  The product
package never imports it.  Semantics implemented (all trivially defined by DGL's docs and pinned
end to end by the reference's shipped golden PDB outputs, see ``oracle/make_golden.py``):

* ``dgl.graph(([], []), idtype=)`` + ``add_nodes`` / ``add_edges`` / ``ndata`` / ``edata``
  (``src/utils/protein_utils.py:331-395``)
* ``dgl.heterograph({canonical_etype: (src, dst)}, num_nodes_dict=)`` (``train_utils.py:61-70``)
* ``dgl.batch`` / ``dgl.unbatch`` / ``batch_num_nodes`` (``train_utils.py:98``,
  ``rigid_docking_model.py:244, 512, 653``)
* ``local_scope`` / ``apply_edges`` (builtin ``u_sub_v`` and UDFs with ``edges.src/dst``) /
  ``update_all(copy_edge, mean)`` with zero in-degree -> 0 (``rigid_docking_model.py:193-283``)
"""
import contextlib

import torch

from . import function  # noqa: F401
from . import backend  # noqa: F401

__version__ = '0.7.0-shim'


class _Frame:
    def __init__(self, store):
        self.data = store


class _NodeView:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _Frame(self._g._ndata[ntype])


class _EdgeView:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _Frame(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, etype=None, form='uv'):
        et = self._g.to_canonical_etype(etype)
        return self._g._edges[et]


class _EdgeBatch:
    """Argument of an ``apply_edges`` UDF: exposes ``.src``, ``.dst``, ``.data`` feature dicts."""

    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _Gathered:
    def __init__(self, store, index):
        self._store, self._index = store, index

    def __getitem__(self, key):
        return self._store[key][self._index]


class DGLHeteroGraph:
    def __init__(self, edges, num_nodes, idtype=torch.int32):
        self._ntypes = list(num_nodes.keys())
        self._etypes = list(edges.keys())
        self._num_nodes = dict(num_nodes)
        self._edges = {et: (torch.as_tensor(s).to(idtype), torch.as_tensor(d).to(idtype))
                       for et, (s, d) in edges.items()}
        self._ndata = {nt: {} for nt in self._ntypes}
        self._edata = {et: {} for et in self._etypes}
        self._batch_num_nodes = None
        self._batch_num_edges = None
        self.idtype = idtype
        self.nodes = _NodeView(self)
        self.edges = _EdgeView(self)

    # -- types -------------------------------------------------------------------------------
    @property
    def ntypes(self):
        return list(self._ntypes)

    @property
    def canonical_etypes(self):
        return list(self._etypes)

    def to_canonical_etype(self, etype):
        if etype is None:
            assert len(self._etypes) == 1, 'etype required on a heterograph'
            return self._etypes[0]
        if isinstance(etype, tuple):
            return etype
        cands = [et for et in self._etypes if et[1] == etype]
        assert len(cands) == 1, f'ambiguous edge type {etype}'
        return cands[0]

    def _ntype(self, ntype):
        if ntype is None:
            assert len(self._ntypes) == 1
            return self._ntypes[0]
        return ntype

    # -- sizes -------------------------------------------------------------------------------
    def num_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None):
        if etype is None:
            return sum(int(s.shape[0]) for s, _ in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    number_of_edges = num_edges

    def batch_num_nodes(self, ntype=None):
        nt = self._ntype(ntype)
        if self._batch_num_nodes is None:
            return torch.tensor([self._num_nodes[nt]], dtype=torch.int64)
        return self._batch_num_nodes[nt]

    def batch_num_edges(self, etype=None):
        et = self.to_canonical_etype(etype)
        if self._batch_num_edges is None:
            return torch.tensor([self.num_edges(et)], dtype=torch.int64)
        return self._batch_num_edges[et]

    @property
    def batch_size(self):
        return int(self.batch_num_nodes(self._ntypes[0]).shape[0])

    # -- homogeneous-graph conveniences ------------------------------------------------------
    @property
    def ndata(self):
        assert len(self._ntypes) == 1
        return self._ndata[self._ntypes[0]]

    @property
    def edata(self):
        assert len(self._etypes) == 1
        return self._edata[self._etypes[0]]

    def add_nodes(self, num, ntype=None):
        nt = self._ntype(ntype)
        assert not self._ndata[nt], 'shim: add_nodes after features were set is not supported'
        self._num_nodes[nt] += int(num)

    def add_edges(self, u, v, etype=None):
        et = self.to_canonical_etype(etype)
        assert not self._edata[et], 'shim: add_edges after features were set is not supported'
        s, d = self._edges[et]
        self._edges[et] = (torch.cat([s, torch.as_tensor(u).to(self.idtype)]),
                           torch.cat([d, torch.as_tensor(v).to(self.idtype)]))

    # -- device ------------------------------------------------------------------------------
    @property
    def device(self):
        return self._edges[self._etypes[0]][0].device

    def to(self, device, **kwargs):
        g = DGLHeteroGraph({et: (s.to(device), d.to(device)) for et, (s, d) in self._edges.items()},
                           self._num_nodes, self.idtype)
        g._ndata = {nt: {k: v.to(device) for k, v in fr.items()} for nt, fr in self._ndata.items()}
        g._edata = {et: {k: v.to(device) for k, v in fr.items()} for et, fr in self._edata.items()}
        if self._batch_num_nodes is not None:
            g._batch_num_nodes = {k: v.to(device) for k, v in self._batch_num_nodes.items()}
            g._batch_num_edges = {k: v.to(device) for k, v in self._batch_num_edges.items()}
        return g

    # -- message passing ---------------------------------------------------------------------
    @contextlib.contextmanager
    def local_scope(self):
        saved_n = {nt: dict(fr) for nt, fr in self._ndata.items()}
        saved_e = {et: dict(fr) for et, fr in self._edata.items()}
        try:
            yield
        finally:
            for nt in self._ntypes:
                self._ndata[nt].clear()
                self._ndata[nt].update(saved_n[nt])
            for et in self._etypes:
                self._edata[et].clear()
                self._edata[et].update(saved_e[et])

    def apply_edges(self, func, edges=None, etype=None):
        et = self.to_canonical_etype(etype)
        src_t, _, dst_t = et
        s, d = self._edges[et]
        s, d = s.long(), d.long()
        if isinstance(func, function.BinaryMessage):
            a = self._ndata[src_t][func.lhs_field][s] if func.lhs == 'u' else self._ndata[dst_t][func.lhs_field][d]
            b = self._ndata[dst_t][func.rhs_field][d] if func.rhs == 'v' else self._ndata[src_t][func.rhs_field][s]
            self._edata[et][func.out] = func.op(a, b)
            return
        out = func(_EdgeBatch(_Gathered(self._ndata[src_t], s), _Gathered(self._ndata[dst_t], d), self._edata[et]))
        self._edata[et].update(out)

    def update_all(self, message_func, reduce_func, apply_node_func=None, etype=None):
        assert apply_node_func is None
        et = self.to_canonical_etype(etype)
        _, _, dst_t = et
        _, d = self._edges[et]
        d = d.long()
        assert isinstance(message_func, function.CopyEdge) and isinstance(reduce_func, function.Reduce)
        assert message_func.out == reduce_func.msg
        m = self._edata[et][message_func.field]
        n = self._num_nodes[dst_t]
        acc = torch.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        acc = acc.index_add(0, d, m)
        if reduce_func.kind == 'mean':
            deg = torch.zeros(n, dtype=m.dtype, device=m.device).index_add(
                0, d, torch.ones(d.shape[0], dtype=m.dtype, device=m.device))
            deg = deg.clamp(min=1).view((n,) + (1,) * (m.dim() - 1))
            acc = acc / deg  # zero in-degree rows stay 0 (DGL semantics)
        else:
            assert reduce_func.kind == 'sum'
        self._ndata[dst_t][reduce_func.out] = acc

    def __repr__(self):
        return f'DGLHeteroGraph-shim(num_nodes={self._num_nodes}, num_edges={ {et: self.num_edges(et) for et in self._etypes} })'


def graph(data, ntype=None, etype=None, num_nodes=None, idtype=torch.int32, device=None, **kwargs):
    src, dst = data
    src = torch.as_tensor(src, dtype=idtype) if not torch.is_tensor(src) else src.to(idtype)
    dst = torch.as_tensor(dst, dtype=idtype) if not torch.is_tensor(dst) else dst.to(idtype)
    if num_nodes is None:
        num_nodes = int(max(src.max().item(), dst.max().item()) + 1) if src.numel() else 0
    return DGLHeteroGraph({('_N', '_E', '_N'): (src, dst)}, {'_N': num_nodes}, idtype)


def heterograph(data_dict, num_nodes_dict=None, idtype=None, device=None):
    first = next(iter(data_dict.values()))
    it = idtype or (first[0].dtype if torch.is_tensor(first[0]) else torch.int32)
    return DGLHeteroGraph(dict(data_dict), dict(num_nodes_dict), it)


def batch(graphs, ndata=None, edata=None):
    graphs = list(graphs)
    g0 = graphs[0]
    num_nodes = {nt: sum(g._num_nodes[nt] for g in graphs) for nt in g0._ntypes}
    edges = {}
    for et in g0._etypes:
        st, _, dt = et
        so = do = 0
        ss, dd = [], []
        for g in graphs:
            s, d = g._edges[et]
            ss.append(s + so)
            dd.append(d + do)
            so += g._num_nodes[st]
            do += g._num_nodes[dt]
        edges[et] = (torch.cat(ss), torch.cat(dd))
    out = DGLHeteroGraph(edges, num_nodes, g0.idtype)
    for nt in g0._ntypes:
        for k in g0._ndata[nt]:
            out._ndata[nt][k] = torch.cat([g._ndata[nt][k] for g in graphs], dim=0)
    for et in g0._etypes:
        for k in g0._edata[et]:
            out._edata[et][k] = torch.cat([g._edata[et][k] for g in graphs], dim=0)
    dev = g0.device
    out._batch_num_nodes = {nt: torch.cat([g.batch_num_nodes(nt) for g in graphs]).to(dev) for nt in g0._ntypes}
    out._batch_num_edges = {et: torch.cat([g.batch_num_edges(et) for g in graphs]).to(dev) for et in g0._etypes}
    return out


def unbatch(g, node_split=None, edge_split=None):
    bnn = {nt: g.batch_num_nodes(nt).tolist() for nt in g._ntypes}
    bne = {et: g.batch_num_edges(et).tolist() for et in g._etypes}
    n_off = {nt: 0 for nt in g._ntypes}
    e_off = {et: 0 for et in g._etypes}
    out = []
    for b in range(len(bnn[g._ntypes[0]])):
        edges = {}
        for et in g._etypes:
            st, _, dt = et
            s, d = g._edges[et]
            lo, hi = e_off[et], e_off[et] + bne[et][b]
            edges[et] = (s[lo:hi] - n_off[st], d[lo:hi] - n_off[dt])
        one = DGLHeteroGraph(edges, {nt: bnn[nt][b] for nt in g._ntypes}, g.idtype)
        for nt in g._ntypes:
            lo, hi = n_off[nt], n_off[nt] + bnn[nt][b]
            one._ndata[nt] = {k: v[lo:hi] for k, v in g._ndata[nt].items()}
        for et in g._etypes:
            lo, hi = e_off[et], e_off[et] + bne[et][b]
            one._edata[et] = {k: v[lo:hi] for k, v in g._edata[et].items()}
        for nt in g._ntypes:
            n_off[nt] += bnn[nt][b]
        for et in g._etypes:
            e_off[et] += bne[et][b]
        out.append(one)
    return out


def save_graphs(filename, g_list, labels=None):
    raise NotImplementedError('dgl shim: save_graphs is outside the hot path')


def load_graphs(filename, idx_list=None):
    raise NotImplementedError('dgl shim: load_graphs is outside the hot path')
