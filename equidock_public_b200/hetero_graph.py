"""DGL-free container for the IEGMN input contract (ligand/receptor residue k-NN graphs).

The reference feeds its model a *batched DGL heterograph* built by
``src/utils/train_utils.py:61-83`` (``hetero_graph_from_sg_l_r_pair``) and
``dgl.batch`` (``train_utils.py:98,107``).  DGL is a third-party dependency that is not
vendored by the reference; the hot path only needs the *data contract* of that object:

* node types ``'ligand'`` / ``'receptor'`` with ``res_feat`` (N,1), ``x`` (N,3),
  ``mu_r_norm`` (N,5) and, for the ligand, ``new_x`` (N,3);
* edge types ``('ligand','ll','ligand')`` / ``('receptor','rr','receptor')`` with ``he`` (E,27)
  and two empty ``'cross'`` relations;
* per-pair node / edge counts (``batch_num_nodes`` / ``batch_num_edges``).

``PairGraphBatch`` holds exactly that and answers the subset of the DGL graph API that the
reference's *callers* of the model touch (``.nodes[nt].data``, ``.edges[et].data``,
``.edges(etype=...)``, ``.batch_num_nodes(nt)``, ``.num_nodes(nt)``, ``.to(device)``), so
the engine in ``rigid_docking_model.py`` can take either a real DGL heterograph or this
object.  It implements no message passing: that lives in the CUDA kernels.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch

LIGAND, RECEPTOR = 'ligand', 'receptor'
LL = (LIGAND, 'll', LIGAND)
RR = (RECEPTOR, 'rr', RECEPTOR)
CROSS_RL = (RECEPTOR, 'cross', LIGAND)
CROSS_LR = (LIGAND, 'cross', RECEPTOR)
CANONICAL_ETYPES = (LL, RR, CROSS_RL, CROSS_LR)


class _Frame:
    """``g.nodes['ligand']`` / ``g.edges['ll']`` view: just carries a ``.data`` dict."""

    __slots__ = ('data',)

    def __init__(self, data: Dict[str, torch.Tensor]):
        self.data = data


class _NodeView:
    def __init__(self, graph: 'PairGraphBatch'):
        self._g = graph

    def __getitem__(self, ntype: str) -> _Frame:
        return _Frame(self._g._ndata[ntype])


class _EdgeView:
    """Indexable (``g.edges['ll'].data``) *and* callable (``g.edges(etype='ll')``), like DGL's."""

    def __init__(self, graph: 'PairGraphBatch'):
        self._g = graph

    def __getitem__(self, etype) -> _Frame:
        return _Frame(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, etype=None, form: str = 'uv'):
        src, dst = self._g._edges[self._g.to_canonical_etype(etype)]
        if form == 'uv':
            return src, dst
        raise ValueError("PairGraphBatch.edges: only form='uv' is supported")


class PairGraphBatch:
    """A batch of (ligand graph, receptor graph) pairs with the reference's field names."""

    ntypes = [LIGAND, RECEPTOR]
    canonical_etypes = list(CANONICAL_ETYPES)

    def __init__(self,
                 num_nodes_dict: Dict[str, int],
                 edges: Dict[Tuple[str, str, str], Tuple[torch.Tensor, torch.Tensor]],
                 batch_num_nodes: Dict[str, torch.Tensor] | None = None,
                 batch_num_edges: Dict[Tuple[str, str, str], torch.Tensor] | None = None):
        self._num_nodes = {nt: int(num_nodes_dict[nt]) for nt in self.ntypes}
        self._edges = {}
        for et in CANONICAL_ETYPES:
            if et in edges:
                s, d = edges[et]
                self._edges[et] = (torch.as_tensor(s), torch.as_tensor(d))
            else:
                z = torch.zeros(0, dtype=torch.int32)
                self._edges[et] = (z, z.clone())
        self._ndata: Dict[str, Dict[str, torch.Tensor]] = {nt: {} for nt in self.ntypes}
        self._edata: Dict[Tuple[str, str, str], Dict[str, torch.Tensor]] = {et: {} for et in CANONICAL_ETYPES}
        if batch_num_nodes is None:
            batch_num_nodes = {nt: torch.tensor([self._num_nodes[nt]], dtype=torch.int64) for nt in self.ntypes}
        if batch_num_edges is None:
            batch_num_edges = {et: torch.tensor([self._edges[et][0].shape[0]], dtype=torch.int64)
                               for et in CANONICAL_ETYPES}
        self._batch_num_nodes = {nt: torch.as_tensor(v, dtype=torch.int64) for nt, v in batch_num_nodes.items()}
        self._batch_num_edges = {et: torch.as_tensor(v, dtype=torch.int64) for et, v in batch_num_edges.items()}
        self.nodes = _NodeView(self)
        self.edges = _EdgeView(self)

    # ---- DGL-compatible accessors -------------------------------------------------------
    def to_canonical_etype(self, etype):
        if etype is None:
            raise ValueError('etype must be given for a heterograph')
        if isinstance(etype, tuple):
            return etype
        matches = [et for et in CANONICAL_ETYPES if et[1] == etype]
        if len(matches) != 1:
            raise KeyError(f'edge type {etype!r} is ambiguous or unknown; use the canonical triple')
        return matches[0]

    def num_nodes(self, ntype: str | None = None) -> int:
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    number_of_nodes = num_nodes

    def num_edges(self, etype=None) -> int:
        if etype is None:
            return sum(int(s.shape[0]) for s, _ in self._edges.values())
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    number_of_edges = num_edges

    def batch_num_nodes(self, ntype: str) -> torch.Tensor:
        return self._batch_num_nodes[ntype]

    def batch_num_edges(self, etype) -> torch.Tensor:
        return self._batch_num_edges[self.to_canonical_etype(etype)]

    @property
    def batch_size(self) -> int:
        return int(self._batch_num_nodes[LIGAND].shape[0])

    @property
    def device(self) -> torch.device:
        for frame in self._ndata.values():
            for t in frame.values():
                return t.device
        return self._edges[LL][0].device

    def to(self, device, non_blocking: bool = False) -> 'PairGraphBatch':
        g = PairGraphBatch(self._num_nodes,
                           {et: (s.to(device, non_blocking=non_blocking), d.to(device, non_blocking=non_blocking))
                            for et, (s, d) in self._edges.items()},
                           self._batch_num_nodes, self._batch_num_edges)  # counts stay on the host: no sync to read
        for nt in self.ntypes:
            g._ndata[nt] = {k: v.to(device, non_blocking=non_blocking) for k, v in self._ndata[nt].items()}
        for et in CANONICAL_ETYPES:
            g._edata[et] = {k: v.to(device, non_blocking=non_blocking) for k, v in self._edata[et].items()}
        return g

    def pin_memory(self) -> 'PairGraphBatch':
        """Page-locks every host tensor in place (for asynchronous H2D copies)."""
        self._edges = {et: (s.pin_memory(), d.pin_memory()) for et, (s, d) in self._edges.items()}
        for fr in list(self._ndata.values()) + list(self._edata.values()):
            for k in list(fr.keys()):
                fr[k] = fr[k].pin_memory()
        return self

    def nbytes(self) -> int:
        tot = sum(s.numel() * s.element_size() + d.numel() * d.element_size() for s, d in self._edges.values())
        for fr in list(self._ndata.values()) + list(self._edata.values()):
            tot += sum(v.numel() * v.element_size() for v in fr.values())
        return int(tot)

    def __repr__(self) -> str:
        return (f'PairGraphBatch(pairs={self.batch_size}, ligand_nodes={self._num_nodes[LIGAND]}, '
                f'receptor_nodes={self._num_nodes[RECEPTOR]}, ll_edges={self.num_edges(LL)}, '
                f'rr_edges={self.num_edges(RR)})')


# ---- constructors mirroring the reference's collate path ------------------------------------

def pair_graph(ligand: Dict[str, torch.Tensor], receptor: Dict[str, torch.Tensor]) -> PairGraphBatch:
    """One (ligand, receptor) pair -> un-batched heterograph.

    Mirrors ``hetero_graph_from_sg_l_r_pair`` (``src/utils/train_utils.py:61-83``).  ``ligand`` /
    ``receptor`` are dicts with ``src``, ``dst`` (E,) int32 -- edge e means "src is one of dst's
    nearest neighbours", ``protein_utils.py:339-346`` -- ``he`` (E,27), ``res_feat`` (N,1),
    ``x`` (N,3), ``mu_r_norm`` (N,5) and, for the ligand, ``new_x`` (N,3).
    """
    n_l = int(ligand['x'].shape[0])
    n_r = int(receptor['x'].shape[0])
    g = PairGraphBatch({LIGAND: n_l, RECEPTOR: n_r},
                       {LL: (ligand['src'].to(torch.int32), ligand['dst'].to(torch.int32)),
                        RR: (receptor['src'].to(torch.int32), receptor['dst'].to(torch.int32))})
    for key in ('res_feat', 'x', 'new_x', 'mu_r_norm'):
        g._ndata[LIGAND][key] = ligand[key] if key in ligand else ligand['x']
    for key in ('res_feat', 'x', 'mu_r_norm'):
        g._ndata[RECEPTOR][key] = receptor[key]
    g._edata[LL]['he'] = ligand['he']
    g._edata[RR]['he'] = receptor['he']
    return g


def batch(graphs: Sequence[PairGraphBatch]) -> PairGraphBatch:
    """``dgl.batch`` for this container: concatenate nodes/edges per type, offset the edge ids."""
    graphs = list(graphs)
    if not graphs:
        raise ValueError('batch() of an empty list')
    num_nodes = {nt: sum(g._num_nodes[nt] for g in graphs) for nt in PairGraphBatch.ntypes}
    edges = {}
    for et in CANONICAL_ETYPES:
        src_t, _, dst_t = et
        so = do = 0
        ss, dd = [], []
        for g in graphs:
            s, d = g._edges[et]
            ss.append(s + so)
            dd.append(d + do)
            so += g._num_nodes[src_t]
            do += g._num_nodes[dst_t]
        edges[et] = (torch.cat(ss), torch.cat(dd))
    bnn = {nt: torch.cat([g._batch_num_nodes[nt] for g in graphs]) for nt in PairGraphBatch.ntypes}
    bne = {et: torch.cat([g._batch_num_edges[et] for g in graphs]) for et in CANONICAL_ETYPES}
    out = PairGraphBatch(num_nodes, edges, bnn, bne)
    for nt in PairGraphBatch.ntypes:
        for key in graphs[0]._ndata[nt]:
            out._ndata[nt][key] = torch.cat([g._ndata[nt][key] for g in graphs], dim=0)
    for et in CANONICAL_ETYPES:
        for key in graphs[0]._edata[et]:
            out._edata[et][key] = torch.cat([g._edata[et][key] for g in graphs], dim=0)
    return out


def unbatch(g: PairGraphBatch) -> List[PairGraphBatch]:
    """``dgl.unbatch``: split back into single-pair graphs (views of the batched tensors)."""
    bnn = {nt: g._batch_num_nodes[nt].tolist() for nt in PairGraphBatch.ntypes}
    bne = {et: g._batch_num_edges[et].tolist() for et in CANONICAL_ETYPES}
    n_off = {nt: 0 for nt in PairGraphBatch.ntypes}
    e_off = {et: 0 for et in CANONICAL_ETYPES}
    out = []
    for b in range(len(bnn[LIGAND])):
        nn = {nt: bnn[nt][b] for nt in PairGraphBatch.ntypes}
        edges = {}
        for et in CANONICAL_ETYPES:
            src_t, _, dst_t = et
            s, d = g._edges[et]
            lo, hi = e_off[et], e_off[et] + bne[et][b]
            edges[et] = (s[lo:hi] - n_off[src_t], d[lo:hi] - n_off[dst_t])
        one = PairGraphBatch(nn, edges)
        for nt in PairGraphBatch.ntypes:
            lo, hi = n_off[nt], n_off[nt] + nn[nt]
            one._ndata[nt] = {k: v[lo:hi] for k, v in g._ndata[nt].items()}
        for et in CANONICAL_ETYPES:
            lo, hi = e_off[et], e_off[et] + bne[et][b]
            one._edata[et] = {k: v[lo:hi] for k, v in g._edata[et].items()}
        for nt in PairGraphBatch.ntypes:
            n_off[nt] += nn[nt]
        for et in CANONICAL_ETYPES:
            e_off[et] += bne[et][b]
        out.append(one)
    return out


def batch_pairs(pairs: Iterable[Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]]) -> PairGraphBatch:
    """Collate ``[(ligand_dict, receptor_dict), ...]`` like ``batchify_and_create_hetero_graphs``
    (``src/utils/train_utils.py:87-100``)."""
    return batch([pair_graph(l, r) for l, r in pairs])
