"""Pipelined host->device->host inference over the reference-facing module: the public "serve a stream of batches"
call.  While the engine works on batch k (compute stream), batch k+1 is copied from pinned host memory on a second
stream and batch k-1's status words / results are read back -- so end-to-end throughput is max(PCIe, compute)
instead of their sum.  Every batch still goes through ``Rigid_Body_Docking_Net.forward_async`` (same kernels, same
status handling as ``model(graph, epoch)``)."""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional

import torch

from .hetero_graph import CANONICAL_ETYPES, PairGraphBatch


def _record_stream(g: PairGraphBatch, stream: torch.cuda.Stream) -> None:
    for s, d in g._edges.values():
        s.record_stream(stream)
        d.record_stream(stream)
    for fr in list(g._ndata.values()) + list(g._edata.values()):
        for t in fr.values():
            t.record_stream(stream)


class PipelinedInference:
    def __init__(self, model, device, on_result: Optional[Callable] = None):
        self.model, self.device = model, torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.on_result = on_result

    def run(self, host_batches: Iterable[PairGraphBatch]) -> Iterator:
        """``host_batches``: pinned ``PairGraphBatch`` objects.  Yields, in order, for every batch a dict of pinned
        host tensors: ``ligand_coors`` (sum N_l, 3), ``rotation`` (B, 3, 3), ``translation`` (B, 1, 3)."""
        compute = torch.cuda.current_stream(self.device)
        it = iter(host_batches)

        def stage(hb):
            with torch.cuda.stream(self.copy_stream):
                g = hb.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            _record_stream(g, compute)
            return g, ev

        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        in_flight = None   # (pending handle, host result buffers)
        while staged is not None:
            g, ev = staged
            compute.wait_event(ev)
            pending = self.model.forward_async(g, 0)
            nxt = next(it, None)
            staged = stage(nxt) if nxt is not None else None     # H2D of the next batch overlaps this batch's kernels
            if in_flight is not None:
                yield self._finish(*in_flight)
            in_flight = (pending, g)
        if in_flight is not None:
            yield self._finish(*in_flight)

    def _finish(self, pending, g):
        coors, _, _, rot, trans = pending.result()
        res = {'ligand_coors': torch.cat(coors), 'rotation': torch.stack(rot), 'translation': torch.stack(trans)}
        out = {}
        for k, t in res.items():
            hbuf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            hbuf.copy_(t, non_blocking=True)
            out[k] = hbuf
        done = torch.cuda.Event()
        done.record()
        out['_event'] = done
        return out
