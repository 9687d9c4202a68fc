"""On-disk / wire formats around the hot path.

* ``RaggedArchive`` -- one flat, memory-mappable file of named ragged arrays: the DGL-free replacement of the reference's
  dataset caches (DGL ``.bin`` graph lists + pickled label dicts, src/utils/db5_data.py:51-63, 137-138).  ``save_pairs`` /
  ``PairArchive`` store a list of (ligand graph, receptor graph[, labels]) in the engine's input contract (CSR edges grouped
  by destination, 27 edge features, node features) so that a batch is a handful of contiguous slices -- no unpickling, no
  per-graph Python objects, no DGL.
* ``write_pdb_with_coords`` / ``read_pdb_atoms`` -- the output side of src/inference_rigid.py:237-239 (biopandas
  ``to_pdb(records=['ATOM'])`` of the transformed ligand): ATOM records with the coordinate columns replaced.
* ``save_checkpoint`` / ``load_checkpoint`` -- the reference's checkpoint dict (src/utils/early_stop.py:178-190: epoch,
  state_dict, optimizer, args without the non-loadable keys), so checkpoints move freely between the two code bases.
"""
from __future__ import annotations

import copy
import json
import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = b'EQDRAGGED1\n'
ALIGN = 64


class RaggedArchive:
    """File = MAGIC | u64 header length | JSON header | 64-byte aligned raw little-endian arrays.
    header = {'arrays': {name: {'dtype', 'shape', 'offset'}}, 'meta': {...}}."""

    def __init__(self, path: str, mmap: bool = True):
        self.path = path
        with open(path, 'rb') as fh:
            if fh.read(len(MAGIC)) != MAGIC:
                raise ValueError(f'{path}: not a RaggedArchive')
            (hlen,) = struct.unpack('<Q', fh.read(8))
            self.header = json.loads(fh.read(hlen).decode())
        self.meta = self.header.get('meta', {})
        self._mm = np.memmap(path, mode='r', dtype=np.uint8) if mmap else np.fromfile(path, dtype=np.uint8)

    def __contains__(self, name):
        return name in self.header['arrays']

    def __getitem__(self, name) -> np.ndarray:
        e = self.header['arrays'][name]
        dt = np.dtype(e['dtype'])
        n = int(np.prod(e['shape'])) if e['shape'] else 1
        return self._mm[e['offset']:e['offset'] + n * dt.itemsize].view(dt).reshape(e['shape'])

    @staticmethod
    def write(path: str, arrays: Dict[str, np.ndarray], meta: Optional[dict] = None):
        entries, blobs = {}, []
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            entries[name] = {'dtype': a.dtype.str, 'shape': list(a.shape), 'offset': 0}
            blobs.append((name, a))
        # two passes: the header length shifts the offsets
        for _ in range(2):
            hdr = json.dumps({'arrays': entries, 'meta': meta or {}}, sort_keys=True).encode()
            off = len(MAGIC) + 8 + len(hdr) + 64          # slack so that the second pass fits
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for name, a in blobs:
                entries[name]['offset'] = off
                off = (off + a.nbytes + ALIGN - 1) // ALIGN * ALIGN
        hdr = json.dumps({'arrays': entries, 'meta': meta or {}}, sort_keys=True).encode()
        tmp = path + '.tmp'
        with open(tmp, 'wb') as fh:
            fh.write(MAGIC)
            fh.write(struct.pack('<Q', len(hdr)))
            fh.write(hdr)
            for name, a in blobs:
                fh.seek(entries[name]['offset'])
                fh.write(a.tobytes())
            fh.truncate(off)
        os.replace(tmp, path)


_NODE_KEYS = (('res_feat', np.uint8, 1), ('x', np.float32, 3), ('mu_r_norm', np.float32, 5))


def save_pairs(path: str, pairs: Sequence[Tuple[dict, dict]], labels: Optional[Sequence[dict]] = None, meta: Optional[dict] = None):
    """pairs: [(ligand dict, receptor dict)] in the fixture / engine format (src, dst, he, res_feat, x, mu_r_norm, ligand
    new_x); labels (training caches): [{'pocket_coors', 'bound_lig', 'bound_rec'}] (db5_data.py:137-138)."""
    arrays = {}
    for side, idx in (('lig', 0), ('rec', 1)):
        nodes = [int(np.asarray(p[idx]['x']).shape[0]) for p in pairs]
        edges = [int(np.asarray(p[idx]['src']).shape[0]) for p in pairs]
        arrays[f'{side}/node_ptr'] = np.concatenate([[0], np.cumsum(nodes)]).astype(np.int64)
        arrays[f'{side}/edge_ptr'] = np.concatenate([[0], np.cumsum(edges)]).astype(np.int64)
        for key, dt, w in _NODE_KEYS:
            arrays[f'{side}/{key}'] = np.concatenate([np.asarray(p[idx][key]).reshape(-1, w) for p in pairs]).astype(dt)
        arrays[f'{side}/src'] = np.concatenate([np.asarray(p[idx]['src']) for p in pairs]).astype(np.int32)
        arrays[f'{side}/dst'] = np.concatenate([np.asarray(p[idx]['dst']) for p in pairs]).astype(np.int32)
        arrays[f'{side}/he'] = np.concatenate([np.asarray(p[idx]['he']).reshape(-1, 27) for p in pairs]).astype(np.float32)
    arrays['lig/new_x'] = np.concatenate([np.asarray(p[0].get('new_x', p[0]['x'])).reshape(-1, 3) for p in pairs]).astype(np.float32)
    if labels is not None:
        pk = [np.asarray(l['pocket_coors'], np.float32).reshape(-1, 3) for l in labels]
        arrays['label/pocket_ptr'] = np.concatenate([[0], np.cumsum([a.shape[0] for a in pk])]).astype(np.int64)
        arrays['label/pocket_coors'] = np.concatenate(pk)
        arrays['label/bound_lig'] = np.concatenate([np.asarray(l['bound_lig'], np.float32).reshape(-1, 3) for l in labels])
        arrays['label/bound_rec'] = np.concatenate([np.asarray(l['bound_rec'], np.float32).reshape(-1, 3) for l in labels])
    RaggedArchive.write(path, arrays, {'kind': 'pairs', 'n_pairs': len(pairs), **(meta or {})})


class PairArchive:
    def __init__(self, path: str, mmap: bool = True):
        self.a = RaggedArchive(path, mmap)
        if self.a.meta.get('kind') != 'pairs':
            raise ValueError(f'{path}: not a pair archive')
        self.n_pairs = int(self.a.meta['n_pairs'])

    def __len__(self):
        return self.n_pairs

    def pair(self, i: int) -> Tuple[dict, dict]:
        out = []
        for side in ('lig', 'rec'):
            n0, n1 = (int(v) for v in self.a[f'{side}/node_ptr'][i:i + 2])
            e0, e1 = (int(v) for v in self.a[f'{side}/edge_ptr'][i:i + 2])
            d = {'src': self.a[f'{side}/src'][e0:e1], 'dst': self.a[f'{side}/dst'][e0:e1], 'he': self.a[f'{side}/he'][e0:e1],
                 'res_feat': self.a[f'{side}/res_feat'][n0:n1].astype(np.float32), 'x': self.a[f'{side}/x'][n0:n1],
                 'mu_r_norm': self.a[f'{side}/mu_r_norm'][n0:n1]}
            if side == 'lig':
                d['new_x'] = self.a['lig/new_x'][n0:n1]
            out.append(d)
        return out[0], out[1]

    def labels(self, i: int) -> dict:
        p0, p1 = (int(v) for v in self.a['label/pocket_ptr'][i:i + 2])
        l0, l1 = (int(v) for v in self.a['lig/node_ptr'][i:i + 2])
        r0, r1 = (int(v) for v in self.a['rec/node_ptr'][i:i + 2])
        return {'pocket_coors': self.a['label/pocket_coors'][p0:p1], 'bound_lig': self.a['label/bound_lig'][l0:l1],
                'bound_rec': self.a['label/bound_rec'][r0:r1]}

    def batch(self, indices: Sequence[int]):
        """-> PairGraphBatch (host tensors) of the selected pairs (``hetero_graph.batch_pairs``)."""
        import torch
        from . import hetero_graph as hg
        tp = [tuple({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()} for d in self.pair(i)) for i in indices]
        return hg.batch_pairs(tp)


# ---- PDB (ATOM records only, fixed columns) --------------------------------------------------------------------------

def read_pdb_atoms(path: str):
    """-> (lines of the ATOM records, (n,3) float64 coordinates) in file order."""
    lines, xyz = [], []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith('ATOM'):
                lines.append(ln.rstrip('\n'))
                xyz.append((float(ln[30:38]), float(ln[38:46]), float(ln[46:54])))
    return lines, np.asarray(xyz, np.float64).reshape(-1, 3)


def write_pdb_with_coords(src_path: str, dst_path: str, coords: np.ndarray):
    """Writes the ATOM records of ``src_path`` with their coordinates replaced (8.3f columns 31-54), what
    ``ppdb.df['ATOM'][[x,y,z]] = coords; ppdb.to_pdb(records=['ATOM'])`` does (inference_rigid.py:237-239)."""
    lines, xyz = read_pdb_atoms(src_path)
    coords = np.asarray(coords, np.float64).reshape(-1, 3)
    if coords.shape[0] != len(lines):
        raise ValueError(f'{src_path}: {len(lines)} ATOM records, {coords.shape[0]} coordinates')
    with open(dst_path, 'w') as fh:
        for ln, (x, y, z) in zip(lines, coords):
            ln = ln.ljust(80)
            fh.write(f'{ln[:30]}{x:8.3f}{y:8.3f}{z:8.3f}{ln[54:]}'.rstrip() + '\n')


def apply_rigid_to_pdb(src_path: str, dst_path: str, rotation, translation):
    """All ligand atoms through (R, t) like inference_rigid.py:205, then written out."""
    _, xyz = read_pdb_atoms(src_path)
    R, t = np.asarray(rotation, np.float64).reshape(3, 3), np.asarray(translation, np.float64).reshape(3)
    write_pdb_with_coords(src_path, dst_path, (R @ xyz.T).T + t)


# ---- checkpoints -----------------------------------------------------------------------------------------------------

NON_LOAD_KEYS = ('device', 'debug', 'worker', 'n_jobs', 'toy')      # early_stop.py:179


def save_checkpoint(path: str, model, optimizer_state: dict, epoch: int, args: dict):
    import torch
    a = copy.deepcopy({k: v for k, v in args.items() if k not in NON_LOAD_KEYS})
    torch.save({'epoch': epoch, 'state_dict': model.state_dict(), 'optimizer': optimizer_state, 'args': a}, path)


def load_checkpoint(path: str, map_location='cpu'):
    import torch
    ck = torch.load(path, map_location=map_location, weights_only=False)
    return ck['args'], ck['state_dict'], ck.get('optimizer'), ck.get('epoch')
