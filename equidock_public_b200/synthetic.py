"""Canonical synthetic "DB5.5-shaped" residue graphs -- the bench / scaling workload.

Per protein: N C-alpha points = cumulative sum of random unit steps x 3.8 A x 0.35 plus N(0, 4^2)
jitter (compact blob); k nearest neighbours by Euclidean distance, edges grouped by destination;
``res_feat`` ~ U{0..20}; ``mu_r_norm`` ~ U(0.05, 1); ``he[:, :15]`` = exp(-d^2/1.5^s) of the edge
length (``protein_utils.py:71-86``), ``he[:, 15:18]`` = a displacement of norm d, ``he[:, 18:27]`` ~
U(-1, 1).  The ligand then gets a random rigid motion (``protein_utils.py:15-23``).  Pure numpy, so
the same generator feeds the CUDA engine, the oracle and the CPU baseline.
"""
from __future__ import annotations

import numpy as np


def synthetic_protein(rng: np.random.Generator, n: int, k: int = 10):
    steps = rng.normal(size=(n, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    x = np.cumsum(steps * 3.8 * 0.35, axis=0) + rng.normal(scale=4.0, size=(n, 3))
    x = x.astype(np.float32)
    d2 = ((x[:, None, :].astype(np.float64) - x[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    kk = min(k, n - 1)
    nbr = np.argsort(d2, axis=1)[:, :kk]                       # (n, kk) sources of each destination
    dst = np.repeat(np.arange(n, dtype=np.int32), kk)
    src = nbr.reshape(-1).astype(np.int32)
    dist = np.sqrt(d2[dst, src])
    sig = 1.5 ** np.arange(15)
    he = np.empty((dst.shape[0], 27), dtype=np.float32)
    he[:, :15] = np.exp(-(dist[:, None] ** 2) / sig[None, :])
    u = rng.normal(size=(dst.shape[0], 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    he[:, 15:18] = u * dist[:, None]
    he[:, 18:27] = rng.uniform(-1, 1, size=(dst.shape[0], 9))
    return {'src': src, 'dst': dst, 'he': he,
            'res_feat': rng.integers(0, 21, size=(n, 1)).astype(np.float32),
            'x': x, 'mu_r_norm': rng.uniform(0.05, 1.0, size=(n, 5)).astype(np.float32)}


def random_rigid(rng: np.random.Generator, translation_interval: float = 5.0, dtype=np.float32):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, a, b, c = q
    R = np.array([[1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)],
                  [2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)],
                  [2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)]])
    t = rng.normal(size=3)
    t = t / np.linalg.norm(t) * rng.uniform(0, translation_interval)
    return R.astype(dtype), t.astype(dtype)


def synthetic_pair(rng: np.random.Generator, n_lig: int = 200, n_rec: int = 200, k: int = 10):
    lig, rec = synthetic_protein(rng, n_lig, k), synthetic_protein(rng, n_rec, k)
    R, t = random_rigid(rng)
    lig['new_x'] = ((R @ lig['x'].T).T + t).astype(np.float32)
    return lig, rec


def synthetic_batch(n_pairs: int, n_lig: int = 200, n_rec: int = 200, k: int = 10, seed: int = 0):
    rng = np.random.default_rng(seed)
    return [synthetic_pair(rng, n_lig, n_rec, k) for _ in range(n_pairs)]


def to_torch_pairs(pairs):
    import torch
    return [tuple({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()} for d in p) for p in pairs]


def synthetic_residue_protein(rng: np.random.Generator, n: int):
    """All-atom stand-in of a protein for the GPU graph builder (graph_build.py): the same C-alpha random walk as
    ``synthetic_protein``, backbone N / C atoms at 1.46 / 1.52 A with a ~110 degree N-CA-C angle, and 1..9 further atoms
    within ~2.5 A of CA (side chain); residue types U{0..20}.  Returns the compact format of oracle/graph_oracle.py."""
    steps = rng.normal(size=(n, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    ca = (np.cumsum(steps * 3.8 * 0.35, axis=0) + rng.normal(scale=4.0, size=(n, 3))).astype(np.float32)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = rng.normal(size=(n, 3)); w -= (w * u).sum(1, keepdims=True) * u; w /= np.linalg.norm(w, axis=1, keepdims=True)
    ang = np.deg2rad(110.0)
    n_at = ca + 1.46 * u
    c_at = ca + 1.52 * (np.cos(ang) * u + np.sin(ang) * w)
    n_side = rng.integers(1, 10, size=n)
    atoms, ptr = [], [0]
    for i in range(n):
        side = ca[i] + rng.normal(scale=1.4, size=(n_side[i], 3))
        a = np.concatenate([n_at[i:i + 1], ca[i:i + 1], c_at[i:i + 1], side]).astype(np.float32)
        atoms.append(a)
        ptr.append(ptr[-1] + a.shape[0])
    return {'atoms': np.concatenate(atoms), 'atom_ptr': np.asarray(ptr, np.int32),
            'nca_c': np.stack([n_at, ca, c_at], axis=1).astype(np.float32),
            'res_feat': rng.integers(0, 21, size=(n, 1)).astype(np.float32)}


def synthetic_residue_pair(rng: np.random.Generator, n_lig: int = 200, n_rec: int = 200):
    lig, rec = synthetic_residue_protein(rng, n_lig), synthetic_residue_protein(rng, n_rec)
    R, t = random_rigid(rng)
    lig['atoms'] = ((R @ lig['atoms'].T).T + t).astype(np.float32)
    lig['nca_c'] = ((lig['nca_c'].reshape(-1, 3) @ R.T) + t).astype(np.float32).reshape(-1, 3, 3)
    return lig, rec
