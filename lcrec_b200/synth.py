"""Deterministic synthetic inputs shared by the tests, the golden-vector generator and bench.py.

Items mimic LLM text embeddings and produce collisions: a fixed low-rank
map ``G`` (64 x dim), ``n_parents`` parent vectors, ``x_i = h_{p(i)} G + noise * n_i`` so roughly
``n / n_parents`` near-duplicates share a parent.  numpy PCG64 streams are stable across versions.
"""
from __future__ import annotations

import numpy as np

RANK = 64


def lowrank_map(dim: int, seed: int = 7) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal((RANK, dim)) / np.sqrt(RANK)).astype(np.float32)


def synth_items(n: int, dim: int, n_parents: int, seed: int, noise: float = 0.05,
                start: int = 0) -> np.ndarray:
    """Items ``start .. start+n`` of the stream identified by (dim, n_parents, seed)."""
    g = lowrank_map(dim)
    parents = np.random.default_rng(seed).standard_normal((max(n_parents, 1), RANK)).astype(np.float32)
    rng = np.random.default_rng([seed, 1, start])
    pid = rng.integers(0, max(n_parents, 1), size=n)
    x = parents[pid] @ g
    x += noise * rng.standard_normal((n, dim), dtype=np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def seeded_weights(dims, n_codes, e_dim: int, seed: int, cb_scale: float = 0.05):
    """Xavier-normal Linear weights (out, in), small non-zero biases, and codebooks whose scale
    halves per level.  Returns (weights, biases, codebooks) as lists of fp32 arrays."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for fi, fo in zip(dims[:-1], dims[1:]):
        std = np.sqrt(2.0 / (fi + fo))
        ws.append((rng.standard_normal((fo, fi), dtype=np.float32) * np.float32(std)))
        bs.append((rng.standard_normal(fo, dtype=np.float32) * np.float32(0.01)))
    cbs = [(rng.standard_normal((k, e_dim), dtype=np.float32) * np.float32(cb_scale * 0.5 ** l))
           for l, k in enumerate(n_codes)]
    return ws, bs, cbs
