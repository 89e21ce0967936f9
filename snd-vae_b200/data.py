"""Host-side synthetic spatial-graph generator (the shapes and semantics of
input_data.py:18-38,54-96 and main.py:305-323, with SURVEY quirk Q6 fixed: the
tiled `features/spatial/rel` rows are aligned with the graph-major `adj` rows).

Random-geometric graphs in the unit square (coords ~ U[0,1)^D, edge iff distance <
sqrt(6/(pi N)), mean degree ~6), node feature ~ U[0,1), and S random spanning
forests per graph via scipy's MST on U[1,2) weights (input_data.py:18-24).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def synthetic_graphs(num_nodes: int, batch: int, sampling_num: int = 10, num_feature: int = 1, spatial_dim: int = 2,
                     seed: int = 1234) -> Dict[str, np.ndarray]:
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree
    rng = np.random.default_rng(seed)
    N, F, D, S, B = num_nodes, num_feature, spatial_dim, sampling_num, batch
    P = rng.random((B, N, D), dtype=np.float32)
    X = rng.random((B, N, F), dtype=np.float32)
    diff = P[:, :, None, :] - P[:, None, :, :]
    rel = np.sqrt((diff ** 2).sum(-1)).astype(np.float32)
    A = (rel < math.sqrt(6.0 / (math.pi * N))).astype(np.float32)
    idx = np.arange(N)
    A[:, idx, idx] = 0.0
    As = np.zeros((B, S, N, N), dtype=np.float32)
    for b in range(B):
        x, y = np.where(A[b])
        if len(x) == 0:
            continue
        for s in range(S):
            cg = csr_matrix((rng.random(len(x)) + 1, (x, y)), shape=(N, N))
            tr, tc = minimum_spanning_tree(cg).nonzero()
            As[b, s, tr, tc] = 1.0
            As[b, s, tc, tr] = 1.0
    return {
        "adj_truth": A, "feature_truth": X, "spatial_truth": P, "rel_truth": rel[..., None],
        "adj": As.reshape(B * S, N, N),
        "features": np.repeat(X, S, axis=0), "spatial": np.repeat(P, S, axis=0),
        "rel": np.repeat(rel, S, axis=0)[..., None],
    }


def tile_pool(pool: Dict[str, np.ndarray], batch: int, pool_graphs: int, sampling_num: int) -> Dict[str, np.ndarray]:
    """Repeat a pool of distinct graphs (with their S sample rows) up to `batch` graphs."""
    reps = (batch + pool_graphs - 1) // pool_graphs
    out = {}
    for k, v in pool.items():
        per = sampling_num if k in ("adj", "features", "spatial", "rel") else 1
        out[k] = np.concatenate([v] * reps, axis=0)[: batch * per]
    return out
