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


# ---- compact feeds (include/sndvae.h sndvae_inputs_compact) ---------------------------------------------------------
def pack_adj_bits(adj: np.ndarray) -> np.ndarray:
    """0/1 adjacency [..., N, N] -> bit rows [..., N, W] uint32, W = ceil(N / 32); bit (j % 32) of word (j / 32) = adj[.., i, j] != 0."""
    N = adj.shape[-1]
    W = (N + 31) // 32
    b = np.zeros(adj.shape[:-1] + (W * 32,), dtype=np.uint8)
    b[..., :N] = adj != 0
    return np.ascontiguousarray(np.packbits(b, axis=-1, bitorder="little")).view("<u4").reshape(adj.shape[:-1] + (W,))


def unpack_adj_bits(bits: np.ndarray, num_nodes: int, dtype=np.int64) -> np.ndarray:
    """Inverse of pack_adj_bits: [..., N, W] uint32 -> [..., N, N]."""
    by = np.ascontiguousarray(bits.astype("<u4")).view(np.uint8)
    return np.unpackbits(by, axis=-1, bitorder="little")[..., :num_nodes].astype(dtype)


def pack_feeds(feeds: Dict[str, np.ndarray], sampling_num: int) -> Dict[str, np.ndarray]:
    """The dense feed_dict arrays of main.py:253-264 -> the compact host feeds: per-graph `features` / `rel` once (the dense
    rows b*S .. b*S+S-1 are copies, main.py:307-309), adjacencies as bit rows.  Raises if an adjacency is not 0/1 or the S
    copies differ -- the compact format cannot carry those."""
    S = sampling_num
    for k in ("adj", "adj_truth"):
        a = feeds[k]
        if not np.array_equal(a, (a != 0).astype(a.dtype)):
            raise ValueError(f"feed '{k}' is not a 0/1 adjacency: use the dense entry point")
    rel = feeds["rel"].reshape(feeds["rel"].shape[:3])
    for k, v in (("rel", rel), ("features", feeds["features"])):
        g = v.reshape((-1, S) + v.shape[1:])
        if not (g == g[:, :1]).all():
            raise ValueError(f"feed '{k}' differs between the samples of a graph: use the dense entry point")
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return {"features": f32(feeds["features"][::S]), "adj_bits": pack_adj_bits(feeds["adj"]), "rel": f32(rel[::S]),
            "adj_truth_bits": pack_adj_bits(feeds["adj_truth"]), "feature_truth": f32(feeds["feature_truth"]),
            "spatial_truth": f32(feeds["spatial_truth"])}
