"""CPU restatement of the device-side synthetic-data generator -- TEST INFRASTRUCTURE ONLY.

What it restates: the *semantics* of the reference's data path for synthetic spatial graphs -- coordinates and node features in
[0,1) (input_data.py:57-58 scale the stored data into that range), `rel` = pairwise Euclidean distance (input_data.py:145-151), a
symmetric zero-diagonal truth adjacency (input_data.py:62-67), and `sampling_num` random spanning forests of it per graph
(input_data.py:18-24,77-82: minimum spanning tree under i.i.d. random edge weights), rows laid out graph-major with aligned
features / spatial / rel (SURVEY quirk Q6 fixed).  The reference draws from numpy's global Mersenne Twister and calls scipy's
MST; a GPU cannot replay that stream, so the generator (device kernel and this file alike) uses a counter-based hash instead:

    u32(stream, idx) = high 32 bits of splitmix64(seed + stream * 0xD1B54A32D192ED03 + idx * 0x9E3779B97F4A7C15)
    uniform          = (u32 >> 8) * 2^-24                                         (float32, exact)
    edge key         = (u32(K_EDGE + sample, i * N + j) << 32) | (i * N + j),  i < j   -> strict total order, unique forest

An MST depends only on the order of the weights, and i.i.d. keys induce the same distribution over orders as the reference's
i.i.d. U[1,2) weights, so the sampled forests have the reference's distribution.  Kruskal with union-find here, Prim on the GPU;
the forest is unique, so the two agree bit for bit.  All float steps are single IEEE operations in a fixed order (no FMA), so the
truth adjacency (a float comparison) is bit-exact as well.
"""
import math

import numpy as np

M64 = (1 << 64) - 1
K_COORD, K_FEAT, K_EDGE = 1, 2, 16


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def u32(seed, stream, idx):
    return splitmix64((seed + stream * 0xD1B54A32D192ED03 + idx * 0x9E3779B97F4A7C15) & M64) >> 32


def uniform(seed, stream, idx):
    return np.float32((u32(seed, stream, idx) >> 8) * (1.0 / 16777216.0))


def radius2(N):
    """r^2 with r = sqrt(6 / (pi N)): mean degree ~ 6 in the unit square; float32, computed once on the host."""
    return np.float32(6.0 / (math.pi * N))


def synth_inputs(N, B, S, F, D, seed):
    """All eight feeds of construct_feed_dict_train (preprocessing.py:32-42) for B graphs, as float32 numpy arrays."""
    P = np.zeros((B, N, D), np.float32); X = np.zeros((B, N, F), np.float32)
    for b in range(B):
        for n in range(N):
            for d in range(D):
                P[b, n, d] = uniform(seed, K_COORD, (b * N + n) * D + d)
            for f in range(F):
                X[b, n, f] = uniform(seed, K_FEAT, (b * N + n) * F + f)
    r2 = radius2(N)
    d2 = np.zeros((B, N, N), np.float32)
    for d in range(D):                                    # d2 = ((dx*dx) + dy*dy) + ...: one rounding per operation, in this order
        diff = (P[:, :, None, d] - P[:, None, :, d]).astype(np.float32)
        d2 = (d2 + (diff * diff).astype(np.float32)).astype(np.float32)
    rel = np.sqrt(d2).astype(np.float32)
    A = (d2 < r2).astype(np.float32)
    idx = np.arange(N)
    A[:, idx, idx] = 0.0
    As = np.zeros((B * S, N, N), np.float32)
    for b in range(B):
        iu, ju = np.nonzero(np.triu(A[b], 1))
        for s in range(S):
            smp = b * S + s
            keys = [((u32(seed, K_EDGE + smp, int(i) * N + int(j)) << 32) | (int(i) * N + int(j)), int(i), int(j)) for i, j in zip(iu, ju)]
            keys.sort()
            parent = list(range(N))

            def find(a):
                while parent[a] != a:
                    parent[a] = parent[parent[a]]; a = parent[a]
                return a
            for _, i, j in keys:
                ri, rj = find(i), find(j)
                if ri != rj:
                    parent[ri] = rj
                    As[smp, i, j] = 1.0; As[smp, j, i] = 1.0
    return {"adj_truth": A, "feature_truth": X, "spatial_truth": P, "rel_truth": rel[..., None], "adj": As,
            "features": np.repeat(X, S, axis=0), "spatial": np.repeat(P, S, axis=0), "rel": np.repeat(rel, S, axis=0)[..., None]}
