// synth.cuh -- device-side synthetic spatial graphs and spanning-forest samples.
//
// Replaces the host loader's work for synthetic data (input_data.py:54-96: coordinates / features in [0,1), pairwise distances,
// symmetric zero-diagonal adjacency; input_data.py:18-24,77-82: `sampling_num` minimum spanning trees per graph under i.i.d. random
// edge weights) with two kernels that fill the eight feeds of construct_feed_dict_train (preprocessing.py:32-42) in HBM, rows
// graph-major and aligned.  Randomness is a counter-based hash (splitmix64 of seed / stream / index), so the CPU
// restatement used by the parity tests reproduces every array bit for bit:
//   * floats: each step is one IEEE operation in a fixed order (__fsub_rn / __fmul_rn / __fadd_rn / __fsqrt_rn, no FMA contraction);
//   * forests: edge keys (hash32, edge id) are a strict total order, so the minimum spanning forest is unique -- Prim here (one CTA
//     per sample, truth adjacency recomputed from the coordinates in shared memory, 64-bit block arg-min per step), Kruskal there.
#pragma once
#include "common.cuh"

#define SY_K_COORD 1ull
#define SY_K_FEAT 2ull
#define SY_K_EDGE 16ull

__host__ __device__ __forceinline__ unsigned long long sy_splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ unsigned int sy_u32(unsigned long long seed, unsigned long long stream, unsigned long long idx) {
  return (unsigned int)(sy_splitmix64(seed + stream * 0xD1B54A32D192ED03ull + idx * 0x9E3779B97F4A7C15ull) >> 32);
}
__device__ __forceinline__ float sy_uniform(unsigned long long seed, unsigned long long stream, unsigned long long idx) {
  return (float)(sy_u32(seed, stream, idx) >> 8) * (1.0f / 16777216.0f);      // exact: 24-bit integer times 2^-24
}
// squared distance with one rounding per operation, in a fixed order (the parity tests replay it on the CPU)
__device__ __forceinline__ float sy_dist2(const float* __restrict__ P, int i, int j, int D) {
  float d2 = 0.f;
  for (int d = 0; d < D; ++d) { const float diff = __fsub_rn(P[i * D + d], P[j * D + d]); d2 = __fadd_rn(d2, __fmul_rn(diff, diff)); }
  return d2;
}

// one CTA per graph: spatial_truth, feature_truth, rel_truth, adj_truth and the S-fold repeated features / spatial rows
__global__ void __launch_bounds__(256) synth_graph_k(unsigned long long seed, int N, int F, int D, int S, float r2,
                                                     float* __restrict__ spatial_truth, float* __restrict__ feature_truth,
                                                     float* __restrict__ rel_truth, float* __restrict__ adj_truth,
                                                     float* __restrict__ features, float* __restrict__ spatial) {
  extern __shared__ float sP[];                     // [N][D]
  const long long b = blockIdx.x;
  for (int t = threadIdx.x; t < N * D; t += blockDim.x) {
    const float v = sy_uniform(seed, SY_K_COORD, (unsigned long long)b * N * D + t);
    sP[t] = v;
    if (spatial_truth) spatial_truth[b * N * D + t] = v;
    if (spatial) for (int s = 0; s < S; ++s) spatial[(b * S + s) * N * D + t] = v;
  }
  for (int t = threadIdx.x; t < N * F; t += blockDim.x) {
    const float v = sy_uniform(seed, SY_K_FEAT, (unsigned long long)b * N * F + t);
    if (feature_truth) feature_truth[b * N * F + t] = v;
    if (features) for (int s = 0; s < S; ++s) features[(b * S + s) * N * F + t] = v;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int i = t / N, j = t - i * N;
    const float d2 = sy_dist2(sP, i, j, D);
    if (rel_truth) rel_truth[b * N * N + t] = __fsqrt_rn(d2);
    if (adj_truth) adj_truth[b * N * N + t] = (i != j && d2 < r2) ? 1.f : 0.f;
  }
}

// one CTA per sample (b, s): minimum spanning forest of graph b's truth adjacency under the sample's edge keys -> adj[b*S+s];
// also writes the sample's copy of rel.  Unreached nodes carry the marker 2^63 + node, so that the arg-min starts a new
// component at the smallest remaining node once the current one is exhausted.
#define SY_THREADS 256
#define SY_MAXN 2048
__global__ void __launch_bounds__(SY_THREADS) synth_sample_k(unsigned long long seed, int N, int D, int S, float r2,
                                                            const float* __restrict__ spatial_truth, float* __restrict__ adj,
                                                            float* __restrict__ rel) {
  extern __shared__ unsigned char sraw[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(sraw);        // [N]
  int* parent = reinterpret_cast<int*>(key + N);                                 // [N]
  float* sP = reinterpret_cast<float*>(parent + N);                              // [N][D]
  unsigned char* intree = reinterpret_cast<unsigned char*>(sP + N * D);          // [N]
  __shared__ unsigned long long wmin[SY_THREADS / 32];
  __shared__ unsigned long long best;
  const long long smp = blockIdx.x, b = smp / S;
  const unsigned long long INF = 1ull << 63;
  for (int t = threadIdx.x; t < N * D; t += blockDim.x) sP[t] = spatial_truth[b * N * D + t];
  for (int t = threadIdx.x; t < N; t += blockDim.x) { key[t] = INF + t; parent[t] = -1; intree[t] = 0; }
  float* A = adj + smp * N * N;
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) A[t] = 0.f;
  __syncthreads();
  if (rel) {
    float* R = rel + smp * N * N;
    for (int t = threadIdx.x; t < N * N; t += blockDim.x) { const int i = t / N, j = t - i * N; R[t] = __fsqrt_rn(sy_dist2(sP, i, j, D)); }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0; it < N; ++it) {
    // arg-min over the nodes outside the forest (keys are unique, markers are ordered by node)
    unsigned long long m = ~0ull;
    for (int u = threadIdx.x; u < N; u += blockDim.x) if (!intree[u] && key[u] < m) m = key[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); if (x < m) m = x; }
    if (lane == 0) wmin[warp] = m;
    __syncthreads();
    if (warp == 0) {
      unsigned long long x = lane < SY_THREADS / 32 ? wmin[lane] : ~0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o); if (y < x) x = y; }
      if (lane == 0) best = x;
    }
    __syncthreads();
    const unsigned long long kb = best;
    // the winner: a marker names its node directly, an edge key names the edge (i < j) whose endpoint outside the forest wins
    int v;
    if (kb >= INF) v = (int)(kb - INF);
    else { const int e = (int)(kb & 0x7FFFFFFFull); const int i = e / N, j = e - i * N; v = intree[i] ? j : i; }
    if (threadIdx.x == 0) {
      if (kb < INF) { const int pv = parent[v]; A[(long long)v * N + pv] = 1.f; A[(long long)pv * N + v] = 1.f; }
      intree[v] = 1;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < N; u += blockDim.x) {
      if (intree[u] || u == v) continue;
      if (sy_dist2(sP, v, u, D) < r2) {
        const int i = min(u, v), j = max(u, v);
        const unsigned long long e = (unsigned long long)i * N + j;
        const unsigned long long k = ((unsigned long long)sy_u32(seed, SY_K_EDGE + (unsigned long long)smp, e) << 31) | e;
        if (k < key[u]) { key[u] = k; parent[u] = v; }
      }
    }
    __syncthreads();
  }
}
