// sgc.cuh -- SpatialGraphConvolution (layers.py:143-198) in its exact factored
// form, on per-sample edge lists.
//
//   m3s_ij = A_ij [ deg_j (P_i + Q_j + phi(r_ij) w4 + b1) + (A Rm)_j + s_j w5 + G_ij w6 ]
//   m2s_i  = deg_i (U_i + b2) + (A V)_i + s_i M2c + (sum_j A_ij phi(m3s_ij)) M2d
//   out_i  = phi([x_i || m2s_i]) M3 + b3
// with P,Q,Rm = phi(x) M1[0:C],[C:2C],[2C:3C]; U,V = phi(x) M2[0:C],[C:2C];
// deg_j = sum_k A_jk, s_j = sum_k A_jk phi(r_jk), G_ij = sum_k A_jk phi(r_ik).
// Every term is multiplied by A_ij, so only stored (non-zero) entries of the
// sampled adjacency matter; the samples are spanning forests
// (input_data.py:18-38: nnz <= 2(N-1)).  sgc_build_edges_k streams `adj` once
// (the HBM-bound stage of the encoder) and gathers the few `rel` entries needed.
//
// Work split.  Everything that is a per-node dense map is written as a product of a per-node
// coefficient row with a small weight block, so that the whole chunk of samples is ONE tall
// library GEMM (rows = samples x N):
//     P    = xphi  . M1a                       xphi  = phi(x)                          [C]
//     Qc   = coefQ . [M1b; M1c; w5; b1]        coefQ = [deg phi(x), apx, s, deg]       [2C+2]
//     m2s  = coef2 . [M2; b2]                  coef2 = [deg phi(x), apx, s, T, deg]    [2C+2+h0]
//     y    = coef3 . [M3; b3]                  coef3 = [phi(x), phi(m2s), 1]           [C+h1+1]
// so that the bracket of m3s is  deg_j P_i + Qc_j + deg_j phi(r_ij) w4 + G_ij w6  (4 FMAs per edge
// and channel).  The kernels below only do the O(E h) edge work and the element-wise glue; the
// backward uses the transposed GEMMs and the same coefficient rows for the parameter gradients.
#pragma once
#include "common.cuh"

struct SgcEdges {        // per-sample edge storage, capacity `cap` entries per sample
  int*   rowstart;       // [samples, N]
  int*   rowcnt;         // [samples, N]
  int*   erow;           // [samples, cap]
  int*   ecol;           // [samples, cap]
  float* ea;             // [samples, cap]  A_ij
  float* epr;            // [samples, cap]  phi(r_ij)
  float* eG;             // [samples, cap]  G_ij
  float* deg;            // [samples, N]
  float* ssum;           // [samples, N]
  int*   nedges;         // [samples]
  int    cap;
};

// One CTA per sample.  Pass A per row: count non-zeros (streams the adjacency
// row, 128-bit loads when N % 4 == 0); allocate a contiguous slot range; pass B
// re-reads the row (L1-hot) and writes (col, a, phi(rel)).  Then G per edge.
__global__ void __launch_bounds__(256) sgc_build_edges_k(const float* __restrict__ adj, const float* __restrict__ rel,
                                                         SgcEdges E, int N, int* __restrict__ err) {
  __shared__ int s_count;
  long long smp = blockIdx.x;
  const float* A = adj + smp * N * N;
  const float* R = rel + smp * N * N;
  int* rowstart = E.rowstart + smp * N;
  int* rowcnt = E.rowcnt + smp * N;
  int* erow = E.erow + smp * E.cap;
  int* ecol = E.ecol + smp * E.cap;
  float* ea = E.ea + smp * E.cap;
  float* epr = E.epr + smp * E.cap;
  float* eG = E.eG + smp * E.cap;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool vec = (N % 4) == 0;
  for (int i = warp; i < N; i += nw) {
    const float* a = A + (size_t)i * N;
    int cnt = 0;
    if (vec) {
      const float4* a4 = reinterpret_cast<const float4*>(a);
      for (int q = lane; q < N / 4; q += 32) {
        float4 v = __ldg(a4 + q);
        cnt += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
      }
    } else {
      for (int j = lane; j < N; j += 32) cnt += (__ldg(a + j) != 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_count, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane == 0) { rowstart[i] = base; rowcnt[i] = (base + cnt <= E.cap) ? cnt : 0; }
    float dsum = 0.f, ss = 0.f;
    if (base + cnt <= E.cap) {
      int pos = base;
      for (int j0 = 0; j0 < N; j0 += 32) {
        int j = j0 + lane;
        float av = j < N ? a[j] : 0.f;
        unsigned m = __ballot_sync(0xffffffffu, av != 0.f);
        if (av != 0.f) {
          int p = pos + __popc(m & ((1u << lane) - 1u));
          float pr = lrelu_f(__ldg(R + (size_t)i * N + j));
          erow[p] = i; ecol[p] = j; ea[p] = av; epr[p] = pr;
          dsum += av; ss += av * pr;
        }
        pos += __popc(m);
      }
    }
    dsum = warp_sum(dsum); ss = warp_sum(ss);
    if (lane == 0) { E.deg[smp * N + i] = dsum; E.ssum[smp * N + i] = ss; }
  }
  __syncthreads();
  int ne = s_count;
  if (ne > E.cap) { if (threadIdx.x == 0) { atomicExch(err, 1); E.nedges[smp] = 0; } return; }
  if (threadIdx.x == 0) E.nedges[smp] = ne;
  // G_ij = sum_{k in nbr(j)} A_jk phi(r_ik)
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    int i = erow[e], j = ecol[e];
    int s0 = rowstart[j], c0 = rowcnt[j];
    float g = 0.f;
    for (int q = s0; q < s0 + c0; ++q) g = fmaf(ea[q], lrelu_f(__ldg(R + (size_t)i * N + ecol[q])), g);
    eG[e] = g;
  }
}

struct SgcDims { int C, h0, h1, h2; };

struct SgcScratch {      // activations of one layer for a chunk of samples (rows = samples x N)
  float* xphi;           // [rows, C]         phi(x)
  float* coefQ;          // [rows, 2C+2]      [deg phi(x), apx, s, deg]
  float* P;              // [rows, h0]
  float* Qc;             // [rows, h0]
  float* coef2;          // [rows, 2C+2+h0]   [deg phi(x), apx, s, T, deg]
  float* m2s;            // [rows, h1]
  float* coef3;          // [rows, C+h1+1]    [phi(x), phi(m2s), 1]
  float* y;              // [rows, h2]        layer output before BN
  // backward
  float* dcoef3;         // [rows, C+h1+1]
  float* dm2s;           // [rows, h1]
  float* dcoef2;         // [rows, 2C+2+h0]
  float* dP;             // [rows, h0]
  float* dQc;            // [rows, h0]
  float* dxphi;          // [rows, C]
  float* dcoefQ;         // [rows, 2C+2]
  float* dpx;            // [rows, C]
  // per-step weight blocks assembled from the arena (and their gradients)
  float* WQ;             // [2C+2, h0]   [M1b; M1c; w5; b1]
  float* W2;             // [2C+2+h0, h1] [M2; b2]
  float* W3;             // [C+h1+1, h2]  [M3; b3]
  float* dWQ;            // gradient of WQ (scattered back into M1 / b1 after the step)
  float* dW2;            // gradient of [M2; b2]  (coef2^T dm2s, accumulated over the chunks of a step)
  float* dW3;            // gradient of [M3; b3]  (coef3^T dy)
  float* w46;            // [2, h0] gradient partial sums of w4, w6 (rows 3C and 3C+2 of M1)
};

// (1) phi(x), apx = A phi(x), coefQ.  One CTA per sample.
__global__ void __launch_bounds__(256) sgc_prep_k(const float* __restrict__ xin, SgcEdges E, SgcDims D, SgcScratch Sx, int N, long long e_off) {
  const long long ls = blockIdx.x, gs = ls + e_off;
  const int C = D.C, KQ = 2 * C + 2;
  const float* x = xin + ls * N * C;
  const int* rowstart = E.rowstart + gs * N; const int* rowcnt = E.rowcnt + gs * N;
  const int* ecol = E.ecol + gs * E.cap; const float* ea = E.ea + gs * E.cap;
  const float* deg = E.deg + gs * N; const float* ssum = E.ssum + gs * N;
  float* xphi = Sx.xphi + ls * N * C; float* cq = Sx.coefQ + ls * (long long)N * KQ;
  for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
    const int i = idx / C, c = idx - i * C;
    const float px = lrelu_f(x[idx]);
    float acc = 0.f;
    for (int q = rowstart[i]; q < rowstart[i] + rowcnt[i]; ++q) acc = fmaf(ea[q], lrelu_f(x[ecol[q] * C + c]), acc);
    xphi[idx] = px;
    cq[i * KQ + c] = deg[i] * px; cq[i * KQ + C + c] = acc;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) { cq[i * KQ + 2 * C] = ssum[i]; cq[i * KQ + 2 * C + 1] = deg[i]; }
}

// (2) T_i = sum_j A_ij phi(m3s_ij) from P, Qc; writes coef2.  One CTA per sample.
__global__ void __launch_bounds__(256) sgc_edge_fwd_k(SgcEdges E, SgcDims D, SgcScratch Sx, const float* __restrict__ w4,
                                                      const float* __restrict__ w6, int N, long long e_off) {
  const long long ls = blockIdx.x, gs = ls + e_off;
  const int C = D.C, h0 = D.h0, KQ = 2 * C + 2, K2 = 2 * C + 2 + h0;
  const int* rowstart = E.rowstart + gs * N; const int* rowcnt = E.rowcnt + gs * N;
  const int* ecol = E.ecol + gs * E.cap;
  const float* ea = E.ea + gs * E.cap; const float* epr = E.epr + gs * E.cap; const float* eG = E.eG + gs * E.cap;
  const float* deg = E.deg + gs * N;
  const float* P = Sx.P + ls * N * h0; const float* Qc = Sx.Qc + ls * N * h0;
  const float* cq = Sx.coefQ + ls * (long long)N * KQ; float* c2 = Sx.coef2 + ls * (long long)N * K2;
  for (int idx = threadIdx.x; idx < N * h0; idx += blockDim.x) {
    const int i = idx / h0, h = idx - i * h0;
    const float pi = P[idx], a4 = w4[h], a6 = w6[h];
    float acc = 0.f;
    for (int q = rowstart[i]; q < rowstart[i] + rowcnt[i]; ++q) {
      const int j = ecol[q]; const float dj = deg[j], a = ea[q];
      const float br = fmaf(dj, pi, Qc[j * h0 + h]) + dj * epr[q] * a4 + eG[q] * a6;
      acc = fmaf(a, lrelu_f(a * br), acc);
    }
    c2[i * K2 + 2 * C + 1 + h] = acc;
  }
  for (int idx = threadIdx.x; idx < N * (2 * C + 1); idx += blockDim.x) {
    const int i = idx / (2 * C + 1), r = idx - i * (2 * C + 1);
    c2[i * K2 + r] = cq[i * KQ + r];                       // [deg phi(x), apx, s]
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) c2[i * K2 + K2 - 1] = deg[i];
}

// (3) coef3 = [phi(x), phi(m2s), 1]   (element-wise over rows)
__global__ void sgc_cat_k(const float* __restrict__ xphi, const float* __restrict__ m2s, float* __restrict__ coef3, long long rows, int C, int h1) {
  const int K3 = C + h1 + 1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K3) return;
  const long long r = idx / K3; const int k = (int)(idx - r * K3);
  coef3[idx] = k < C ? xphi[r * C + k] : (k < C + h1 ? lrelu_f(m2s[r * h1 + k - C]) : 1.f);
}

// (4) backward through the lrelu of the concat: dx_direct = dcoef3[:, :C] phi'(x), dm2s = dcoef3[:, C:C+h1] phi'(m2s)
__global__ void sgc_cat_bwd_k(const float* __restrict__ dcoef3, const float* __restrict__ x, const float* __restrict__ m2s,
                              float* __restrict__ dx, float* __restrict__ dm2s, long long rows, int C, int h1) {
  const int K3 = C + h1 + 1, W = C + h1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * W) return;
  const long long r = idx / W; const int k = (int)(idx - r * W);
  const float g = dcoef3[r * K3 + k];
  if (k < C) { if (dx) dx[r * C + k] = g * lrelu_g(x[r * C + k]); }
  else dm2s[r * h1 + k - C] = g * lrelu_g(m2s[r * h1 + k - C]);
}

// (5) edge backward: ee = A^2 dT phi'(m3s);  dP_i = sum_j deg_j ee,  dQc_j += ee (atomics),  w4 / w6 gradient sums.
// One CTA per sample.  dQc must be zeroed by the caller.
__global__ void __launch_bounds__(256) sgc_edge_bwd_k(SgcEdges E, SgcDims D, SgcScratch Sx, const float* __restrict__ w4,
                                                      const float* __restrict__ w6, int N, long long e_off) {
  extern __shared__ float sw[];          // [2][h0] partial sums of the w4 / w6 gradients
  const long long ls = blockIdx.x, gs = ls + e_off;
  const int C = D.C, h0 = D.h0, K2 = 2 * C + 2 + h0;
  const int* rowstart = E.rowstart + gs * N; const int* rowcnt = E.rowcnt + gs * N;
  const int* ecol = E.ecol + gs * E.cap;
  const float* ea = E.ea + gs * E.cap; const float* epr = E.epr + gs * E.cap; const float* eG = E.eG + gs * E.cap;
  const float* deg = E.deg + gs * N;
  const float* P = Sx.P + ls * N * h0; const float* Qc = Sx.Qc + ls * N * h0;
  const float* dc2 = Sx.dcoef2 + ls * (long long)N * K2;
  float* dP = Sx.dP + ls * N * h0; float* dQc = Sx.dQc + ls * N * h0;
  for (int t = threadIdx.x; t < 2 * h0; t += blockDim.x) sw[t] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * h0; idx += blockDim.x) {
    const int i = idx / h0, h = idx - i * h0;
    const float pi = P[idx], a4 = w4[h], a6 = w6[h], dT = dc2[i * K2 + 2 * C + 1 + h];
    float accP = 0.f, acc4 = 0.f, acc6 = 0.f;
    for (int q = rowstart[i]; q < rowstart[i] + rowcnt[i]; ++q) {
      const int j = ecol[q]; const float dj = deg[j], a = ea[q];
      const float br = fmaf(dj, pi, Qc[j * h0 + h]) + dj * epr[q] * a4 + eG[q] * a6;
      const float ee = a * a * dT * lrelu_g(a * br);
      accP = fmaf(dj, ee, accP);
      acc4 = fmaf(dj * epr[q], ee, acc4); acc6 = fmaf(eG[q], ee, acc6);
      atomicAdd(dQc + j * h0 + h, ee);
    }
    dP[idx] = accP;
    atomicAdd(sw + h, acc4); atomicAdd(sw + h0 + h, acc6);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * h0; t += blockDim.x) atomicAdd(Sx.w46 + t, sw[t]);
}

// (6) node backward: d phi(x) = dxphi + deg (dcoefQ[:, :C] + dcoef2[:, :C]);  d apx = dcoefQ[:, C:2C] + dcoef2[:, C:2C];
//     apx_j = sum_k A_jk phi(x_k)  =>  d phi(x_k) += A_jk d apx_j;  dx += d phi(x) phi'(x).   One CTA per sample.
__global__ void __launch_bounds__(256) sgc_node_bwd_k(const float* __restrict__ xin, float* __restrict__ dxio, SgcEdges E, SgcDims D,
                                                      SgcScratch Sx, int N, long long e_off) {
  const long long ls = blockIdx.x, gs = ls + e_off;
  const int C = D.C, h0 = D.h0, KQ = 2 * C + 2, K2 = 2 * C + 2 + h0;
  const int* erow = E.erow + gs * E.cap; const int* ecol = E.ecol + gs * E.cap; const float* ea = E.ea + gs * E.cap;
  const float* deg = E.deg + gs * N;
  const int ne = E.nedges[gs];
  const float* x = xin + ls * N * C; float* dx = dxio + ls * N * C;
  const float* dxphi = Sx.dxphi + ls * N * C;
  const float* dcq = Sx.dcoefQ + ls * (long long)N * KQ; const float* dc2 = Sx.dcoef2 + ls * (long long)N * K2;
  float* dpx = Sx.dpx + ls * N * C;
  for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) {
    const int i = idx / C, c = idx - i * C;
    dpx[idx] = dxphi[idx] + deg[i] * (dcq[i * KQ + c] + dc2[i * K2 + c]);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < ne * C; idx += blockDim.x) {
    const int e = idx / C, c = idx - e * C; const int j = erow[e];
    atomicAdd(dpx + ecol[e] * C + c, ea[e] * (dcq[j * KQ + C + c] + dc2[j * K2 + C + c]));
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * C; idx += blockDim.x) dx[idx] += dpx[idx] * lrelu_g(x[idx]);
}

// assemble the per-step weight blocks from the arena
__global__ void sgc_pack_weights_k(const float* __restrict__ M1, const float* __restrict__ b1, const float* __restrict__ M2,
                                   const float* __restrict__ b2, const float* __restrict__ M3, const float* __restrict__ b3,
                                   float* __restrict__ WQ, float* __restrict__ W2, float* __restrict__ W3, SgcDims D) {
  const int C = D.C, h0 = D.h0, h1 = D.h1, h2 = D.h2;
  const int nQ = (2 * C + 2) * h0, n2 = (2 * C + 2 + h0) * h1, n3 = (C + h1 + 1) * h2;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nQ) {
    const int r = idx / h0, h = idx - r * h0;
    WQ[idx] = r < 2 * C ? M1[(C + r) * h0 + h] : (r == 2 * C ? M1[(3 * C + 1) * h0 + h] : b1[h]);
  } else if (idx < nQ + n2) {
    const int q = idx - nQ; const int r = q / h1, h = q - r * h1;
    W2[q] = r < 2 * C + 1 + h0 ? M2[r * h1 + h] : b2[h];
  } else if (idx < nQ + n2 + n3) {
    const int q = idx - nQ - n2; const int r = q / h2, h = q - r * h2;
    W3[q] = r < C + h1 ? M3[r * h2 + h] : b3[h];
  }
}
// scatter the gradients of the packed blocks [M2; b2], [M3; b3] back into the arena (once per step and layer)
__global__ void sgc_unpack_w23_k(const float* __restrict__ dW2, const float* __restrict__ dW3, float* __restrict__ gM2,
                                 float* __restrict__ gb2, float* __restrict__ gM3, float* __restrict__ gb3, SgcDims D) {
  const int C = D.C, h0 = D.h0, h1 = D.h1, h2 = D.h2;
  const int K2 = 2 * C + 2 + h0, K3 = C + h1 + 1, n2 = K2 * h1, n3 = K3 * h2;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n2) {
    const int r = idx / h1, h = idx - r * h1;
    if (r < K2 - 1) gM2[idx] += dW2[idx]; else gb2[h] += dW2[idx];
  } else if (idx < n2 + n3) {
    const int q = idx - n2; const int r = q / h2, h = q - r * h2;
    if (r < K3 - 1) gM3[q] += dW3[q]; else gb3[h] += dW3[q];
  }
}
// scatter the gradient of WQ and the w4 / w6 sums back into M1 / b1 (once per step and layer)
__global__ void sgc_unpack_grads_k(const float* __restrict__ dWQ, const float* __restrict__ w46, float* __restrict__ gM1,
                                   float* __restrict__ gb1, SgcDims D) {
  const int C = D.C, h0 = D.h0;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (2 * C + 4) * h0) return;
  const int r = idx / h0, h = idx - r * h0;
  if (r < 2 * C) gM1[(C + r) * h0 + h] += dWQ[idx];
  else if (r == 2 * C) gM1[(3 * C + 1) * h0 + h] += dWQ[idx];
  else if (r == 2 * C + 1) gb1[h] += dWQ[idx];
  else if (r == 2 * C + 2) gM1[(3 * C) * h0 + h] += w46[h];
  else gM1[(3 * C + 2) * h0 + h] += w46[h0 + h];
}
