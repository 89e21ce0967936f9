// sgc3.cuh -- SpatialGraphConvolution_3D (layers.py:200-277; the `protein` / `mnist` branch of the joint encoder,
// model.py:139-140, flag blocks main.py:218-241) in its exact factored form.
//
// As written the layer materialises a [B, N, N, N, N, 4C+5] tensor.  Nothing nonlinear sits between Matrix0 and the
// adjacency-weighted p-sum, so the N^4 level collapses to closed forms per (i, j, k); the k-sum keeps one N^3 h0 pointwise term
// and the j-sum one N^2 h1 term.  With phi = lrelu, px = phi(x), pr = phi(rel), deg_k = sum_p A_kp, s_k = sum_p A_kp pr_kp,
// G_ik = sum_p A_kp pr_ip, a1..a5 = the five scalar rows of Matrix0 (r_ij, r_jk, r_kp, r_ik, r_ip), c1..c3 those of Matrix1:
//   alpha_k  = deg_k (px_k M0c + b0) + (A (px M0d))_k + s_k a3
//   m4s_ijk  = A_ij A_jk [ deg_k (px_i M0a + px_j M0b + pr_ij a1 + pr_jk a2 + pr_ik a4) + alpha_k + G_ik a5 ]
//   T3_ij    = sum_k A_jk phi(m4s_ijk)
//   beta_j   = deg_j (px_j M1b + b1) + (A (px M1c))_j + s_j c2
//   m3s_ij   = A_ij [ deg_j (px_i M1a + pr_ij c1) + beta_j + G_ij c3 + T3_ij M1e ]
//   T2_i     = sum_j A_ij phi(m3s_ij)
//   m2s_i    = deg_i (px_i M2a + b2) + (A (px M2b))_i + s_i M2c + T2_i M2d
//   y_i      = phi([x_i || m2s_i]) M3 + b3
// (factored == literal N^4 form to 1e-12 in fp64: tests/test_sgc3d_factored_equals_literal).  Holds for any real A -- the protein contact maps
// are not forests -- so these kernels take the dense [N, N] adjacency.
//
// One CTA per sample, phases separated by block barriers, every intermediate in a per-sample slice of a global workspace
// (L1 / L2 resident at these sizes: the reference's own data sets have N <= 50).  The backward kernel recomputes the forward
// intermediates of its sample and walks the four levels back: parameter gradients leave as one atomicAdd per parameter and
// sample, the input gradient as dx.  O(N^3 h0 + N^2 h0 h1) per sample either way.
#pragma once
#include "common.cuh"

struct Sgc3Dims { int C, h0, h1, h2, h3; };
struct Sgc3Params {       // pointers into the parameter (or gradient) arena
  float *M0, *b0, *M1, *b1, *M2, *b2, *M3, *b3;
};

// floats of workspace per sample
static inline long long sgc3_ws_floats(int N, Sgc3Dims d) {
  const long long n = N, hm = (long long)(d.h0 > d.h1 ? d.h0 : d.h1) > d.h2 ? (d.h0 > d.h1 ? d.h0 : d.h1) : d.h2;
  return n * n * (2 + 2 * d.h0 + d.h1) + n * (2 * d.C + 2 + 8 * d.h0 + 8 * d.h1 + 5 * d.h2) + 16 * (d.h0 + d.h1) + hm + 64;
}

#define S3_FOR(var, n) for (int var = threadIdx.x; var < (n); var += blockDim.x)

template <bool BWD>
__global__ void __launch_bounds__(256) sgc3_k(const float* __restrict__ xin, const float* __restrict__ adj, const float* __restrict__ rel,
                                              Sgc3Params W, Sgc3Params dW, Sgc3Dims D, int N, float* __restrict__ yout,
                                              const float* __restrict__ dyin, float* __restrict__ dxout, float* __restrict__ ws, long long ws_stride) {
  const long long smp = blockIdx.x;
  const int C = D.C, h0 = D.h0, h1 = D.h1, h2 = D.h2, h3 = D.h3;
  const float* x = xin + smp * N * C;
  const float* A = adj + smp * N * N;
  const float* R = rel + smp * N * N;
  float* w = ws + smp * ws_stride;
  // workspace carve-up
  float* px = w;            w += N * C;
  float* pr = w;            w += N * N;
  float* G = w;             w += N * N;
  float* deg = w;           w += N;
  float* ssum = w;          w += N;
  float* P0 = w;            w += N * h0;
  float* Q0 = w;            w += N * h0;
  float* S0 = w;            w += N * h0;
  float* alpha = w;         w += N * h0;
  float* T3 = w;            w += (long long)N * N * h0;
  float* P1 = w;            w += N * h1;
  float* R1 = w;            w += N * h1;
  float* beta = w;          w += N * h1;
  float* T2 = w;            w += N * h1;
  float* V = w;             w += N * h2;
  float* m2s = w;           w += N * h2;
  // backward-only
  float* dm2s = w;          w += N * h2;
  float* dV = w;            w += N * h2;
  float* dT2 = w;           w += N * h1;
  float* dP1 = w;           w += N * h1;
  float* dbeta = w;         w += N * h1;
  float* dR1 = w;           w += N * h1;
  float* g3 = w;            w += (long long)N * N * h1;
  float* dT3 = w;           w += (long long)N * N * h0;
  float* dP0 = w;           w += N * h0;
  float* dQ0 = w;           w += N * h0;
  float* dalpha = w;        w += N * h0;
  float* dS0 = w;           w += N * h0;
  float* dpx = w;           w += N * C;
  float* dvec = w;          w += 16 * (h0 + h1);      // da1, da2, da3, da4, da5 [h0 each]; dc1, dc2, dc3 [h1 each]

  const float* M0 = W.M0; const float* M1 = W.M1; const float* M2 = W.M2; const float* M3 = W.M3;
  const float* a1 = M0 + (size_t)(4 * C) * h0; const float* a2 = a1 + h0; const float* a3 = a2 + h0; const float* a4 = a3 + h0; const float* a5 = a4 + h0;
  const float* c1 = M1 + (size_t)(3 * C) * h1; const float* c2 = c1 + h1; const float* c3 = c2 + h1; const float* M1e = c3 + h1;
  const float* M2c = M2 + (size_t)(2 * C) * h2; const float* M2d = M2c + h2;

  // ---- F1-F3: phi(x), phi(rel), degrees, s, G --------------------------------------------------------------------
  S3_FOR(t, N * C) px[t] = lrelu_f(x[t]);
  S3_FOR(t, N * N) pr[t] = lrelu_f(R[t]);
  __syncthreads();
  S3_FOR(k, N) {
    float d = 0.f, s = 0.f;
    for (int p = 0; p < N; ++p) { const float a = A[k * N + p]; d += a; s = fmaf(a, pr[k * N + p], s); }
    deg[k] = d; ssum[k] = s;
  }
  S3_FOR(t, N * N) {
    const int i = t / N, k = t - i * N;
    float g = 0.f;
    for (int p = 0; p < N; ++p) g = fmaf(A[k * N + p], pr[i * N + p], g);
    G[t] = g;
  }
  // ---- F4: node products of level 4 ------------------------------------------------------------------------------
  S3_FOR(t, N * h0) {
    const int n = t / h0, h = t - n * h0;
    float p = 0.f, q = 0.f, r = 0.f, s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = px[n * C + c];
      p = fmaf(v, M0[(size_t)c * h0 + h], p); q = fmaf(v, M0[(size_t)(C + c) * h0 + h], q);
      r = fmaf(v, M0[(size_t)(2 * C + c) * h0 + h], r); s = fmaf(v, M0[(size_t)(3 * C + c) * h0 + h], s);
    }
    P0[t] = p; Q0[t] = q; alpha[t] = r; S0[t] = s;            // alpha holds R0 until the next phase
  }
  __syncthreads();
  S3_FOR(t, N * h0) {
    const int k = t / h0, h = t - k * h0;
    float as = 0.f;
    for (int p = 0; p < N; ++p) as = fmaf(A[k * N + p], S0[p * h0 + h], as);
    alpha[t] = deg[k] * (alpha[t] + W.b0[h]) + as + ssum[k] * a3[h];
  }
  __syncthreads();
  // ---- F6: T3_ij = sum_k A_jk phi(m4s_ijk) ------------------------------------------------------------------------
  for (int t = threadIdx.x; t < N * N * h0; t += blockDim.x) {
    const int h = t % h0, ij = t / h0, i = ij / N, j = ij - i * N;
    const float aij = A[ij];
    float acc = 0.f;
    if (aij != 0.f) {
      const float base = P0[i * h0 + h] + Q0[j * h0 + h] + pr[ij] * a1[h];
      for (int k = 0; k < N; ++k) {
        const float ajk = A[j * N + k];
        if (ajk == 0.f) continue;
        const float in4 = deg[k] * (base + pr[j * N + k] * a2[h] + pr[i * N + k] * a4[h]) + alpha[k * h0 + h] + G[i * N + k] * a5[h];
        acc = fmaf(ajk, lrelu_f(aij * ajk * in4), acc);
      }
    }
    T3[t] = acc;
  }
  // ---- F7-F8: node products of level 3 ----------------------------------------------------------------------------
  S3_FOR(t, N * h1) {
    const int n = t / h1, h = t - n * h1;
    float p = 0.f, q = 0.f, r = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = px[n * C + c];
      p = fmaf(v, M1[(size_t)c * h1 + h], p); q = fmaf(v, M1[(size_t)(C + c) * h1 + h], q); r = fmaf(v, M1[(size_t)(2 * C + c) * h1 + h], r);
    }
    P1[t] = p; beta[t] = q; R1[t] = r;                        // beta holds Q1 until the next phase
  }
  __syncthreads();
  S3_FOR(t, N * h1) {
    const int j = t / h1, h = t - j * h1;
    float ar = 0.f;
    for (int k = 0; k < N; ++k) ar = fmaf(A[j * N + k], R1[k * h1 + h], ar);
    beta[t] = deg[j] * (beta[t] + W.b1[h]) + ar + ssum[j] * c2[h];
  }
  __syncthreads();
  // ---- F9: T2_i = sum_j A_ij phi(m3s_ij)   (backward: g3_ij = A_ij^2 dT2_i phi'(m3s_ij) is formed in the same loop later) ----
  S3_FOR(t, N * h1) {
    const int i = t / h1, h = t - i * h1;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) {
      const float aij = A[i * N + j];
      if (aij == 0.f) continue;
      float tm = 0.f;
      const float* t3 = T3 + (size_t)(i * N + j) * h0;
      for (int q = 0; q < h0; ++q) tm = fmaf(t3[q], M1e[(size_t)q * h1 + h], tm);
      const float in3 = deg[j] * (P1[t] + pr[i * N + j] * c1[h]) + beta[j * h1 + h] + G[i * N + j] * c3[h] + tm;
      acc = fmaf(aij, lrelu_f(aij * in3), acc);
    }
    T2[t] = acc;
  }
  // ---- F10: m2s ---------------------------------------------------------------------------------------------------
  S3_FOR(t, N * h2) {
    const int n = t / h2, h = t - n * h2;
    float v = 0.f;
    for (int c = 0; c < C; ++c) v = fmaf(px[n * C + c], M2[(size_t)(C + c) * h2 + h], v);
    V[t] = v;
  }
  __syncthreads();
  S3_FOR(t, N * h2) {
    const int i = t / h2, h = t - i * h2;
    float u = 0.f, av = 0.f, td = 0.f;
    for (int c = 0; c < C; ++c) u = fmaf(px[i * C + c], M2[(size_t)c * h2 + h], u);
    for (int j = 0; j < N; ++j) av = fmaf(A[i * N + j], V[j * h2 + h], av);
    for (int q = 0; q < h1; ++q) td = fmaf(T2[i * h1 + q], M2d[(size_t)q * h2 + h], td);
    m2s[t] = deg[i] * (u + W.b2[h]) + av + ssum[i] * M2c[h] + td;
  }
  __syncthreads();
  if (!BWD) {
    // ---- F11: y = phi([x || m2s]) M3 + b3 ------------------------------------------------------------------------
    float* y = yout + smp * N * h3;
    S3_FOR(t, N * h3) {
      const int i = t / h3, h = t - i * h3;
      float acc = W.b3[h];
      for (int c = 0; c < C; ++c) acc = fmaf(px[i * C + c], M3[(size_t)c * h3 + h], acc);
      for (int q = 0; q < h2; ++q) acc = fmaf(lrelu_f(m2s[i * h2 + q]), M3[(size_t)(C + q) * h3 + h], acc);
      y[t] = acc;
    }
    return;
  }

  // ================================================== backward ==================================================
  const float* dy = dyin + smp * N * h3;
  S3_FOR(t, 16 * (h0 + h1)) dvec[t] = 0.f;
  S3_FOR(t, N * h0) { dP0[t] = 0.f; dQ0[t] = 0.f; dalpha[t] = 0.f; }
  S3_FOR(t, N * C) dpx[t] = 0.f;
  float* da1 = dvec; float* da2 = da1 + h0; float* da3 = da2 + h0; float* da4 = da3 + h0; float* da5 = da4 + h0;
  float* dc1 = da5 + h0; float* dc2 = dc1 + h1; float* dc3 = dc2 + h1;
  // ---- B1: Matrix3 / bias3, the concat ---------------------------------------------------------------------------------
  S3_FOR(t, (C + h2) * h3) {
    const int r = t / h3, h = t - r * h3;
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc = fmaf(r < C ? px[i * C + r] : lrelu_f(m2s[i * h2 + r - C]), dy[i * h3 + h], acc);
    atomicAdd(dW.M3 + t, acc);
  }
  S3_FOR(h, h3) { float acc = 0.f; for (int i = 0; i < N; ++i) acc += dy[i * h3 + h]; atomicAdd(dW.b3 + h, acc); }
  S3_FOR(t, N * h2) {
    const int i = t / h2, q = t - i * h2;
    float acc = 0.f;
    for (int h = 0; h < h3; ++h) acc = fmaf(dy[i * h3 + h], M3[(size_t)(C + q) * h3 + h], acc);
    dm2s[t] = acc * lrelu_g(m2s[t]);
  }
  __syncthreads();
  // ---- B2: level 2 ------------------------------------------------------------------------------------------------------
  S3_FOR(t, N * h2) {          // dV_j = sum_i A_ij dm2s_i
    const int j = t / h2, h = t - j * h2;
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc = fmaf(A[i * N + j], dm2s[i * h2 + h], acc);
    dV[t] = acc;
  }
  S3_FOR(t, N * h1) {          // dT2 = dm2s M2d^T
    const int i = t / h1, q = t - i * h1;
    float acc = 0.f;
    for (int h = 0; h < h2; ++h) acc = fmaf(dm2s[i * h2 + h], M2d[(size_t)q * h2 + h], acc);
    dT2[t] = acc;
  }
  S3_FOR(h, h2) {
    float sb = 0.f, sc = 0.f;
    for (int i = 0; i < N; ++i) { sb = fmaf(deg[i], dm2s[i * h2 + h], sb); sc = fmaf(ssum[i], dm2s[i * h2 + h], sc); }
    atomicAdd(dW.b2 + h, sb); atomicAdd(dW.M2 + (size_t)(2 * C) * h2 + h, sc);
  }
  S3_FOR(t, h1 * h2) {         // dM2d = T2^T dm2s
    const int q = t / h2, h = t - q * h2;
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc = fmaf(T2[i * h1 + q], dm2s[i * h2 + h], acc);
    atomicAdd(dW.M2 + (size_t)(2 * C + 1 + q) * h2 + h, acc);
  }
  __syncthreads();
  S3_FOR(t, C * h2) {          // dM2a = (deg px)^T dm2s, dM2b = px^T dV
    const int c = t / h2, h = t - c * h2;
    float sa = 0.f, sb = 0.f;
    for (int i = 0; i < N; ++i) { sa = fmaf(deg[i] * px[i * C + c], dm2s[i * h2 + h], sa); sb = fmaf(px[i * C + c], dV[i * h2 + h], sb); }
    atomicAdd(dW.M2 + (size_t)c * h2 + h, sa); atomicAdd(dW.M2 + (size_t)(C + c) * h2 + h, sb);
  }
  S3_FOR(t, N * C) {
    const int n = t / C, c = t - n * C;
    float acc = 0.f;
    for (int h = 0; h < h2; ++h) acc += deg[n] * dm2s[n * h2 + h] * M2[(size_t)c * h2 + h] + dV[n * h2 + h] * M2[(size_t)(C + c) * h2 + h];
    dpx[t] += acc;
  }
  // ---- B3: level 3.  g3_ij = A_ij^2 dT2_i phi'(m3s_ij) ---------------------------------------------------------------------
  for (int t = threadIdx.x; t < N * N * h1; t += blockDim.x) {
    const int h = t % h1, ij = t / h1, i = ij / N, j = ij - i * N;
    const float aij = A[ij];
    float g = 0.f;
    if (aij != 0.f) {
      float tm = 0.f;
      const float* t3 = T3 + (size_t)ij * h0;
      for (int q = 0; q < h0; ++q) tm = fmaf(t3[q], M1e[(size_t)q * h1 + h], tm);
      const float in3 = deg[j] * (P1[i * h1 + h] + pr[ij] * c1[h]) + beta[j * h1 + h] + G[ij] * c3[h] + tm;
      g = aij * aij * dT2[i * h1 + h] * lrelu_g(aij * in3);
    }
    g3[t] = g;
  }
  __syncthreads();
  S3_FOR(t, N * h1) {
    const int n = t / h1, h = t - n * h1;
    float sp = 0.f, sb = 0.f;
    for (int j = 0; j < N; ++j) sp = fmaf(deg[j], g3[(size_t)(n * N + j) * h1 + h], sp);         // dP1_n = sum_j deg_j g3_nj
    for (int i = 0; i < N; ++i) sb += g3[(size_t)(i * N + n) * h1 + h];                           // dbeta_n = sum_i g3_in
    dP1[t] = sp; dbeta[t] = sb;
  }
  S3_FOR(h, h1) {
    float s1 = 0.f, s3 = 0.f;
    for (int ij = 0; ij < N * N; ++ij) {
      const float g = g3[(size_t)ij * h1 + h];
      s1 = fmaf(deg[ij % N] * pr[ij], g, s1); s3 = fmaf(G[ij], g, s3);
    }
    dc1[h] = s1; dc3[h] = s3;
  }
  for (int t = threadIdx.x; t < N * N * h0; t += blockDim.x) {      // dT3_ij = g3_ij M1e^T
    const int q = t % h0, ij = t / h0;
    float acc = 0.f;
    const float* g = g3 + (size_t)ij * h1;
    for (int h = 0; h < h1; ++h) acc = fmaf(g[h], M1e[(size_t)q * h1 + h], acc);
    dT3[t] = acc;
  }
  S3_FOR(t, h0 * h1) {         // dM1e = sum_ij T3_ij^T g3_ij
    const int q = t / h1, h = t - q * h1;
    float acc = 0.f;
    for (int ij = 0; ij < N * N; ++ij) acc = fmaf(T3[(size_t)ij * h0 + q], g3[(size_t)ij * h1 + h], acc);
    atomicAdd(dW.M1 + (size_t)(3 * C + 3 + q) * h1 + h, acc);
  }
  __syncthreads();
  S3_FOR(t, N * h1) {          // dR1_k = sum_j A_jk dbeta_j
    const int k = t / h1, h = t - k * h1;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(A[j * N + k], dbeta[j * h1 + h], acc);
    dR1[t] = acc;
  }
  S3_FOR(h, h1) {
    float sb = 0.f, s2 = 0.f;
    for (int j = 0; j < N; ++j) { sb = fmaf(deg[j], dbeta[j * h1 + h], sb); s2 = fmaf(ssum[j], dbeta[j * h1 + h], s2); }
    atomicAdd(dW.b1 + h, sb); dc2[h] = s2;
  }
  __syncthreads();
  S3_FOR(t, C * h1) {
    const int c = t / h1, h = t - c * h1;
    float sa = 0.f, sb = 0.f, sc = 0.f;
    for (int n = 0; n < N; ++n) {
      const float v = px[n * C + c];
      sa = fmaf(v, dP1[n * h1 + h], sa); sb = fmaf(v * deg[n], dbeta[n * h1 + h], sb); sc = fmaf(v, dR1[n * h1 + h], sc);
    }
    atomicAdd(dW.M1 + (size_t)c * h1 + h, sa); atomicAdd(dW.M1 + (size_t)(C + c) * h1 + h, sb); atomicAdd(dW.M1 + (size_t)(2 * C + c) * h1 + h, sc);
  }
  S3_FOR(h, h1) {
    atomicAdd(dW.M1 + (size_t)(3 * C) * h1 + h, dc1[h]); atomicAdd(dW.M1 + (size_t)(3 * C + 1) * h1 + h, dc2[h]);
    atomicAdd(dW.M1 + (size_t)(3 * C + 2) * h1 + h, dc3[h]);
  }
  S3_FOR(t, N * C) {
    const int n = t / C, c = t - n * C;
    float acc = 0.f;
    for (int h = 0; h < h1; ++h)
      acc += dP1[n * h1 + h] * M1[(size_t)c * h1 + h] + deg[n] * dbeta[n * h1 + h] * M1[(size_t)(C + c) * h1 + h] + dR1[n * h1 + h] * M1[(size_t)(2 * C + c) * h1 + h];
    dpx[t] += acc;
  }
  // ---- B4: level 4.  g4_ijk = A_ij A_jk^2 dT3_ij phi'(m4s_ijk) --------------------------------------------------------------
  for (int t = threadIdx.x; t < N * N * h0; t += blockDim.x) {
    const int h = t % h0, ij = t / h0, i = ij / N, j = ij - i * N;
    const float aij = A[ij], d3 = dT3[t];
    if (aij == 0.f || d3 == 0.f) continue;
    const float base = P0[i * h0 + h] + Q0[j * h0 + h] + pr[ij] * a1[h];
    float sPQ = 0.f, s1 = 0.f, s2 = 0.f, s4 = 0.f, s5 = 0.f;
    for (int k = 0; k < N; ++k) {
      const float ajk = A[j * N + k];
      if (ajk == 0.f) continue;
      const float in4 = deg[k] * (base + pr[j * N + k] * a2[h] + pr[i * N + k] * a4[h]) + alpha[k * h0 + h] + G[i * N + k] * a5[h];
      const float g = aij * ajk * ajk * d3 * lrelu_g(aij * ajk * in4);
      const float dg = deg[k] * g;
      sPQ += dg; s1 = fmaf(pr[ij], dg, s1); s2 = fmaf(pr[j * N + k], dg, s2); s4 = fmaf(pr[i * N + k], dg, s4); s5 = fmaf(G[i * N + k], g, s5);
      atomicAdd(dalpha + k * h0 + h, g);
    }
    atomicAdd(dP0 + i * h0 + h, sPQ); atomicAdd(dQ0 + j * h0 + h, sPQ);
    atomicAdd(da1 + h, s1); atomicAdd(da2 + h, s2); atomicAdd(da4 + h, s4); atomicAdd(da5 + h, s5);
  }
  __syncthreads();
  S3_FOR(t, N * h0) {          // dS0_p = sum_k A_kp dalpha_k
    const int p = t / h0, h = t - p * h0;
    float acc = 0.f;
    for (int k = 0; k < N; ++k) acc = fmaf(A[k * N + p], dalpha[k * h0 + h], acc);
    dS0[t] = acc;
  }
  S3_FOR(h, h0) {
    float sb = 0.f, s3 = 0.f;
    for (int k = 0; k < N; ++k) { sb = fmaf(deg[k], dalpha[k * h0 + h], sb); s3 = fmaf(ssum[k], dalpha[k * h0 + h], s3); }
    atomicAdd(dW.b0 + h, sb); da3[h] = s3;
  }
  __syncthreads();
  S3_FOR(t, C * h0) {
    const int c = t / h0, h = t - c * h0;
    float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f;
    for (int n = 0; n < N; ++n) {
      const float v = px[n * C + c];
      sa = fmaf(v, dP0[n * h0 + h], sa); sb = fmaf(v, dQ0[n * h0 + h], sb); sc = fmaf(v * deg[n], dalpha[n * h0 + h], sc); sd = fmaf(v, dS0[n * h0 + h], sd);
    }
    atomicAdd(dW.M0 + (size_t)c * h0 + h, sa); atomicAdd(dW.M0 + (size_t)(C + c) * h0 + h, sb);
    atomicAdd(dW.M0 + (size_t)(2 * C + c) * h0 + h, sc); atomicAdd(dW.M0 + (size_t)(3 * C + c) * h0 + h, sd);
  }
  S3_FOR(h, h0) {
    float* r = dW.M0 + (size_t)(4 * C) * h0 + h;
    atomicAdd(r, da1[h]); atomicAdd(r + h0, da2[h]); atomicAdd(r + 2 * h0, da3[h]); atomicAdd(r + 3 * h0, da4[h]); atomicAdd(r + 4 * h0, da5[h]);
  }
  if (dxout) {
    float* dx = dxout + smp * N * C;
    S3_FOR(t, N * C) {
      const int n = t / C, c = t - n * C;
      float acc = 0.f;
      for (int h = 0; h < h0; ++h)
        acc += dP0[n * h0 + h] * M0[(size_t)c * h0 + h] + dQ0[n * h0 + h] * M0[(size_t)(C + c) * h0 + h] +
               deg[n] * dalpha[n * h0 + h] * M0[(size_t)(2 * C + c) * h0 + h] + dS0[n * h0 + h] * M0[(size_t)(3 * C + c) * h0 + h];
      float dd = 0.f;                                           // direct path through the concat of the last map
      for (int h = 0; h < h3; ++h) dd = fmaf(dy[n * h3 + h], M3[(size_t)c * h3 + h], dd);
      dx[t] = (dd + dpx[t] + acc) * lrelu_g(x[t]);
    }
  }
}
#undef S3_FOR
