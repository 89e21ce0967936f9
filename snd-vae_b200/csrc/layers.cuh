// layers.cuh -- node-level layers of the SND-VAE encoders / decoders (fp32 SIMT).
// These stages are O(N * C^2) per graph and together well under 1% of a train
// step at N=256; they are written for clarity and coalescing, the
// heavy N^2 stages live in sgc.cuh / edge.cuh / e2e_tc.cuh.
// All activations are row-major [rows, C] with rows = (graph, node).
#pragma once
#include "common.cuh"

// ---- per-node linear:  out[r, co] = act(bias[co] + sum_ci in[r, ci] W[ci, co]) ---------
// (layers.py:566-576 `linear` on [B*N, C] rows; GraphConvolution's x.w, layers.py:121)
__global__ void rowlin_fwd_k(const float* __restrict__ in, int ldi, const float* __restrict__ W,
                             const float* __restrict__ bias, float* __restrict__ out, int ldo,
                             long long rows, int Ci, int Co, int act) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Co) return;
  long long r = idx / Co; int co = (int)(idx - r * Co);
  float acc = bias ? bias[co] : 0.f;
  const float* x = in + r * ldi;
  for (int ci = 0; ci < Ci; ++ci) acc = fmaf(x[ci], W[ci * Co + co], acc);
  out[r * ldo + co] = act_f(act, acc);
}

// din[r, ci] (+)= sum_co dout[r, co] W[ci, co]
__global__ void rowlin_bwd_in_k(const float* __restrict__ dout, int ldd, const float* __restrict__ W,
                                float* __restrict__ din, int ldi, long long rows, int Ci, int Co, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Ci) return;
  long long r = idx / Ci; int ci = (int)(idx - r * Ci);
  const float* d = dout + r * ldd;
  float acc = 0.f;
  for (int co = 0; co < Co; ++co) acc = fmaf(d[co], W[ci * Co + co], acc);
  if (accumulate) din[r * ldi + ci] += acc; else din[r * ldi + ci] = acc;
}

// ---- conv1d over the node axis, k taps, SAME zero padding (model.py:122,191,216) -------
// out[b, n, co] = bias[co] + sum_t sum_ci in[b, n + t - pb, ci] K[t, ci, co],  pb = (k-1)/2
__global__ void conv1d_fwd_k(const float* __restrict__ in, const float* __restrict__ K,
                             const float* __restrict__ bias, float* __restrict__ out,
                             long long rows, int N, int Ci, int Co, int ktaps) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Co) return;
  long long r = idx / Co; int co = (int)(idx - r * Co);
  int n = (int)(r % N);
  int pb = (ktaps - 1) / 2;
  float acc = bias[co];
  for (int t = 0; t < ktaps; ++t) {
    int nn = n + t - pb;
    if (nn < 0 || nn >= N) continue;
    const float* x = in + (r + t - pb) * Ci;
    const float* w = K + (size_t)t * Ci * Co + co;
    for (int ci = 0; ci < Ci; ++ci) acc = fmaf(x[ci], w[ci * Co], acc);
  }
  out[idx] = acc;
}

// din[b, n', ci] = sum_t sum_co dout[b, n' - t + pb, co] K[t, ci, co]
__global__ void conv1d_bwd_in_k(const float* __restrict__ dout, const float* __restrict__ K,
                                float* __restrict__ din, long long rows, int N, int Ci, int Co, int ktaps) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Ci) return;
  long long r = idx / Ci; int ci = (int)(idx - r * Ci);
  int n = (int)(r % N);
  int pb = (ktaps - 1) / 2;
  float acc = 0.f;
  for (int t = 0; t < ktaps; ++t) {
    int nn = n - t + pb;
    if (nn < 0 || nn >= N) continue;
    const float* d = dout + (r - t + pb) * Co;
    const float* w = K + ((size_t)t * Ci + ci) * Co;
    for (int co = 0; co < Co; ++co) acc = fmaf(d[co], w[co], acc);
  }
  din[idx] = acc;
}

// ---- weight gradients:  dW[t, ci, co] += sum_r X[r + t - pb, ci] dY[r, co]  ------------
// (ktaps = 1: plain X^T dY for `linear` / graph-conv weights).  Each block owns a
// slab of rows, each thread a strided set of (t, ci, co); one atomicAdd per
// (block, parameter).
#define XTDY_SLAB 512
// grid of a row-slab reduction kernel: slabs of at most `max_slab` rows, but at least ~4 blocks per SM when the problem is small
// (at N = 25, B = 32 a fixed 512-row slab left 2 blocks on 148 SMs)
static inline unsigned slab_grid(long long rows, int max_slab) {
  long long slab = rows / (148 * 4); if (slab < 8) slab = 8; if (slab > max_slab) slab = max_slab;
  return (unsigned)((rows + slab - 1) / slab);
}
__global__ void xtdy_k(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy,
                       float* __restrict__ dW, long long rows, int N, int Ci, int Co, int ktaps) {
  const long long slab = (rows + gridDim.x - 1) / gridDim.x;      // rows per block: the launch picks the grid (slab_grid)
  long long r0 = (long long)blockIdx.x * slab;
  long long r1 = r0 + slab; if (r1 > rows) r1 = rows;
  int pb = (ktaps - 1) / 2;
  int np = ktaps * Ci * Co;
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    int co = p % Co; int q = p / Co; int ci = q % Ci; int t = q / Ci;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) {
      int nn = (int)(r % N) + t - pb;
      if (nn < 0 || nn >= N) continue;
      acc = fmaf(X[(r + t - pb) * ldx + ci], dY[r * ldy + co], acc);
    }
    atomicAdd(dW + p, acc);
  }
}

// col[r, t*Ci + ci] = X[r + t - pb, ci] (zero outside the graph): the conv1d weight gradient is then
// the library GEMM  dK[(t,ci), co] += col^T . dY
__global__ void im2col_k(const float* __restrict__ X, float* __restrict__ col, long long rows, int N, int Ci, int ktaps) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = ktaps * Ci;
  if (idx >= rows * W) return;
  long long r = idx / W; int q = (int)(idx - r * W); int t = q / Ci, ci = q - t * Ci;
  int pb = (ktaps - 1) / 2;
  int nn = (int)(r % N) + t - pb;
  col[idx] = (nn >= 0 && nn < N) ? X[(r + t - pb) * Ci + ci] : 0.f;
}

// transpose of im2col_k: dX[r, ci] = sum_t dcol[r - t + pb, t, ci] over the taps whose source row is in the same graph
__global__ void col2im_k(const float* __restrict__ dcol, float* __restrict__ dX, long long rows, int N, int Ci, int ktaps) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Ci) return;
  long long r = idx / Ci; int ci = (int)(idx - r * Ci);
  const int n = (int)(r % N), pb = (ktaps - 1) / 2, W = ktaps * Ci;
  float acc = 0.f;
  for (int t = 0; t < ktaps; ++t) {
    const int nn = n - t + pb;
    if (nn >= 0 && nn < N) acc += dcol[(r - t + pb) * W + t * Ci + ci];
  }
  dX[idx] = acc;
}

// db[c] += sum_r dY[r, c];  grid = (row slabs, column tiles of blockDim.x)
__global__ void colsum_k(const float* __restrict__ dY, int ldy, float* __restrict__ db, long long rows, int C) {
  const long long slab = (rows + gridDim.x - 1) / gridDim.x;      // rows per block: the launch picks the grid (slab_grid)
  long long r0 = (long long)blockIdx.x * slab;
  long long r1 = r0 + slab; if (r1 > rows) r1 = rows;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += dY[r * ldy + c];
  atomicAdd(db + c, acc);
}

// ---- Keras BN (inference affine, SURVEY finding 3) fused with an activation ------------
// order 0: out = act(bn(in))   (model.py:122-123,146: relu/lrelu after BN)
// order 1: out = bn(act(in))   (model.py:107: BN of GraphConvolution's lrelu output)
__global__ void bn_act_fwd_k(const float* __restrict__ in, int ldi, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ out, int ldo,
                             long long rows, int C, int act, int order) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  long long r = idx / C; int c = (int)(idx - r * C);
  float x = in[r * ldi + c];
  float g = gamma ? gamma[c] * BN_RS : 1.f, b = beta ? beta[c] : 0.f;
  float y = order == 0 ? act_f(act, fmaf(x, g, b)) : fmaf(act_f(act, x), g, b);
  out[r * ldo + c] = y;
}

// backward of the above.  din may alias dout.  dgamma/dbeta accumulate (atomics).
#define BN_SLAB 256
__global__ void bn_act_bwd_k(const float* __restrict__ dout, int ldd, const float* __restrict__ in, int ldi,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             float* din, int ldn, float* __restrict__ dgamma, float* __restrict__ dbeta,
                             long long rows, int C, int act, int order) {
  const long long slab = (rows + gridDim.x - 1) / gridDim.x;
  long long r0 = (long long)blockIdx.x * slab;
  long long r1 = r0 + slab; if (r1 > rows) r1 = rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float g = gamma ? gamma[c] * BN_RS : 1.f, b = beta ? beta[c] : 0.f;
    float sg = 0.f, sb = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float x = in[r * ldi + c];
      float d = dout[r * ldd + c];
      float dx;
      if (order == 0) {
        float pre = fmaf(x, g, b);
        float dd = d * act_g(act, pre);
        sg += dd * x; sb += dd; dx = dd * g;
      } else {
        sg += d * act_f(act, x); sb += d; dx = d * g * act_g(act, x);
      }
      if (din) din[r * ldn + c] = dx;
    }
    if (dgamma) { atomicAdd(dgamma + c, sg * BN_RS); atomicAdd(dbeta + c, sb); }
  }
}

// copy a column block:  out[r, oc + c] = in[r, ic + c]  (concat / split)
__global__ void copy_cols_k(const float* __restrict__ in, int ldi, int ic, float* __restrict__ out, int ldo, int oc,
                            long long rows, int C, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  long long r = idx / C; int c = (int)(idx - r * C);
  float v = in[r * ldi + ic + c];
  if (accumulate) out[r * ldo + oc + c] += v; else out[r * ldo + oc + c] = v;
}

// ---- GraphConvolution propagate (layers.py:122): c[b,i,h] = sum_j A[b,i,j] t[b,j,h] ----
// one warp per (graph,row); lanes stride over j (coalesced A reads), zeros skipped.
template <int HMAX>
__global__ void graph_prop_fwd_k(const float* __restrict__ A, const float* __restrict__ t,
                                 float* __restrict__ c, long long rows, int N, int H) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long b = row / N;
  const float* a = A + row * N;
  const float* tb = t + b * N * H;
  float acc[HMAX];
#pragma unroll
  for (int h = 0; h < HMAX; ++h) acc[h] = 0.f;
  for (int j = lane; j < N; j += 32) {
    float av = a[j];
    if (av != 0.f) {
      const float* tj = tb + (size_t)j * H;
#pragma unroll
      for (int h = 0; h < HMAX; ++h) if (h < H) acc[h] = fmaf(av, tj[h], acc[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < HMAX; ++h) {
    if (h < H) {
      float v = warp_sum(acc[h]);
      if (lane == 0) c[row * H + h] = v;
    }
  }
}

// dt[b,j,h] += sum_i A[b,i,j] dc[b,i,h]   (dt must be zeroed by the caller)
__global__ void graph_prop_bwd_k(const float* __restrict__ A, const float* __restrict__ dc,
                                 float* __restrict__ dt, long long rows, int N, int H) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long b = row / N;
  const float* a = A + row * N;
  const float* d = dc + row * H;
  float* tb = dt + b * N * H;
  for (int j = lane; j < N; j += 32) {
    float av = a[j];
    if (av != 0.f) {
      for (int h = 0; h < H; ++h) atomicAdd(tb + (size_t)j * H + h, av * d[h]);
    }
  }
}

// ---- reparameterisation + KL (model.py:153-161, optimizer.py:160-162) -----------------
// z = mu + eps * exp(ls);  kl_sum += -0.5 * sum(1 + 2 ls - mu^2 - exp(ls)^2)
__global__ void reparam_kl_k(const float* __restrict__ mu, const float* __restrict__ ls,
                             const float* __restrict__ eps, float* __restrict__ z,
                             float* __restrict__ kl_sum, long long n) {
  __shared__ float red[32];
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float k = 0.f;
  if (idx < n) {
    float m = mu[idx], l = ls[idx];
    float e = expf(l);
    z[idx] = fmaf(eps[idx], e, m);
    k = -0.5f * (1.f + 2.f * l - m * m - e * e);
  }
  k = block_sum(k, red);
  if (threadIdx.x == 0) atomicAdd(kl_sum, k);
}

// dmu = dz + w*mu;  dls = dz*eps*exp(ls) + w*(exp(2 ls) - 1),  w = beta/(rows_global*L)
// dz_rep > 1: dz row index = row / dz_rep and scaled by 1/dz_rep (the S-mean, model.py:180)
// gate_sum != NULL: the KL weight is w only while the batch-mean KL  gate_sum[0] * gate_inv  exceeds gate_C (the relu of the
// 'disentangled_C' capacity loss, optimizer.py:173), otherwise 0
__global__ void reparam_kl_bwd_k(const float* __restrict__ mu, const float* __restrict__ ls,
                                 const float* __restrict__ eps, const float* __restrict__ dz,
                                 float* __restrict__ dmu, float* __restrict__ dls,
                                 long long rows, int L, int dz_rep, float w,
                                 const float* __restrict__ gate_sum = nullptr, float gate_inv = 0.f, float gate_C = 0.f) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * L) return;
  if (gate_sum && !(gate_sum[0] * gate_inv > gate_C)) w = 0.f;
  long long r = idx / L; int c = (int)(idx - r * L);
  float g = dz[(r / dz_rep) * L + c] / (float)dz_rep;
  float m = mu[idx], l = ls[idx];
  float e = expf(l);
  dmu[idx] = g + w * m;
  dls[idx] = g * eps[idx] * e + w * (e * e - 1.f);
}

// zbar[b, c] = mean_s z[b*S + s, c]   (S-mean hoisted before d_sg_lin1; exact, linear)
__global__ void smean_k(const float* __restrict__ z, float* __restrict__ zbar, long long B, int S, int L) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * L) return;
  long long b = idx / L; int c = (int)(idx - b * L);
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += z[(b * S + s) * L + c];
  zbar[idx] = acc / (float)S;
}

// ---- sigmoid + MSE head (model.py:193,218; optimizer.py:149,153) -----------------------
// pred = sigmoid(pre); loss_sum += sum (target - pred)^2; dpre = 2 (pred - target) pred (1 - pred) * scale
__global__ void sigmoid_mse_k(const float* __restrict__ pre, const float* __restrict__ target,
                              float* __restrict__ pred, float* __restrict__ dpre,
                              float* __restrict__ loss_sum, long long n, float scale) {
  __shared__ float red[32];
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (idx < n) {
    float p = 1.f / (1.f + expf(-pre[idx]));
    pred[idx] = p;
    if (target) {
      float d = p - target[idx];
      l = d * d;
      if (dpre) dpre[idx] = 2.f * d * p * (1.f - p) * scale;
    }
  }
  l = block_sum(l, red);
  if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, l);
}

__global__ void add_inplace_k(float* __restrict__ a, const float* __restrict__ b, long long n) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) a[idx] += b[idx];
}

// ---- DIP-VAE-I regulariser on a batch of posterior means (optimizer.py:7-21) ---------------------------------------
// cov = E[mu mu^T] - m m^T (E over the rows);  reg = ld sum_i (cov_ii - 1)^2 + lod sum_{i != j} cov_ij^2
// in : S = mu^T mu [L, L] (unnormalised), msum [L] column sums.  out: G = d reg / d cov [L, L], m [L], reg_sum += reg
__global__ void dip_cov_k(const float* __restrict__ S, const float* __restrict__ msum, float* __restrict__ G, float* __restrict__ m,
                          float* __restrict__ reg_sum, int L, float inv_rows, float lod, float ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float r = 0.f;
  if (idx < L * L) {
    const int i = idx / L, j = idx - i * L;
    const float mi = msum[i] * inv_rows, mj = msum[j] * inv_rows;
    const float cov = S[idx] * inv_rows - mi * mj;
    if (i == j) { G[idx] = 2.f * ld * (cov - 1.f); r = ld * (cov - 1.f) * (cov - 1.f); m[i] = mi; }
    else { G[idx] = 2.f * lod * cov; r = lod * cov * cov; }
  }
  r = warp_sum(r);
  if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(reg_sum, r);
}
// dmu[r, :] -= alpha * v   (the mean term of d reg / d mu = (2 / rows) (mu - m) G, with v = m G)
__global__ void sub_row_k(float* __restrict__ dmu, const float* __restrict__ v, long long rows, int L, float alpha) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * L) dmu[idx] -= alpha * v[idx % L];
}

// ---- minibatch total correlation (beta-TCVAE; optimizer.py:23-63,185-190) ----------------------------------------
// lq(j,i,l) = -0.5 ((z_jl - mu_il)^2 p_il + 2 ls_il + log 2pi),  p = exp(-2 ls)   (log-variance = log(exp(ls)^2), optimizer.py:43)
// TC = mean_j [ lse_i sum_l lq(j,i,l)  -  sum_l lse_i lq(j,i,l) ]
// One CTA per row of the [rows, rows] pair grid, 4 warps striding over the other index; lane owns l = lane + 32 k.
#define TCOR_THREADS 128
#define TCOR_LK 4                      // L <= 128
#define TCOR_LOG2PI 1.8378770664093453f

__global__ void tc_prec_k(const float* __restrict__ ls, float* __restrict__ p, long long n) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) p[idx] = expf(-2.f * ls[idx]);
}

__device__ __forceinline__ void lse_push(float& mx, float& sm, float q) {
  if (q > mx) { sm = sm * expf(mx - q) + 1.f; mx = q; } else sm += expf(q - mx);
}
// The joint logit sum_l lq is ~ -1.5 L: its fp32 rounding (1e-5 at L = 100) would be the relative error of every softmax
// weight, so the sum over the warp and the running maximum are carried in fp64 (differences are rounded to fp32 for expf).
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void lse_push_d(double& mx, float& sm, double q) {
  if (q > mx) { sm = sm * expf((float)(mx - q)) + 1.f; mx = q; } else sm += expf((float)(q - mx));
}

// forward: lseJ[j], lseL[j, l], tc_sum += scale * (lseJ[j] - sum_l lseL[j, l])
__global__ void __launch_bounds__(TCOR_THREADS) tc_fwd_k(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ ls,
                                                        const float* __restrict__ p, double* __restrict__ lseJ, float* __restrict__ lseL,
                                                        float* __restrict__ tc_sum, long long rows, int L, float scale) {
  __shared__ float smx[4][TCOR_THREADS], ssm[4][TCOR_THREADS], sjs[4], sred[4];
  __shared__ double sjm[4];
  const long long j = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float zj[TCOR_LK], mx[TCOR_LK], sm[TCOR_LK];
#pragma unroll
  for (int k = 0; k < TCOR_LK; ++k) { const int l = lane + 32 * k; zj[k] = l < L ? z[j * L + l] : 0.f; mx[k] = -INFINITY; sm[k] = 0.f; }
  double mJ = -INFINITY; float sJ = 0.f;
  for (long long i = warp; i < rows; i += 4) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < TCOR_LK; ++k) {
      const int l = lane + 32 * k;
      if (l < L) {
        const float d = zj[k] - mu[i * L + l];
        const float q = -0.5f * (d * d * p[i * L + l] + 2.f * ls[i * L + l] + TCOR_LOG2PI);
        lse_push(mx[k], sm[k], q); t += q;
      }
    }
    lse_push_d(mJ, sJ, warp_sum_d((double)t));
  }
#pragma unroll
  for (int k = 0; k < TCOR_LK; ++k) { smx[warp][lane + 32 * k] = mx[k]; ssm[warp][lane + 32 * k] = sm[k]; }
  if (lane == 0) { sjm[warp] = mJ; sjs[warp] = sJ; }
  __syncthreads();
  const int l = threadIdx.x;
  float v = 0.f;
  if (l < L) {
    float M = fmaxf(fmaxf(smx[0][l], smx[1][l]), fmaxf(smx[2][l], smx[3][l])), s = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) if (ssm[w][l] > 0.f) s += ssm[w][l] * expf(smx[w][l] - M);
    v = M + logf(s);
    lseL[j * L + l] = v;
  }
  v = warp_sum(v);
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double M = fmax(fmax(sjm[0], sjm[1]), fmax(sjm[2], sjm[3])); float s = 0.f;
    for (int w = 0; w < 4; ++w) if (sjs[w] > 0.f) s += sjs[w] * expf((float)(sjm[w] - M));
    const double J = M + (double)logf(s);
    lseJ[j] = J;
    atomicAdd(tc_sum, scale * (float)(J - (double)(sred[0] + sred[1] + sred[2] + sred[3])));
  }
}

// backward.  d TC / d lq(j,i,l) = (W_ji - V_jil) / rows,  W = softmax_i of the joint logits, V = softmax_i per latent.
// BY_I = false: CTA = j, sums over i  -> dz[j, l]   = scale sum_i g (-(z - mu) p)
// BY_I = true : CTA = i, sums over j  -> dmu[i, l]  = scale sum_j g (z - mu) p,   dls[i, l] = scale sum_j g ((z - mu)^2 p - 1)
template <bool BY_I>
__global__ void __launch_bounds__(TCOR_THREADS) tc_bwd_k(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ ls,
                                                        const float* __restrict__ p, const double* __restrict__ lseJ, const float* __restrict__ lseL,
                                                        float* __restrict__ o0, float* __restrict__ o1, long long rows, int L, float scale) {
  __shared__ float s0[4][TCOR_THREADS], s1[4][TCOR_THREADS];
  const long long own = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a[TCOR_LK], b[TCOR_LK], c[TCOR_LK], acc0[TCOR_LK], acc1[TCOR_LK];       // own row: (z, lseL) or (mu, p, ls)
#pragma unroll
  for (int k = 0; k < TCOR_LK; ++k) {
    const int l = lane + 32 * k; acc0[k] = acc1[k] = 0.f; a[k] = b[k] = c[k] = 0.f;
    if (l < L) {
      if (BY_I) { a[k] = mu[own * L + l]; b[k] = p[own * L + l]; c[k] = ls[own * L + l]; }
      else { a[k] = z[own * L + l]; b[k] = lseL[own * L + l]; }
    }
  }
  const double ownJ = BY_I ? 0.0 : lseJ[own];
  for (long long o = warp; o < rows; o += 4) {
    float q[TCOR_LK], d[TCOR_LK], pp[TCOR_LK], lL[TCOR_LK], t = 0.f;
#pragma unroll
    for (int k = 0; k < TCOR_LK; ++k) {
      const int l = lane + 32 * k; q[k] = d[k] = pp[k] = lL[k] = 0.f;
      if (l < L) {
        float lsv;
        if (BY_I) { d[k] = z[o * L + l] - a[k]; pp[k] = b[k]; lsv = c[k]; lL[k] = lseL[o * L + l]; }
        else { d[k] = a[k] - mu[o * L + l]; pp[k] = p[o * L + l]; lsv = ls[o * L + l]; lL[k] = b[k]; }
        q[k] = -0.5f * (d[k] * d[k] * pp[k] + 2.f * lsv + TCOR_LOG2PI);
        t += q[k];
      }
    }
    const float W = expf((float)(warp_sum_d((double)t) - (BY_I ? lseJ[o] : ownJ)));
#pragma unroll
    for (int k = 0; k < TCOR_LK; ++k) {
      const int l = lane + 32 * k;
      if (l < L) {
        const float g = W - expf(q[k] - lL[k]);
        const float dp = d[k] * pp[k];
        if (BY_I) { acc0[k] += g * dp; acc1[k] += g * (d[k] * dp - 1.f); }
        else acc0[k] -= g * dp;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < TCOR_LK; ++k) { s0[warp][lane + 32 * k] = acc0[k]; s1[warp][lane + 32 * k] = acc1[k]; }
  __syncthreads();
  const int l = threadIdx.x;
  if (l < L) {
    o0[own * L + l] = scale * (s0[0][l] + s0[1][l] + s0[2][l] + s0[3][l]);
    if (BY_I) o1[own * L + l] = scale * (s1[0][l] + s1[1][l] + s1[2][l] + s1[3][l]);
  }
}

// dmu += tdz + tdmu;  dls += tdz eps exp(ls) + tdls      (z = mu + eps exp(ls))
__global__ void tc_apply_k(float* __restrict__ dmu, float* __restrict__ dls, const float* __restrict__ tdz, const float* __restrict__ tdmu,
                           const float* __restrict__ tdls, const float* __restrict__ eps, const float* __restrict__ ls, long long n) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float g = tdz[idx];
  dmu[idx] += g + tdmu[idx];
  dls[idx] += g * eps[idx] * expf(ls[idx]) + tdls[idx];
}
