// spectral.cuh -- e2e layer 1 (layers.py:431-450 at model.py:202) in the frequency domain.
//
// The width-N "cross" convolution of e2e is a full-width 1-D correlation along one axis of the [N, N] grid:
//     O[line, s, q] = sum_{s', c} Y[line, s', c] w1[s' - s + p, c, q]         (line = (direction, graph, i or j))
// i.e. a circular convolution of the zero-padded line with g[m] = w1[p - m] (m in [-q, p], q = N-1-p) as soon as
// the transform length L >= N + q.  N = 256 gives L = 384 = 3 * 2^7, exactly.  In the frequency domain
//     O^[f, q]  = sum_c Y^[f, c] G^[f, c, q]                 (fwd)
//     dY^[f, c] = sum_q dO^[f, q] conj(G^[f, c, q])          (dgrad)
//     dG^[f, c, q] = sum_lines conj(Y^[f, c]) dO^[f, q]      (wgrad; dw1[t] = irfft(dG^)[p - t])
// so per frequency the channel mix is a small dense GEMM shared by every line of the batch: K = 2*50 real
// (re | im), N = 2*20.  That is ~30x fewer tensor flops than the block-Toeplitz GEMM (e2e_tc.cuh), and the stage
// becomes HBM-bound.  Pipeline per micro-batch (all hand-written):
//     spec_fft_fwd_k   fp32 lines (optionally through BN+relu) -> Stockham FFT in shared memory (two real channels
//                      per complex transform) -> bf16 hi/lo planes A^[f][line][re c.. | im c..]
//     spec_gemm_k      tcgen05 kind::f16, 3-pass split-bf16, one [128 lines x K] x [K x N] product per (f, line tile);
//                      TMA-fed ring, TMEM slot per item, epilogue staged in smem and written with a bulk copy
//     spec_fft_inv_k   fp32 spectra -> inverse Stockham FFT -> fp32 lines (first N positions, scaled by 1/L)
//     spec_wgrad_k     per frequency  P[f] += A^[f]^T . dO^[f]  (MN-major operands straight from the same planes),
//                      two-level (TMEM -> fp32 register) accumulation, atomics into P; folded to dw1 once per step
// spec_fft_fwd_k / spec_fft_inv_k are the runtime-plan transforms (any N); where a compile-time plan matches (L = 384,
// 192, 48 with 50 / 20 channels) the persistent spec_fft_fwd_fast_k / spec_fft_inv_fast_k / spec_fft_inv_fast2_k run
// instead: next line staged asynchronously (cp.async / cp.async.bulk) while the current one is transformed, fixed channel
// pair per thread, lines walked in an order chosen for DRAM sector merging and L2 reuse (LineWalk).
#pragma once
#include "e2e_tc.cuh"
#include "fft2p.cuh"
#include <vector>
#include <cmath>

#define SP_KA1 104     /* row stride (bf16) of the Y^ planes: 25 x [re c, re c+1, im c, im c+1] + 4 pad = 208 B (TMA: 16 B multiples) */
#define SP_KA2 40      /* row stride (bf16) of the dO^ planes: 10 x [re q, re q+1, im q, im q+1] = 80 B */
#define SP_NF 48       /* fwd GEMM N: 40 -> 48 (N % 16 == 0 at M = 128) */
#define SP_ND 112      /* dgrad GEMM N: 100 -> 112 */
#define SP_FFT_THREADS_MAX 512
// Spectrum rows hold the channels in pairs, [re c, re c+1, im c, im c+1] for even c: the transforms pack two real channels
// into one complex signal, so one thread owns exactly these four values of a row -- one 8-byte bf16 store per plane (forward) or
// one 16-byte load (inverse) instead of two accesses half a row apart.  The GEMM operands (spec_stage_weights_k) and the
// weight-gradient fold (spec_wgrad_finalize_k) index the rows through the same two functions.
__host__ __device__ __forceinline__ int sp_kre(int c) { return ((c >> 1) << 2) | (c & 1); }
__host__ __device__ __forceinline__ int sp_kim(int c) { return sp_kre(c) + 2; }
__device__ __forceinline__ uint2 sp_pack_hi_lo(float x1r, float x2r, float x1i, float x2i, uint2& lo) {     // -> [re c, re c+1, im c, im c+1] hi, lo
  const __nv_bfloat162 hr = __floats2bfloat162_rn(x1r, x2r), hi = __floats2bfloat162_rn(x1i, x2i);
  const __nv_bfloat162 lr = __floats2bfloat162_rn(x1r - __low2float(hr), x2r - __high2float(hr));
  const __nv_bfloat162 li = __floats2bfloat162_rn(x1i - __low2float(hi), x2i - __high2float(hi));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&lr), *reinterpret_cast<const uint32_t*>(&li));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&hr), *reinterpret_cast<const uint32_t*>(&hi));
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulnegi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)

// ---- small forward DFTs on registers (natural order in and out; e^{-2 pi i / R}) -----------------------------
template <int R> __device__ __forceinline__ void dft_r(float2* v);
template <> __device__ __forceinline__ void dft_r<2>(float2* v) {
  const float2 a = v[0], b = v[1]; v[0] = caddf(a, b); v[1] = csubf(a, b);
}
template <> __device__ __forceinline__ void dft_r<4>(float2* v) {
  const float2 t0 = caddf(v[0], v[2]), t1 = csubf(v[0], v[2]), t2 = caddf(v[1], v[3]), t3 = cmulnegi(csubf(v[1], v[3]));
  v[0] = caddf(t0, t2); v[1] = caddf(t1, t3); v[2] = csubf(t0, t2); v[3] = csubf(t1, t3);
}
template <> __device__ __forceinline__ void dft_r<8>(float2* v) {
  float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
  dft_r<4>(e); dft_r<4>(o);
  const float s = 0.70710678118654752f;
  const float2 o1 = make_float2(s * (o[1].x + o[1].y), s * (o[1].y - o[1].x));        // * (s, -s)
  const float2 o2 = cmulnegi(o[2]);
  const float2 o3 = make_float2(s * (o[3].y - o[3].x), -s * (o[3].x + o[3].y));       // * (-s, -s)
  v[0] = caddf(e[0], o[0]); v[4] = csubf(e[0], o[0]);
  v[1] = caddf(e[1], o1);   v[5] = csubf(e[1], o1);
  v[2] = caddf(e[2], o2);   v[6] = csubf(e[2], o2);
  v[3] = caddf(e[3], o3);   v[7] = csubf(e[3], o3);
}
template <> __device__ __forceinline__ void dft_r<3>(float2* v) {
  const float h = 0.86602540378443865f;
  const float2 s = caddf(v[1], v[2]), d = csubf(v[1], v[2]);
  const float2 m = make_float2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
  const float2 n = make_float2(h * d.y, -h * d.x);                                    // (-i sqrt(3)/2) d
  v[0] = caddf(v[0], s); v[1] = caddf(m, n); v[2] = csubf(m, n);
}
template <> __device__ __forceinline__ void dft_r<6>(float2* v) {
  float2 e[3] = {v[0], v[2], v[4]}, o[3] = {v[1], v[3], v[5]};
  dft_r<3>(e); dft_r<3>(o);
  const float h = 0.86602540378443865f;
  const float2 o1 = cmulf(o[1], make_float2(0.5f, -h)), o2 = cmulf(o[2], make_float2(-0.5f, -h));
  v[0] = caddf(e[0], o[0]); v[3] = csubf(e[0], o[0]);
  v[1] = caddf(e[1], o1);   v[4] = csubf(e[1], o1);
  v[2] = caddf(e[2], o2);   v[5] = csubf(e[2], o2);
}

// ---- sources / sinks of a Stockham pass ------------------------------------------------------------------------
struct SmemIO {                       // [pos][g] complex buffer in shared memory
  float2* p; int g;
  __device__ __forceinline__ float2 ld(int pos, int cp) const { return p[pos * g + cp]; }
  __device__ __forceinline__ void st(int pos, int cp, float2 v) const { p[pos * g + cp] = v; }
};
struct LineSrc {                      // fp32 line in global memory, two real channels per complex value, zero padded past N
  const float* base; long long pstride; int N; const float* sg; const float* sb;   // sg/sb: BN scale/shift in smem (or NULL)
  __device__ __forceinline__ float2 ld(int pos, int cp) const {
    if (pos >= N) return make_float2(0.f, 0.f);
    float2 x = *reinterpret_cast<const float2*>(base + (long long)pos * pstride + 2 * cp);
    if (sg) { x.x = fmaxf(fmaf(x.x, sg[2 * cp], sb[2 * cp]), 0.f); x.y = fmaxf(fmaf(x.y, sg[2 * cp + 1], sb[2 * cp + 1]), 0.f); }
    return x;
  }
};
struct LineDst {                      // inverse transform tail: (re, im) swapped back, scaled, first N positions only
  float* base; long long pstride; int N; float scale;
  __device__ __forceinline__ void st(int pos, int cp, float2 v) const {
    if (pos < N) *reinterpret_cast<float2*>(base + (long long)pos * pstride + 2 * cp) = make_float2(v.y * scale, v.x * scale);
  }
};

// one Stockham autosort pass of radix R over T = L / R butterflies x g channel pairs:
//   v[r] = in[j + r T] * tw^(r k L / (Ns R)),  k = j mod Ns;  out[(j - k) R + k + r Ns] = DFT_R(v)[r]
template <int R, class Src, class Dst>
__device__ __forceinline__ void fft_pass(const Src& in, const Dst& out, const float2* __restrict__ tw, int L, int Ns, int g) {
  const int T = L / R, tstep = L / (Ns * R);
  for (int w = threadIdx.x; w < T * g; w += blockDim.x) {
    const int j = w / g, cp = w - j * g;
    const int k = j % Ns;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in.ld(j + r * T, cp);
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmulf(v[r], tw[r * k * tstep]);
    }
    dft_r<R>(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out.st(j0 + r * Ns, cp, v[r]);
  }
}
template <class Src, class Dst>
__device__ __forceinline__ void fft_pass_any(int R, const Src& in, const Dst& out, const float2* tw, int L, int Ns, int g) {
  switch (R) {
    case 2: fft_pass<2>(in, out, tw, L, Ns, g); break;
    case 3: fft_pass<3>(in, out, tw, L, Ns, g); break;
    case 4: fft_pass<4>(in, out, tw, L, Ns, g); break;
    case 6: fft_pass<6>(in, out, tw, L, Ns, g); break;
    default: fft_pass<8>(in, out, tw, L, Ns, g); break;
  }
}

struct FftPlan { int L, F, npass, rad[8]; };

// ---- forward: lines -> bf16 hi/lo spectra planes ------------------------------------------------------------------
struct FftFwdArgs {
  const float* in;
  long long lines0;       // lines [0, lines0): contiguous, base = line * N * C, position stride C
  long long lines;        // lines [lines0, lines): line' = line - lines0
  int dir1_strided;       //   1: line' = (b, j) of a [b, N, N, C] tensor read along its first N axis (base = (b N N + j) C, stride N C)
                          //   0: contiguous like direction 0 (base = line * N * C)
                          //   2: graph-tiled pair of tensors (y_producer_tc_k): line = (b, x) lives at in (dir 0) / in1 (dir 1)
                          //      + ((((b / 128) N + x) 128 + b % 128) N C, contiguous
  const float* in1;
  const float* gam; const float* bet;     // optional: x <- relu(x * gam * BN_RS + bet)  (BN_e1 + relu, model.py:201-202)
  const float* bias0; const float* bias1; const float* b0;   // optional (with gam): x <- x + bias{0,1}[(b N + pos of the line), :] + 2 b0
                                                              // before the BN (direction 0 / 1 lines; see y_producer_tc_k)
  __nv_bfloat16* oh; __nv_bfloat16* ol;   // [F][RA][KA]
  long long RA; int KA;
  int N, C, G;            // C real channels (even), G channel pairs per shared-memory group
  const float2* tw;       // exp(-2 pi i k / L), k < L
  FftPlan pl;
  int order;              // fast kernels: order in which the persistent grid walks the lines (LineWalk)
  int bulk;               // fast kernels: contiguous lines staged by one cp.async.bulk instead of per-thread cp.async
};
__global__ void __launch_bounds__(SP_FFT_THREADS_MAX, 1) spec_fft_fwd_k(FftFwdArgs A) {
  extern __shared__ __align__(16) float2 fft_sm[];
  __shared__ float s_g[64], s_b[64];
  const int L = A.pl.L, F = A.pl.F, CP = A.C / 2;
  float2* tw = fft_sm; float2* buf0 = tw + L; float2* buf1 = buf0 + (size_t)L * A.G;
  for (int t = threadIdx.x; t < L; t += blockDim.x) tw[t] = A.tw[t];
  if (A.gam) for (int t = threadIdx.x; t < A.C; t += blockDim.x) { s_g[t] = A.gam[t] * BN_RS; s_b[t] = A.bet[t]; }
  for (long long line = blockIdx.x; line < A.lines; line += gridDim.x) {
    const float* base; long long pstride;
    if (A.dir1_strided == 2) {
      const bool d1 = line >= A.lines0; const long long l1 = d1 ? line - A.lines0 : line; const long long b = l1 / A.N; const int x = (int)(l1 - b * A.N);
      base = (d1 ? A.in1 : A.in) + (((b / 128) * A.N + x) * 128 + (b % 128)) * A.N * A.C; pstride = A.C;
    } else if (line < A.lines0 || !A.dir1_strided) { base = A.in + line * A.N * A.C; pstride = A.C; }
    else { const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
           base = A.in + (b * A.N * A.N + j) * A.C; pstride = (long long)A.N * A.C; }
    if (A.gam && A.bias0) {            // per-line shift: (bias + 2 b0) * g + beta
      __syncthreads();
      const bool d1 = line >= A.lines0; const float* br = (d1 ? A.bias1 + (line - A.lines0) * A.C : A.bias0 + line * A.C);
      for (int t = threadIdx.x; t < A.C; t += blockDim.x) s_b[t] = fmaf(br[t] + 2.f * A.b0[t], s_g[t], A.bet[t]);
    }
    for (int cp0 = 0; cp0 < CP; cp0 += A.G) {
      const int g = min(A.G, CP - cp0);
      __syncthreads();
      LineSrc src; src.base = base + 2 * cp0; src.pstride = pstride; src.N = A.N;
      src.sg = A.gam ? s_g + 2 * cp0 : nullptr; src.sb = s_b + 2 * cp0;
      SmemIO cur; cur.p = buf0; cur.g = g; SmemIO oth; oth.p = buf1; oth.g = g;
      fft_pass_any(A.pl.rad[0], src, cur, tw, L, 1, g);
      int Ns = A.pl.rad[0];
      for (int ps = 1; ps < A.pl.npass; ++ps) {
        __syncthreads();
        fft_pass_any(A.pl.rad[ps], cur, oth, tw, L, Ns, g);
        Ns *= A.pl.rad[ps];
        float2* t = cur.p; cur.p = oth.p; oth.p = t;
      }
      __syncthreads();
      // separate the two real channels packed in each complex transform: X1 = (Z[f] + conj Z[L-f]) / 2,
      // X2 = (Z[f] - conj Z[L-f]) / (2i); split to bf16 hi / lo and store [re | im]
      for (int w = threadIdx.x; w < F * g; w += blockDim.x) {
        const int f = w / g, cp = w - f * g;
        const float2 z1 = cur.p[f * g + cp], z2 = cur.p[(f == 0 ? 0 : L - f) * g + cp];
        const float x1r = 0.5f * (z1.x + z2.x), x1i = 0.5f * (z1.y - z2.y);
        const float x2r = 0.5f * (z1.y + z2.y), x2i = 0.5f * (z2.x - z1.x);
        const long long o = ((long long)f * A.RA + line) * A.KA + 4 * (cp0 + cp);
        uint2 lo; const uint2 hi = sp_pack_hi_lo(x1r, x2r, x1i, x2i, lo);
        *reinterpret_cast<uint2*>(A.oh + o) = hi; *reinterpret_cast<uint2*>(A.ol + o) = lo;
        if (cp0 + cp == CP - 1) for (int e = 4; e < A.KA - 4 * (CP - 1); e += 2) {   // the row's pad: no partly written sectors
          *reinterpret_cast<uint32_t*>(A.oh + o + e) = 0u; *reinterpret_cast<uint32_t*>(A.ol + o + e) = 0u; }
      }
    }
  }
}

// ---- inverse: fp32 spectra [F][RA][2C] ([re c.. | im c..]) -> fp32 lines [line][N][C] -----------------------------
struct FftInvArgs {
  const float* in; long long RA;
  float* out; long long lines;
  long long lines0; float* out1;     // out1 != NULL: lines >= lines0 are (b, j) columns written TRANSPOSED into out1[b][pos][j][C]
                                     // (position stride N C), so that both directions share the [b, i, j, C] layout
  int N, C, G;
  const float2* tw;
  FftPlan pl;
  int order;                         // fast kernels: LineWalk order
  int bulk;                          // fast kernels: spectrum rows staged by one cp.async.bulk per row (experimental, off by default)
};
__global__ void __launch_bounds__(SP_FFT_THREADS_MAX, 1) spec_fft_inv_k(FftInvArgs A) {
  extern __shared__ __align__(16) float2 fft_sm[];
  const int L = A.pl.L, F = A.pl.F, CP = A.C / 2, W = 2 * A.C;
  float2* tw = fft_sm; float2* buf0 = tw + L; float2* buf1 = buf0 + (size_t)L * A.G;
  for (int t = threadIdx.x; t < L; t += blockDim.x) tw[t] = A.tw[t];
  const float scale = 1.f / (float)L;
  for (long long line = blockIdx.x; line < A.lines; line += gridDim.x) {
    for (int cp0 = 0; cp0 < CP; cp0 += A.G) {
      const int g = min(A.G, CP - cp0);
      __syncthreads();
      // Z = X1 + i X2 over the full circle (X[L-f] = conj X[f]); inverse DFT = swap . forward DFT . swap
      for (int w = threadIdx.x; w < F * g; w += blockDim.x) {
        const int f = w / g, cp = w - f * g;
        const float4 ab = *reinterpret_cast<const float4*>(A.in + ((long long)f * A.RA + line) * W + 4 * (cp0 + cp));
        const float2 a = make_float2(ab.x, ab.y);
        float2 b = make_float2(ab.z, ab.w);
        const bool selfc = (f == 0) || (2 * f == L);
        if (selfc) b = make_float2(0.f, 0.f);
        buf0[f * g + cp] = make_float2(b.x + a.y, a.x - b.y);
        if (!selfc) buf0[(L - f) * g + cp] = make_float2(a.y - b.x, a.x + b.y);
      }
      SmemIO cur; cur.p = buf0; cur.g = g; SmemIO oth; oth.p = buf1; oth.g = g;
      LineDst dst; dst.base = A.out + line * A.N * A.C + 2 * cp0; dst.pstride = A.C; dst.N = A.N; dst.scale = scale;
      if (A.out1 && line >= A.lines0) {
        const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
        dst.base = A.out1 + (b * A.N * A.N + j) * A.C + 2 * cp0; dst.pstride = (long long)A.N * A.C;
      }
      int Ns = 1;
      for (int ps = 0; ps < A.pl.npass; ++ps) {
        __syncthreads();
        if (ps == A.pl.npass - 1) fft_pass_any(A.pl.rad[ps], cur, dst, tw, L, Ns, g);
        else fft_pass_any(A.pl.rad[ps], cur, oth, tw, L, Ns, g);
        Ns *= A.pl.rad[ps];
        float2* t = cur.p; cur.p = oth.p; oth.p = t;
      }
    }
  }
}

// ---- specialised transforms: compile-time plan (R0 x R1 x R2), all G channel pairs of a line in one group, next line prefetched
//      with cp.async into a staging buffer while the current one is transformed --------------------------------------------
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
// one bulk (TMA, no tensor map) copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Work item w = it NT + tid is butterfly j = w / G of channel pair cp = w % G.  NT is a multiple of G, so cp and
// jb = tid / G are fixed per thread (computed once per kernel) and j = it (NT / G) + jb: no division in the loops, and the
// sources address  tid + (compile-time offset) G  relative to their base.
template <int R, int NS, int L, int G, int NT, class Src, class Dst>
__device__ __forceinline__ void fft_pass_ct(const Src& in, const Dst& out, const float2* __restrict__ tw, const int jb, const int cp) {
  constexpr int T = L / R, TSTEP = L / (NS * R), JS = NT / G;
  static_assert(NT % G == 0, "the thread count must be a multiple of the channel pairs");
  const int tid = threadIdx.x;
#pragma unroll
  for (int j00 = 0; j00 < T; j00 += JS) {
    const int j = j00 + jb;
    if (T % JS == 0 || j < T) {
      const int k = j % NS;
      float2 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = in.ld(jb, tid, j00 + r * T);
      if (NS > 1) {
        // one table read per butterfly; the other powers by multiplication (shared-memory instructions, not FP32 issue, bound
        // these kernels): w^2 = w w, w^3 = w^2 w, w^4 = w^2 w^2, ... each at most three products deep
        float2 w[R];
        w[1] = tw[k * TSTEP];
#pragma unroll
        for (int r = 2; r < R; ++r) w[r] = (r & 1) ? cmulf(w[r - 1], w[1]) : cmulf(w[r / 2], w[r / 2]);
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmulf(v[r], w[r]);
      }
      dft_r<R>(v);
      const int j0 = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) out.st(j0 + r * NS, cp, v[r]);
    }
  }
}
template <int G> struct SmemCT {      // [pos][G] complex buffer, compile-time pitch; loads at position jb + off, pair cp: index tid + off G
  float2* p;
  __device__ __forceinline__ float2 ld(int, int tid, int off) const { return p[tid + off * G]; }
  __device__ __forceinline__ void st(int pos, int cp, float2 v) const { p[pos * G + cp] = v; }
};
template <int G> struct StageSrc {    // staged fp32 line [pos][2G] in shared memory, BN + relu on the fly, zero padded past N
  const float2* s; int N; bool bn; float gx, gy, bx, by;      // the thread's channel pair is fixed: its BN scale / shift live in registers
  __device__ __forceinline__ float2 ld(int jb, int tid, int off) const {
    if (jb + off >= N) return make_float2(0.f, 0.f);
    float2 x = s[tid + off * G];
    if (bn) { x.x = fmaxf(fmaf(x.x, gx, bx), 0.f); x.y = fmaxf(fmaf(x.y, gy, by), 0.f); }
    return x;
  }
};
struct LineDstCT {                    // inverse transform tail for a fixed channel pair: base already points at the pair
  float* base; int pstride; int N; float scale;
  __device__ __forceinline__ void st(int pos, int, float2 v) const {
    if (pos < N) *reinterpret_cast<float2*>(base + (size_t)(unsigned)(pos * pstride)) = make_float2(v.y * scale, v.x * scale);
  }
};
// Order in which a persistent CTA walks the lines.  The spectrum rows of a line are 80 .. 400 bytes (2.5 .. 12.5 sectors) and
// the transposed column-line chunks 80 / 200 bytes, so two neighbouring lines share a 32-byte sector; when they are processed
// by different CTAs that drift apart, the half-written sector leaves L2 before its other half arrives (a DRAM read-modify-write).
//   order 0: line = cta + it * grid (the plain grid stride)
//   order 1: each CTA takes PAIRS of neighbouring lines back to back: position o = 2 (cta + (it / 2) grid) + (it & 1), line = o
//   order 2: pairs, and the positions run graph by graph (N row lines of graph b, then its N column lines): a tensor that
//            both directions read (dO, 20 N^2 B per graph) is fetched from DRAM once and found in L2 the second time
struct LineWalk {
  long long lines, lines0; int N, order;
  __device__ __forceinline__ long long at(long long it) const {      // >= lines: past the end (positions grow with it)
    if (order == 0) return (long long)blockIdx.x + it * gridDim.x;
    const long long o = 2 * ((long long)blockIdx.x + (it >> 1) * gridDim.x) + (it & 1);
    if (order == 1 || o >= lines) return o;
    const long long b = o / (2 * N); const int r = (int)(o - b * (2 * N));
    return r < N ? b * N + r : lines0 + b * N + (r - N);
  }
};
template <int R0, int R1, int R2, int G>
struct FastCfg {
  static constexpr int L = R0 * R1 * (R2 ? R2 : 1), F = L / 2 + 1;
  static constexpr size_t BUF = (size_t)F * G * 16 > (size_t)L * G * 8 ? (size_t)F * G * 16 : (size_t)L * G * 8;   // bytes of one ping-pong buffer
  static size_t smem_fwd(int N) { return (size_t)L * 8 + 2 * BUF + (size_t)N * 2 * G * 4; }
  static size_t smem_inv(bool alias) { return (size_t)L * 8 + (alias ? 2 : 3) * BUF; }
};

template <int R0, int R1, int R2, int G, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) spec_fft_fwd_fast_k(FftFwdArgs A) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  constexpr int L = Cfg::L, F = Cfg::F, C = 2 * G;
  extern __shared__ __align__(16) uint8_t fsm[];
  __shared__ float s_g[2 * G], s_b[2 * G], s_raw[2 * G];
  __shared__ uint64_t stage_bar;             // completion of a contiguous line's bulk copy into the staging tile
  if (threadIdx.x == 0) { mbar_init(&stage_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();                           // nobody polls the barrier before it exists
  uint32_t stage_ph = 0;
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float2* bufB = reinterpret_cast<float2*>(fsm + (size_t)L * 8 + Cfg::BUF);
  float* stage = reinterpret_cast<float*>(fsm + (size_t)L * 8 + 2 * Cfg::BUF);
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  if (A.gam) for (int t = threadIdx.x; t < C; t += NT) { s_g[t] = A.gam[t] * BN_RS; s_b[t] = A.bet[t]; }
  // thread t < C owns channel t of the per-line shift (bias + 2 b0) g + beta: the constant part lives in a register and the line's
  // bias element arrives with the line (cp.async by the same thread, so no barrier between its arrival and its use)
  const bool shifter = A.gam && A.bias0 && threadIdx.x < C;
  float sh_g = 0.f, sh_c = 0.f;
  if (shifter) { sh_g = A.gam[threadIdx.x] * BN_RS; sh_c = fmaf(2.f * A.b0[threadIdx.x], sh_g, A.bet[threadIdx.x]); }
  const int N = A.N;
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;      // this thread's channel pair and first butterfly / frequency
  const bool vec16 = ((long long)N * C) % 4 == 0;
  auto prefetch = [&](long long line) {
    if (line >= A.lines) return;
    if (shifter) cp_async4(s_raw + threadIdx.x, (line >= A.lines0 ? A.bias1 + (line - A.lines0) * C : A.bias0 + line * C) + threadIdx.x);
    if (line < A.lines0 || A.dir1_strided != 1) {
      const float* base = A.in + line * N * C;
      if (A.dir1_strided == 2) {
        const bool d1 = line >= A.lines0; const long long l1 = d1 ? line - A.lines0 : line; const long long b = l1 / N; const int x = (int)(l1 - b * N);
        base = (d1 ? A.in1 : A.in) + (((b / 128) * N + x) * 128 + (b % 128)) * N * C;
      }
      // a contiguous line is ONE bulk copy issued by one thread: no LDGSTS instructions, no LSU wavefronts for the staging
      // writes (they were 14 % of the kernel's shared-memory wavefronts); the others wait on the mbarrier's phase
      if (A.bulk && vec16) { if (threadIdx.x == 0) { mbar_expect_tx(&stage_bar, (uint32_t)(N * C * 4)); bulk_load(stage, base, (uint32_t)(N * C * 4), &stage_bar); } }
      else if (vec16) { for (int t = threadIdx.x; t < N * C / 4; t += NT) cp_async16(stage + 4 * t, base + 4 * t); }
      else { for (int t = threadIdx.x; t < N * G; t += NT) cp_async8(stage + 2 * t, base + 2 * t); }
    } else {
      const long long l1 = line - A.lines0; const long long b = l1 / N; const int j = (int)(l1 - b * N);
      const float* base = A.in + (b * N * N + j) * C; const long long ps = (long long)N * C;
      if constexpr (G % 2 == 0) {        // a position's C floats are whole 16-byte pieces (and every chunk starts 16-byte aligned)
        constexpr int Q = G / 2;
        for (int t = threadIdx.x; t < N * Q; t += NT) { const int pos = t / Q, k = t - pos * Q; cp_async16(stage + pos * C + 4 * k, base + pos * ps + 4 * k); }
      } else {
        for (int pos = jb; pos < N; pos += NT / G) cp_async8(stage + 2 * (pos * G + cp), base + pos * ps + 2 * cp);
      }
    }
  };
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = N; walk.order = A.order;
  long long line = walk.at(0);
  prefetch(line);
  cp_async_commit();
  for (long long it = 0; line < A.lines; ++it) {
    const long long next = walk.at(it + 1);
    cp_async_wait_all();
    if (A.bulk && vec16 && (line < A.lines0 || A.dir1_strided != 1)) { mbar_wait(&stage_bar, stage_ph); stage_ph ^= 1; }
    if (shifter) s_b[threadIdx.x] = fmaf(s_raw[threadIdx.x], sh_g, sh_c);     // pass 1 of the previous line is long done with s_b
    __syncthreads();                       // staged line visible; previous line's readers of bufA / bufB are done
    StageSrc<G> src; src.s = reinterpret_cast<const float2*>(stage); src.N = N; src.bn = A.gam != nullptr;
    if (src.bn) { src.gx = s_g[2 * cp]; src.gy = s_g[2 * cp + 1]; src.bx = s_b[2 * cp]; src.by = s_b[2 * cp + 1]; }
    SmemCT<G> a; a.p = bufA; SmemCT<G> b; b.p = bufB;
    fft_pass_ct<R0, 1, L, G, NT>(src, a, tw, jb, cp);
    __syncthreads();
    prefetch(next);                        // the staging buffer is free: overlap the next line's loads with passes 2, 3 and the store
    cp_async_commit();
    fft_pass_ct<R1, R0, L, G, NT>(a, b, tw, jb, cp);
    __syncthreads();
    const float2* res = bufB;
    if (R2) { fft_pass_ct<(R2 ? R2 : 2), R0 * R1, L, G, NT>(b, a, tw, jb, cp); __syncthreads(); res = bufA; }
    // spectra of the pair's two real channels, split into bf16 hi / lo rows of the GEMM operand: f = jb, jb + NT / G, ...
    __nv_bfloat16* oh = A.oh + ((long long)jb * A.RA + line) * A.KA + 4 * cp;
    __nv_bfloat16* ol = A.ol + ((long long)jb * A.RA + line) * A.KA + 4 * cp;
    const long long fstep = (long long)(NT / G) * A.RA * A.KA;
    const int padw = cp == G - 1 ? A.KA - 2 * C : 0;       // the last pair's thread also writes the row's pad (bf16 elements)
#pragma unroll 4
    for (int f = jb; f < F; f += NT / G, oh += fstep, ol += fstep) {
      const float2 z1 = res[f * G + cp], z2 = res[(f == 0 ? 0 : L - f) * G + cp];
      const float x1r = 0.5f * (z1.x + z2.x), x1i = 0.5f * (z1.y - z2.y);
      const float x2r = 0.5f * (z1.y + z2.y), x2i = 0.5f * (z2.x - z1.x);
      uint2 lo; const uint2 hi = sp_pack_hi_lo(x1r, x2r, x1i, x2i, lo);
      *reinterpret_cast<uint2*>(oh) = hi; *reinterpret_cast<uint2*>(ol) = lo;
      // the row's pad (KA - 2C elements the tensor maps never read) is written too: a sector that leaves L2 partly written
      // costs a DRAM read-modify-write (measured: +1.7 GB of reads per 256 graphs, one sector per row and plane)
      if (padw == 4) { *reinterpret_cast<uint2*>(oh + 4) = make_uint2(0u, 0u); *reinterpret_cast<uint2*>(ol + 4) = make_uint2(0u, 0u); }
      else for (int e = 0; e < padw; e += 2) { *reinterpret_cast<uint32_t*>(oh + 4 + e) = 0u; *reinterpret_cast<uint32_t*>(ol + 4 + e) = 0u; }
    }
    line = next;
  }
  cp_async_wait_all();
}

// inverse: the raw spectrum rows of a line ([f][2C] fp32) are staged with cp.async; ALIAS = the staging buffer is bufB
// (3-pass plans at G = 25 do not fit a third buffer; the prefetch then overlaps only the last pass)
template <int R0, int R1, int R2, int G, int NT, int MINB, bool ALIAS>
__global__ void __launch_bounds__(NT, MINB) spec_fft_inv_fast_k(FftInvArgs A) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  constexpr int L = Cfg::L, F = Cfg::F, C = 2 * G, W = 4 * G;
  static_assert(!ALIAS || R2 != 0, "aliasing the staging buffer needs a 3-pass plan");
  extern __shared__ __align__(16) uint8_t fsm[];
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float2* bufB = reinterpret_cast<float2*>(fsm + (size_t)L * 8 + Cfg::BUF);
  float* stage = ALIAS ? reinterpret_cast<float*>(bufB) : reinterpret_cast<float*>(fsm + (size_t)L * 8 + 2 * Cfg::BUF);
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  __shared__ uint64_t stage_bar;             // bulk staging: completion of a line's F row copies
  if (threadIdx.x == 0) { mbar_init(&stage_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t stage_ph = 0;
  const float scale = 1.f / (float)L;
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;      // this thread's channel pair and first butterfly / frequency
  auto prefetch = [&](long long line) {
    if (line >= A.lines) return;
    if (A.bulk) {                              // one bulk copy per spectrum row (W floats), all counted on one mbarrier phase
      if (threadIdx.x == 0) mbar_expect_tx(&stage_bar, (uint32_t)(F * W * 4));
      for (int f = threadIdx.x; f < F; f += NT) bulk_load(stage + f * W, A.in + ((long long)f * A.RA + line) * W, (uint32_t)(W * 4), &stage_bar);
      return;
    }
    const float* src = A.in + ((long long)jb * A.RA + line) * W + 4 * cp;     // 16 bytes = two channel pairs' worth of one half row
    const long long fstep = (long long)(NT / G) * A.RA * W;
    for (int f = jb; f < F; f += NT / G, src += fstep) cp_async16(stage + f * W + 4 * cp, src);
  };
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = A.N; walk.order = A.order;
  long long line = walk.at(0);
  prefetch(line);
  cp_async_commit();
  for (long long it = 0; line < A.lines; ++it) {
    const long long next = walk.at(it + 1);
    cp_async_wait_all();
    if (A.bulk) { mbar_wait(&stage_bar, stage_ph); stage_ph ^= 1; }
    __syncthreads();
#pragma unroll 4
    for (int f = jb; f < F; f += NT / G) {
      const float4 ab = *reinterpret_cast<const float4*>(stage + f * W + 4 * cp);
      const float2 a = make_float2(ab.x, ab.y);
      float2 b = make_float2(ab.z, ab.w);
      const bool selfc = (f == 0) || (2 * f == L);
      if (selfc) b = make_float2(0.f, 0.f);
      bufA[f * G + cp] = make_float2(b.x + a.y, a.x - b.y);
      if (!selfc) bufA[(L - f) * G + cp] = make_float2(a.y - b.x, a.x + b.y);
    }
    __syncthreads();
    SmemCT<G> a; a.p = bufA; SmemCT<G> b; b.p = bufB;
    LineDstCT dst; dst.base = A.out + line * A.N * C + 2 * cp; dst.pstride = C; dst.N = A.N; dst.scale = scale;
    if (A.out1 && line >= A.lines0) {
      const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
      dst.base = A.out1 + (b * A.N * A.N + j) * C + 2 * cp; dst.pstride = A.N * C;
    }
    if (!ALIAS) { prefetch(next); cp_async_commit(); }
    fft_pass_ct<R0, 1, L, G, NT>(a, b, tw, jb, cp);
    __syncthreads();
    if (R2) {
      fft_pass_ct<R1, R0, L, G, NT>(b, a, tw, jb, cp);
      __syncthreads();
      if (ALIAS) { prefetch(next); cp_async_commit(); }
      fft_pass_ct<(R2 ? R2 : 2), R0 * R1, L, G, NT>(a, dst, tw, jb, cp);
    } else {
      fft_pass_ct<R1, R0, L, G, NT>(b, dst, tw, jb, cp);
    }
    line = next;
  }
  cp_async_wait_all();
}

// inverse, 3-pass plans, second form: the staged spectrum rows are read by pass 1 directly (Hermitian extension and the
// swap of the inverse on the fly: no extension phase, one block barrier less per line), and the staging area is split in two
// halves -- rows [0, FH) in a dedicated buffer that pass 1 frees, rows [FH, F) aliased onto the ping-pong buffer that the
// line's last pass does not touch -- so the next line's loads overlap passes 2 AND 3 (the ALIAS form overlaps only pass 3
// because three full buffers do not fit 227 KB at G = 25), in 196 KB (G = 25) / 80 KB (G = 10).
template <int L, int G> struct SpecHalfSrc {
  const float* h1; const float* h2; int cp;
  static constexpr int F = L / 2 + 1, FH = (F + 1) / 2, W = 4 * G, C = 2 * G;
  __device__ __forceinline__ float2 ld(int jb, int, int off) const {
    const int p = jb + off; const bool up = p >= F; const int f = up ? L - p : p;
    const float4 ab = *reinterpret_cast<const float4*>((f < FH ? h1 + f * W : h2 + (f - FH) * W) + 4 * cp);
    const float2 a = make_float2(ab.x, ab.y);
    float2 b = make_float2(ab.z, ab.w);
    if (f == 0 || 2 * f == L) b = make_float2(0.f, 0.f);
    return up ? make_float2(a.y - b.x, a.x + b.y) : make_float2(b.x + a.y, a.x - b.y);
  }
};
template <int R0, int R1, int R2, int G, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) spec_fft_inv_fast2_k(FftInvArgs A) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  constexpr int L = Cfg::L, F = Cfg::F, C = 2 * G, W = 4 * G, FH = (F + 1) / 2, JS = NT / G;
  static_assert(R2 != 0, "3-pass plans only");
  static_assert((size_t)(F - FH) * W * 4 <= Cfg::BUF, "the second half of the staged rows must fit a ping-pong buffer");
  extern __shared__ __align__(16) uint8_t fsm[];
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bp = reinterpret_cast<float2*>(fsm + (size_t)L * 8);                       // holds rows [FH, F) of the staged line
  float2* bq = reinterpret_cast<float2*>(fsm + (size_t)L * 8 + Cfg::BUF);
  float* h1 = reinterpret_cast<float*>(fsm + (size_t)L * 8 + 2 * Cfg::BUF);         // rows [0, FH)
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  __shared__ uint64_t stage_bar;             // bulk staging: completion of a line's F row copies (both halves on one phase)
  if (threadIdx.x == 0) { mbar_init(&stage_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t stage_ph = 0;
  const float scale = 1.f / (float)L;
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;
  auto prefetch = [&](long long line, int part, float* h2) {      // 16 bytes = two channel pairs' worth of one half row
    if (line >= A.lines) return;
    if (A.bulk) {
      // the phase's single arrival (with the whole line's byte count) comes with part 0; part 1's copies only complete bytes
      if (part == 0 && threadIdx.x == 0) mbar_expect_tx(&stage_bar, (uint32_t)(F * W * 4));
      const int f0 = part == 0 ? 0 : FH, f1 = part == 0 ? FH : F;
      for (int f = f0 + threadIdx.x; f < f1; f += NT)
        bulk_load((part == 0 ? h1 + f * W : h2 + (f - FH) * W), A.in + ((long long)f * A.RA + line) * W, (uint32_t)(W * 4), &stage_bar);
      return;
    }
    const float* src = A.in + ((long long)jb * A.RA + line) * W + 4 * cp;
    const long long fstep = (long long)JS * A.RA * W;
    for (int f = jb; f < F; f += JS, src += fstep)
      if ((f < FH) == (part == 0)) cp_async16((part == 0 ? h1 + f * W : h2 + (f - FH) * W) + 4 * cp, src);
  };
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = A.N; walk.order = A.order;
  long long line = walk.at(0);
  prefetch(line, 0, nullptr); prefetch(line, 1, reinterpret_cast<float*>(bp));
  cp_async_commit();
  for (long long it = 0; line < A.lines; ++it) {
    const long long next = walk.at(it + 1);
    cp_async_wait_all();
    if (A.bulk) { mbar_wait(&stage_bar, stage_ph); stage_ph ^= 1; }
    __syncthreads();                       // both halves of the line have landed; the previous line's last pass is done with bq
    SpecHalfSrc<L, G> src; src.h1 = h1; src.h2 = reinterpret_cast<const float*>(bp); src.cp = cp;
    SmemCT<G> p; p.p = bp; SmemCT<G> q; q.p = bq;
    LineDstCT dst; dst.base = A.out + line * A.N * C + 2 * cp; dst.pstride = C; dst.N = A.N; dst.scale = scale;
    if (A.out1 && line >= A.lines0) {
      const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
      dst.base = A.out1 + (b * A.N * A.N + j) * C + 2 * cp; dst.pstride = A.N * C;
    }
    fft_pass_ct<R0, 1, L, G, NT>(src, q, tw, jb, cp);
    __syncthreads();                       // h1 and bp are free
    prefetch(next, 0, nullptr); cp_async_commit();
    fft_pass_ct<R1, R0, L, G, NT>(q, p, tw, jb, cp);
    __syncthreads();                       // bq is free
    prefetch(next, 1, reinterpret_cast<float*>(bq)); cp_async_commit();
    fft_pass_ct<R2, R0 * R1, L, G, NT>(p, dst, tw, jb, cp);
    float2* t = bp; bp = bq; bq = t;       // the next line's second half sits in what was bq
    line = next;
  }
  cp_async_wait_all();
}

// ---- two-pass transforms (fft2p.cuh): L = 48 M as (3M) x 16 forward / 16 x (3M) inverse, 16 G threads ---------------------
template <int M, int G> struct Fft2Cfg {
  // JB = butterfly slots per channel pair.  L = 384: 16 (pass 1 has 16 butterflies, the paired pass 13 units).  L = 192: the paired pass
  // has only 7 units, so 8 slots (the 16 butterflies of the other pass take two rounds) and twice the CTAs per SM instead of 9 idle slots.
  static constexpr int L = 48 * M, F = L / 2 + 1, JB = M == 4 ? 8 : 16, NT = JB * G;
  static constexpr size_t BUF = (size_t)L * G * 8;
  __host__ __device__ static size_t stage_fwd(int N) { return ((size_t)N * 2 * G * 4 + 127) & ~(size_t)127; }
  static size_t smem_fwd(int N) { return (size_t)L * 8 + BUF + 2 * stage_fwd(N); }
  static size_t smem_inv() { return (size_t)L * 8 + BUF + (size_t)F * 4 * G * 4; }
};
struct SpecEmitFwd {                   // X1 / X2 of a frequency -> bf16 hi / lo rows [re c.. | im c..] of the GEMM operand planes
  __nv_bfloat16* oh; __nv_bfloat16* ol; long long fstride; int C, padw;
  __device__ __forceinline__ void operator()(int f, float2 x1, float2 x2) const {
    __nv_bfloat16* ph = oh + f * fstride; __nv_bfloat16* pl = ol + f * fstride;
    uint2 lo; const uint2 hi = sp_pack_hi_lo(x1.x, x2.x, x1.y, x2.y, lo);
    *reinterpret_cast<uint2*>(ph) = hi; *reinterpret_cast<uint2*>(pl) = lo;
    // the row's pad is written too (a partly written sector costs a DRAM read-modify-write, see spec_fft_fwd_fast_k)
    if (padw == 4) { *reinterpret_cast<uint2*>(ph + 4) = make_uint2(0u, 0u); *reinterpret_cast<uint2*>(pl + 4) = make_uint2(0u, 0u); }
    else if (padw) for (int e = 0; e < padw; e += 2) { *reinterpret_cast<uint32_t*>(ph + 4 + e) = 0u; *reinterpret_cast<uint32_t*>(pl + 4 + e) = 0u; }
  }
};
template <int M, int G, int MINB, bool TABLE>
__global__ void __launch_bounds__(Fft2Cfg<M, G>::NT, MINB) spec_fft_fwd2_k(FftFwdArgs A) {
  using Cfg = Fft2Cfg<M, G>;
  constexpr int L = Cfg::L, C = 2 * G, NT = Cfg::NT, R0 = 3 * M, JB = Cfg::JB;
  extern __shared__ __align__(128) uint8_t fsm[];
  __shared__ float s_g[C], s_b[C], s_raw[2][C];
  __shared__ uint64_t stage_bar[2];          // completion of a contiguous line's bulk copy into staging tile 0 / 1
  if (threadIdx.x == 0) { mbar_init(&stage_bar[0], 1); mbar_init(&stage_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  const int N = A.N;                         // N C is a multiple of 4 (checked by the launcher): every line is whole 16-byte pieces
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float* stage0 = reinterpret_cast<float*>(fsm + (size_t)L * 8 + Cfg::BUF);
  const size_t sstride = Cfg::stage_fwd(N) / 4;
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  if (A.gam) for (int t = threadIdx.x; t < C; t += NT) { s_g[t] = A.gam[t] * BN_RS; s_b[t] = A.bet[t]; }
  const bool shifter = A.gam && A.bias0 && threadIdx.x < C;       // see spec_fft_fwd_fast_k: the line's bias element arrives with the line
  float sh_g = 0.f, sh_c = 0.f;
  if (shifter) { sh_g = A.gam[threadIdx.x] * BN_RS; sh_c = fmaf(2.f * A.b0[threadIdx.x], sh_g, A.bet[threadIdx.x]); }
  __syncthreads();
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = N; walk.order = A.order;
  uint32_t phases = 0;
  const int padw = cp == G - 1 ? A.KA - 2 * C : 0;
  const long long fstride = A.RA * A.KA;
  // One loop body for the pipeline's prologue and steady state (the code is long; a second copy of it costs instruction-cache
  // misses): iteration `it` transforms line it (when it >= 0) and, between its two passes, starts the copy of line it + 2 into the
  // staging tile that pass 1 has just freed.  Iterations -2 and -1 only start the copies of lines 0 and 1.
  long long line = -1, line1 = -1;           // lines it and it + 1 (line = -1: nothing to transform yet)
  for (long long it = -2;; ++it) {
    const int slot = (int)(it & 1);
    const long long line2 = walk.at(it + 2);
    if (it >= 0) {
      if (line >= A.lines) break;
      asm volatile("cp.async.wait_group 1;" ::: "memory");       // this line's per-thread copies have landed (the next line's may be in flight)
      if (A.bulk && (line < A.lines0 || A.dir1_strided != 1)) { mbar_wait(&stage_bar[slot], (phases >> slot) & 1u); phases ^= 1u << slot; }
      if (shifter) s_b[threadIdx.x] = fmaf(s_raw[slot][threadIdx.x], sh_g, sh_c);
      __syncthreads();                       // staged line and shift visible; the previous line's pass 2 is done with bufA
      FpBn bn; bn.on = A.gam != nullptr; bn.gx = bn.gy = 1.f; bn.bx = bn.by = 0.f;
      if (bn.on) { bn.gx = s_g[2 * cp]; bn.gy = s_g[2 * cp + 1]; bn.bx = s_b[2 * cp]; bn.by = s_b[2 * cp + 1]; }
#pragma unroll 1
      for (int j = jb; j < 16; j += JB) fp_fwd_pass1<M, G>(reinterpret_cast<const float2*>(stage0 + slot * sstride), bufA, j, cp, N, bn);
      __syncthreads();                       // bufA complete; this staging tile is free
    }
    if (line2 < A.lines) {                   // prefetch line it + 2
      float* stage = stage0 + slot * sstride;
      if (shifter) cp_async4(&s_raw[slot][threadIdx.x], (line2 >= A.lines0 ? A.bias1 + (line2 - A.lines0) * C : A.bias0 + line2 * C) + threadIdx.x);
      if (line2 < A.lines0 || A.dir1_strided != 1) {
        const float* base = A.in + line2 * N * C;
        if (A.dir1_strided == 2) {
          const bool d1 = line2 >= A.lines0; const long long l1 = d1 ? line2 - A.lines0 : line2; const long long b = l1 / N; const int x = (int)(l1 - b * N);
          base = (d1 ? A.in1 : A.in) + (((b / 128) * N + x) * 128 + (b % 128)) * N * C;
        }
        // a contiguous line is ONE bulk copy issued by one thread (no LDGSTS instructions); the others wait on the mbarrier's phase
        if (A.bulk) { if (threadIdx.x == 0) { mbar_expect_tx(&stage_bar[slot], (uint32_t)(N * C * 4)); bulk_load(stage, base, (uint32_t)(N * C * 4), &stage_bar[slot]); } }
        else { for (int t = threadIdx.x; t < N * C / 4; t += NT) cp_async16(stage + 4 * t, base + 4 * t); }
      } else {                               // column line of a [b, N, N, C] tensor: one C-float piece per position
        const long long l1 = line2 - A.lines0; const long long b = l1 / N; const int j = (int)(l1 - b * N);
        const float* base = A.in + (b * N * N + j) * C; const long long ps = (long long)N * C;
        if constexpr (G % 2 == 0) {
          constexpr int Q = G / 2;
          for (int t = threadIdx.x; t < N * Q; t += NT) { const int pos = t / Q, k = t - pos * Q; cp_async16(stage + pos * C + 4 * k, base + pos * ps + 4 * k); }
        } else {
          for (int pos = jb; pos < N; pos += JB) cp_async8(stage + 2 * (pos * G + cp), base + pos * ps + 2 * cp);
        }
      }
    }
    cp_async_commit();
    if (it >= 0 && jb <= R0 / 2) {
      SpecEmitFwd em; em.oh = A.oh + line * A.KA + 4 * cp; em.ol = A.ol + line * A.KA + 4 * cp; em.fstride = fstride; em.C = C; em.padw = padw;
      fp_fwd_pass2<M, G, TABLE>(bufA, tw, jb, cp, em);
    }
    line = line1; line1 = line2;
  }
  cp_async_wait_all();
}
struct SpecEmitInv {                   // inverse tail: (re, im) swapped back, scaled, first N positions only
  float* base; int pstride; int N; float scale;
  __device__ __forceinline__ void operator()(int pos, float2 v) const {
    if (pos < N) *reinterpret_cast<float2*>(base + (size_t)(unsigned)(pos * pstride)) = make_float2(v.y * scale, v.x * scale);
  }
};
template <int M, int G, int MINB, bool TABLE>
__global__ void __launch_bounds__(Fft2Cfg<M, G>::NT, MINB) spec_fft_inv2_k(FftInvArgs A) {
  using Cfg = Fft2Cfg<M, G>;
  constexpr int L = Cfg::L, F = Cfg::F, C = 2 * G, W = 4 * G, NT = Cfg::NT, R1 = 3 * M, JB = Cfg::JB;
  extern __shared__ __align__(128) uint8_t fsm[];
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float* stage = reinterpret_cast<float*>(fsm + (size_t)L * 8 + Cfg::BUF);
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  __shared__ uint64_t stage_bar;
  if (threadIdx.x == 0) { mbar_init(&stage_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t stage_ph = 0;
  const float scale = 1.f / (float)L;
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;
  auto prefetch = [&](long long line) {
    if (line >= A.lines) return;
    if (A.bulk) {                              // one bulk copy per spectrum row (W floats), all counted on one mbarrier phase
      if (threadIdx.x == 0) mbar_expect_tx(&stage_bar, (uint32_t)(F * W * 4));
      for (int f = threadIdx.x; f < F; f += NT) bulk_load(stage + f * W, A.in + ((long long)f * A.RA + line) * W, (uint32_t)(W * 4), &stage_bar);
      return;
    }
    const float* src = A.in + ((long long)jb * A.RA + line) * W + 4 * cp;     // 16 bytes = two channel pairs' worth of one half row
    const long long fstep = (long long)JB * A.RA * W;
    for (int f = jb; f < F; f += JB, src += fstep) cp_async16(stage + f * W + 4 * cp, src);
  };
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = A.N; walk.order = A.order;
  long long line = walk.at(0);
  prefetch(line);
  cp_async_commit();
  for (long long it = 0; line < A.lines; ++it) {
    const long long next = walk.at(it + 1);
    cp_async_wait_all();
    if (A.bulk) { mbar_wait(&stage_bar, stage_ph); stage_ph ^= 1; }
    __syncthreads();                       // the line's rows have landed; the previous line's pass 2 is done with bufA
    if (jb <= R1 / 2) fp_inv_pass1<M, G>(stage, bufA, jb, cp);
    __syncthreads();                       // bufA complete; the staging area is free
    prefetch(next); cp_async_commit();
    SpecEmitInv dst; dst.base = A.out + line * A.N * C + 2 * cp; dst.pstride = C; dst.N = A.N; dst.scale = scale;
    if (A.out1 && line >= A.lines0) {
      const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
      dst.base = A.out1 + (b * A.N * A.N + j) * C + 2 * cp; dst.pstride = A.N * C;
    }
#pragma unroll 1
    for (int k = jb; k < 16; k += JB) fp_inv_pass2<M, G, TABLE>(bufA, tw, k, cp, dst);
    line = next;
  }
  cp_async_wait_all();
}

// ---- three-pass transforms for L = 1536 (N = 769..1024): 24 (pruned) x 8 x 8 (paired) forward, 8 (paired) x 8 x 24 (pruned) inverse.
// A line is 12 KB of shared memory per channel pair, so a CTA works on (line, group of G = 5 channel pairs) items: 64 G threads,
// ping-pong buffers, the group's 40-byte (forward) / 80-byte (inverse) pieces of the line staged by cp.async one / two items ahead.
template <int G> struct Fft3Cfg {
  static constexpr int L = 1536, F = L / 2 + 1, NT = 64 * G;
  static constexpr size_t BUF = (size_t)L * G * 8;
  __host__ __device__ static size_t stage_fwd(int N) { return ((size_t)N * 2 * G * 4 + 127) & ~(size_t)127; }
  static size_t smem_fwd(int N) { return (size_t)L * 8 + 2 * BUF + 2 * stage_fwd(N); }
  static size_t smem_inv() { return (size_t)L * 8 + 2 * BUF + (size_t)F * 4 * G * 4; }
};
template <int G>
__global__ void __launch_bounds__(64 * G, 1) spec_fft_fwd3_k(FftFwdArgs A) {
  using Cfg = Fft3Cfg<G>;
  constexpr int L = Cfg::L, NT = Cfg::NT;
  extern __shared__ __align__(128) uint8_t fsm[];
  __shared__ float s_g[64], s_c[64], s_b[2 * G], s_raw[2][2 * G];
  const int N = A.N, C = A.C, NG = C / (2 * G);          // C is a multiple of 2G (checked by the launcher)
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float2* bufB = reinterpret_cast<float2*>(fsm + (size_t)L * 8 + Cfg::BUF);
  float* stage0 = reinterpret_cast<float*>(fsm + (size_t)L * 8 + 2 * Cfg::BUF);
  const size_t sstride = Cfg::stage_fwd(N) / 4;
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  const bool bn = A.gam != nullptr, shift = bn && A.bias0 != nullptr;
  if (bn) for (int t = threadIdx.x; t < C; t += NT) {        // x <- relu(x s_g + s_c (+ line bias * s_g))
    s_g[t] = A.gam[t] * BN_RS;
    s_c[t] = shift ? fmaf(2.f * A.b0[t], A.gam[t] * BN_RS, A.bet[t]) : A.bet[t];
  }
  __syncthreads();
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = N; walk.order = A.order;
  const long long fstride = A.RA * A.KA;
  // items it = (line index it / NG, group it % NG); iteration `it` transforms item it and, after its first pass, starts the copies of
  // item it + 2 into the staging tile that pass has just freed (iterations -2, -1 only start the copies of items 0, 1)
  for (long long it = -2;; ++it) {
    const int slot = (int)(it & 1);
    const long long it2 = it + 2, line2 = walk.at(it2 / NG), line = it >= 0 ? walk.at(it / NG) : -1;
    const int grp = it >= 0 ? (int)(it % NG) : 0, grp2 = (int)(it2 % NG);
    if (it >= 0) {
      if (line >= A.lines) break;
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (shift && threadIdx.x < 2 * G) s_b[threadIdx.x] = fmaf(s_raw[slot][threadIdx.x], s_g[2 * G * grp + threadIdx.x], s_c[2 * G * grp + threadIdx.x]);
      __syncthreads();                       // staged piece and shift visible; the previous item's last pass is done with bufB
      FpBn p; p.on = bn; p.gx = p.gy = 1.f; p.bx = p.by = 0.f;
      if (bn) {
        const int c = 2 * (G * grp + cp);
        p.gx = s_g[c]; p.gy = s_g[c + 1];
        p.bx = shift ? s_b[2 * cp] : s_c[c]; p.by = shift ? s_b[2 * cp + 1] : s_c[c + 1];
      }
      fp_fwd_pass1_t<8, G, 64>(reinterpret_cast<const float2*>(stage0 + slot * sstride), bufA, jb, cp, N, p);
      __syncthreads();                       // bufA complete; this staging tile is free
    }
    if (line2 < A.lines) {                   // prefetch item it + 2: the group's 2G floats of every position
      float* stage = stage0 + slot * sstride;
      const float* base; long long ps = C;
      if (line2 < A.lines0 || A.dir1_strided != 1) {
        base = A.in + line2 * N * C;
        if (A.dir1_strided == 2) {
          const bool d1 = line2 >= A.lines0; const long long l1 = d1 ? line2 - A.lines0 : line2; const long long b = l1 / N; const int x = (int)(l1 - b * N);
          base = (d1 ? A.in1 : A.in) + (((b / 128) * N + x) * 128 + (b % 128)) * N * C;
        }
      } else {
        const long long l1 = line2 - A.lines0; const long long b = l1 / N; const int j = (int)(l1 - b * N);
        base = A.in + (b * N * N + j) * C; ps = (long long)N * C;
      }
      base += 2 * G * grp2;
      if (shift && threadIdx.x < 2 * G)
        cp_async4(&s_raw[slot][threadIdx.x], (line2 >= A.lines0 ? A.bias1 + (line2 - A.lines0) * C : A.bias0 + line2 * C) + 2 * G * grp2 + threadIdx.x);
      for (int pos = jb; pos < N; pos += 64) cp_async8(stage + 2 * (pos * G + cp), base + pos * ps + 2 * cp);
    }
    cp_async_commit();
    if (it >= 0) {
#pragma unroll 1
      for (int i = 0; i < 3; ++i) fp_mid8<G, 24, 192, 8>(bufA, bufB, tw, jb + 64 * i, cp);
      __syncthreads();                       // bufB complete
      const int cg = G * grp + cp;           // channel pair of the whole line
      SpecEmitFwd em; em.oh = A.oh + line * A.KA + 4 * cg; em.ol = A.ol + line * A.KA + 4 * cg; em.fstride = fstride; em.C = C;
      em.padw = 2 * cg + 2 == C ? A.KA - 2 * C : 0;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) { const int u = jb + 64 * i; if (u <= 96) fp_fwd_last_t<8, 192, G, true>(bufB, tw, u, cp, em); }
    }
  }
  cp_async_wait_all();
}
template <int G>
__global__ void __launch_bounds__(64 * G, 1) spec_fft_inv3_k(FftInvArgs A) {
  using Cfg = Fft3Cfg<G>;
  constexpr int L = Cfg::L, F = Cfg::F, NT = Cfg::NT;
  extern __shared__ __align__(128) uint8_t fsm[];
  float2* tw = reinterpret_cast<float2*>(fsm);
  float2* bufA = reinterpret_cast<float2*>(fsm + (size_t)L * 8);
  float2* bufB = reinterpret_cast<float2*>(fsm + (size_t)L * 8 + Cfg::BUF);
  float* stage = reinterpret_cast<float*>(fsm + (size_t)L * 8 + 2 * Cfg::BUF);
  for (int t = threadIdx.x; t < L; t += NT) tw[t] = A.tw[t];
  __syncthreads();
  const int C = A.C, W = 2 * C, NG = C / (2 * G);
  const float scale = 1.f / (float)L;
  const int jb = threadIdx.x / G, cp = threadIdx.x - jb * G;
  LineWalk walk; walk.lines = A.lines; walk.lines0 = A.lines0; walk.N = A.N; walk.order = A.order;
  // iteration `it` transforms item it = (line index it / NG, group it % NG); the rows of item it + 1 load under its passes 2 and 3
  for (long long it = -1;; ++it) {
    const long long line = it >= 0 ? walk.at(it / NG) : -1, next = walk.at((it + 1) / NG);
    const int grp = it >= 0 ? (int)(it % NG) : 0, grpn = (int)((it + 1) % NG);
    if (it >= 0) {
      if (line >= A.lines) break;
      cp_async_wait_all();
      __syncthreads();                       // the item's rows have landed; the previous item's last pass is done with bufB
#pragma unroll 1
      for (int i = 0; i < 2; ++i) { const int u = jb + 64 * i; if (u <= 96) fp_inv_first_t<8, 192, G>(stage, bufA, u, cp); }
      __syncthreads();                       // bufA complete; the staging area is free
    }
    if (next < A.lines) {                    // 16 bytes = one channel pair's [re, re, im, im] of one row
      const float* src = A.in + ((long long)jb * A.RA + next) * W + 4 * (G * grpn + cp);
      const long long fstep = 64LL * A.RA * W;
      for (int f = jb; f < F; f += 64, src += fstep) cp_async16(stage + (f * G + cp) * 4, src);
    }
    cp_async_commit();
    if (it >= 0) {
#pragma unroll 1
      for (int i = 0; i < 3; ++i) fp_mid8<G, 8, 192, 24>(bufA, bufB, tw, jb + 64 * i, cp);
      __syncthreads();                       // bufB complete
      const int cg = G * grp + cp;
      SpecEmitInv dst; dst.base = A.out + line * A.N * C + 2 * cg; dst.pstride = C; dst.N = A.N; dst.scale = scale;
      if (A.out1 && line >= A.lines0) {
        const long long l1 = line - A.lines0; const long long b = l1 / A.N; const int j = (int)(l1 - b * A.N);
        dst.base = A.out1 + (b * A.N * A.N + j) * C + 2 * cg; dst.pstride = A.N * C;
      }
      fp_inv_last_t<8, G, 64, false>(bufB, tw, jb, cp, dst);
    }
  }
  cp_async_wait_all();
}

// ---- per-step weight spectra and GEMM operand staging -----------------------------------------------------------
// G^[f][c][q] = sum_t w1[t][c][q] exp(-2 pi i f (p - t) / L)   (double accumulation over the exact table)
//   fwd   B (K-major, [f][n][k], k < 128): n = q: (k=c: Gr, k=C1+c: -Gi);  n = C2+q: (k=c: Gi, k=C1+c: Gr)
//   dgrad B ([f][n][k], k < 64):           n = c: (k=q: Gr, k=C2+q: Gi);   n = C1+c: (k=q: -Gi, k=C2+q: Gr)
__global__ void spec_stage_weights_k(const float* __restrict__ w1, const double2* __restrict__ twd, __nv_bfloat16* __restrict__ Bfh,
                                     __nv_bfloat16* __restrict__ Bfl, __nv_bfloat16* __restrict__ Bdh, __nv_bfloat16* __restrict__ Bdl,
                                     int N, int L, int F, int C1, int C2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)F * C1 * C2) return;
  const int q = (int)(idx % C2), c = (int)((idx / C2) % C1), f = (int)(idx / ((long long)C1 * C2));
  const int p = (N - 1) / 2;
  double gr = 0.0, gi = 0.0;
  for (int t = 0; t < N; ++t) {
    int m = (p - t) % L; if (m < 0) m += L;
    const double2 e = twd[(int)(((long long)f * m) % L)];
    const double w = (double)w1[((size_t)t * C1 + c) * C2 + q];
    gr += w * e.x; gi += w * e.y;
  }
  const float Gr = (float)gr, Gi = (float)gi;
  auto put = [](__nv_bfloat16* hi, __nv_bfloat16* lo, size_t o, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v); hi[o] = h; lo[o] = __float2bfloat16_rn(v - __bfloat162float(h)); };
  const size_t bf = (size_t)f * SP_NF * 128, bd = (size_t)f * SP_ND * 64;
  const int cr = sp_kre(c), ci = sp_kim(c), qr = sp_kre(q), qi = sp_kim(q);      // row positions of (re, im) of channel c / q
  put(Bfh, Bfl, bf + (size_t)qr * 128 + cr, Gr);          put(Bfh, Bfl, bf + (size_t)qr * 128 + ci, -Gi);
  put(Bfh, Bfl, bf + (size_t)qi * 128 + cr, Gi);          put(Bfh, Bfl, bf + (size_t)qi * 128 + ci, Gr);
  put(Bdh, Bdl, bd + (size_t)cr * 64 + qr, Gr);           put(Bdh, Bdl, bd + (size_t)cr * 64 + qi, Gi);
  put(Bdh, Bdl, bd + (size_t)ci * 64 + qr, -Gi);          put(Bdh, Bdl, bd + (size_t)ci * 64 + qi, Gr);
}
// dw1[t][c][q] += (1/L) sum_f kappa_f Re(dG^[f][c][q] e^{+2 pi i f (p - t) / L}),  dG^ = conj(Y^) dO^ from the 2x2 blocks of P:
//   dGr = P[c][q] + P[C1+c][C2+q],  dGi = P[c][C2+q] - P[C1+c][q]      (P[f]: [128][SP_NF])
__global__ void spec_wgrad_finalize_k(const float* __restrict__ P, const double2* __restrict__ twd, float* __restrict__ dw1,
                                      int N, int L, int F, int C1, int C2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * C1 * C2) return;
  const int q = (int)(idx % C2), c = (int)((idx / C2) % C1), t = (int)(idx / ((long long)C1 * C2));
  const int p = (N - 1) / 2;
  int m = (p - t) % L; if (m < 0) m += L;
  double acc = 0.0;
  for (int f = 0; f < F; ++f) {
    const float* Pf = P + (size_t)f * 128 * SP_NF;
    const int cr = sp_kre(c), ci = sp_kim(c), qr = sp_kre(q), qi = sp_kim(q);
    const double gr = (double)Pf[cr * SP_NF + qr] + (double)Pf[ci * SP_NF + qi];
    const double gi = (double)Pf[cr * SP_NF + qi] - (double)Pf[ci * SP_NF + qr];
    const double2 e = twd[(int)(((long long)f * m) % L)];        // e^{-i theta}: cos = e.x, sin(theta) = -e.y
    const double kap = (f == 0 || 2 * f == L) ? 1.0 : 2.0;
    acc += kap * (gr * e.x + gi * e.y);                           // Re((gr + i gi)(cos + i sin)) = gr cos - gi sin
  }
  dw1[idx] += (float)(acc / (double)L);
}

// ---- per-frequency channel-mix GEMM on the tensor cores -------------------------------------------------------------
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct SpecGemmArgs {
  float* out;            // [F][RA][WOUT] fp32
  long long RA, rows;    // allocated / valid lines per frequency
  int F, MT;             // frequencies, 128-line tiles per frequency
  long long items;       // F * MT
};
// BN: MMA N (48 fwd / 112 dgrad); KSTEPS: 16-wide K steps actually multiplied (7 / 3); ABOX: 64-wide K boxes per plane (2 / 1);
// WOUT: valid output columns (40 / 100).  192 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
template <int BN, int KSTEPS, int ABOX, int WOUT, int NSTAGE>
__global__ void __launch_bounds__(192, 1) spec_gemm_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                      const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                      SpecGemmArgs P) {
  constexpr int A_BOX = 128 * 64 * 2;              // 16 KB
  constexpr int A_PLANE = ABOX * A_BOX;
  constexpr int A_STAGE = 2 * A_PLANE;
  constexpr int B_BOX = BN * 64 * 2;
  constexpr int B_PLANE = ABOX * B_BOX;
  constexpr int B_BYTES = 2 * B_PLANE;
  // Two accumulators per item: a short-N tcgen05.mma into the tile the previous one is still updating costs ~170-300 cycles of
  // latency (measured), so the 3 x KSTEPS accumulates of an item are split into two independent chains (summed by the
  // epilogue) instead of one.
  constexpr int ACC_COLS = BN <= 64 ? 64 : 128;
  constexpr int SLOT_COLS = 2 * ACC_COLS;
  constexpr int NSLOT = 512 / SLOT_COLS;
  constexpr int STG_BYTES = 128 * WOUT * 4;
  constexpr int NLD = (WOUT + 7) / 8;              // tcgen05.ld x8 per row
  static_assert(NLD * 8 <= BN, "epilogue reads past the accumulator");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem + (size_t)NSTAGE * A_STAGE;
  float* stg = reinterpret_cast<float*>(sB + B_BYTES);
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], bfull, bempty, acc_full[NSLOT], acc_empty[NSLOT];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (P.items + gridDim.x - 1) / gridDim.x;
  const long long item0 = (long long)blockIdx.x * per;
  long long item1 = item0 + per; if (item1 > P.items) item1 = P.items;
  const int nit = item1 > item0 ? (int)(item1 - item0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&bfull, 1); mbar_init(&bempty, 1);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int cur_f = -1, nb = 0;
      for (int it = 0; it < nit; ++it) {
        const long long id = item0 + it; const int f = (int)(id / P.MT), mt = (int)(id - (long long)f * P.MT);
        if (f != cur_f) {
          mbar_wait(&bempty, (nb & 1) ^ 1);            // every MMA that read the previous frequency's B has retired
          mbar_expect_tx(&bfull, B_BYTES);
#pragma unroll
          for (int b = 0; b < ABOX; ++b) {
            tma_load_3d(sB + b * B_BOX, &tmBh, &bfull, 64 * b, 0, f);
            tma_load_3d(sB + B_PLANE + b * B_BOX, &tmBl, &bfull, 64 * b, 0, f);
          }
          cur_f = f; ++nb;
        }
        const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * A_STAGE;
        mbar_expect_tx(&full_bar[s], A_STAGE);
#pragma unroll
        for (int b = 0; b < ABOX; ++b) {
          tma_load_3d(st + b * A_BOX, &tmAh, &full_bar[s], 64 * b, mt * 128, f);
          tma_load_3d(st + A_PLANE + b * A_BOX, &tmAl, &full_bar[s], 64 * b, mt * 128, f);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, BN, 0, 0);
      int cur_f = -1, nb = 0;
      const uint32_t sb = smem_u32(sB);
      for (int it = 0; it < nit; ++it) {
        const long long id = item0 + it; const int f = (int)(id / P.MT);
        if (f != cur_f) { mbar_wait(&bfull, nb & 1); ++nb; cur_f = f; }
        const int slot = it % NSLOT;
        mbar_wait(&acc_empty[slot], ((it / NSLOT) & 1) ^ 1);
        const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * A_STAGE);
        const uint32_t td = tmem_d + slot * SLOT_COLS;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const int box = ks >> 2, kk = ks & 3;
          const uint64_t ah = umma_desc(sa + box * A_BOX + kk * 32, 16, 1024, 2ull);
          const uint64_t al = umma_desc(sa + A_PLANE + box * A_BOX + kk * 32, 16, 1024, 2ull);
          const uint64_t bh = umma_desc(sb + box * B_BOX + kk * 32, 16, 1024, 2ull);
          const uint64_t bl = umma_desc(sb + B_PLANE + box * B_BOX + kk * 32, 16, 1024, 2ull);
          // chain 0: hi.hi (+ hi.lo on even k steps);  chain 1: lo.hi (+ hi.lo on odd k steps)
          umma_bf16(td, ah, bh, idesc, ks ? 1u : 0u);
          umma_bf16(td + ACC_COLS, al, bh, idesc, ks ? 1u : 0u);
          umma_bf16((ks & 1) ? td + ACC_COLS : td, ah, bl, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
        umma_commit(&acc_full[slot]);
        const bool last_of_f = (it == nit - 1) || ((int)((id + 1) / P.MT) != f);
        if (last_of_f) umma_commit(&bempty);
      }
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;                  // row of the tile
    for (int it = 0; it < nit; ++it) {
      const long long id = item0 + it; const int f = (int)(id / P.MT), mt = (int)(id - (long long)f * P.MT);
      const int slot = it % NSLOT;
      mbar_wait(&acc_full[slot], (it / NSLOT) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) bulk_wait_read();    // the previous bulk store has finished reading the staging tile
      epi_bar();
      const uint32_t ta = tmem_d + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
      float* srow = stg + (size_t)r * WOUT;
#pragma unroll
      for (int c0 = 0; c0 < NLD * 8; c0 += 8) {
        uint32_t v[8], v2[8];
        tmem_ld8(ta + c0, v);
        tmem_ld8(ta + ACC_COLS + c0, v2);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        if (c0 + 8 <= WOUT) {
          *reinterpret_cast<float4*>(srow + c0) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          *reinterpret_cast<float4*>(srow + c0 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) if (c0 + i < WOUT) srow[c0 + i] = __uint_as_float(v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      fence_async_smem();
      epi_bar();
      if (threadIdx.x == 64) {
        const long long m0 = (long long)mt * 128;
        long long nr = P.rows - m0; if (nr > 128) nr = 128;
        if (nr > 0) bulk_store(P.out + ((long long)f * P.RA + m0) * WOUT, stg, (uint32_t)(nr * WOUT * 4));
      }
    }
    if (threadIdx.x == 64) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

// ---- wgrad: P[f][m][n] += sum_lines A^[f][line][m] dO^[f][line][n]  (m < 128 over [re c | im c], n < 48 over [re q | im q]) ----
struct SpecWgradArgs {
  float* P;              // [F][128][SP_NF]
  long long rows;        // valid lines
  int F, KS;             // frequencies, K splits per frequency
  int MV, NV;            // valid m / n extents (2*C1, 2*C2)
  int stages;            // TMA ring depth (<= SP_WSTAGES_MAX), 48 KB per stage
};
#define SP_WKC 64
#define SP_WSTAGES 3
#define SP_WSTAGES_MAX 4
#define SP_WGROUP 8      /* K chunks per TMEM accumulation group: 8 * 4 k-steps * 3 passes = 96 accumulates over two chains */
__global__ void __launch_bounds__(TC_THREADS, 1) spec_wgrad_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                              const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                              SpecWgradArgs P) {
  constexpr int BLK = SP_WKC * 128;                 // one [SP_WKC k-rows x 64 mn] box
  constexpr int A_PLANE = 2 * BLK, B_PLANE = BLK;
  constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;    // 48 KB (35.8 KB of them real data: 100 / 40 of the 128 / 64 columns)
  constexpr int HC = SP_NF / 2;                     // accumulator columns per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[SP_WSTAGES_MAX], empty_bar[SP_WSTAGES_MAX], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long units = (long long)P.F * P.KS;
  const long long upc = (units + gridDim.x - 1) / gridDim.x;
  const long long u0 = (long long)blockIdx.x * upc;
  long long u1 = u0 + upc; if (u1 > units) u1 = units;
  const long long kchunks = (P.rows + SP_WKC - 1) / SP_WKC;
  const long long cper = (kchunks + P.KS - 1) / P.KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SP_WSTAGES_MAX; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);     // 2 slots x 2 accumulate chains (one per 16-line K step of a chunk) x 64 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (long long u = u0; u < u1; ++u) {
        const int f = (int)(u / P.KS), ks = (int)(u - (long long)f * P.KS);
        const long long c_lo = (long long)ks * cper; long long c_hi = c_lo + cper; if (c_hi > kchunks) c_hi = kchunks;
        for (long long ch = c_lo; ch < c_hi; ++ch, ++it) {
          const int s = it % P.stages; const uint32_t ph = (it / P.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          const int r0 = (int)(ch * SP_WKC);
          tma_load_3d(st, &tmAh, &full_bar[s], 0, r0, f);
          tma_load_3d(st + BLK, &tmAh, &full_bar[s], 64, r0, f);
          tma_load_3d(st + A_PLANE, &tmAl, &full_bar[s], 0, r0, f);
          tma_load_3d(st + A_PLANE + BLK, &tmAl, &full_bar[s], 64, r0, f);
          tma_load_3d(st + 2 * A_PLANE, &tmBh, &full_bar[s], 0, r0, f);
          tma_load_3d(st + 2 * A_PLANE + B_PLANE, &tmBl, &full_bar[s], 0, r0, f);
        }
      }
    }
  } else if (warp == 1) {
    {     // the whole warp runs the issue loop converged (umma_bf16_conv elects the issuing lane); ring slot and phase by counters
      constexpr uint32_t idesc = umma_idesc(128, SP_NF, 1, 1);
      int s = 0, gcount = 0; uint32_t ph = 0;
      for (long long u = u0; u < u1; ++u) {
        const int ks = (int)(u % P.KS);
        const long long c_lo = (long long)ks * cper; long long c_hi = c_lo + cper; if (c_hi > kchunks) c_hi = kchunks;
        const int nk = c_hi > c_lo ? (int)(c_hi - c_lo) : 0;
        for (int i = 0; i < nk; ++i) {
          const int gi = i % SP_WGROUP;
          const int slot = gcount & 1;
          if (gi == 0) { mbar_wait(&acc_empty[slot], ((gcount >> 1) & 1) ^ 1); tc_fence_after(); }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint32_t td0 = tmem_d + slot * 128;
#pragma unroll
          for (int k = 0; k < SP_WKC / 16; ++k) {
            const uint64_t ah = umma_desc_sw128(sa + k * 2048, BLK, 1024);
            const uint64_t al = umma_desc_sw128(sa + A_PLANE + k * 2048, BLK, 1024);
            const uint64_t bh = umma_desc_sw128(sa + 2 * A_PLANE + k * 2048, BLK, 1024);
            const uint64_t bl = umma_desc_sw128(sa + 2 * A_PLANE + B_PLANE + k * 2048, BLK, 1024);
            const uint32_t td = td0 + (k & 1) * 64;      // two independent chains: no back-to-back accumulates into one tile
            umma_bf16_conv(td, ah, bh, idesc, (gi | (k >> 1)) ? 1u : 0u);
            umma_bf16_conv(td, ah, bl, idesc, 1u);
            umma_bf16_conv(td, al, bh, idesc, 1u);
          }
          umma_commit_conv(&empty_bar[s]);
          if (gi == SP_WGROUP - 1 || i == nk - 1) { umma_commit_conv(&acc_full[slot]); ++gcount; }
          if (++s == P.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    float acc[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) acc[i] = 0.f;
    int gcount = 0;
    for (long long u = u0; u < u1; ++u) {
      const int f = (int)(u / P.KS), ks = (int)(u - (long long)f * P.KS);
      const long long c_lo = (long long)ks * cper; long long c_hi = c_lo + cper; if (c_hi > kchunks) c_hi = kchunks;
      const int nk = c_hi > c_lo ? (int)(c_hi - c_lo) : 0;
      const int ng = (nk + SP_WGROUP - 1) / SP_WGROUP;
      for (int g = 0; g < ng; ++g, ++gcount) {
        const int slot = gcount & 1;
        mbar_wait(&acc_full[slot], (gcount >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem_d + slot * 128 + ((uint32_t)(q * 32) << 16) + half * HC;
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 8) {
          uint32_t v[8], v2[8];
          tmem_ld8(ta + c0, v);
          tmem_ld8(ta + 64 + c0, v2);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c0 + i] += __uint_as_float(v[i]) + __uint_as_float(v2[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
      }
      // flush when the next unit belongs to another frequency (or this CTA is done)
      const bool flush = (u == u1 - 1) || ((int)((u + 1) / P.KS) != f);
      if (flush) {
        if (m < P.MV) {
          float* Pf = P.P + ((size_t)f * 128 + m) * SP_NF + half * HC;
#pragma unroll
          for (int i = 0; i < HC; ++i) if (half * HC + i < P.NV) atomicAdd(Pf + i, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < HC; ++i) acc[i] = 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 256); }
}

// ------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------
struct SpecState {
  int ready, N, C1, C2, G1, G2;
  FftPlan pl;
  long long RA;                          // allocated lines per frequency (2 * chunk graphs * N)
  float2* tw; double2* twd;
  __nv_bfloat16 *Ah, *Al, *Dh, *Dl;      // Y^ planes [F][RA][SP_KA1], dO^ planes [F][RA][SP_KA2]
  float *Oc, *dYc;                       // O^ [F][RA][2 C2], dY^ [F][RA][2 C1]
  __nv_bfloat16 *Bfh, *Bfl, *Bdh, *Bdl;  // staged weight spectra
  float* P;                              // wgrad accumulator [F][128][SP_NF]
  CUtensorMap mBfh, mBfl, mBdh, mBdl;
  int fft_threads, fft_threads_generic, generic_only, grid_sms, fft_order, fft_order_inv, fft_inv2, wgrad_stages, fft_bulk, fft_bulk_inv, fft_2pass;
};
static size_t spec_fft_smem(int L, int G) { return (size_t)(L + 2 * (size_t)L * G) * sizeof(float2); }
static constexpr int SPF_STAGES = 2, SPD_STAGES = 4;
static size_t spec_gemm_smem(int BN, int ABOX, int WOUT, int NSTAGE) {
  return (size_t)NSTAGE * 2 * ABOX * 128 * 64 * 2 + (size_t)2 * ABOX * BN * 64 * 2 + (size_t)128 * WOUT * 4 + 1024;
}
static const size_t SP_WGRAD_SMEM = (size_t)SP_WSTAGES_MAX * (2 * 2 * SP_WKC * 128 + 2 * SP_WKC * 128) + 1024;

// transform length: smallest even L in {2^a, 3 * 2^a} with L >= N + q, q = N - 1 - (N-1)/2
static int spec_pick_L(int N) {
  const int need = N + (N - 1 - (N - 1) / 2);
  int best = 0;
  for (int a = 1; a < 20; ++a) {
    const int c2 = 1 << a, c3 = 3 << a;
    if (c2 >= need && (!best || c2 < best)) best = c2;
    if (c3 >= need && (!best || c3 < best)) best = c3;
    if (best && c2 > best) break;
  }
  return best < 4 ? 4 : best;
}
static void spec_make_plan(FftPlan& pl, int L) {
  pl.L = L; pl.F = L / 2 + 1; pl.npass = 0;
  int n = L;
  if (n % 3 == 0) { n /= 3; if (n % 2 == 0) { n /= 2; pl.rad[pl.npass++] = 6; } else pl.rad[pl.npass++] = 3; }
  while (n % 8 == 0) { pl.rad[pl.npass++] = 8; n /= 8; }
  if (n > 1) pl.rad[pl.npass++] = n;    // 2 or 4
}
static int spec_pick_G(int L, int CP) {
  int g = CP;
  while (g > 1 && spec_fft_smem(L, g) > 200 * 1024) --g;
  return g;
}
static int spec_enc3(CUtensorMap* tm, const void* base, long long d0, long long d1, long long d2, long long s1_b, long long s2_b, int b0, int b1) {
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t str[2] = {(cuuint64_t)s1_b, (cuuint64_t)s2_b};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
  return tc_encode(tm, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
static void spec_destroy(SpecState& s) {
  void* ps[] = {s.tw, s.twd, s.Ah, s.Al, s.Dh, s.Dl, s.Oc, s.dYc, s.Bfh, s.Bfl, s.Bdh, s.Bdl, s.P};
  for (void* p : ps) if (p) cudaFree(p);
  memset(&s, 0, sizeof s);
}
static int spec_init(SpecState& s, int N, long long rows_alloc, cudaStream_t st) {
  memset(&s, 0, sizeof s);
  if (tc_global_init()) return -1;
  s.N = N; s.C1 = TC_C1; s.C2 = TC_C2; s.RA = rows_alloc;
  const int L = spec_pick_L(N);
  spec_make_plan(s.pl, L);
  const int F = s.pl.F;
  s.G1 = spec_pick_G(L, s.C1 / 2); s.G2 = spec_pick_G(L, s.C2 / 2);
  std::vector<float2> tw(L); std::vector<double2> twd(L);
  for (int k = 0; k < L; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)L;
    twd[k].x = cos(a); twd[k].y = sin(a); tw[k].x = (float)twd[k].x; tw[k].y = (float)twd[k].y;
  }
  const size_t nA = (size_t)F * rows_alloc * SP_KA1 + 256, nD = (size_t)F * rows_alloc * SP_KA2 + 256;
  const size_t nO = (size_t)F * rows_alloc * 2 * s.C2, nY = (size_t)F * rows_alloc * 2 * s.C1;
  const size_t nBf = (size_t)F * SP_NF * 128, nBd = (size_t)F * SP_ND * 64, nP = (size_t)F * 128 * SP_NF;
  if (cudaMalloc(&s.tw, L * sizeof(float2)) || cudaMalloc(&s.twd, L * sizeof(double2)) || cudaMalloc(&s.Ah, nA * 2) || cudaMalloc(&s.Al, nA * 2) ||
      cudaMalloc(&s.Dh, nD * 2) || cudaMalloc(&s.Dl, nD * 2) || cudaMalloc(&s.Oc, nO * 4) || cudaMalloc(&s.dYc, nY * 4) ||
      cudaMalloc(&s.Bfh, nBf * 2) || cudaMalloc(&s.Bfl, nBf * 2) || cudaMalloc(&s.Bdh, nBd * 2) || cudaMalloc(&s.Bdl, nBd * 2) ||
      cudaMalloc(&s.P, nP * 4)) {
    snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of spectral buffers failed (%s)", cudaGetErrorString(cudaGetLastError())); return -1;
  }
  cudaMemcpyAsync(s.tw, tw.data(), L * sizeof(float2), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(s.twd, twd.data(), L * sizeof(double2), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);               // the host vectors go out of scope
  cudaMemsetAsync(s.Ah, 0, nA * 2, st); cudaMemsetAsync(s.Al, 0, nA * 2, st); cudaMemsetAsync(s.Dh, 0, nD * 2, st); cudaMemsetAsync(s.Dl, 0, nD * 2, st);
  cudaMemsetAsync(s.Bfh, 0, nBf * 2, st); cudaMemsetAsync(s.Bfl, 0, nBf * 2, st); cudaMemsetAsync(s.Bdh, 0, nBd * 2, st); cudaMemsetAsync(s.Bdl, 0, nBd * 2, st);
  cudaMemsetAsync(s.P, 0, nP * 4, st);
  if (spec_enc3(&s.mBfh, s.Bfh, 128, SP_NF, F, 128 * 2, (long long)SP_NF * 128 * 2, 64, SP_NF) ||
      spec_enc3(&s.mBfl, s.Bfl, 128, SP_NF, F, 128 * 2, (long long)SP_NF * 128 * 2, 64, SP_NF) ||
      spec_enc3(&s.mBdh, s.Bdh, 64, SP_ND, F, 64 * 2, (long long)SP_ND * 64 * 2, 64, SP_ND) ||
      spec_enc3(&s.mBdl, s.Bdl, 64, SP_ND, F, 64 * 2, (long long)SP_ND * 64 * 2, 64, SP_ND)) return -1;
  cudaFuncSetAttribute(spec_fft_fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(spec_fft_inv_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(spec_gemm_k<SP_NF, 7, 2, 40, SPF_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec_gemm_smem(SP_NF, 2, 40, SPF_STAGES));
  cudaFuncSetAttribute(spec_gemm_k<SP_ND, 3, 1, 100, SPD_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec_gemm_smem(SP_ND, 1, 100, SPD_STAGES));
  cudaFuncSetAttribute(spec_wgrad_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_WGRAD_SMEM);
  s.fft_threads = getenv("SNDVAE_FFT_THREADS") ? atoi(getenv("SNDVAE_FFT_THREADS")) : 400;
  s.fft_threads_generic = L == 384 ? 400 : (L >= 768 ? 512 : 256);      // measured at N = 1024 (L = 1536): 512 threads 186 graphs/s, 256 threads 167
  if (getenv("SNDVAE_FFT_THREADS_GENERIC")) { const int t = atoi(getenv("SNDVAE_FFT_THREADS_GENERIC")); if (t >= 64 && t <= SP_FFT_THREADS_MAX) s.fft_threads_generic = t; }
  if (getenv("SNDVAE_FFT_G")) { const int g = atoi(getenv("SNDVAE_FFT_G")); if (g >= 1 && g <= s.G1) s.G1 = g; if (g >= 1 && g <= s.G2) s.G2 = g; }
  s.generic_only = getenv("SNDVAE_FFT_GENERIC") ? 1 : 0;     // force the runtime-plan kernels (any N)
  s.fft_bulk = getenv("SNDVAE_FFT_BULK") ? atoi(getenv("SNDVAE_FFT_BULK")) : 1;      // measured: forward Y 5.67 -> 5.54 ms per 256 graphs
  s.fft_bulk_inv = getenv("SNDVAE_FFT_BULK_INV") ? atoi(getenv("SNDVAE_FFT_BULK_INV")) : 0;    // not measured yet (round 2)
  // bit 0: forward, bit 1: inverse two-pass transforms (fft2p.cuh) where L = 384 / 192; measured per 256 graphs at N = 256 against the
  // three-pass kernels: forward Y 5.20 -> 4.72 ms, inverse O 1.74 -> 1.34, forward dO 2.20 -> 2.25, inverse dY 4.53 -> 3.58
  s.fft_2pass = getenv("SNDVAE_FFT_2PASS") ? atoi(getenv("SNDVAE_FFT_2PASS")) : 3;
  s.fft_inv2 = getenv("SNDVAE_FFT_INV2") ? atoi(getenv("SNDVAE_FFT_INV2")) : 1;
  s.wgrad_stages = getenv("SNDVAE_WGRAD_STAGES") ? atoi(getenv("SNDVAE_WGRAD_STAGES")) : SP_WSTAGES;
  if (s.wgrad_stages < 2 || s.wgrad_stages > SP_WSTAGES_MAX) s.wgrad_stages = SP_WSTAGES;
  s.fft_order = getenv("SNDVAE_FFT_ORDER") ? atoi(getenv("SNDVAE_FFT_ORDER")) : -1;              // LineWalk order of the forward / inverse
  s.fft_order_inv = getenv("SNDVAE_FFT_ORDER_INV") ? atoi(getenv("SNDVAE_FFT_ORDER_INV")) : -1;  // fast transforms (-1: default)
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  s.grid_sms = sms;
  s.ready = 1;
  return 0;
}
// per step: weight spectra -> GEMM operand planes
static int spec_stage_weights(SpecState& s, const float* w1, cudaStream_t st) {
  const long long n = (long long)s.pl.F * s.C1 * s.C2;
  spec_stage_weights_k<<<cdiv(n, 128), 128, 0, st>>>(w1, s.twd, s.Bfh, s.Bfl, s.Bdh, s.Bdl, s.N, s.pl.L, s.pl.F, s.C1, s.C2);
  return tc_check_launch("spec_stage_weights_k");
}
// fast-path launchers: one instantiation per (plan, channel pairs); 0 = launched, 1 = no matching instantiation
template <int R0, int R1, int R2, int G, int NT, int MINB>
static int spec_launch_fwd_fast(SpecState& s, const FftFwdArgs& a, cudaStream_t st) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  const size_t smem = Cfg::smem_fwd(a.N);
  if (smem > 225 * 1024) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_fwd_fast_k<R0, R1, R2, G, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024); attr = true; }
  const long long want = (long long)s.grid_sms * MINB, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_fwd_fast_k<R0, R1, R2, G, NT, MINB><<<(unsigned)(units < want ? units : want), NT, smem, st>>>(a);
  return 0;
}
template <int R0, int R1, int R2, int G, int NT, int MINB, bool ALIAS>
static int spec_launch_inv_fast(SpecState& s, const FftInvArgs& a, cudaStream_t st) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  const size_t smem = Cfg::smem_inv(ALIAS);
  if (smem > 225 * 1024) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_inv_fast_k<R0, R1, R2, G, NT, MINB, ALIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024); attr = true; }
  const long long want = (long long)s.grid_sms * MINB, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_inv_fast_k<R0, R1, R2, G, NT, MINB, ALIAS><<<(unsigned)(units < want ? units : want), NT, smem, st>>>(a);
  return 0;
}
template <int R0, int R1, int R2, int G, int NT, int MINB>
static int spec_launch_inv_fast2(SpecState& s, const FftInvArgs& a, cudaStream_t st) {
  using Cfg = FastCfg<R0, R1, R2, G>;
  const size_t smem = (size_t)Cfg::L * 8 + 2 * Cfg::BUF + (size_t)((Cfg::F + 1) / 2) * 4 * G * 4;
  if (smem > 225 * 1024) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_inv_fast2_k<R0, R1, R2, G, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024); attr = true; }
  const long long want = (long long)s.grid_sms * MINB, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_inv_fast2_k<R0, R1, R2, G, NT, MINB><<<(unsigned)(units < want ? units : want), NT, smem, st>>>(a);
  return 0;
}
template <int M, int G, int MINB, bool TABLE>
static int spec_launch_fwd2(SpecState& s, const FftFwdArgs& a, cudaStream_t st) {
  using Cfg = Fft2Cfg<M, G>;
  const size_t smem = Cfg::smem_fwd(a.N);
  if (smem > 226 * 1024 || a.N > 32 * M || ((long long)a.N * 2 * G) % 4 != 0) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_fwd2_k<M, G, MINB, TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); attr = true; }
  const long long want = (long long)s.grid_sms * MINB, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_fwd2_k<M, G, MINB, TABLE><<<(unsigned)(units < want ? units : want), Cfg::NT, smem, st>>>(a);
  return 0;
}
template <int M, int G, int MINB, bool TABLE>
static int spec_launch_inv2(SpecState& s, const FftInvArgs& a, cudaStream_t st) {
  using Cfg = Fft2Cfg<M, G>;
  const size_t smem = Cfg::smem_inv();
  if (smem > 226 * 1024 || a.N > 32 * M) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_inv2_k<M, G, MINB, TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); attr = true; }
  const long long want = (long long)s.grid_sms * MINB, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_inv2_k<M, G, MINB, TABLE><<<(unsigned)(units < want ? units : want), Cfg::NT, smem, st>>>(a);
  return 0;
}
template <int G>
static int spec_launch_fwd3(SpecState& s, const FftFwdArgs& a, cudaStream_t st) {
  using Cfg = Fft3Cfg<G>;
  const size_t smem = Cfg::smem_fwd(a.N);
  if (smem > 226 * 1024 || a.N > 1024 || a.C % (2 * G) != 0 || a.C > 64) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_fwd3_k<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); attr = true; }
  const long long want = s.grid_sms, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_fwd3_k<G><<<(unsigned)(units < want ? units : want), Cfg::NT, smem, st>>>(a);
  return 0;
}
template <int G>
static int spec_launch_inv3(SpecState& s, const FftInvArgs& a, cudaStream_t st) {
  using Cfg = Fft3Cfg<G>;
  const size_t smem = Cfg::smem_inv();
  if (smem > 226 * 1024 || a.N > 1024 || a.C % (2 * G) != 0) return 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(spec_fft_inv3_k<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); attr = true; }
  const long long want = s.grid_sms, units = a.order ? (a.lines + 1) / 2 : a.lines;
  spec_fft_inv3_k<G><<<(unsigned)(units < want ? units : want), Cfg::NT, smem, st>>>(a);
  return 0;
}
static bool spec_plan_is(const FftPlan& pl, int r0, int r1, int r2) {
  return pl.npass == (r2 ? 3 : 2) && pl.rad[0] == r0 && pl.rad[1] == r1 && (!r2 || pl.rad[2] == r2);
}
static int spec_fft_fwd(SpecState& s, const float* in, const float* in1, long long lines0, long long lines, int dir1_strided, const float* gam, const float* bet,
                        const float* bias0, const float* bias1, const float* b0,
                        __nv_bfloat16* oh, __nv_bfloat16* ol, int KA, int C, int G, cudaStream_t st) {
  FftFwdArgs a; a.in = in; a.in1 = in1; a.bias0 = bias0; a.bias1 = bias1; a.b0 = b0; a.lines0 = lines0; a.lines = lines; a.dir1_strided = dir1_strided; a.gam = gam; a.bet = bet; a.oh = oh; a.ol = ol;
  a.RA = s.RA; a.KA = KA; a.N = s.N; a.C = C; a.G = G; a.tw = s.tw; a.pl = s.pl;
  // a tensor read by both directions (dir1_strided == 1) is walked graph by graph, otherwise neighbouring lines in pairs (LineWalk)
  a.order = s.fft_order >= 0 ? s.fft_order : (dir1_strided == 1 ? 2 : 1);
  if (a.order == 2 && (lines != 2 * lines0 || lines0 % s.N != 0)) a.order = 1;
  a.bulk = s.fft_bulk;
  int r = 1;
  if (!s.generic_only && (s.fft_2pass & 1)) {      // twiddles from the table (measured: forward Y 5.32 -> 4.95 ms against the product chain)
    if (s.pl.L == 384 && C == 50) r = spec_launch_fwd2<8, 25, 1, true>(s, a, st);
    else if (s.pl.L == 384 && C == 20) r = spec_launch_fwd2<8, 10, 2, true>(s, a, st);      // (three CTAs per SM at 124 registers: 2.60 ms against 2.25)
    else if (s.pl.L == 192 && C == 50) r = spec_launch_fwd2<4, 25, 2, true>(s, a, st);
    else if (s.pl.L == 192 && C == 20) r = spec_launch_fwd2<4, 10, 4, true>(s, a, st);
    else if (s.pl.L == 1536) r = spec_launch_fwd3<5>(s, a, st);                              // N = 769..1024
    if (r == 0) return tc_check_launch("spec_fft_fwd2_k");
  }
  if (!s.generic_only) {
    if (spec_plan_is(s.pl, 6, 8, 8) && C == 50) r = s.fft_threads == 800 ? spec_launch_fwd_fast<6, 8, 8, 25, 800, 1>(s, a, st) : spec_launch_fwd_fast<6, 8, 8, 25, 400, 1>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 8) && C == 20) r = spec_launch_fwd_fast<6, 8, 8, 10, 320, 2>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 4) && C == 50) r = spec_launch_fwd_fast<6, 8, 4, 25, 400, 2>(s, a, st);      // N = 65..128 (L = 192)
    else if (spec_plan_is(s.pl, 6, 8, 4) && C == 20) r = spec_launch_fwd_fast<6, 8, 4, 10, 160, 4>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 0) && C == 50) r = spec_launch_fwd_fast<6, 8, 0, 25, 200, 2>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 0) && C == 20) r = spec_launch_fwd_fast<6, 8, 0, 10, 160, 2>(s, a, st);
  }
  if (r == 0) return tc_check_launch("spec_fft_fwd_fast_k");
  const unsigned grid = (unsigned)(lines < s.grid_sms ? lines : s.grid_sms);
  spec_fft_fwd_k<<<grid, s.fft_threads_generic, spec_fft_smem(s.pl.L, G), st>>>(a);
  return tc_check_launch("spec_fft_fwd_k");
}
static int spec_fft_inv(SpecState& s, const float* in, float* out, float* out1, long long lines0, long long lines, int C, int G, cudaStream_t st) {
  FftInvArgs a; a.in = in; a.RA = s.RA; a.out = out; a.out1 = out1; a.lines0 = lines0; a.lines = lines; a.N = s.N; a.C = C; a.G = G; a.tw = s.tw; a.pl = s.pl;
  a.bulk = s.fft_bulk_inv;
  a.order = s.fft_order_inv >= 0 ? s.fft_order_inv : 0;      // measured: pairs do not help the inverse (its rows are read, not written)
  if (a.order == 2 && (lines != 2 * lines0 || lines0 % s.N != 0)) a.order = 1;
  int r = 1;
  if (!s.generic_only && (s.fft_2pass & 2)) {      // twiddles by product chain (measured: inverse dY 3.70 ms against 3.77 with table reads)
    if (s.pl.L == 384 && C == 50) r = spec_launch_inv2<8, 25, 1, false>(s, a, st);
    else if (s.pl.L == 384 && C == 20) r = spec_launch_inv2<8, 10, 2, false>(s, a, st);
    else if (s.pl.L == 192 && C == 50) r = spec_launch_inv2<4, 25, 2, false>(s, a, st);
    else if (s.pl.L == 192 && C == 20) r = spec_launch_inv2<4, 10, 4, false>(s, a, st);
    else if (s.pl.L == 1536) r = spec_launch_inv3<5>(s, a, st);
    if (r == 0) return tc_check_launch("spec_fft_inv2_k");
  }
  if (!s.generic_only && s.fft_inv2 && C == 50) {      // 3-pass plans at G = 25: split staging, extension fused into pass 1
    // (measured at N = 256: 4.88 -> 4.69 ms per 256 graphs; the G = 10 transforms, which have room for a separate staging
    // buffer anyway, lose 3 % with it and stay on the first form)
    if (spec_plan_is(s.pl, 6, 8, 8)) r = spec_launch_inv_fast2<6, 8, 8, 25, 400, 1>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 4)) r = spec_launch_inv_fast2<6, 8, 4, 25, 400, 2>(s, a, st);
    if (r == 0) return tc_check_launch("spec_fft_inv_fast2_k");
  }
  if (!s.generic_only) {
    if (spec_plan_is(s.pl, 6, 8, 8) && C == 50) r = s.fft_threads == 800 ? spec_launch_inv_fast<6, 8, 8, 25, 800, 1, true>(s, a, st) : spec_launch_inv_fast<6, 8, 8, 25, 400, 1, true>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 8) && C == 20) r = spec_launch_inv_fast<6, 8, 8, 10, 320, 2, false>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 4) && C == 50) r = spec_launch_inv_fast<6, 8, 4, 25, 400, 2, true>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 4) && C == 20) r = spec_launch_inv_fast<6, 8, 4, 10, 160, 4, false>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 0) && C == 50) r = spec_launch_inv_fast<6, 8, 0, 25, 200, 2, false>(s, a, st);
    else if (spec_plan_is(s.pl, 6, 8, 0) && C == 20) r = spec_launch_inv_fast<6, 8, 0, 10, 160, 2, false>(s, a, st);
  }
  if (r == 0) return tc_check_launch("spec_fft_inv_fast_k");
  const unsigned grid = (unsigned)(lines < s.grid_sms ? lines : s.grid_sms);
  spec_fft_inv_k<<<grid, s.fft_threads_generic, spec_fft_smem(s.pl.L, G), st>>>(a);
  return tc_check_launch("spec_fft_inv_k");
}
// forward: O12[2 rows][N C2] = e2e-layer-1 row / column products of Y = relu(BN_e1(E1));  rows = bc * N
// E1 / E1T: graph-tiled layer-0 outputs of y_producer_tc_k (row and column lines both contiguous)
// Sa / Rc / b0: the line-constant bias rows left out of E1 / E1T
static int spec_forward(SpecState& s, const float* E1, const float* E1T, const float* gam1, const float* bet1, const float* Sa, const float* Rc,
                        const float* b0, long long rows, float* O12, cudaStream_t st) {
  const long long lines = 2 * rows; const int F = s.pl.F;
  if (spec_fft_fwd(s, E1, E1T, rows, lines, 2, gam1, bet1, Sa, Rc, b0, s.Ah, s.Al, SP_KA1, s.C1, s.G1, st)) return -1;
  CUtensorMap ah, al;
  if (spec_enc3(&ah, s.Ah, 2 * s.C1, lines, F, SP_KA1 * 2, s.RA * SP_KA1 * 2, 64, 128) ||
      spec_enc3(&al, s.Al, 2 * s.C1, lines, F, SP_KA1 * 2, s.RA * SP_KA1 * 2, 64, 128)) return -1;
  SpecGemmArgs g; g.out = s.Oc; g.RA = s.RA; g.rows = lines; g.F = F; g.MT = (int)((lines + 127) / 128); g.items = (long long)F * g.MT;
  const unsigned grid = (unsigned)(g.items < s.grid_sms ? g.items : s.grid_sms);
  spec_gemm_k<SP_NF, 7, 2, 40, SPF_STAGES><<<grid, 192, spec_gemm_smem(SP_NF, 2, 40, SPF_STAGES), st>>>(ah, al, s.mBfh, s.mBfl, g);
  if (tc_check_launch("spec_gemm_k(fwd)")) return -1;
  // both directions land in the [b, i, j, C2] layout: O12 plane 0 = row products, plane 1 = column products
  return spec_fft_inv(s, s.Oc, O12, O12 + rows * s.N * s.C2, rows, lines, s.C2, s.G2, st);
}
// backward: dY12[2][rows][N C1] (both planes in the [b, i, j, C1] layout) from dO [rows][N C2] (fp32, one layout: the
// column lines are read strided), and P += Y^^T dO^
static int spec_backward(SpecState& s, const float* dO, long long rows, float* dY12, cudaStream_t st) {
  const long long lines = 2 * rows; const int F = s.pl.F;
  if (spec_fft_fwd(s, dO, nullptr, rows, lines, 1, nullptr, nullptr, nullptr, nullptr, nullptr, s.Dh, s.Dl, SP_KA2, s.C2, s.G2, st)) return -1;
  CUtensorMap dh, dl;
  if (spec_enc3(&dh, s.Dh, 2 * s.C2, lines, F, SP_KA2 * 2, s.RA * SP_KA2 * 2, 64, 128) ||
      spec_enc3(&dl, s.Dl, 2 * s.C2, lines, F, SP_KA2 * 2, s.RA * SP_KA2 * 2, 64, 128)) return -1;
  SpecGemmArgs g; g.out = s.dYc; g.RA = s.RA; g.rows = lines; g.F = F; g.MT = (int)((lines + 127) / 128); g.items = (long long)F * g.MT;
  const unsigned grid = (unsigned)(g.items < s.grid_sms ? g.items : s.grid_sms);
  spec_gemm_k<SP_ND, 3, 1, 100, SPD_STAGES><<<grid, 192, spec_gemm_smem(SP_ND, 1, 100, SPD_STAGES), st>>>(dh, dl, s.mBdh, s.mBdl, g);
  if (tc_check_launch("spec_gemm_k(dgrad)")) return -1;
  if (spec_fft_inv(s, s.dYc, dY12, dY12 + rows * s.N * s.C1, rows, lines, s.C1, s.G1, st)) return -1;
  // wgrad: MN-major views of the same planes, SP_WKC lines per K chunk
  CUtensorMap ah, al, bh, bl;
  if (spec_enc3(&ah, s.Ah, 2 * s.C1, lines, F, SP_KA1 * 2, s.RA * SP_KA1 * 2, 64, SP_WKC) ||
      spec_enc3(&al, s.Al, 2 * s.C1, lines, F, SP_KA1 * 2, s.RA * SP_KA1 * 2, 64, SP_WKC) ||
      spec_enc3(&bh, s.Dh, 2 * s.C2, lines, F, SP_KA2 * 2, s.RA * SP_KA2 * 2, 64, SP_WKC) ||
      spec_enc3(&bl, s.Dl, 2 * s.C2, lines, F, SP_KA2 * 2, s.RA * SP_KA2 * 2, 64, SP_WKC)) return -1;
  SpecWgradArgs w; w.P = s.P; w.rows = lines; w.F = F; w.MV = 2 * s.C1; w.NV = 2 * s.C2; w.stages = s.wgrad_stages;
  const long long kchunks = (lines + SP_WKC - 1) / SP_WKC;
  // K splits per frequency: the kernel hands each CTA ceil(F KS / grid) consecutive (frequency, split) units, so KS is chosen
  // to level the CTAs' chunk counts (F = 193 over 148 CTAs: KS = 7 left 12 SMs idle and the rest at 10 units, 91 %)
  int ks = 1; long long best = -1;
  for (int cand = 1; cand <= 32 && cand <= kchunks; ++cand) {
    const long long units = (long long)F * cand, grid = units < s.grid_sms ? units : s.grid_sms;
    const long long upc = (units + grid - 1) / grid, cper = (kchunks + cand - 1) / cand;
    long long worst = 0;
    for (long long c = 0; c < grid; ++c) {
      long long load = 0;
      for (long long u = c * upc; u < (c + 1) * upc && u < units; ++u) {
        const long long lo = (u % cand) * cper; long long hi = lo + cper; if (hi > kchunks) hi = kchunks;
        if (hi > lo) load += hi - lo;
      }
      if (load > worst) worst = load;
    }
    worst += 2 * upc;        // a unit's pipeline fill / drain and its share of the flush atomics, in chunk units
    if (best < 0 || worst < best) { best = worst; ks = cand; }
  }
  if (getenv("SNDVAE_WGRAD_KS")) ks = atoi(getenv("SNDVAE_WGRAD_KS"));
  if (ks > kchunks) ks = (int)kchunks; if (ks < 1) ks = 1;
  w.KS = ks;
  const long long units = (long long)F * ks;
  const unsigned wgrid = (unsigned)(units < s.grid_sms ? units : s.grid_sms);
  spec_wgrad_k<<<wgrid, TC_THREADS, SP_WGRAD_SMEM, st>>>(ah, al, bh, bl, w);
  return tc_check_launch("spec_wgrad_k");
}
static int spec_zero_wgrad(SpecState& s, cudaStream_t st) {
  return cudaMemsetAsync(s.P, 0, (size_t)s.pl.F * 128 * SP_NF * 4, st) == cudaSuccess ? 0 : -1;
}
static int spec_finalize_wgrad(SpecState& s, float* dw1, cudaStream_t st) {
  const long long n = (long long)s.N * s.C1 * s.C2;
  spec_wgrad_finalize_k<<<cdiv(n, 128), 128, 0, st>>>(s.P, s.twd, dw1, s.N, s.pl.L, s.pl.F, s.C1, s.C2);
  return tc_check_launch("spec_wgrad_finalize_k");
}

// ------------------------------------------------------------------------------------------------------------------------
// e2e layer 0 (collapsed) on the tensor cores, batch-major:
//     E1[b, i, j, :] = [a[b,i,:] | c[b,j,:]] . [WSa[j]; WSc[i]] + Rc[b,j,:] + Sa[b,i,:] + 2 b0
// (the kernel stores E1 - Sa[b,i,:] - 2 b0: that part is constant along the line the CTA sweeps and is added by the readers)
// For a fixed (i, j) this is a GEMM over the graphs of the micro-batch: M = 128 graphs (TMEM lanes), K = 2 * 2H (3 + 3
// k-steps of 16, zero padded by TMA), N = 64 (50 valid).  A CTA owns (graph tile, i): a[.,i,:] and WSc[i] stay in shared
// memory while it sweeps j with a TMA ring of (c[.,j,:], WSa[j]) stages; positions are processed in pairs so that each
// thread (= graph) writes 400 contiguous bytes of E1.  3-pass split-bf16, fp32 accumulation in TMEM.
// ------------------------------------------------------------------------------------------------------------------------
struct YtcArgs {
  float* E1;             // tiled [nbt][N (fixed pos)][128 graphs][N (swept pos)][C1]; WITHOUT the fixed position's bias row and 2 b0
                         // (constant along a line: the consumers -- forward FFT, l0_combine_planes_k -- add them)
  const float* Rc;       // [Bc*N, C1] bias rows of the swept position
  const float* Sa;       // unused (bias rows of the fixed position are added by the consumers)
  const float* b0;       // unused
  int Bc, N, C1, nbt;    // graphs in the micro-batch, nodes, channels (<= 56), graph tiles of 128
  int JS;                // work items per (graph tile, fixed position): the swept positions are cut into JS ranges (tail balance)
};
#define YTC_STAGES 2
__global__ void __launch_bounds__(192, 1) y_producer_tc_k(const __grid_constant__ CUtensorMap tmAah, const __grid_constant__ CUtensorMap tmAal,
                                                          const __grid_constant__ CUtensorMap tmAch, const __grid_constant__ CUtensorMap tmAcl,
                                                          const __grid_constant__ CUtensorMap tmWah, const __grid_constant__ CUtensorMap tmWal,
                                                          const __grid_constant__ CUtensorMap tmWch, const __grid_constant__ CUtensorMap tmWcl,
                                                          YtcArgs P) {
  constexpr int A_BYTES = 128 * 64 * 2;     // one plane of an activation tile: 128 graphs x 64 (K, 40 valid)
  constexpr int W_BYTES = 64 * 64 * 2;      // one plane of a weight tile: 64 (o) x 64 (K)
  constexpr int FIX_BYTES = 2 * A_BYTES + 2 * W_BYTES;       // a[., i] hi/lo + WSc[i] hi/lo
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;     // c[., j] hi/lo + WSa[j] hi/lo
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* fix = smem; uint8_t* ring = smem + FIX_BYTES;
  float* xbuf = reinterpret_cast<float*>(ring + (size_t)YTC_STAGES * STAGE_BYTES);     // 4 epilogue warps x 12.8 KB
  __shared__ uint64_t fix_full, fix_empty, full_bar[YTC_STAGES], empty_bar[YTC_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = P.N;
  const int JS = P.JS, NJ = N / JS;          // NJ swept positions per work item (even when JS > 1)
  const long long nwork = (long long)P.nbt * N * JS;

  if (threadIdx.x == 0) {
    mbar_init(&fix_full, 1); mbar_init(&fix_empty, 1);
    for (int s = 0; s < YTC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, nw = 0;
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++nw) {
        const int bt = (int)(w / (N * JS)), rem = (int)(w - (long long)bt * N * JS), i = rem / JS, jlo = (rem - i * JS) * NJ;
        mbar_wait(&fix_empty, (nw & 1) ^ 1);
        mbar_expect_tx(&fix_full, FIX_BYTES);
        tma_load_3d(fix, &tmAah, &fix_full, 0, i, bt * 128);
        tma_load_3d(fix + A_BYTES, &tmAal, &fix_full, 0, i, bt * 128);
        tma_load_3d(fix + 2 * A_BYTES, &tmWch, &fix_full, 0, 0, i);
        tma_load_3d(fix + 2 * A_BYTES + W_BYTES, &tmWcl, &fix_full, 0, 0, i);
        for (int j = jlo; j < jlo + NJ; ++j, ++it) {
          const int s = it % YTC_STAGES; const uint32_t ph = (it / YTC_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = ring + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_3d(st, &tmAch, &full_bar[s], 0, j, bt * 128);
          tma_load_3d(st + A_BYTES, &tmAcl, &full_bar[s], 0, j, bt * 128);
          tma_load_3d(st + 2 * A_BYTES, &tmWah, &full_bar[s], 0, 0, j);
          tma_load_3d(st + 2 * A_BYTES + W_BYTES, &tmWal, &full_bar[s], 0, 0, j);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, 64, 0, 0);
      int it = 0, nw = 0, pc = 0;
      const uint32_t sf = smem_u32(fix);
      for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++nw) {
        mbar_wait(&fix_full, nw & 1);
        tc_fence_after();
        for (int j = 0; j < NJ; ++j, ++it) {
          const int slot = pc & 1, half = j & 1;
          if (half == 0) { mbar_wait(&acc_empty[slot], ((pc >> 1) & 1) ^ 1); tc_fence_after(); }
          const int s = it % YTC_STAGES; const uint32_t ph = (it / YTC_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t ss = smem_u32(ring + (size_t)s * STAGE_BYTES);
          // two independent accumulate chains per position (a-part / c-part), summed by the epilogue
          const uint32_t td = tmem_d + slot * 256 + half * 64, td2 = td + 128;
#pragma unroll
          for (int k = 0; k < 3; ++k) {      // a[., i] . WSa[j]
            const uint64_t ah = umma_desc(sf + k * 32, 16, 1024, 2ull), al = umma_desc(sf + A_BYTES + k * 32, 16, 1024, 2ull);
            const uint64_t bh = umma_desc(ss + 2 * A_BYTES + k * 32, 16, 1024, 2ull), bl = umma_desc(ss + 2 * A_BYTES + W_BYTES + k * 32, 16, 1024, 2ull);
            umma_bf16(td, ah, bh, idesc, k ? 1u : 0u); umma_bf16(td, ah, bl, idesc, 1u); umma_bf16(td, al, bh, idesc, 1u);
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {      // c[., j] . WSc[i]
            const uint64_t ah = umma_desc(ss + k * 32, 16, 1024, 2ull), al = umma_desc(ss + A_BYTES + k * 32, 16, 1024, 2ull);
            const uint64_t bh = umma_desc(sf + 2 * A_BYTES + k * 32, 16, 1024, 2ull), bl = umma_desc(sf + 2 * A_BYTES + W_BYTES + k * 32, 16, 1024, 2ull);
            umma_bf16(td2, ah, bh, idesc, k ? 1u : 0u); umma_bf16(td2, ah, bl, idesc, 1u); umma_bf16(td2, al, bh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (half == 1 || j == NJ - 1) { umma_commit(&acc_full[slot]); ++pc; }
        }
        umma_commit(&fix_empty);
      }
    }
  } else {
    // epilogue: thread = graph (TMEM lane).  Output layout is tiled so that the 128 graphs of a tile are 51 KB apart, not a
    // whole [N,N,C1] plane apart:  out[((bt N + x) 128 + bl) N + y][C1]  (x = fixed position of the CTA, y = swept position)
    constexpr int C1 = TC_C1;                   // 50: a pair of positions is 100 floats = 25 x 16 bytes
    const int q = warp & 3;
    int pc = 0;
    const bool v4 = (N % 2) == 0;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int bt = (int)(w / (N * JS)), rem = (int)(w - (long long)bt * N * JS), x = rem / JS, jlo = (rem - x * JS) * NJ;
      const int bl = q * 32 + lane, b = bt * 128 + bl;
      const bool valid = b < P.Bc;
      const int bb = valid ? b : 0;
      float* erow = P.E1 + (((long long)bt * N + x) * 128 + bl) * N * C1;
      const float* rrow = P.Rc + (long long)bb * N * C1;
      if (v4) {
        // Even N: a pair of positions is one 16-byte aligned 400-byte run per graph.  The warp moves its 32 x 400 B block
        // cooperatively (consecutive lanes = consecutive 16-byte chunks, so every global instruction touches a few lines
        // instead of 32) through a private shared-memory tile; each thread then adds its TMEM row in place.
        float4* xw = reinterpret_cast<float4*>(xbuf) + (size_t)(warp - 2) * 800;      // [32 rows][25 x 16 B]
        float* etile = P.E1 + (((long long)bt * N + x) * 128 + q * 32) * N * C1;
        const float* rtile = P.Rc + (long long)(bt * 128 + q * 32) * N * C1;          // same row pitch as the output tile
        int odst[25]; uint32_t vmask = 0;
#pragma unroll
        for (int t = 0; t < 25; ++t) {
          const int k = lane + 32 * t, row = k / 25, col = k - row * 25;
          odst[t] = row * N * C1 + 4 * col;
          if (bt * 128 + q * 32 + row < P.Bc) vmask |= 1u << t;
        }
        float4 pf[25];
#pragma unroll
        for (int t = 0; t < 25; ++t) pf[t] = (vmask >> t & 1) ? __ldg(reinterpret_cast<const float4*>(rtile + odst[t] + jlo * C1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = jlo; j0 < jlo + NJ; j0 += 2, ++pc) {
          const int slot = pc & 1;
          mbar_wait(&acc_full[slot], (pc >> 1) & 1);
          tc_fence_after();
          const uint32_t ta = tmem_d + slot * 256 + ((uint32_t)(q * 32) << 16);
          float2* xr = reinterpret_cast<float2*>(xw + lane * 25);
          // TMEM columns [0, 50) and [64, 114) -> floats [0, 50) and [50, 100) of this graph's row of the tile
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {      // 16 columns of both accumulators per tcgen05.wait::ld (8 for the tail)
              uint32_t v[16], v2[16];
              if (c0 < 48) { tmem_ld16_nowait(ta + jj * 64 + c0, v); tmem_ld16_nowait(ta + 128 + jj * 64 + c0, v2); }
              else { tmem_ld8_nowait(ta + jj * 64 + c0, v); tmem_ld8_nowait(ta + 128 + jj * 64 + c0, v2); }
              tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (c0 + 2 * u < C1) xr[(jj * C1 + c0) / 2 + u] = make_float2(__uint_as_float(v[2 * u]) + __uint_as_float(v2[2 * u]),
                                                                             __uint_as_float(v[2 * u + 1]) + __uint_as_float(v2[2 * u + 1]));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);
          // cooperative phase: tile chunk + prefetched bias chunk -> global; then prefetch the next pair's bias rows
#pragma unroll
          for (int t = 0; t < 25; ++t) {
            float4 v = xw[lane + 32 * t];
            v.x += pf[t].x; v.y += pf[t].y; v.z += pf[t].z; v.w += pf[t].w;
            if (vmask >> t & 1) *reinterpret_cast<float4*>(etile + odst[t] + j0 * C1) = v;
          }
          if (j0 + 2 < jlo + NJ) {
#pragma unroll
            for (int t = 0; t < 25; ++t) if (vmask >> t & 1) pf[t] = __ldg(reinterpret_cast<const float4*>(rtile + odst[t] + (j0 + 2) * C1));
          }
          __syncwarp();
        }
        continue;
      }
      for (int j0 = 0; j0 < N; j0 += 2, ++pc) {
        const int slot = pc & 1;
        const int nj = N - j0 < 2 ? N - j0 : 2;
        float r[2 * C1];
        {
          const float2* rc = reinterpret_cast<const float2*>(rrow + (long long)j0 * C1);
#pragma unroll
          for (int u = 0; u < C1; ++u) { float2 t = make_float2(0.f, 0.f); if (u < nj * (C1 / 2)) t = __ldg(rc + u); r[2 * u] = t.x; r[2 * u + 1] = t.y; }
        }
        mbar_wait(&acc_full[slot], (pc >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem_d + slot * 256 + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (jj < nj) {
#pragma unroll
            for (int c0 = 0; c0 < 56; c0 += 8) {
              uint32_t v[8], v2[8];
              tmem_ld8(ta + jj * 64 + c0, v);
              tmem_ld8(ta + 128 + jj * 64 + c0, v2);
#pragma unroll
              for (int u = 0; u < 8; ++u) if (c0 + u < C1) r[jj * C1 + c0 + u] += __uint_as_float(v[u]) + __uint_as_float(v2[u]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        if (valid) {
          float2* dst = reinterpret_cast<float2*>(erow + (long long)j0 * C1);
#pragma unroll
          for (int u = 0; u < C1; ++u) if (u < nj * (C1 / 2)) dst[u] = make_float2(r[2 * u], r[2 * u + 1]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

// WS fp32 [N][C1][Ch] -> bf16 hi / lo planes of the same shape (row stride CS = Ch padded to 16 bytes)
__global__ void ytc_stage_ws_k(const float* __restrict__ WS, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long rows, int Ch, int CS) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Ch) return;
  const long long r = idx / Ch; const int ch = (int)(idx - r * Ch);
  const float v = WS[idx];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[r * CS + ch] = h; lo[r * CS + ch] = __float2bfloat16_rn(v - __bfloat162float(h));
}
struct YtcState {
  int ready, N, Ch, CS, C1;
  __nv_bfloat16 *Wah, *Wal, *Wch, *Wcl;      // [N][C1][CS]
  CUtensorMap mWah, mWal, mWch, mWcl;
  int grid_sms;
};
static const size_t YTC_SMEM = (size_t)(1 + YTC_STAGES) * (2 * 128 * 64 * 2 + 2 * 64 * 64 * 2) + 4 * 12800 + 1024;
static void ytc_destroy(YtcState& y) {
  void* ps[] = {y.Wah, y.Wal, y.Wch, y.Wcl};
  for (void* p : ps) if (p) cudaFree(p);
  memset(&y, 0, sizeof y);
}
static int ytc_enc3(CUtensorMap* tm, const void* base, long long d0, long long d1, long long d2, long long s1_b, long long s2_b, int b0, int b1, int b2) {
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t str[2] = {(cuuint64_t)s1_b, (cuuint64_t)s2_b};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  return tc_encode(tm, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
static int ytc_init(YtcState& y, int N, int Ch, int C1, cudaStream_t st) {
  memset(&y, 0, sizeof y);
  if (tc_global_init()) return -1;
  if (C1 != TC_C1 || Ch > 48) { snprintf(g_tc_err, sizeof g_tc_err, "y_producer_tc: C1 = %d and 2H <= 48 required", TC_C1); return -1; }
  y.N = N; y.Ch = Ch; y.CS = tc_pad16(Ch); y.C1 = C1;
  const size_t n = (size_t)N * C1 * y.CS + 256;
  if (cudaMalloc(&y.Wah, n * 2) || cudaMalloc(&y.Wal, n * 2) || cudaMalloc(&y.Wch, n * 2) || cudaMalloc(&y.Wcl, n * 2)) {
    snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of y_producer_tc weights failed"); return -1;
  }
  cudaMemsetAsync(y.Wah, 0, n * 2, st); cudaMemsetAsync(y.Wal, 0, n * 2, st); cudaMemsetAsync(y.Wch, 0, n * 2, st); cudaMemsetAsync(y.Wcl, 0, n * 2, st);
  // dims (k = Ch, o = C1, position = N): out-of-bounds k / o are zero-filled up to the 64 x 64 box
  if (ytc_enc3(&y.mWah, y.Wah, Ch, C1, N, y.CS * 2, (long long)C1 * y.CS * 2, 64, 64, 1) || ytc_enc3(&y.mWal, y.Wal, Ch, C1, N, y.CS * 2, (long long)C1 * y.CS * 2, 64, 64, 1) ||
      ytc_enc3(&y.mWch, y.Wch, Ch, C1, N, y.CS * 2, (long long)C1 * y.CS * 2, 64, 64, 1) || ytc_enc3(&y.mWcl, y.Wcl, Ch, C1, N, y.CS * 2, (long long)C1 * y.CS * 2, 64, 64, 1)) return -1;
  cudaFuncSetAttribute(y_producer_tc_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)YTC_SMEM);
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  y.grid_sms = sms;
  y.ready = 1;
  return 0;
}
static int ytc_stage(YtcState& y, const float* WSa, const float* WSc, cudaStream_t st) {
  const long long rows = (long long)y.N * y.C1;
  ytc_stage_ws_k<<<cdiv(rows * y.Ch, 256), 256, 0, st>>>(WSa, y.Wah, y.Wal, rows, y.Ch, y.CS);
  ytc_stage_ws_k<<<cdiv(rows * y.Ch, 256), 256, 0, st>>>(WSc, y.Wch, y.Wcl, rows, y.Ch, y.CS);
  return tc_check_launch("ytc_stage_ws_k");
}
// ah/al, ch/cl: bf16 planes of a / c for the graphs of this micro-batch, [bc*N][CS]
// fixed-side planes (fh/fl: a for E1, c for the transposed copy), swept-side planes (rh/rl); swap = 0: fixed weights WSc, swept WSa;
// Rc = bias rows of the swept position, Sa = bias rows of the fixed position
static int ytc_run(YtcState& y, const __nv_bfloat16* ah, const __nv_bfloat16* al, const __nv_bfloat16* ch, const __nv_bfloat16* cl, int swap,
                   const float* Rc, const float* Sa, const float* b0, float* E1, int bc, cudaStream_t st) {
  CUtensorMap mah, mal, mch, mcl;
  const int N = y.N;
  if (ytc_enc3(&mah, ah, y.Ch, N, bc, y.CS * 2, (long long)N * y.CS * 2, 64, 1, 128) || ytc_enc3(&mal, al, y.Ch, N, bc, y.CS * 2, (long long)N * y.CS * 2, 64, 1, 128) ||
      ytc_enc3(&mch, ch, y.Ch, N, bc, y.CS * 2, (long long)N * y.CS * 2, 64, 1, 128) || ytc_enc3(&mcl, cl, y.Ch, N, bc, y.CS * 2, (long long)N * y.CS * 2, 64, 1, 128)) return -1;
  YtcArgs a; a.E1 = E1; a.Rc = Rc; a.Sa = Sa; a.b0 = b0; a.Bc = bc; a.N = N; a.C1 = y.C1; a.nbt = (bc + 127) / 128;
  // cut the swept positions into JS ranges per (tile, fixed position) when that shortens the last wave of the persistent grid
  // (nbt N items over 148 CTAs: 512 items = 4 waves at N = 256, 1024 half-items = 7 half-waves); each range re-loads the 48 KB
  // fixed operands, hence the small penalty per split
  a.JS = 1;
  { double best = 1e30;
    for (int js = 1; js <= 8; js *= 2) {
      if (N % (2 * js) != 0) break;
      const long long items = (long long)a.nbt * N * js;
      const double cost = (double)((items + y.grid_sms - 1) / y.grid_sms) / js * (1.0 + 0.01 * js);
      if (cost < best - 1e-9) { best = cost; a.JS = js; }
    } }
  const long long nwork = (long long)a.nbt * N * a.JS;
  const unsigned grid = (unsigned)(nwork < y.grid_sms ? nwork : y.grid_sms);
  if (swap) y_producer_tc_k<<<grid, 192, YTC_SMEM, st>>>(mah, mal, mch, mcl, y.mWch, y.mWcl, y.mWah, y.mWal, a);
  else y_producer_tc_k<<<grid, 192, YTC_SMEM, st>>>(mah, mal, mch, mcl, y.mWah, y.mWal, y.mWch, y.mWcl, a);
  return tc_check_launch("y_producer_tc_k");
}
