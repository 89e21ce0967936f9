// tsgemm.cuh -- the node-level contractions of the SND-VAE step on the 5th-gen tensor cores.
//
// Every dense product of the node-level path -- `linear` (layers.py:566-576 at model.py:113-115,127-129,149-151,177-179),
// tf.layers.conv1d as im2col rows x kernel (model.py:122,191,216), the coefficient-row products of the factored
// SpatialGraphConvolution (layers.py:171-196; sgc.cuh) and all of their input / weight gradients -- is one row-major fp32 GEMM
//     C[M, N] = alpha * op(A)[M, K] * op(B)[K, N] + beta * C (+ bias[N])
// with shapes that are tall and skinny (M = samples x nodes in the millions, K, N <= 256), or short with a huge reduction
// (weight gradients: K = samples x nodes), or plain (latent heads: K = N * channels).  One kernel covers them:
//
//  * operands stay fp32 in HBM; four producer warps read the tile with coalesced loads (along whichever index is contiguous
//    in memory -- a transposed operand is transposed by the index math of the loader, not by a copy), split every value into
//    bf16 hi + lo and write the K-major, un-swizzled canonical UMMA tiles (8 x 16-byte core matrices) by hand;
//  * one thread issues the 3-pass split product  hi.hi + hi.lo + lo.hi  as tcgen05.mma kind::f16 (M = 128, N <= 256) with fp32
//    accumulation in TMEM: fp32-grade results (~1e-6 relative) at tensor-core speed, so these GEMMs run at the speed their
//    operands stream from HBM;
//  * K loops longer than TG_GROUP chunks alternate between two TMEM slots while the epilogue warps drain the finished slot
//    into round-to-nearest fp32 registers (tcgen05.mma accumulates with truncation: see e2e_tc.cuh);
//  * weight gradients split the reduction over CTAs (grid.z) and add their partial tiles with red.global.add.f32.
#pragma once
#include "e2e_tc.cuh"

#define TG_BM 128        /* rows per CTA = MMA M */
#define TG_BK 32         /* K elements per pipeline stage (two MMA K steps of 16) */
#define TG_STAGES 3
#define TG_GROUP 16      /* K chunks per accumulation group: 96 tcgen05.mma accumulates */
#define TG_PROD 128      /* producer threads (warps 0..3); warp 4 issues the MMAs; warps 5..8 are the epilogue */
#define TG_THREADS (TG_PROD + 32 + 128)

struct TgArgs {
  const float* A; const float* B; float* C; const float* bias;
  long long M; int N, K;
  long long a_rs, a_ks;      // op(A)[m, k] = A[m * a_rs + k * a_ks]
  long long b_ns, b_ks;      // op(B)[k, n] = B[n * b_ns + k * b_ks]
  long long ldc;
  float alpha, beta;
  int chunks_per_split;      // K chunks handled by one CTA along grid.z
  int atomic;                // add the partial tile into C with atomics (split K; beta is taken as 1)
};

// 8 consecutive K values -> 8 bf16 "hi" (round to nearest) + 8 bf16 "lo" (the rounded remainder), lowest K at the lowest address
__device__ __forceinline__ void tg_split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Stage one operand tile: R rows (r0 ...) x 32 K values (k0 ...) of the operand whose element (r, k) lives at base[r * rs + k * ks];
// rows >= rmax and k >= kmax read as zero.  Canonical K-major tile without swizzle: core matrix (8 rows x 8 K) = 128 contiguous
// bytes, the 4 core matrices of a row group along K 128 B apart (LBO), row groups 512 B apart (SBO).
__device__ __forceinline__ void tg_load_tile(const float* __restrict__ base, long long rs, long long ks, long long r0, long long rmax,
                                             int k0, int kmax, int R, uint8_t* sh, uint8_t* sl, int t, bool vec_ok) {
  const int items = R * (TG_BK / 8);
  if (ks == 1) {                    // K contiguous in memory: 4 neighbouring threads read one row's 128 bytes
#pragma unroll 4
    for (int idx = t; idx < items; idx += TG_PROD) {
      const int row = idx >> 2, g = idx & 3;
      const long long r = r0 + row; const int k = k0 + 8 * g;
      float v[8];
      if (r < rmax && k < kmax) {
        const float* p = base + r * rs + k;
        if (vec_ok && k + 8 <= kmax) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (k + j < kmax) ? __ldg(p + j) : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      uint4 hi, lo; tg_split8(v, hi, lo);
      const int off = (row >> 3) * 512 + g * 128 + (row & 7) * 16;
      *reinterpret_cast<uint4*>(sh + off) = hi; *reinterpret_cast<uint4*>(sl + off) = lo;
    }
  } else {                          // rows contiguous in memory (a transposed operand): a warp reads 32 neighbouring rows per K value
#pragma unroll 2
    for (int idx = t; idx < items; idx += TG_PROD) {
      const int g = idx / R, row = idx - g * R;
      const long long r = r0 + row; const int k = k0 + 8 * g;
      float v[8];
      const float* p = base + r * rs + (long long)k * ks;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < rmax && k + j < kmax) ? __ldg(p + (long long)j * ks) : 0.f;
      uint4 hi, lo; tg_split8(v, hi, lo);
      const int off = (row >> 3) * 512 + g * 128 + (row & 7) * 16;
      *reinterpret_cast<uint4*>(sh + off) = hi; *reinterpret_cast<uint4*>(sl + off) = lo;
    }
  }
}

// BNMAX: widest N tile (TMEM columns per slot, accumulator registers of the grouped form); GROUPED: two-level accumulation
template <int BNMAX, bool GROUPED>
__global__ void __launch_bounds__(TG_THREADS, GROUPED ? 1 : (BNMAX <= 64 ? 3 : 2)) tsgemm_k(TgArgs P) {
  constexpr int A_BYTES = TG_BM * TG_BK * 2;          // one bf16 plane of the A tile: 8 KB
  constexpr int B_BYTES = BNMAX * TG_BK * 2;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = GROUPED ? 2 * BNMAX : BNMAX;
  static_assert(BNMAX == 64 || BNMAX == 128 || BNMAX == 256, "TMEM allocations are powers of two");
  static_assert(TMEM_COLS <= 512, "TMEM");
  extern __shared__ __align__(128) uint8_t tg_smem[];
  __shared__ uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * TG_BM;
  const int n0 = blockIdx.y * BNMAX;
  int bn = P.N - n0; if (bn > BNMAX) bn = BNMAX;
  const int bnc = (bn + 15) & ~15;                    // MMA N: a multiple of 16
  const int total_chunks = (P.K + TG_BK - 1) / TG_BK;
  const int c_lo = blockIdx.z * P.chunks_per_split;
  int nk = total_chunks - c_lo; if (nk > P.chunks_per_split) nk = P.chunks_per_split; if (nk < 0) nk = 0;
  const int ngroups = (nk + TG_GROUP - 1) / TG_GROUP;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&full_bar[s], TG_PROD); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp < 4) {
    // ---- producers: global fp32 -> bf16 hi / lo canonical tiles ----
    const int t = threadIdx.x;
    const bool a_vec = (P.a_ks == 1) && (P.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0);
    const bool b_vec = (P.b_ks == 1) && (P.b_ns % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0);
    for (int it = 0; it < nk; ++it) {
      const int s = it % TG_STAGES; const uint32_t ph = (it / TG_STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = tg_smem + (size_t)s * STAGE_BYTES;
      const int k0 = (c_lo + it) * TG_BK;
      tg_load_tile(P.A, P.a_rs, P.a_ks, m0, P.M, k0, P.K, TG_BM, st, st + A_BYTES, t, a_vec);
      tg_load_tile(P.B, P.b_ns, P.b_ks, n0, P.N, k0, P.K, bnc, st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, t, b_vec);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&full_bar[s]);
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(TG_BM, bnc, 0, 0);
      for (int it = 0; it < nk; ++it) {
        const int g = it / TG_GROUP, gi = it - g * TG_GROUP, slot = GROUPED ? (g & 1) : 0;
        if (GROUPED && gi == 0) { mbar_wait(&acc_empty[slot], ((g >> 1) & 1) ^ 1); tc_fence_after(); }
        const int s = it % TG_STAGES; const uint32_t ph = (it / TG_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tg_smem + (size_t)s * STAGE_BYTES);
        const uint32_t td = tmem_d + slot * BNMAX;
#pragma unroll
        for (int k = 0; k < TG_BK / 16; ++k) {        // one K step = two core matrices = 256 bytes further along the row group
          const uint64_t ah = umma_desc(sa + k * 256, 128, 512, 0ull);
          const uint64_t al = umma_desc(sa + A_BYTES + k * 256, 128, 512, 0ull);
          const uint64_t bh = umma_desc(sa + 2 * A_BYTES + k * 256, 128, 512, 0ull);
          const uint64_t bl = umma_desc(sa + 2 * A_BYTES + B_BYTES + k * 256, 128, 512, 0ull);
          umma_bf16(td, ah, bh, idesc, (gi | k) ? 1u : 0u);
          umma_bf16(td, ah, bl, idesc, 1u);
          umma_bf16(td, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
        if (gi == TG_GROUP - 1 || it == nk - 1) umma_commit(&acc_full[slot]);
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes [32 (w % 4), +32) = rows of the tile ----
    const int q = warp & 3;
    const long long row = m0 + q * 32 + lane;
    const bool rok = row < P.M;
    float* crow = P.C + row * P.ldc + n0;
    const bool c_vec = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) && (n0 % 4 == 0);
    if (nk == 0) {
      // nothing to multiply in this split: only the first split owns beta * C + bias
      if (!P.atomic && rok) for (int c = 0; c < bn; ++c) crow[c] = (P.beta != 0.f ? P.beta * crow[c] : 0.f) + (P.bias ? P.bias[n0 + c] : 0.f);
    } else if (!GROUPED) {
      mbar_wait(&acc_full[0], 0);
      tc_fence_after();
      const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < bnc; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(ta + c0, r);
        if (!rok) continue;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = P.alpha * __uint_as_float(r[i]);
        if (P.atomic) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (c0 + i < bn) atomicAdd(crow + c0 + i, v[i]);
        } else {
          if (P.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c0 + i < bn) v[i] += __ldg(P.bias + n0 + c0 + i);
          }
          if (c_vec && c0 + 16 <= bn) {
            float4* dst = reinterpret_cast<float4*>(crow + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (P.beta != 0.f) { const float4 c = dst[i]; o.x += P.beta * c.x; o.y += P.beta * c.y; o.z += P.beta * c.z; o.w += P.beta * c.w; }
              dst[i] = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c0 + i < bn) crow[c0 + i] = v[i] + (P.beta != 0.f ? P.beta * crow[c0 + i] : 0.f);
          }
        }
      }
    } else {
      float acc[BNMAX];
#pragma unroll
      for (int i = 0; i < BNMAX; ++i) acc[i] = 0.f;
      for (int g = 0; g < ngroups; ++g) {
        const int slot = g & 1;
        mbar_wait(&acc_full[slot], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem_d + slot * BNMAX + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < BNMAX; c0 += 16) {
          if (c0 < bnc) {
            uint32_t r[16];
            tmem_ld16(ta + c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[c0 + i] += __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
      }
      if (rok) {
#pragma unroll
        for (int i = 0; i < BNMAX; ++i) {
          if (i < bn) {
            const float v = P.alpha * acc[i];
            if (P.atomic) atomicAdd(crow + i, v);
            else crow[i] = v + (P.bias ? __ldg(P.bias + n0 + i) : 0.f) + (P.beta != 0.f ? P.beta * crow[i] : 0.f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_d, TMEM_COLS); }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <int BNMAX, bool GROUPED>
static cudaError_t tg_launch_t(const TgArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int smem = TG_STAGES * (2 * TG_BM * TG_BK * 2 + 2 * BNMAX * TG_BK * 2);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tsgemm_k<BNMAX, GROUPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  tsgemm_k<BNMAX, GROUPED><<<grid, TG_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

// row-major  C[M, N] = alpha op(A) op(B) + beta C (+ bias);  op(A) = A^T when tA (A stored [K, M]), op(B) = B^T when tB (B stored [N, K]).
// `launches` is incremented by the number of kernels / memsets enqueued.
static cudaError_t tsgemm(cudaStream_t st, bool tA, bool tB, long long M, int N, int K, float alpha, const float* A, long long lda,
                          const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, long long* launches) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  TgArgs a;
  a.A = A; a.B = B; a.C = C; a.bias = bias; a.M = M; a.N = N; a.K = K;
  a.a_rs = tA ? 1 : lda; a.a_ks = tA ? lda : 1;
  a.b_ns = tB ? ldb : 1; a.b_ks = tB ? 1 : ldb;
  a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.atomic = 0;
  const int chunks = (K + TG_BK - 1) / TG_BK;
  const int bnmax = N <= 64 ? 64 : 128;
  const long long mt = (M + TG_BM - 1) / TG_BM; const int nt = (N + bnmax - 1) / bnmax;
  int ksplit = 1;
  // a short output with a long reduction (weight gradients): split K over CTAs, partial tiles added with atomics
  if (bias == nullptr && chunks > TG_GROUP && mt * nt < 2 * 148) {
    long long want = (4 * 148 + mt * nt - 1) / (mt * nt);
    const long long most = (chunks + 3) / 4;                         // at least 4 chunks per CTA
    if (want > most) want = most;
    const long long least = (chunks + TG_GROUP - 1) / TG_GROUP;      // at most one accumulation group per CTA
    if (want < least) want = least;
    ksplit = (int)want;
  }
  a.chunks_per_split = (chunks + ksplit - 1) / ksplit;
  if (a.chunks_per_split < 1) a.chunks_per_split = 1;
  ksplit = chunks > 0 ? (chunks + a.chunks_per_split - 1) / a.chunks_per_split : 1;
  if (ksplit > 1) {
    a.atomic = 1;
    if (beta == 0.f) { cudaError_t e = cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st); if (e != cudaSuccess) return e; if (launches) ++*launches; }
    else if (beta != 1.f) return cudaErrorInvalidValue;              // not needed by any caller
  }
  if (mt > 0x7fffffffLL || ksplit > 65535 || nt > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)ksplit);
  const bool grouped = a.chunks_per_split > TG_GROUP;
  if (launches) ++*launches;
  if (bnmax == 64) return grouped ? tg_launch_t<64, true>(a, grid, st) : tg_launch_t<64, false>(a, grid, st);
  return grouped ? tg_launch_t<128, true>(a, grid, st) : tg_launch_t<128, false>(a, grid, st);
}
