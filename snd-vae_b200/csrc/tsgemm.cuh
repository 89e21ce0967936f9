// tsgemm.cuh -- the node-level contractions of the SND-VAE step on the 5th-gen tensor cores.
//
// Every dense product of the node-level path -- `linear` (layers.py:566-576 at model.py:113-115,127-129,149-151,177-179),
// tf.layers.conv1d as im2col rows x kernel (model.py:122,191,216), the coefficient-row products of the factored
// SpatialGraphConvolution (layers.py:171-196; sgc.cuh) and all of their input / weight gradients -- is one row-major fp32 GEMM
//     C[M, N] = alpha * op(A)[M, K] * op(B)[K, N] + beta * C (+ bias[N])
// with shapes that are tall and skinny (M = samples x nodes in the millions, K, N <= 256), or short with a huge reduction
// (weight gradients: K = samples x nodes), or plain (latent heads: K = N * channels).
//
//  * Operands stay fp32 in HBM.  Producer warps read the tile with coalesced loads (along whichever index is contiguous in
//    memory -- a transposed operand is transposed by the index math of the loader, not by a copy), split every value into
//    three bf16 planes  x = hi + mid + lo  (24 mantissa bits: the split is exact up to fp32 rounding) and write the K-major,
//    SWIZZLE_64B UMMA tiles (rows of 32 bf16, 16-byte groups XOR-ed with the row pair) by hand.
//  * One thread issues the six products of total order <= 2  (hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid)  as tcgen05.mma
//    kind::f16, M = 128, N <= 128, fp32 accumulation in TMEM: the dropped terms are <= 2^-24 of |a||b|, i.e. the result is an
//    fp32-grade product (weight gradients are cancelling sums over millions of rows -- a 3-pass bf16 split, 2^-17 per term,
//    was measured to leave 8e-3 relative error on a bias gradient).  The tensor pipe has the room: these products are bound by
//    the HBM stream of their operands.
//  * tsgemm_tall_k (M in the millions, K <= 320): persistent CTAs, op(B) converted once and resident in shared memory, 16 producer
//    warps software-pipelined one chunk ahead, two TMEM accumulator slots so that the epilogue of tile i (TMEM -> registers ->
//    padded shared tile -> coalesced global stores) runs under the loads of tile i + 1.
//  * tsgemm_k (everything else): one CTA per (M tile, N tile, K split); K loops longer than TG_GROUP chunks alternate between two
//    TMEM slots while the epilogue warps drain the finished slot into round-to-nearest fp32 registers (tcgen05.mma accumulates
//    with truncation, see e2e_tc.cuh); weight gradients split the reduction over grid.z and add their partial tiles with
//    red.global.add.f32.
#pragma once
#include "e2e_tc.cuh"

#define TG_BM 128        /* rows per tile = MMA M */
#define TG_BK 32         /* K elements per pipeline stage (two MMA K steps of 16) */
#define TG_NP 3          /* bf16 planes per operand */
#define TG_GROUP 8       /* K chunks per accumulation group: 96 tcgen05.mma accumulates */
#define TG_PROD 128      /* tsgemm_k: producer threads (warps 0..3); warp 4 issues the MMAs; warps 5..8 are the epilogue */
#define TG_THREADS (TG_PROD + 32 + 128 + 32)    /* + warp 9: bulk-copy loader of the slab form */
#define TG_PROD_SLAB 256 /* slab form: eight producer warps */
#define TG_THREADS_SLAB (TG_PROD_SLAB + 32 + 128 + 32)
#define TG_RAW 4          /* raw fp32 slab ring of the slab form */
#define TG_APLANE (TG_BM * TG_BK * 2)      /* one bf16 plane of an A chunk: 8 KB */
#define TT_SMEM_MAX (225 * 1024)           /* dynamic shared memory limit (227 KB per CTA minus the static part) */

struct TgArgs {
  const float* A; const float* B; float* C; const float* bias;
  long long M; int N, K;
  long long a_rs, a_ks;      // op(A)[m, k] = A[m * a_rs + k * a_ks]
  long long b_ns, b_ks;      // op(B)[k, n] = B[n * b_ns + k * b_ks]
  long long ldc;
  float alpha, beta;
  int chunks_per_split;      // tsgemm_k: K chunks handled by one CTA along grid.z
  int atomic;                // tsgemm_k: add the partial tile into C with atomics (split K; beta is taken as 1)
  int tiles_per_cta;         // tsgemm_tall_k: consecutive M tiles per CTA
  int bn_tile;               // tsgemm_tall_k: columns per N tile (a multiple of 16, <= TT_BN)
  long long split_stride;    // tsgemm_k, deterministic split K: split z writes its partial tile to C + z * split_stride (a workspace)
  int lda, ldb;              // slab form: row lengths (floats) of the K-major slabs of A and B
};

// 8 consecutive K values -> three packed bf16x8 planes, lowest K at the lowest address.  Pairs are converted with the packed
// cvt.rn.bf16x2.f32 (one F2FP for two values; the scalar F2F runs on the 16-lane conversion pipe and was measured to bound the
// producers) and the remainder is taken against the bf16 bits widened by a shift / mask.
__device__ __forceinline__ void tg_split8(const float* v, uint4* pl) {
  uint32_t w[TG_NP][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r0 = v[2 * i], r1 = v[2 * i + 1];
#pragma unroll
    for (int p = 0; p < TG_NP; ++p) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);          // .x = r0 (low half), .y = r1 (high half)
      const uint32_t bits = *reinterpret_cast<const uint32_t*>(&b);
      w[p][i] = bits;
      if (p + 1 < TG_NP) { r0 -= __uint_as_float(bits << 16); r1 -= __uint_as_float(bits & 0xffff0000u); }
    }
  }
#pragma unroll
  for (int p = 0; p < TG_NP; ++p) pl[p] = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}
// 4 consecutive K values -> three packed bf16x4 planes (8 bytes each)
__device__ __forceinline__ void tg_split4(const float* v, uint2* pl) {
  uint32_t w[TG_NP][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float r0 = v[2 * i], r1 = v[2 * i + 1];
#pragma unroll
    for (int p = 0; p < TG_NP; ++p) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(r0, r1);
      const uint32_t bits = *reinterpret_cast<const uint32_t*>(&b);
      w[p][i] = bits;
      if (p + 1 < TG_NP) { r0 -= __uint_as_float(bits << 16); r1 -= __uint_as_float(bits & 0xffff0000u); }
    }
  }
#pragma unroll
  for (int p = 0; p < TG_NP; ++p) pl[p] = make_uint2(w[p][0], w[p][1]);
}
// 4 K values of row r starting at k (K contiguous in memory)
__device__ __forceinline__ void tg_load4_kmajor(const float* __restrict__ base, long long rs, long long r, long long rmax, int k, int kmax,
                                                bool vec_ok, float* v) {
  if (r < rmax && k < kmax) {
    const float* p = base + r * rs + k;
    if (vec_ok && k + 4 <= kmax) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (k + j < kmax) ? __ldg(p + j) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = 0.f;
  }
}
// byte offset of (row, 8-wide K group g) inside one plane of a chunk tile.  K-major rows of 32 bf16 = 64 bytes in the
// SWIZZLE_64B layout of the UMMA descriptor: the 16-byte group g of row r sits at position g ^ ((r >> 1) & 3) of its row
// (Swizzle<2,4,3> on the byte address; tiles are 512-byte aligned).  Whatever way the threads of a warp are spread over (row, g)
// -- 4 groups of one row next to each other (K-contiguous operands) or 8 rows of one group (transposed operands) -- the 8 lanes
// of a 16-byte store phase then hit 8 different bank groups.  (The first version used the un-swizzled interleaved layout:
// 68 % of the kernel's shared-memory wavefronts were bank conflicts.)
__device__ __forceinline__ int tg_tile_off(int row, int g) { return row * 64 + ((g ^ ((row >> 1) & 3)) << 4); }
#define TG_DESC_ZERO umma_desc(0u, 16, 512, 4ull)     /* SWIZZLE_64B, 8-row groups 512 B apart; + (address >> 4) gives a tile's descriptor */
#define TG_KSTEP_BYTES 32                             /* one MMA K step (16 bf16) further along the swizzled row */

// 8 K values of row r (global row index) starting at k of an operand with K contiguous in memory
__device__ __forceinline__ void tg_load8_kmajor(const float* __restrict__ base, long long rs, long long r, long long rmax, int k, int kmax,
                                                bool vec_ok, float* v) {
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
}

// 8 K values of item (row, g) of a chunk for either orientation of the operand (ks == 1: K contiguous; else rows contiguous)
__device__ __forceinline__ void tg_load_item(const float* __restrict__ base, long long rs, long long ks, long long r, long long rmax,
                                             int k, int kmax, bool vec_ok, float* v) {
  if (ks == 1) tg_load8_kmajor(base, rs, r, rmax, k, kmax, vec_ok, v);
  else {
    const float* p8 = base + r * rs + (long long)k * ks;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (r < rmax && k + j < kmax) ? __ldg(p8 + (long long)j * ks) : 0.f;
  }
}
__device__ __forceinline__ void tg_store_item(const float* v, uint8_t* st, int pstride, int row, int g) {
  uint4 pl[TG_NP]; tg_split8(v, pl);
  const int off = tg_tile_off(row, g);
#pragma unroll
  for (int p = 0; p < TG_NP; ++p) *reinterpret_cast<uint4*>(st + p * pstride + off) = pl[p];
}

// Stage one operand chunk: R rows (r0 ...) x 32 K values (k0 ...) of the operand whose element (r, k) lives at base[r * rs + k * ks];
// rows >= rmax and k >= kmax read as zero.  `st` = plane 0, the other planes follow at `pstride` bytes.  nt threads, this one is t.
__device__ __forceinline__ void tg_load_tile(const float* __restrict__ base, long long rs, long long ks, long long r0, long long rmax,
                                             int k0, int kmax, int R, uint8_t* st, int pstride, int t, int nt, bool vec_ok) {
  const int items = R * (TG_BK / 8);
  if (ks == 1) {                    // K contiguous in memory: 4 neighbouring threads read one row's 128 bytes
#pragma unroll 2
    for (int idx = t; idx < items; idx += nt) {
      const int row = idx >> 2, g = idx & 3;
      float v[8];
      tg_load8_kmajor(base, rs, r0 + row, rmax, k0 + 8 * g, kmax, vec_ok, v);
      uint4 pl[TG_NP]; tg_split8(v, pl);
      const int off = tg_tile_off(row, g);
#pragma unroll
      for (int p = 0; p < TG_NP; ++p) *reinterpret_cast<uint4*>(st + p * pstride + off) = pl[p];
    }
  } else {                          // rows contiguous in memory (a transposed operand): a warp reads 32 neighbouring rows per K value
#pragma unroll 2
    for (int idx = t; idx < items; idx += nt) {
      const int g = idx / R, row = idx - g * R;
      const long long r = r0 + row; const int k = k0 + 8 * g;
      float v[8];
      const float* p8 = base + r * rs + (long long)k * ks;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < rmax && k + j < kmax) ? __ldg(p8 + (long long)j * ks) : 0.f;
      uint4 pl[TG_NP]; tg_split8(v, pl);
      const int off = tg_tile_off(row, g);
#pragma unroll
      for (int p = 0; p < TG_NP; ++p) *reinterpret_cast<uint4*>(st + p * pstride + off) = pl[p];
    }
  }
}

// the six split products of one 32-wide K chunk: A planes at sa + p * TG_APLANE, B planes at sb + p * bplane.  `dz` is the
// descriptor of shared address 0 (TG_DESC_ZERO): the start-address field is the low 14 bits in 16-byte units
// and every shared address is < 256 KB, so a tile's descriptor is dz + (address >> 4)
__device__ __forceinline__ void tg_issue_chunk(uint32_t td, uint32_t sa, uint32_t sb, uint32_t bplane, uint32_t idesc, bool first, uint64_t dz) {
#pragma unroll
  for (int k = 0; k < TG_BK / 16; ++k) {          // one K step = 32 bytes further along the swizzled rows
    uint64_t a[TG_NP], b[TG_NP];
#pragma unroll
    for (int p = 0; p < TG_NP; ++p) {
      a[p] = dz + ((sa + p * TG_APLANE + k * TG_KSTEP_BYTES) >> 4);
      b[p] = dz + ((sb + p * bplane + k * TG_KSTEP_BYTES) >> 4);
    }
    umma_bf16(td, a[0], b[0], idesc, (first && k == 0) ? 0u : 1u);
    umma_bf16(td, a[0], b[1], idesc, 1u);
    umma_bf16(td, a[1], b[0], idesc, 1u);
    umma_bf16(td, a[0], b[2], idesc, 1u);
    umma_bf16(td, a[2], b[0], idesc, 1u);
    umma_bf16(td, a[1], b[1], idesc, 1u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// tsgemm_k: one CTA per (M tile, N tile, K split).  BNMAX: widest N tile; GROUPED: two-level accumulation for long K loops.
// ---------------------------------------------------------------------------------------------------------------------------
// SLAB (weight gradients: both operands stored [K, *] with short rows): a 32-row K chunk of A and of B is one contiguous slab
// in memory, so warp 9 streams raw fp32 slabs into a ring with one cp.async.bulk each (TG_RAW chunks in flight, no registers
// tied up) and the producer warps transpose / split from shared memory instead of from global.
template <int BNMAX, bool GROUPED, bool SLAB>
__global__ void __launch_bounds__(SLAB ? TG_THREADS_SLAB : TG_THREADS, (SLAB || (GROUPED && BNMAX > 64)) ? 1 : 2) tsgemm_k(TgArgs P) {
  // warp roles: [0, PW) producers, PW issues the MMAs, PW + 1 .. PW + 4 epilogue (PW is a multiple of 4, so warp & 3 is still the
  // TMEM lane quarter), PW + 5 the slab loader.  The slab form is one CTA per SM (its raw ring fills the shared memory), so it
  // carries twice the producer warps: with four, the fp32 -> three-plane conversion of a chunk bounded it at ~1.2 TB/s.
  constexpr int PROD = SLAB ? TG_PROD_SLAB : TG_PROD, PW = PROD / 32;
  constexpr int STAGES = SLAB ? 2 : ((BNMAX <= 64 || GROUPED) ? 3 : 2);
  constexpr int B_PLANE = BNMAX * TG_BK * 2;
  constexpr int STAGE_BYTES = TG_NP * (TG_APLANE + B_PLANE);
  constexpr uint32_t TMEM_COLS = GROUPED ? 2 * BNMAX : BNMAX;
  static_assert(BNMAX == 64 || BNMAX == 128, "TMEM allocations are powers of two");
  extern __shared__ __align__(128) uint8_t tg_smem_raw[];
  uint8_t* tg_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 512-byte aligned tiles
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], raw_full[TG_RAW], raw_empty[TG_RAW];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * TG_BM;
  float* raw = reinterpret_cast<float*>(tg_smem + (size_t)STAGES * STAGE_BYTES);     // slab ring: [TG_RAW][32 lda + 32 ldb]
  const int raw_floats = SLAB ? 32 * (P.lda + P.ldb) : 0;
  const int n0 = blockIdx.y * BNMAX;
  int bn = P.N - n0; if (bn > BNMAX) bn = BNMAX;
  const int bnc = (bn + 15) & ~15;                    // MMA N: a multiple of 16
  const int total_chunks = (P.K + TG_BK - 1) / TG_BK;
  const int c_lo = blockIdx.z * P.chunks_per_split;
  int nk = total_chunks - c_lo; if (nk > P.chunks_per_split) nk = P.chunks_per_split; if (nk < 0) nk = 0;
  const int ngroups = (nk + TG_GROUP - 1) / TG_GROUP;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], PW); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    for (int s = 0; s < TG_RAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], PW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PW) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (SLAB && warp == PW + 5) {
    if (lane == 0) {
      int rc = 0;
      for (int it = 0; it < nk; ++it) {
        const long long k0 = (long long)(c_lo + it) * TG_BK;
        if (k0 + TG_BK > P.K) continue;                      // ragged last chunk: the producers read it from global
        const int rs = rc % TG_RAW;
        mbar_wait(&raw_empty[rs], ((rc / TG_RAW) & 1) ^ 1);
        float* dst = raw + (size_t)rs * raw_floats;
        mbar_expect_tx(&raw_full[rs], 128u * (uint32_t)(P.lda + P.ldb));
        bulk_load(dst, P.A + k0 * P.lda, 128u * (uint32_t)P.lda, &raw_full[rs]);
        bulk_load(dst + 32 * P.lda, P.B + k0 * P.ldb, 128u * (uint32_t)P.ldb, &raw_full[rs]);
        ++rc;
      }
    }
  } else if (SLAB && warp < PW) {
    const int t = threadIdx.x;
    constexpr int NA = TG_BM * 4 / PROD, NB = (BNMAX * 4 + PROD - 1) / PROD;
    int rc = 0;
    for (int it = 0; it < nk; ++it) {
      const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
      const int k0 = (c_lo + it) * TG_BK;
      const bool partial = k0 + TG_BK > P.K;
      const int rs = rc % TG_RAW;
      const float* rA = raw + (size_t)rs * raw_floats; const float* rB = rA + 32 * P.lda;
      if (lane == 0) { if (!partial) mbar_wait(&raw_full[rs], (rc / TG_RAW) & 1); mbar_wait(&empty_bar[s], ph ^ 1); }
      __syncwarp();
      uint8_t* st = tg_smem + (size_t)s * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int idx = t + q * PROD, row = idx & (TG_BM - 1), g = idx >> 7;
        if (m0 + row >= P.M) continue;           // rows past M only feed output rows that are never stored: left as they are
        float v[8];
        if (partial) tg_load_item(P.A, 1, P.lda, m0 + row, P.M, k0 + 8 * g, P.K, false, v);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rA[(8 * g + j) * P.lda + m0 + row];
        }
        tg_store_item(v, st, TG_APLANE, row, g);
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int idx = t + q * PROD;
        if (idx < bnc * 4) {
          const int g = idx / bnc, row = idx - g * bnc;
          if (n0 + row >= P.N) continue;         // likewise: columns past N are never stored
          float v[8];
          if (partial) tg_load_item(P.B, 1, P.ldb, n0 + row, P.N, k0 + 8 * g, P.K, false, v);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = rB[(8 * g + j) * P.ldb + n0 + row];
          }
          tg_store_item(v, st + TG_NP * TG_APLANE, B_PLANE, row, g);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&full_bar[s]); if (!partial) mbar_arrive(&raw_empty[rs]); }
      if (!partial) ++rc;
    }
  } else if (warp < PW) {
    // ---- producers: global fp32 -> bf16 planes in canonical tiles ----
    const int t = threadIdx.x;
    const bool a_vec = (P.a_ks == 1) && (P.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0);
    const bool b_vec = (P.b_ks == 1) && (P.b_ns % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0);
    // item (row, g) of thread t: K-contiguous operands give 4 neighbouring threads one row's 128 bytes; row-contiguous
    // (transposed) operands give a warp 32 neighbouring rows per K value.  Every load of the chunk is issued before the first
    // conversion (and before the wait for the stage), so that a thread has 48 - 64 values in flight.
    constexpr int NA = TG_BM * 4 / PROD, NB = (BNMAX * 4 + PROD - 1) / PROD;
    for (int it = 0; it < nk; ++it) {
      const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
      const int k0 = (c_lo + it) * TG_BK;
      float va[NA][8], vb[NB][8];
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int idx = t + q * PROD;
        const int row = P.a_ks == 1 ? idx >> 2 : idx & (TG_BM - 1), g = P.a_ks == 1 ? idx & 3 : idx >> 7;
        tg_load_item(P.A, P.a_rs, P.a_ks, m0 + row, P.M, k0 + 8 * g, P.K, a_vec, va[q]);
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int idx = t + q * PROD;
        if (idx < bnc * 4) {
          const int g = P.b_ks == 1 ? idx & 3 : idx / bnc, row = P.b_ks == 1 ? idx >> 2 : idx - g * bnc;
          tg_load_item(P.B, P.b_ns, P.b_ks, n0 + row, P.N, k0 + 8 * g, P.K, b_vec, vb[q]);
        }
      }
      if (lane == 0) mbar_wait(&empty_bar[s], ph ^ 1);      // one poller per warp; __syncwarp orders the others behind it
      __syncwarp();
      uint8_t* st = tg_smem + (size_t)s * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int idx = t + q * PROD;
        const int row = P.a_ks == 1 ? idx >> 2 : idx & (TG_BM - 1), g = P.a_ks == 1 ? idx & 3 : idx >> 7;
        tg_store_item(va[q], st, TG_APLANE, row, g);
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const int idx = t + q * PROD;
        if (idx < bnc * 4) {
          const int g = P.b_ks == 1 ? idx & 3 : idx / bnc, row = P.b_ks == 1 ? idx >> 2 : idx - g * bnc;
          tg_store_item(vb[q], st + TG_NP * TG_APLANE, B_PLANE, row, g);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);                        // one arrival per producer warp
    }
  } else if (warp == PW + 5) {
    // (idle in the plain form)
  } else if (warp == PW) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(TG_BM, bnc, 0, 0);
      const uint64_t dz = TG_DESC_ZERO;
      for (int it = 0; it < nk; ++it) {
        const int g = it / TG_GROUP, gi = it - g * TG_GROUP, slot = GROUPED ? (g & 1) : 0;
        if (GROUPED && gi == 0) { mbar_wait(&acc_empty[slot], ((g >> 1) & 1) ^ 1); tc_fence_after(); }
        const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tg_smem + (size_t)s * STAGE_BYTES);
        tg_issue_chunk(tmem_d + slot * BNMAX, sa, sa + TG_NP * TG_APLANE, B_PLANE, idesc, gi == 0, dz);
        umma_commit(&empty_bar[s]);
        if (gi == TG_GROUP - 1 || it == nk - 1) umma_commit(&acc_full[slot]);
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes [32 (w % 4), +32) = rows of the tile ----
    const int q = warp & 3;
    const long long row = m0 + q * 32 + lane;
    const bool rok = row < P.M;
    float* crow = P.C + (long long)blockIdx.z * P.split_stride + row * P.ldc + n0;
    const bool c_vec = (P.ldc % 4 == 0) && (P.split_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) && (n0 % 4 == 0);
    if (nk == 0) {
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
  if (warp == PW) { tc_fence_after(); tmem_dealloc(tmem_d, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// tsgemm_tall_k: persistent CTAs over consecutive M tiles, A with K contiguous (not transposed), op(B) resident.
// 16 producer warps (one 8-value item per thread and chunk, loads issued one chunk ahead of the conversion), warp 16 issues the
// MMAs, warps 17..20 are the epilogue.  Shared memory: B planes [chunk][plane][BNMAX x 32], A ring, padded output tile.
// ---------------------------------------------------------------------------------------------------------------------------
#define TT_PROD 512
#define TT_THREADS (TT_PROD + 32 + 128)
#define TT_STAGES 3
#define TT_BN 128         /* widest N tile of the persistent kernel: two B planes stacked along N make one MMA of N <= 256 */
#define TT_AHEAD 3        /* producer register prefetch distance (units of one 8-value item); 5 measured no faster */
// MMA shape.  The three bf16 planes of op(B) lie back to back along N, so that products sharing their A plane are one
// instruction over two stacked B planes; the six split products of a K step are four instructions into an accumulator of
// 2 bnc columns, and the epilogue adds the two column blocks (which block a product lands in does not matter):
//     A_hi  x [B_hi | B_mid]   (N = 2 bnc)      columns [0, bnc) += hi.hi,  [bnc, 2 bnc) += hi.mid
//     A_mid x [B_hi | B_mid]   (N = 2 bnc)      columns [0, bnc) += mid.hi, [bnc, 2 bnc) += mid.mid
//     A_lo  x [B_hi]           (N = bnc)        columns [0, bnc) += lo.hi
//     A_hi  x [B_lo]           (N = bnc)        columns [bnc, 2 bnc) += hi.lo
// (Measured: 6 -> 3 instructions per K step left the kernel's time unchanged -- it is not paced by the MMA count.)
__global__ void __launch_bounds__(TT_THREADS, 1) tsgemm_tall_k(TgArgs P) {
  constexpr int BNMAX = TT_BN;
  constexpr int A_STAGE = TG_NP * TG_APLANE;
  constexpr uint32_t SLOT_COLS = 256;                  // one accumulator slot: 2 bnc <= 256 columns
  constexpr uint32_t TMEM_COLS = 2 * SLOT_COLS;
  extern __shared__ __align__(128) uint8_t tg_smem_raw[];
  uint8_t* tg_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TT_STAGES], empty_bar[TT_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * P.bn_tile;
  int bn = P.N - n0; if (bn > P.bn_tile) bn = P.bn_tile;
  const int bnc = (bn + 15) & ~15;
  const int B_PLANE = bnc * TG_BK * 2;                  // planes back to back: rows p * bnc + n of the stacked operand
  const int SROW = P.bn_tile + 1;                       // padded row of the output staging tile (floats)
  const int nkc = (P.K + TG_BK - 1) / TG_BK;
  const long long mtiles = (P.M + TG_BM - 1) / TG_BM;
  const long long t_lo = (long long)blockIdx.x * P.tiles_per_cta;
  long long nt = mtiles - t_lo; if (nt > P.tiles_per_cta) nt = P.tiles_per_cta; if (nt < 0) nt = 0;
  uint8_t* sB = tg_smem;
  uint8_t* sA = tg_smem + (((size_t)nkc * TG_NP * B_PLANE + 1023) & ~(size_t)1023);
  float* sC = reinterpret_cast<float*>(sA + (size_t)TT_STAGES * A_STAGE);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TT_STAGES; ++s) { mbar_init(&full_bar[s], TT_PROD / 32); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc(&tmem_base_s, TMEM_COLS);
  {   // op(B): converted once per CTA by every thread
    const bool b_vec = (P.b_ks == 1) && (P.b_ns % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0);
    for (int c = 0; c < nkc; ++c)
      tg_load_tile(P.B, P.b_ns, P.b_ks, n0, P.N, c * TG_BK, P.K, bnc, sB + (size_t)c * TG_NP * B_PLANE, B_PLANE, threadIdx.x, TT_THREADS, b_vec);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const long long units = nt * nkc;

  if (warp < 16) {
    const int t = threadIdx.x, row = t >> 2, g = t & 3;
    const bool a_vec = (P.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0);
    // Each thread loads the K quads 4 g .. 4 g + 3 and 16 + 4 g .. 16 + 4 g + 3 of row t / 4 (g = t % 4): the four threads of a row
    // read 64 contiguous bytes per load instruction.  Neighbouring threads then swap one quad (shuffle) so that the even one holds
    // the 8 consecutive K values of group g / 2 and the odd one those of group 2 + g / 2, and each stores three 16-byte plane pieces.
    // Register ring: the loads of unit u + TT_AHEAD are issued before unit u is converted (TT_AHEAD x 32 bytes in flight per thread).
    const int off = tg_tile_off(row, (g & 1) ? 2 + (g >> 1) : (g >> 1));
    float buf[TT_AHEAD + 1][8];
    long long lt = 0; int lc = 0;                                     // (tile, chunk) of the next unit to load
#define TT_LOAD_(dst) do { const long long r_ = (t_lo + lt) * TG_BM + row; const int k_ = lc * TG_BK + 4 * g; \
      tg_load4_kmajor(P.A, P.a_rs, r_, P.M, k_, P.K, a_vec, dst); tg_load4_kmajor(P.A, P.a_rs, r_, P.M, k_ + 16, P.K, a_vec, dst + 4); } while (0)
#pragma unroll
    for (int d = 0; d < TT_AHEAD; ++d) {
      if (d < units) TT_LOAD_(buf[d]);
      if (++lc == nkc) { lc = 0; ++lt; }
    }
    for (long long u0 = 0; u0 < units; u0 += TT_AHEAD + 1) {
#pragma unroll
      for (int d = 0; d <= TT_AHEAD; ++d) {
        const long long u = u0 + d;
        if (u < units) {             // `units` is the same for every thread of the CTA: the shuffles below are warp-uniform
          if (u + TT_AHEAD < units) TT_LOAD_(buf[(d + TT_AHEAD) % (TT_AHEAD + 1)]);
          if (++lc == nkc) { lc = 0; ++lt; }
          const int s = (int)(u % TT_STAGES); const uint32_t ph = (uint32_t)((u / TT_STAGES) & 1);
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // even thread: keeps its first quad, receives the partner's first quad; odd thread: receives the partner's second, keeps its second
            const float send = (g & 1) ? buf[d][j] : buf[d][4 + j];
            const float got = __shfl_xor_sync(0xffffffffu, send, 1);
            v[j] = (g & 1) ? got : buf[d][j];
            v[4 + j] = (g & 1) ? buf[d][4 + j] : got;
          }
          uint4 pl[TG_NP]; tg_split8(v, pl);
          if (lane == 0) mbar_wait(&empty_bar[s], ph ^ 1);
          __syncwarp();
          uint8_t* st = sA + (size_t)s * A_STAGE + off;
#pragma unroll
          for (int p = 0; p < TG_NP; ++p) *reinterpret_cast<uint4*>(st + p * TG_APLANE) = pl[p];
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[s]);
        }
      }
    }
#undef TT_LOAD_
  } else if (warp == 16) {
    if (lane == 0) {
      const uint32_t idesc2 = umma_idesc(TG_BM, 2 * bnc, 0, 0), idesc1 = umma_idesc(TG_BM, bnc, 0, 0);
      const uint64_t dz = TG_DESC_ZERO;
      const uint32_t sb0 = smem_u32(sB), sa0 = smem_u32(sA);
      long long u = 0;
      for (long long ti = 0; ti < nt; ++ti) {
        const int slot = (int)(ti & 1);
        mbar_wait(&acc_empty[slot], (uint32_t)(((ti >> 1) & 1) ^ 1));
        tc_fence_after();
        for (int c = 0; c < nkc; ++c, ++u) {
          const int s = (int)(u % TT_STAGES); const uint32_t ph = (uint32_t)((u / TT_STAGES) & 1);
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t td = tmem_d + slot * SLOT_COLS, sa = sa0 + s * A_STAGE, sb = sb0 + c * TG_NP * B_PLANE;
#pragma unroll
          for (int k = 0; k < TG_BK / 16; ++k) {
            const uint64_t b = dz + ((sb + k * TG_KSTEP_BYTES) >> 4), bl = dz + ((sb + 2 * B_PLANE + k * TG_KSTEP_BYTES) >> 4);
            const uint64_t ah = dz + ((sa + k * TG_KSTEP_BYTES) >> 4);
            umma_bf16(td, ah, b, idesc2, (c == 0 && k == 0) ? 0u : 1u);
            umma_bf16(td, dz + ((sa + TG_APLANE + k * TG_KSTEP_BYTES) >> 4), b, idesc2, 1u);
            umma_bf16(td, dz + ((sa + 2 * TG_APLANE + k * TG_KSTEP_BYTES) >> 4), b, idesc1, 1u);
            umma_bf16(td + bnc, ah, bl, idesc1, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[slot]);
      }
    }
  } else {
    const int q = warp & 3;
    float* srow = sC + (size_t)(q * 32) * SROW;         // this warp's 32 rows of the staging tile (only this warp touches them)
    const bool flat = (P.ldc == bn) && (n0 == 0);         // the warp's 32 x bn block is contiguous in C
    // contiguous block, nothing to read back: stage it in C's own layout and hand it to one bulk store per tile
    const bool bulk = flat && P.beta == 0.f && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    const int sr = bulk ? bn : SROW;                      // staging row pitch (floats)
    bool pending = false;
    for (long long ti = 0; ti < nt; ++ti) {
      const int slot = (int)(ti & 1);
      mbar_wait(&acc_full[slot], (uint32_t)((ti >> 1) & 1));
      tc_fence_after();
      if (pending) { if (lane == 0) bulk_wait_read(); __syncwarp(); pending = false; }     // the previous store has left the staging rows
      const uint32_t ta = tmem_d + slot * SLOT_COLS + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < bnc; c0 += 16) {
        uint32_t r0[16], r1[16];
        tmem_ld16_nowait(ta + c0, r0); tmem_ld16_nowait(ta + bnc + c0, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < bn)
            srow[lane * sr + c0 + i] = P.alpha * (__uint_as_float(r1[i]) + __uint_as_float(r0[i])) + (P.bias ? __ldg(P.bias + n0 + c0 + i) : 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);        // the accumulator slot is free for tile ti + 2
      const long long r0 = (t_lo + ti) * TG_BM + q * 32;
      long long nr = P.M - r0; if (nr > 32) nr = 32;
      if (nr == 32 && bulk) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bulk_store(P.C + r0 * P.ldc, srow, 32u * (uint32_t)bn * 4u);
        pending = true;
      } else if (nr > 0) {
        if (flat) {
          float* dst = P.C + r0 * P.ldc;
          const int total = (int)nr * bn;
          for (int e0 = 0; e0 < total; e0 += 128) {
            float v[4]; int ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              ee[u] = e0 + u * 32 + lane;
              const int rr = ee[u] / bn, cc = ee[u] - rr * bn;
              v[u] = ee[u] < total ? srow[rr * sr + cc] : 0.f;
            }
            if (P.beta != 0.f) {
#pragma unroll
              for (int u = 0; u < 4; ++u) if (ee[u] < total) v[u] += P.beta * dst[ee[u]];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (ee[u] < total) dst[ee[u]] = v[u];
          }
        } else {
          for (int rr = 0; rr < (int)nr; rr += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (rr + u < (int)nr) {
                float* dst = P.C + (r0 + rr + u) * P.ldc + n0;
                for (int cc = lane; cc < bn; cc += 32) {
                  float v = srow[(rr + u) * sr + cc];
                  if (P.beta != 0.f) v += P.beta * dst[cc];
                  dst[cc] = v;
                }
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (pending && lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_d, TMEM_COLS); }
}

// C[m, n] = beta C + bias[n] + sum_z ws[z][m][n]   (fixed summation order: the deterministic half of a split-K product)
__global__ void tg_reduce_k(const float* __restrict__ ws, int nsplit, long long MN, int N, float* __restrict__ C, long long ldc,
                            const float* __restrict__ bias, float beta) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= MN) return;
  const long long m = idx / N; const int n = (int)(idx - m * N);
  float acc = 0.f;
  for (int z = 0; z < nsplit; ++z) acc += ws[(long long)z * MN + idx];
  float* c = C + m * ldc + n;
  *c = acc + (bias ? __ldg(bias + n) : 0.f) + (beta != 0.f ? beta * *c : 0.f);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <int BNMAX, bool GROUPED, bool SLAB>
static cudaError_t tg_launch_t(const TgArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = SLAB ? 2 : ((BNMAX <= 64 || GROUPED) ? 3 : 2);
  const int smem = 1024 + STAGES * TG_NP * (TG_APLANE + BNMAX * TG_BK * 2) + (SLAB ? TG_RAW * 128 * (a.lda + a.ldb) : 0);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tsgemm_k<BNMAX, GROUPED, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  tsgemm_k<BNMAX, GROUPED, SLAB><<<grid, SLAB ? TG_THREADS_SLAB : TG_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
template <int BNMAX>
static cudaError_t tg_launch(const TgArgs& a, dim3 grid, cudaStream_t st, bool grouped, bool slab) {
  if (slab) return grouped ? tg_launch_t<BNMAX, true, true>(a, grid, st) : tg_launch_t<BNMAX, false, true>(a, grid, st);
  return grouped ? tg_launch_t<BNMAX, true, false>(a, grid, st) : tg_launch_t<BNMAX, false, false>(a, grid, st);
}
static size_t tt_smem_bytes(int K, int bnc) {
  const int nkc = (K + TG_BK - 1) / TG_BK;
  return 1024 + (((size_t)nkc * TG_NP * bnc * TG_BK * 2 + 1023) & ~(size_t)1023) + (size_t)TT_STAGES * TG_NP * TG_APLANE + (size_t)TG_BM * (bnc + 1) * 4;
}
static cudaError_t tt_launch(const TgArgs& a, dim3 grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tsgemm_tall_k, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  tsgemm_tall_k<<<grid, TT_THREADS, tt_smem_bytes(a.K, a.bn_tile), st>>>(a);
  return cudaGetLastError();
}

// row-major  C[M, N] = alpha op(A) op(B) + beta C (+ bias);  op(A) = A^T when tA (A stored [K, M]), op(B) = B^T when tB (B stored [N, K]).
// Results are deterministic except for weight-gradient shapes (tA with a long reduction), whose K splits are added with atomics.
// `launches` is incremented by the number of kernels / memsets enqueued.
// `ws` (ws_floats floats, may be NULL): workspace for the deterministic split of long-K products with few output tiles.
static cudaError_t tsgemm(cudaStream_t st, bool tA, bool tB, long long M, int N, int K, float alpha, const float* A, long long lda,
                          const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, long long* launches,
                          float* ws = nullptr, size_t ws_floats = 0) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  TgArgs a;
  a.A = A; a.B = B; a.C = C; a.bias = bias; a.M = M; a.N = N; a.K = K;
  a.a_rs = tA ? 1 : lda; a.a_ks = tA ? lda : 1;
  a.b_ns = tB ? ldb : 1; a.b_ks = tB ? 1 : ldb;
  a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.atomic = 0; a.tiles_per_cta = 1; a.chunks_per_split = 1; a.split_stride = 0; a.bn_tile = 0;
  const int chunks = (K + TG_BK - 1) / TG_BK;
  const int bnmax = N <= 64 ? 64 : 128;
  const long long mt = (M + TG_BM - 1) / TG_BM; const int nt = (N + bnmax - 1) / bnmax;
  if (mt > 0x7fffffffLL || nt > 65535) return cudaErrorInvalidValue;
  // tall and skinny with op(B) small enough to stay in shared memory: the persistent kernel, N tiles of equal width <= TT_BN
  const int tnt = (N + TT_BN - 1) / TT_BN;
  const int tbn = (((N + tnt - 1) / tnt) + 15) & ~15;
  if (!tA && K > 0 && mt >= 148 && tbn <= TT_BN && tt_smem_bytes(K, tbn) <= TT_SMEM_MAX) {
    int device = 0, sms = 148; cudaGetDevice(&device); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int tnt2 = (N + tbn - 1) / tbn;
    long long ctas = sms / tnt2; if (ctas < 1) ctas = 1; if (ctas > mt) ctas = mt;
    a.tiles_per_cta = (int)((mt + ctas - 1) / ctas);
    ctas = (mt + a.tiles_per_cta - 1) / a.tiles_per_cta;
    a.bn_tile = tbn;
    dim3 grid((unsigned)ctas, (unsigned)tnt2, 1);
    if (launches) ++*launches;
    return tt_launch(a, grid, st);
  }
  int ksplit = 1;
  // a long reduction with few output tiles: split K over CTAs.  Weight gradients (tA: the reduction runs over rows) add their
  // partial tiles with atomics; forward / input-gradient products (the latent heads at small batch) must stay deterministic, so
  // their partial tiles go to the workspace and are summed in a fixed order by tg_reduce_k.
  bool split_ws = false;
  if (chunks >= 2 * TG_GROUP && mt * nt < 2 * 148) {
    long long want = (2 * 148) / (mt * nt);                            // one wave of two CTAs per SM
    if (want < 1) want = 1;
    const long long most = chunks / TG_GROUP;                        // at least one accumulation group per CTA
    if (want > most) want = most;
    if (want > 65535) want = 65535;
    if (!tA || bias != nullptr) {
      const long long room = ws ? (long long)(ws_floats / (size_t)(M * N)) : 0;
      if (want > room) want = room;
      split_ws = want > 1;
    }
    if (want > 1) ksplit = (int)want;
  }
  a.chunks_per_split = (chunks + ksplit - 1) / ksplit;
  if (a.chunks_per_split < 1) a.chunks_per_split = 1;
  ksplit = chunks > 0 ? (chunks + a.chunks_per_split - 1) / a.chunks_per_split : 1;
  if (ksplit > 1 && split_ws) {
    a.C = ws; a.ldc = N; a.split_stride = M * N; a.bias = nullptr; a.beta = 0.f;
  } else if (ksplit > 1) {
    a.atomic = 1;
    if (beta == 0.f) { cudaError_t e = cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st); if (e != cudaSuccess) return e; if (launches) ++*launches; }
    else if (beta != 1.f) return cudaErrorInvalidValue;              // not needed by any caller
  }
  if (ksplit > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)ksplit);
  const bool grouped = a.chunks_per_split > TG_GROUP;
  if (launches) ++*launches;
  // slab form: both operands K-major with rows short enough for a ring of raw slabs, 16-byte aligned bases
  a.lda = (int)lda; a.ldb = (int)ldb;
  const bool slab = tA && !tB && chunks >= 4 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 &&
                    1024 + 2 * TG_NP * (TG_APLANE + bnmax * TG_BK * 2) + TG_RAW * 128 * (lda + ldb) <= TT_SMEM_MAX;
  cudaError_t e = bnmax == 64 ? tg_launch<64>(a, grid, st, grouped, slab) : tg_launch<128>(a, grid, st, grouped, slab);
  if (e != cudaSuccess || !(ksplit > 1 && split_ws)) return e;
  tg_reduce_k<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(ws, ksplit, M * N, N, C, ldc, bias, beta);
  if (launches) ++*launches;
  return cudaGetLastError();
}
