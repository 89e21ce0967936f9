// e2e_tc.cuh -- e2e layer 1 (layers.py:431-450 at model.py:202) on the 5th-gen tensor cores.
//
// The width-N "cross" convolution is a GEMM against a block-Toeplitz matrix shared by the
// whole batch:
//     fwd   : O [rows, (j,q)]  = Y [rows, (j',c)] . T          T[(j',c),(j,q)] = w1[j'-j+p, c, q]
//     dgrad : dY[rows, (j',c)] = dO[rows, (j,q)]  . T^T
//     wgrad : dT[(j',c),(j,q)] = Y^T . dO, summed along block diagonals into dw1[t,c,q]
// rows = (direction, graph, line): both conv directions are stacked into one GEMM.
//
// fp32-grade parity (rtol 1e-4 on logits, K up to N*50) rules out a single bf16 pass, so
// every product is the 3-pass split   x.y ~= xh.yh + xh.yl + xl.yh   (bf16 hi/lo planes,
// fp32 accumulation in TMEM):  executed MMA flops = 3 x algorithmic flops.
//
// T is never materialised (262 MB at N=256).  The B operand of fwd/dgrad is a TMA *window
// view* of a zero-padded, channel-major copy of w1:  row (ch, jr) of the tile is the
// contiguous slice  Wp[ch][x + jr*CS : +64]  (jr = N-1-j, CS = padded channel stride), i.e. a
// 3-D tensor map (x, jr, ch) whose jr-stride (CS elements) is smaller than the x extent.
//
// Kernels: one CTA (192 threads) per 128 x BN output tile; warp 0 = TMA producer, warp 1 =
// TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld).
// smem ring of 2 stages (K chunk 64, SWIZZLE_128B), mbarrier full/empty pipeline.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_runtime.h>

#define TC_C1 50      /* e2e layer-1 input channels  */
#define TC_C2 20      /* e2e layer-1 output channels */
#define TC_CP 56      /* channel stride of Y planes   (112 B: TMA strides need 16 B multiples) */
#define TC_OP 24      /* channel stride of dO planes  (48 B) */
#define TC_BM 128
#define TC_KC 64      /* K elements per stage = one 128-byte swizzle row */
#define TC_STAGES 2

static char g_tc_err[256] = "";
static const char* tc_last_error() { return g_tc_err; }

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// fwd / dgrad: K-major Toeplitz GEMM.  out[row, jout*Cout + ch] for the tile's (ch, jr) columns.
// ------------------------------------------------------------------------------------------
struct ToepArgs {
  float* out;          // [rows, N*Cout]
  long long rows;      // valid rows
  int N;               // nodes
  int Cout;            // output channels per position (20 fwd, 50 dgrad)
  int CS;              // channel stride of the K index (56 fwd, 24 dgrad)
  int CT, JT;          // tile = CT channels x JT positions  (BN = CT*JT)
  int n_ctiles;        // Cout / CT
  int n_jtiles;        // ceil(N / JT)
  int pad_rows;        // zero rows in front of the padded weights (N-1-p fwd, p dgrad)
  int KA;              // K extent = N*CS
};

template <int BN>
__global__ void __launch_bounds__(192, 1) toep_gemm_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                      const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                      ToepArgs P) {
  constexpr int A_BYTES = TC_BM * TC_KC * 2;       // 16 KB
  constexpr int B_BYTES = BN * TC_KC * 2;          // 30 KB at BN = 240
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = blockIdx.x % (P.n_ctiles * P.n_jtiles);
  const int mtile = blockIdx.x / (P.n_ctiles * P.n_jtiles);
  const int ct = ntile % P.n_ctiles, jt = ntile / P.n_ctiles;
  const int ch0 = ct * P.CT, jr0 = jt * P.JT;
  const int m0 = mtile * TC_BM;
  // K chunks that can touch a non-zero weight: x + jr*CS in [pad_rows*CS, (pad_rows+N)*CS)
  long long xlo = (long long)(P.pad_rows - jr0 - P.JT + 1) * P.CS - (TC_KC - 1);
  long long xhi = (long long)(P.pad_rows + P.N - jr0) * P.CS;     // exclusive
  if (xlo < 0) xlo = 0;
  if (xhi > P.KA) xhi = P.KA;
  const int kc_lo = (int)(xlo / TC_KC);
  const int kc_hi = (int)((xhi + TC_KC - 1) / TC_KC);
  const int nk = kc_hi > kc_lo ? kc_hi - kc_lo : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_STAGES; const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int x = (kc_lo + it) * TC_KC;
        tma_load_2d(st, &tmAh, &full_bar[s], x, m0);
        tma_load_2d(st + A_BYTES, &tmAl, &full_bar[s], x, m0);
        tma_load_3d(st + 2 * A_BYTES, &tmBh, &full_bar[s], x, jr0, ch0);
        tma_load_3d(st + 2 * A_BYTES + B_BYTES, &tmBl, &full_bar[s], x, jr0, ch0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(TC_BM, BN, 0, 0);
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_STAGES; const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < TC_KC / 16; ++k) {
          // K-major SWIZZLE_128B: 8-row groups 1024 B apart; +32 B per 16-element K step
          const uint64_t ah = umma_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t al = umma_desc_sw128(sa + A_BYTES + k * 32, 16, 1024);
          const uint64_t bh = umma_desc_sw128(sa + 2 * A_BYTES + k * 32, 16, 1024);
          const uint64_t bl = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES + k * 32, 16, 1024);
          umma_bf16(tmem_d, ah, bh, idesc, (it | k) ? 1u : 0u);
          umma_bf16(tmem_d, ah, bl, idesc, 1u);
          umma_bf16(tmem_d, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);          // smem slot free once these MMAs retire
      }
      umma_commit(&done_bar);                // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    if (nk > 0) { mbar_wait(&done_bar, 0); tc_fence_after(); }
    const int q = warp & 3;
    const long long row = (long long)m0 + q * 32 + lane;
    float* orow = P.out + row * (long long)P.N * P.Cout;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      if (nk > 0) tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0u;
      }
      if (row < P.rows) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = c0 + i;
          const int chl = n / P.JT, jr = jr0 + n - chl * P.JT;
          const int j = P.N - 1 - jr;
          if (j >= 0) orow[(long long)j * P.Cout + ch0 + chl] = __uint_as_float(r[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// wgrad: MN-major GEMM  dT[m, n] = sum_rows Y[row, m] dO[row, n]; the epilogue folds the
// block diagonals:  dw1[t, c, q] += dT[(j',c),(j,q)],  t = j'-j+p  
// ------------------------------------------------------------------------------------------
struct WgradArgs {
  float* dw1;          // [N, C1, C2] gradient slot (atomicAdd)
  long long rows;
  int N;
  int n_ntiles;        // ceil(N*OP / 256)
  int ksplit;          // CTAs per output tile along K
};
#define TC_WN 256
__global__ void __launch_bounds__(192, 1) wgrad_gemm_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                       const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                       WgradArgs P) {
  constexpr int A_BYTES = TC_KC * TC_BM * 2;       // [2 x (64 k-rows x 128 B)] = 16 KB
  constexpr int B_BYTES = TC_KC * TC_WN * 2;       // [4 x (64 k-rows x 128 B)] = 32 KB
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
  constexpr int BLK = TC_KC * 128;                 // one 64(mn) x 64(k) box = 8 KB
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / P.ksplit, ks = blockIdx.x % P.ksplit;
  const int nt = tile % P.n_ntiles, mt = tile / P.n_ntiles;
  const int m0 = mt * TC_BM, n0 = nt * TC_WN;
  const int p = (P.N - 1) / 2;
  // band test: does any (j', j) in this tile have 0 <= j'-j+p < N ?
  const int jp_lo = m0 / TC_CP, jp_hi = min(P.N - 1, (m0 + TC_BM - 1) / TC_CP);
  const int j_lo = n0 / TC_OP, j_hi = min(P.N - 1, (n0 + TC_WN - 1) / TC_OP);
  if (jp_lo > P.N - 1 || j_lo > P.N - 1) return;
  if (jp_hi - j_lo + p < 0 || jp_lo - j_hi + p > P.N - 1) return;
  const long long kchunks = (P.rows + TC_KC - 1) / TC_KC;
  const long long per = (kchunks + P.ksplit - 1) / P.ksplit;
  const long long kc_lo = (long long)ks * per;
  long long kc_hi = kc_lo + per; if (kc_hi > kchunks) kc_hi = kchunks;
  const int nk = kc_hi > kc_lo ? (int)(kc_hi - kc_lo) : 0;
  if (nk == 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_STAGES; const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int r0 = (int)((kc_lo + it) * TC_KC);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          tma_load_2d(st + b * BLK, &tmAh, &full_bar[s], m0 + 64 * b, r0);
          tma_load_2d(st + A_BYTES + b * BLK, &tmAl, &full_bar[s], m0 + 64 * b, r0);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          tma_load_2d(st + 2 * A_BYTES + b * BLK, &tmBh, &full_bar[s], n0 + 64 * b, r0);
          tma_load_2d(st + 2 * A_BYTES + B_BYTES + b * BLK, &tmBl, &full_bar[s], n0 + 64 * b, r0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(TC_BM, TC_WN, 1, 1);
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_STAGES; const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < TC_KC / 16; ++k) {
          // MN-major SWIZZLE_128B: 64-element MN blocks BLK bytes apart (LBO), 8-row K groups
          // 1024 B apart (SBO); +2048 B per 16-row K step
          const uint64_t ah = umma_desc_sw128(sa + k * 2048, BLK, 1024);
          const uint64_t al = umma_desc_sw128(sa + A_BYTES + k * 2048, BLK, 1024);
          const uint64_t bh = umma_desc_sw128(sa + 2 * A_BYTES + k * 2048, BLK, 1024);
          const uint64_t bl = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES + k * 2048, BLK, 1024);
          umma_bf16(tmem_d, ah, bh, idesc, (it | k) ? 1u : 0u);
          umma_bf16(tmem_d, ah, bl, idesc, 1u);
          umma_bf16(tmem_d, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&done_bar);
    }
  } else {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const int jp = m / TC_CP, c = m - jp * TC_CP;
    const bool mok = jp < P.N && c < TC_C1;
#pragma unroll 1
    for (int c0 = 0; c0 < TC_WN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
      if (mok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + c0 + i;
          const int j = n / TC_OP, qq = n - j * TC_OP;
          const int t = jp - j + p;
          if (j < P.N && qq < TC_C2 && t >= 0 && t < P.N)
            atomicAdd(P.dw1 + ((size_t)t * TC_C1 + c) * TC_C2 + qq, __uint_as_float(r[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 256); }
}

// ------------------------------------------------------------------------------------------
// per-step weight staging: zero-padded, channel-major, bf16 hi/lo copies of w1 [N, C1, C2]
//   Wf[q][rho*CP + c] = w1[rho - (N-1-p), c, q]      rho in [0, 2N-1)   (forward B operand)
//   Wd[c][rho*OP + q] = w1[N-1+p - rho,  c, q]                           (dgrad  B operand)
// ------------------------------------------------------------------------------------------
__global__ void tc_stage_weights_k(const float* __restrict__ w1, __nv_bfloat16* __restrict__ Wfh, __nv_bfloat16* __restrict__ Wfl,
                                   __nv_bfloat16* __restrict__ Wdh, __nv_bfloat16* __restrict__ Wdl, int N) {
  const int p = (N - 1) / 2;
  const long long LF = (long long)(2 * N - 1) * TC_CP, LD = (long long)(2 * N - 1) * TC_OP;
  const long long nf = (long long)TC_C2 * LF, nd = (long long)TC_C1 * LD;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nf) {
    int q = (int)(idx / LF); long long x = idx - q * LF; int rho = (int)(x / TC_CP), c = (int)(x - (long long)rho * TC_CP);
    int t = rho - (N - 1 - p);
    float v = (c < TC_C1 && t >= 0 && t < N) ? w1[((size_t)t * TC_C1 + c) * TC_C2 + q] : 0.f;
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Wfh[idx] = hi; Wfl[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  } else if (idx < nf + nd) {
    idx -= nf;
    int c = (int)(idx / LD); long long x = idx - c * LD; int rho = (int)(x / TC_OP), q = (int)(x - (long long)rho * TC_OP);
    int t = N - 1 + p - rho;
    float v = (q < TC_C2 && t >= 0 && t < N) ? w1[((size_t)t * TC_C1 + c) * TC_C2 + q] : 0.f;
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Wdh[idx] = hi; Wdl[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  int ready;
  int N;
  long long max_rows;
  __nv_bfloat16 *Wfh, *Wfl, *Wdh, *Wdl;
  PFN_encodeTiled encode;
  // tensor maps over the weights are fixed; maps over Y / dO planes are rebuilt when pointers change
  const void *yh, *yl, *dh, *dl;
  const void *wyh, *wdh; long long wrows;
  CUtensorMap fA_h, fA_l, fB_h, fB_l;      // forward: A = Y planes (K-major), B = Wf window view
  CUtensorMap dA_h, dA_l, dB_h, dB_l;      // dgrad:   A = dO planes,          B = Wd window view
  CUtensorMap wA_h, wA_l, wB_h, wB_l;      // wgrad:   A = Y planes (MN-major boxes), B = dO planes
};

static int tc_encode(TcState& s, CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                     const cuuint32_t* box) {
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = s.encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled failed: %d (rank %d)", (int)r, rank); return -1; }
  return 0;
}

static int tc_init(TcState& s, int N, long long max_rows, cudaStream_t st) {
  memset(&s, 0, sizeof s);
  s.N = N; s.max_rows = max_rows;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled entry point not found"); return -1;
  }
  s.encode = (PFN_encodeTiled)fn;
  const long long LF = (long long)(2 * N - 1) * TC_CP, LD = (long long)(2 * N - 1) * TC_OP;
  // + one K chunk of slack so that window reads past the last row stay inside the allocation
  size_t nf = (size_t)TC_C2 * LF + 2 * TC_KC, nd = (size_t)TC_C1 * LD + 2 * TC_KC;
  if (cudaMalloc(&s.Wfh, nf * 2) || cudaMalloc(&s.Wfl, nf * 2) || cudaMalloc(&s.Wdh, nd * 2) || cudaMalloc(&s.Wdl, nd * 2)) {
    snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of staged weights failed"); return -1;
  }
  cudaMemsetAsync(s.Wfh, 0, nf * 2, st); cudaMemsetAsync(s.Wfl, 0, nf * 2, st);
  cudaMemsetAsync(s.Wdh, 0, nd * 2, st); cudaMemsetAsync(s.Wdl, 0, nd * 2, st);
  // window views: dims (x, jr, ch), strides (CS*2, L*2) bytes
  {
    cuuint64_t dims[3] = {(cuuint64_t)N * TC_CP, (cuuint64_t)N, (cuuint64_t)TC_C2};
    cuuint64_t str[2] = {(cuuint64_t)TC_CP * 2, (cuuint64_t)LF * 2};
    cuuint32_t box[3] = {TC_KC, 12, 20};
    if (tc_encode(s, &s.fB_h, s.Wfh, 3, dims, str, box) || tc_encode(s, &s.fB_l, s.Wfl, 3, dims, str, box)) return -1;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)N * TC_OP, (cuuint64_t)N, (cuuint64_t)TC_C1};
    cuuint64_t str[2] = {(cuuint64_t)TC_OP * 2, (cuuint64_t)LD * 2};
    cuuint32_t box[3] = {TC_KC, 24, 10};
    if (tc_encode(s, &s.dB_h, s.Wdh, 3, dims, str, box) || tc_encode(s, &s.dB_l, s.Wdl, 3, dims, str, box)) return -1;
  }
  cudaFuncSetAttribute(toep_gemm_k<240>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_STAGES * (2 * TC_BM * TC_KC * 2 + 2 * 240 * TC_KC * 2) + 1024);
  cudaFuncSetAttribute(wgrad_gemm_k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_STAGES * (2 * TC_KC * TC_BM * 2 + 2 * TC_KC * TC_WN * 2) + 1024);
  s.ready = 1;
  return 0;
}

static void tc_destroy(TcState& s) {
  if (s.Wfh) cudaFree(s.Wfh); if (s.Wfl) cudaFree(s.Wfl); if (s.Wdh) cudaFree(s.Wdh); if (s.Wdl) cudaFree(s.Wdl);
  s.Wfh = s.Wfl = s.Wdh = s.Wdl = nullptr; s.ready = 0;
}

// (re)build the tensor maps over the activation planes
static int tc_bind_planes(TcState& s, const __nv_bfloat16* yh, const __nv_bfloat16* yl, const __nv_bfloat16* dh, const __nv_bfloat16* dl) {
  const int N = s.N;
  if (yh && (yh != s.yh || yl != s.yl)) {
    cuuint64_t dims[2] = {(cuuint64_t)N * TC_CP, (cuuint64_t)s.max_rows};
    cuuint64_t str[1] = {(cuuint64_t)N * TC_CP * 2};
    cuuint32_t boxk[2] = {TC_KC, TC_BM};
    if (tc_encode(s, &s.fA_h, yh, 2, dims, str, boxk) || tc_encode(s, &s.fA_l, yl, 2, dims, str, boxk)) return -1;
    s.yh = yh; s.yl = yl;
  }
  if (dh && (dh != s.dh || dl != s.dl)) {
    cuuint64_t dims[2] = {(cuuint64_t)N * TC_OP, (cuuint64_t)s.max_rows};
    cuuint64_t str[1] = {(cuuint64_t)N * TC_OP * 2};
    cuuint32_t boxk[2] = {TC_KC, TC_BM};
    if (tc_encode(s, &s.dA_h, dh, 2, dims, str, boxk) || tc_encode(s, &s.dA_l, dl, 2, dims, str, boxk)) return -1;
    s.dh = dh; s.dl = dl;
  }
  return 0;
}
// wgrad reduces over the rows, so its maps are bounded by the exact row count (TMA zero-fills beyond)
static int tc_bind_wgrad(TcState& s, const __nv_bfloat16* yh, const __nv_bfloat16* yl, const __nv_bfloat16* dh, const __nv_bfloat16* dl,
                         long long rows) {
  const int N = s.N;
  if (yh == s.wyh && dh == s.wdh && rows == s.wrows) return 0;
  cuuint32_t boxm[2] = {64, TC_KC};
  { cuuint64_t dims[2] = {(cuuint64_t)N * TC_CP, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)N * TC_CP * 2};
    if (tc_encode(s, &s.wA_h, yh, 2, dims, str, boxm) || tc_encode(s, &s.wA_l, yl, 2, dims, str, boxm)) return -1; }
  { cuuint64_t dims[2] = {(cuuint64_t)N * TC_OP, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)N * TC_OP * 2};
    if (tc_encode(s, &s.wB_h, dh, 2, dims, str, boxm) || tc_encode(s, &s.wB_l, dl, 2, dims, str, boxm)) return -1; }
  s.wyh = yh; s.wdh = dh; s.wrows = rows;
  return 0;
}

static int tc_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "%s launch: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}

static int tc_prepare_weights(TcState& s, const float* w1, int N, cudaStream_t st) {
  if (!s.ready) { snprintf(g_tc_err, sizeof g_tc_err, "tc state not initialised"); return -1; }
  long long total = (long long)TC_C2 * (2 * N - 1) * TC_CP + (long long)TC_C1 * (2 * N - 1) * TC_OP;
  tc_stage_weights_k<<<cdiv(total, 256), 256, 0, st>>>(w1, s.Wfh, s.Wfl, s.Wdh, s.Wdl, N);
  return tc_check_launch("tc_stage_weights_k");
}

static int tc_fwd(TcState& s, const __nv_bfloat16* Yhi, const __nv_bfloat16* Ylo, float* O12, long long rows, int N, cudaStream_t st) {
  if (tc_bind_planes(s, Yhi, Ylo, nullptr, nullptr)) return -1;
  ToepArgs a; a.out = O12; a.rows = rows; a.N = N; a.Cout = TC_C2; a.CS = TC_CP; a.CT = 20; a.JT = 12; a.n_ctiles = 1;
  a.n_jtiles = (N + 11) / 12; a.pad_rows = N - 1 - (N - 1) / 2; a.KA = N * TC_CP;
  unsigned grid = (unsigned)(((rows + TC_BM - 1) / TC_BM) * a.n_ctiles * a.n_jtiles);
  size_t smem = TC_STAGES * (2 * TC_BM * TC_KC * 2 + 2 * 240 * TC_KC * 2) + 1024;
  toep_gemm_k<240><<<grid, 192, smem, st>>>(s.fA_h, s.fA_l, s.fB_h, s.fB_l, a);
  return tc_check_launch("toep_gemm_k(fwd)");
}

static int tc_dgrad(TcState& s, const __nv_bfloat16* dOhi, const __nv_bfloat16* dOlo, float* dY12, long long rows, int N, cudaStream_t st) {
  if (tc_bind_planes(s, nullptr, nullptr, dOhi, dOlo)) return -1;
  ToepArgs a; a.out = dY12; a.rows = rows; a.N = N; a.Cout = TC_C1; a.CS = TC_OP; a.CT = 10; a.JT = 24; a.n_ctiles = 5;
  a.n_jtiles = (N + 23) / 24; a.pad_rows = (N - 1) / 2; a.KA = N * TC_OP;
  unsigned grid = (unsigned)(((rows + TC_BM - 1) / TC_BM) * a.n_ctiles * a.n_jtiles);
  size_t smem = TC_STAGES * (2 * TC_BM * TC_KC * 2 + 2 * 240 * TC_KC * 2) + 1024;
  toep_gemm_k<240><<<grid, 192, smem, st>>>(s.dA_h, s.dA_l, s.dB_h, s.dB_l, a);
  return tc_check_launch("toep_gemm_k(dgrad)");
}

static int tc_wgrad(TcState& s, const __nv_bfloat16* Yhi, const __nv_bfloat16* Ylo, const __nv_bfloat16* dOhi, const __nv_bfloat16* dOlo,
                    float* gw1, long long rows, int N, cudaStream_t st) {
  if (tc_bind_wgrad(s, Yhi, Ylo, dOhi, dOlo, rows)) return -1;
  WgradArgs a; a.dw1 = gw1; a.rows = rows; a.N = N; a.n_ntiles = (N * TC_OP + TC_WN - 1) / TC_WN;
  int n_mtiles = (N * TC_CP + TC_BM - 1) / TC_BM;
  long long tiles = (long long)n_mtiles * a.n_ntiles;
  long long kchunks = (rows + TC_KC - 1) / TC_KC;
  int ks = 1;
  while (tiles * ks < 2 * 148 && ks * 8 < kchunks) ks *= 2;     // fill the 148 SMs when N is small
  a.ksplit = ks;
  size_t smem = TC_STAGES * (2 * TC_KC * TC_BM * 2 + 2 * TC_KC * TC_WN * 2) + 1024;
  wgrad_gemm_k<<<(unsigned)(tiles * ks), 192, smem, st>>>(s.wA_h, s.wA_l, s.wB_h, s.wB_l, a);
  return tc_check_launch("wgrad_gemm_k");
}
