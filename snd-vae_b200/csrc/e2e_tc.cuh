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
// Kernels: one CTA (320 threads) per 128 x BN output tile; warp 0 = TMA producer, warp 1 =
// TMEM allocator + single-thread tcgen05.mma issuer, warps 2..9 = accumulate/epilogue (tcgen05.ld).
// smem ring: 2 stages of K chunk 64 (SWIZZLE_128B rows) for the K-major kernel, 4 stages of 32 K-rows for the
// MN-major one; mbarrier full/empty pipeline.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_runtime.h>

#define TC_C1 50      /* e2e layer-1 input channels  */
#define TC_C2 20      /* e2e layer-1 output channels */
#define TC_CP 56      /* padded channel stride of Y planes  (112 B: TMA strides need 16 B multiples); compact = 50 */
#define TC_OP 24      /* padded channel stride of dO planes (48 B); compact = 20 */
#define TC_BM 128
#define TC_KC 64      /* K elements per stage of the K-major (fwd / dgrad) kernel = one 128-byte swizzle row.  (32 with
                         SWIZZLE_64B and 4 stages also works but measured 20 % slower: fwd 66 -> 82 ms per 512 graphs.) */
#define TC_STAGES 2   /* 92 KB per stage */
#define TC_WKC 32     /* K rows per stage of the MN-major (wgrad) kernel */
#define TC_WSTAGES 4

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
// The same two instructions for a warp that runs its issue loop CONVERGED (all 32 lanes, warp-uniform operands): the election sits
// inside the asm statement, so the compiler keeps descriptors, TMEM addresses and barrier addresses in uniform registers and emits one
// predicated UTCHMMA -- under `if (lane == 0)` it has to move every operand into uniform registers (R2UR) and wrap each instruction in an
// ELECT / BRA.U.ANY loop.  Measured (tools/mma_probe.cu): 73 instead of 83 cycles per back-to-back MMA with nothing else in the loop; the
// scalar work between the MMAs is what the short-N products of this library are bound by (~180 cycles per MMA before).
__device__ __forceinline__ void umma_bf16_conv(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_conv(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same loads without the wait: issue several, then one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return umma_desc(saddr, lbo_bytes, sbo_bytes, 2ull);     // SWIZZLE_128B
}
// K-major operand whose rows are TC_KC bf16 wide: 128-byte rows -> SWIZZLE_128B (atoms of 8 rows = 1024 B),
// 64-byte rows -> SWIZZLE_64B (layout 4, atoms of 512 B)
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
  return TC_KC == 64 ? umma_desc(saddr, 16, 1024, 2ull) : umma_desc(saddr, 16, 512, 4ull);
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
  int accumulate;      // out += result instead of out = result
  int nsh;             // shifted weight copies: tile row n = (r, ch, u), position jr = jr0 + u*nsh + r
};
struct TMapSet { CUtensorMap h[4]; CUtensorMap l[4]; };   // window views of the nsh shifted copies (hi / lo)

// Two-level accumulation.  tcgen05.mma adds into its fp32 TMEM accumulator with truncation
// (measured here: a systematic shrink of ~1.7e-8 per accumulate, -4.5e-5 over the 2688 accumulates
// of one N=256 output), so the K loop is cut into groups of TC_GROUP chunks: each group accumulates
// into one of two TMEM slots, and the epilogue warps drain the finished slot into round-to-nearest
// fp32 registers while the tensor pipe fills the other slot.
#ifndef TC_GROUP
#define TC_GROUP (512 / TC_KC)   /* K chunks per accumulation group: 512 K elements = 96 accumulates */
#endif
#define TC_EPI_WARPS 8
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)
template <int BN, int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS, 1) toep_gemm_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                             const __grid_constant__ TMapSet tmB, ToepArgs P) {
  constexpr int A_BYTES = TC_BM * TC_KC * 2;       // 16 KB
  constexpr int B_BYTES = BN * TC_KC * 2;          // 30 KB at BN = 240
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 512;              // two accumulator slots of 256 columns
  constexpr int HC = BN / 2;                       // columns per epilogue warp (two warps per lane quarter)
  static_assert(BN <= 256 && HC % 8 == 0, "tile shape");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], acc_full[2], acc_empty[2];
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
  const int ngroups = (nk + TC_GROUP - 1) / TC_GROUP;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TC_EPI_WARPS); }
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
        const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int x = (kc_lo + it) * TC_KC;
        tma_load_2d(st, &tmAh, &full_bar[s], x, m0);
        tma_load_2d(st + A_BYTES, &tmAl, &full_bar[s], x, m0);
        // B tile rows = (shift r, channel, u): one sub-box per shifted copy, 128-byte rows, contiguous
        const int sub = B_BYTES / P.nsh, u0 = jr0 / P.nsh;
        for (int r = 0; r < P.nsh; ++r) {
          tma_load_3d(st + 2 * A_BYTES + r * sub, &tmB.h[r], &full_bar[s], x, u0, ch0);
          tma_load_3d(st + 2 * A_BYTES + B_BYTES + r * sub, &tmB.l[r], &full_bar[s], x, u0, ch0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(TC_BM, BN, 0, 0);
      for (int it = 0; it < nk; ++it) {
        const int g = it / TC_GROUP, gi = it - g * TC_GROUP, slot = g & 1;
        if (gi == 0) { mbar_wait(&acc_empty[slot], ((g >> 1) & 1) ^ 1); tc_fence_after(); }   // slot drained by the epilogue
        const int s = it % NSTAGE; const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t td = tmem_d + slot * 256;
#pragma unroll
        for (int k = 0; k < TC_KC / 16; ++k) {
          // K-major swizzled rows: 8-row groups one atom apart; +32 B per 16-element K step
          const uint64_t ah = umma_desc_kmajor(sa + k * 32);
          const uint64_t al = umma_desc_kmajor(sa + A_BYTES + k * 32);
          const uint64_t bh = umma_desc_kmajor(sa + 2 * A_BYTES + k * 32);
          const uint64_t bl = umma_desc_kmajor(sa + 2 * A_BYTES + B_BYTES + k * 32);
          umma_bf16(td, ah, bh, idesc, (gi | k) ? 1u : 0u);
          umma_bf16(td, ah, bl, idesc, 1u);
          umma_bf16(td, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);          // smem slot free once these MMAs retire
        if (gi == TC_GROUP - 1 || it == nk - 1) umma_commit(&acc_full[slot]);   // group complete
      }
    }
  } else {
    // epilogue warps: warp w may touch TMEM lanes [32*(w%4), +32); two warps share a quarter, HC columns each
    const int q = warp & 3, half = (warp - 2) >> 2;
    float acc[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) acc[i] = 0.f;
    for (int g = 0; g < ngroups; ++g) {
      const int slot = g & 1;
      mbar_wait(&acc_full[slot], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_d + slot * 256 + ((uint32_t)(q * 32) << 16) + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(ta + c0, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c0 + i] += __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
    }
    const long long row = (long long)m0 + q * 32 + lane;
    if (row < P.rows) {
      float* orow = P.out + row * (long long)P.N * P.Cout;
#pragma unroll
      for (int i = 0; i < HC; ++i) {
        const int n = half * HC + i;
        const int jts = P.JT / P.nsh, per = P.CT * jts;
        const int r = n / per, w = n - r * per;
        const int chl = w / jts, u = w - chl * jts;
        const int jr = jr0 + u * P.nsh + r;
        const int j = P.N - 1 - jr;
        if (j >= 0 && ch0 + chl < P.Cout) {
          float* dst = orow + (long long)j * P.Cout + ch0 + chl;
          *dst = P.accumulate ? *dst + acc[i] : acc[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster share one 256 x BN tile.  Each CTA loads its own
// 128 rows of A and HALF of the B tile; the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads
// the B halves from both CTAs' shared memory, and each CTA's TMEM receives its own 128 rows.  Per CTA this
// halves the B fill traffic and shrinks a stage to 62 KB (3-stage ring), relieving the shared-memory
// bandwidth that bounds the single-CTA kernel (SS-mode MMA reads + TMA fills > 128 B/cycle/SM).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

#define TC2_STAGES 3
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
toep_gemm2_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ TMapSet tmB, ToepArgs P) {
  constexpr int A_BYTES = TC_BM * TC_KC * 2;       // 16 KB: this CTA's 128 rows
  constexpr int BH_BYTES = (BN / 2) * TC_KC * 2;   // 15 KB: this CTA's half of the B tile
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int HC = BN / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC2_STAGES], empty_bar[TC2_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cl = blockIdx.x >> 1;
  const int ntile = cl % (P.n_ctiles * P.n_jtiles);
  const int mpair = cl / (P.n_ctiles * P.n_jtiles);
  const int ct = ntile % P.n_ctiles, jt = ntile / P.n_ctiles;
  const int ch0 = ct * P.CT, jr0 = jt * P.JT;
  const int m0 = mpair * 2 * TC_BM + (int)rank * TC_BM;
  long long xlo = (long long)(P.pad_rows - jr0 - P.JT + 1) * P.CS - (TC_KC - 1);
  long long xhi = (long long)(P.pad_rows + P.N - jr0) * P.CS;
  if (xlo < 0) xlo = 0;
  if (xhi > P.KA) xhi = P.KA;
  const int kc_lo = (int)(xlo / TC_KC);
  const int kc_hi = (int)((xhi + TC_KC - 1) / TC_KC);
  const int nk = kc_hi > kc_lo ? kc_hi - kc_lo : 0;
  const int ngroups = (nk + TC_GROUP - 1) / TC_GROUP;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // this CTA's rows of the B tile: sub-boxes [rank * nsh/2, (rank+1) * nsh/2) of the nsh shifted copies
      const int sub = (BN * TC_KC * 2) / P.nsh, u0 = jr0 / P.nsh, r_lo = (int)rank * (P.nsh / 2);
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC2_STAGES; const uint32_t ph = (it / TC2_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);      // the leader's full barrier
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);        // bytes of both CTAs
        const int x = (kc_lo + it) * TC_KC;
        tma_load_2d_2sm(st, &tmAh, lead_full, x, m0);
        tma_load_2d_2sm(st + A_BYTES, &tmAl, lead_full, x, m0);
        for (int r = 0; r < P.nsh / 2; ++r) {
          tma_load_3d_2sm(st + 2 * A_BYTES + r * sub, &tmB.h[r_lo + r], lead_full, x, u0, ch0);
          tma_load_3d_2sm(st + 2 * A_BYTES + BH_BYTES + r * sub, &tmB.l[r_lo + r], lead_full, x, u0, ch0);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc(2 * TC_BM, BN, 0, 0);
      for (int it = 0; it < nk; ++it) {
        const int g = it / TC_GROUP, gi = it - g * TC_GROUP, slot = g & 1;
        if (gi == 0) { mbar_wait(&acc_empty[slot], ((g >> 1) & 1) ^ 1); tc_fence_after(); }   // drained by both CTAs
        const int s = it % TC2_STAGES; const uint32_t ph = (it / TC2_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t td = tmem_d + slot * 256;
#pragma unroll
        for (int k = 0; k < TC_KC / 16; ++k) {
          const uint64_t ah = umma_desc_kmajor(sa + k * 32);
          const uint64_t al = umma_desc_kmajor(sa + A_BYTES + k * 32);
          const uint64_t bh = umma_desc_kmajor(sa + 2 * A_BYTES + k * 32);
          const uint64_t bl = umma_desc_kmajor(sa + 2 * A_BYTES + BH_BYTES + k * 32);
          umma_bf16_2sm(td, ah, bh, idesc, (gi | k) ? 1u : 0u);
          umma_bf16_2sm(td, ah, bl, idesc, 1u);
          umma_bf16_2sm(td, al, bh, idesc, 1u);
        }
        umma_commit_2sm(&empty_bar[s]);                                          // frees the stage in both CTAs
        if (gi == TC_GROUP - 1 || it == nk - 1) umma_commit_2sm(&acc_full[slot]); // group complete, both CTAs
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    float acc[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) acc[i] = 0.f;
    for (int g = 0; g < ngroups; ++g) {
      const int slot = g & 1;
      mbar_wait(&acc_full[slot], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_d + slot * 256 + ((uint32_t)(q * 32) << 16) + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(ta + c0, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c0 + i] += __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[slot]), 0));   // the leader collects 16 arrivals
    }
    const long long row = (long long)m0 + q * 32 + lane;
    if (row < P.rows) {
      float* orow = P.out + row * (long long)P.N * P.Cout;
#pragma unroll
      for (int i = 0; i < HC; ++i) {
        const int n = half * HC + i;
        const int jts = P.JT / P.nsh, per = P.CT * jts;
        const int r = n / per, w = n - r * per;
        const int chl = w / jts, u = w - chl * jts;
        const int jr = jr0 + u * P.nsh + r;
        const int j = P.N - 1 - jr;
        if (j >= 0 && ch0 + chl < P.Cout) {
          float* dst = orow + (long long)j * P.Cout + ch0 + chl;
          *dst = P.accumulate ? *dst + acc[i] : acc[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no CTA of the pair exits while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: MN-major GEMM  dT[m, n] = sum_rows Y[row, m] dO[row, n]; the epilogue folds the
// block diagonals:  dw1[t, c, q] += dT[(j',c),(j,q)],  t = j'-j+p  
// ------------------------------------------------------------------------------------------
struct WgradArgs {
  float* dw;           // [N, Ctot, Cout] gradient slot (atomicAdd); this product owns channels [coff, coff+Cin)
  long long rows;
  int N;
  int CSi, Cin;        // channel stride / count of the M-side planes
  int CSo, Cout;       // channel stride / count of the N-side planes
  int Ctot, coff;
  int n_ntiles;        // ceil(N*CSo / 256)
  int ksplit;          // CTAs per output tile along K
  int plain;           // 1: no Toeplitz folding: dw[(m / CSi) * Cin + m % CSi][n] += D[m, n] for n < Cout (row stride Cout)
};
#define TC_WN 256
#define TC_WGROUP (512 / TC_WKC)
__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_gemm_k(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                                                              const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                                                              WgradArgs P) {
  constexpr int A_BYTES = TC_WKC * TC_BM * 2;      // [2 x (TC_WKC k-rows x 128 B)]
  constexpr int B_BYTES = TC_WKC * TC_WN * 2;      // [4 x (TC_WKC k-rows x 128 B)]
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 48 KB at TC_WKC = 32
  constexpr int BLK = TC_WKC * 128;                // one 64(mn) x TC_WKC(k) box
  constexpr int HC = TC_WN / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC_WSTAGES], empty_bar[TC_WSTAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / P.ksplit, ks = blockIdx.x % P.ksplit;
  const int nt = tile % P.n_ntiles, mt = tile / P.n_ntiles;
  const int m0 = mt * TC_BM, n0 = nt * TC_WN;
  const int p = (P.N - 1) / 2;
  // band test: does any (j', j) in this tile have 0 <= j'-j+p < N ?
  const int jp_lo = m0 / P.CSi, jp_hi = min(P.N - 1, (m0 + TC_BM - 1) / P.CSi);
  const int j_lo = n0 / P.CSo, j_hi = min(P.N - 1, (n0 + TC_WN - 1) / P.CSo);
  if (!P.plain) {
    if (jp_lo > P.N - 1 || j_lo > P.N - 1) return;
    if (jp_hi - j_lo + p < 0 || jp_lo - j_hi + p > P.N - 1) return;
  }
  const long long kchunks = (P.rows + TC_WKC - 1) / TC_WKC;
  const long long per = (kchunks + P.ksplit - 1) / P.ksplit;
  const long long kc_lo = (long long)ks * per;
  long long kc_hi = kc_lo + per; if (kc_hi > kchunks) kc_hi = kchunks;
  const int nk = kc_hi > kc_lo ? (int)(kc_hi - kc_lo) : 0;
  if (nk == 0) return;
  const int ngroups = (nk + TC_WGROUP - 1) / TC_WGROUP;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_WSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % TC_WSTAGES; const uint32_t ph = (it / TC_WSTAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int r0 = (int)((kc_lo + it) * TC_WKC);
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
        const int g = it / TC_WGROUP, gi = it - g * TC_WGROUP, slot = g & 1;
        if (gi == 0) { mbar_wait(&acc_empty[slot], ((g >> 1) & 1) ^ 1); tc_fence_after(); }
        const int s = it % TC_WSTAGES; const uint32_t ph = (it / TC_WSTAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t td = tmem_d + slot * 256;
#pragma unroll
        for (int k = 0; k < TC_WKC / 16; ++k) {
          // MN-major SWIZZLE_128B: 64-element MN blocks BLK bytes apart (LBO), 8-row K groups
          // 1024 B apart (SBO); +2048 B per 16-row K step
          const uint64_t ah = umma_desc_sw128(sa + k * 2048, BLK, 1024);
          const uint64_t al = umma_desc_sw128(sa + A_BYTES + k * 2048, BLK, 1024);
          const uint64_t bh = umma_desc_sw128(sa + 2 * A_BYTES + k * 2048, BLK, 1024);
          const uint64_t bl = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES + k * 2048, BLK, 1024);
          umma_bf16(td, ah, bh, idesc, (gi | k) ? 1u : 0u);
          umma_bf16(td, ah, bl, idesc, 1u);
          umma_bf16(td, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
        if (gi == TC_WGROUP - 1 || it == nk - 1) umma_commit(&acc_full[slot]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    float acc[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) acc[i] = 0.f;
    for (int g = 0; g < ngroups; ++g) {
      const int slot = g & 1;
      mbar_wait(&acc_full[slot], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t ta = tmem_d + slot * 256 + ((uint32_t)(q * 32) << 16) + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(ta + c0, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c0 + i] += __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
    }
    const int m = m0 + q * 32 + lane;
    const int jp = m / P.CSi, c = m - jp * P.CSi;
    if (P.plain) {
      if (jp < P.N && c < P.Cin) {
#pragma unroll
        for (int i = 0; i < HC; ++i) {
          const int n = n0 + half * HC + i;
          if (n < P.Cout) atomicAdd(P.dw + ((size_t)jp * P.Cin + c) * P.Cout + n, acc[i]);
        }
      }
    } else if (jp < P.N && c < P.Cin) {
#pragma unroll
      for (int i = 0; i < HC; ++i) {
        const int n = n0 + half * HC + i;
        const int j = n / P.CSo, qq = n - j * P.CSo;
        const int t = jp - j + p;
        if (j < P.N && qq < P.Cout && t >= 0 && t < P.N)
          atomicAdd(P.dw + ((size_t)t * P.Ctot + P.coff + c) * P.Cout + qq, acc[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

// ------------------------------------------------------------------------------------------
// per-step weight staging: zero-padded, channel-major, bf16 hi/lo copies of w[N, Ctot, Cout]
// (channels [coff, coff+Cin) of the input side):
//   Wf[q][rho*CSi + c] = w[rho - (N-1-p), coff + c, q]      rho in [0, 2N-1)   (forward B operand)
//   Wd[c][rho*CSo + q] = w[N-1+p - rho,  coff + c, q]                           (dgrad  B operand)
// ------------------------------------------------------------------------------------------
__global__ void tc_stage_weights_k(const float* __restrict__ w, __nv_bfloat16* __restrict__ Wfh, __nv_bfloat16* __restrict__ Wfl,
                                   __nv_bfloat16* __restrict__ Wdh, __nv_bfloat16* __restrict__ Wdl, int N, int Ctot, int coff,
                                   int Cin, int Cout, int CSi, int CSo, int nsf, int nsd, long long LFp, long long LDp) {
  // copy r of the forward weights is shifted by r positions: Wf_r[q][y] = Wf[q][y + r*CSi] (16-byte aligned window starts)
  const int p = (N - 1) / 2;
  const long long nf = (long long)nsf * Cout * LFp, nd = (long long)nsd * Cin * LDp;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nf) {
    int r = (int)(idx / (Cout * LFp)); long long rem = idx - (long long)r * Cout * LFp;
    int q = (int)(rem / LFp); long long x = rem - q * LFp + (long long)r * CSi;
    int rho = (int)(x / CSi), c = (int)(x - (long long)rho * CSi);
    int t = rho - (N - 1 - p);
    float v = (c < Cin && t >= 0 && t < N) ? w[((size_t)t * Ctot + coff + c) * Cout + q] : 0.f;
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Wfh[idx] = hi; Wfl[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  } else if (idx < nf + nd) {
    idx -= nf;
    int r = (int)(idx / (Cin * LDp)); long long rem = idx - (long long)r * Cin * LDp;
    int c = (int)(rem / LDp); long long x = rem - c * LDp + (long long)r * CSo;
    int rho = (int)(x / CSo), q = (int)(x - (long long)rho * CSo);
    int t = N - 1 + p - rho;
    float v = (q < Cout && t >= 0 && t < N) ? w[((size_t)t * Ctot + coff + c) * Cout + q] : 0.f;
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Wdh[idx] = hi; Wdl[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// fp32 rows [rows, C] -> bf16 hi / lo planes [rows, CS] (pad channels stay zero from allocation)
__global__ void tc_split_planes_k(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  long long rows, int C, int CS) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  long long r = idx / C; int c = (int)(idx - r * C);
  float v = src[idx];
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[r * CS + c] = h; lo[r * CS + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------
// host side: one ToepPlan per weight block  w[N, Ctot, Cout] restricted to Cin input channels
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_tc_encode = nullptr;
static size_t tc_toep2_smem(int bn) { return (size_t)TC2_STAGES * (2 * TC_BM * TC_KC * 2 + 2 * (bn / 2) * TC_KC * 2) + 1024; }
static int g_tc_2cta = -1;
static size_t tc_toep_smem(int bn, int stages) { return (size_t)stages * (2 * TC_BM * TC_KC * 2 + 2 * bn * TC_KC * 2) + 1024; }

struct ToepPlan {
  int ready, N, Cin, Cout, CSi, CSo;
  int fCT, fJT, dCT, dJT;                  // tile shapes (channels x positions); product = BN in {240, 160}
  int nsf, nsd;                            // shifted weight copies (window starts must be 16-byte aligned)
  long long LFp, LDp;                      // padded row lengths of the staged weights
  __nv_bfloat16 *Wfh, *Wfl, *Wdh, *Wdl;    // staged weights [nsh][channels][Lp]
  TMapSet fB, dB;                          // window views of the staged weights
};
struct TcState {                           // all tensor-core products of the edge decoder
  int ready;
  ToepPlan l1;                             // e2e layer 1:  50 -> 20
  ToepPlan l0a, l0c;                       // e2e layer 0 vector terms: a-half / c-half of w0 (Ch -> 50)
  __nv_bfloat16 *ah, *al, *ch, *cl;        // a / c planes      [B, N*CSi]
  __nv_bfloat16 *dsh, *dsl, *drh, *drl;    // dSa / dRc planes  [B, N*CSo]
};

static int tc_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b, const cuuint32_t* box,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_tc_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled failed: %d (rank %d)", (int)r, rank); return -1; }
  return 0;
}
#define TC_KSWZ (TC_KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)
static int tc_encode_rows(CUtensorMap* tm, const void* base, long long width, long long rows, int box0, int box1,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)width * 2};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  return tc_encode(tm, base, 2, dims, str, box, swz);
}
static int tc_global_init() {
  if (g_tc_encode) return 0;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled entry point not found"); return -1;
  }
  g_tc_encode = (PFN_encodeTiled)fn;
  cudaFuncSetAttribute(toep_gemm_k<240, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_toep_smem(240, 2));
  cudaFuncSetAttribute(toep_gemm_k<160, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_toep_smem(160, 3));
  cudaFuncSetAttribute(toep_gemm2_k<240>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_toep2_smem(240));
  cudaFuncSetAttribute(toep_gemm_k<48, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_toep_smem(48, 4));
  cudaFuncSetAttribute(wgrad_gemm_k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_WSTAGES * (2 * TC_WKC * TC_BM * 2 + 2 * TC_WKC * TC_WN * 2) + 1024);
  return 0;
}
static int tc_pad16(int c) { return (c * 2) % 16 == 0 ? c : (c + 7) / 8 * 8; }   // channel stride with 16-byte rows
static int tc_gcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// channel stride of a plane with C channels: compact (C) when the rows stay 16-byte aligned and the tile's position
// count is a multiple of the number of shifted weight copies that compact windows need; padded to 8 otherwise
static int tc_stride(int N, int C, int JT, int compact) {
  if (compact && ((long long)N * C * 2) % 16 == 0) { int nsh = 16 / tc_gcd(16, C * 2); if (nsh <= 4 && JT % nsh == 0) return C; }
  return tc_pad16(C);
}

static int tc_plan_init(ToepPlan& pl, int N, int Cin, int Cout, int fCT, int fJT, int dCT, int dJT, int compact, cudaStream_t st) {
  memset(&pl, 0, sizeof pl);
  pl.N = N; pl.Cin = Cin; pl.Cout = Cout;
  pl.CSi = tc_stride(N, Cin, fJT, compact); pl.CSo = tc_stride(N, Cout, dJT, compact);
  pl.nsf = 16 / tc_gcd(16, pl.CSi * 2); pl.nsd = 16 / tc_gcd(16, pl.CSo * 2);
  pl.fCT = fCT; pl.fJT = fJT; pl.dCT = dCT; pl.dJT = dJT;
  if ((fCT * fJT != 240 && fCT * fJT != 160) || (dCT * dJT != 240 && dCT * dJT != 160) || Cout % fCT || Cin % dCT || fJT % pl.nsf || dJT % pl.nsd) { snprintf(g_tc_err, sizeof g_tc_err, "bad tile shape"); return -1; }
  pl.LFp = ((long long)(2 * N - 1 + 4) * pl.CSi + 7) / 8 * 8; pl.LDp = ((long long)(2 * N - 1 + 4) * pl.CSo + 7) / 8 * 8;
  size_t nf = (size_t)pl.nsf * Cout * pl.LFp + 2 * TC_KC, nd = (size_t)pl.nsd * Cin * pl.LDp + 2 * TC_KC;
  if (cudaMalloc(&pl.Wfh, nf * 2) || cudaMalloc(&pl.Wfl, nf * 2) || cudaMalloc(&pl.Wdh, nd * 2) || cudaMalloc(&pl.Wdl, nd * 2)) {
    snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of staged weights failed"); return -1;
  }
  cudaMemsetAsync(pl.Wfh, 0, nf * 2, st); cudaMemsetAsync(pl.Wfl, 0, nf * 2, st);
  cudaMemsetAsync(pl.Wdh, 0, nd * 2, st); cudaMemsetAsync(pl.Wdl, 0, nd * 2, st);
  // window views: dims (x, u, ch) with position jr = u*nsh + r; strides (nsh*CS*2, Lp*2) bytes -- the position stride is
  // smaller than the x extent, so consecutive tile rows are overlapping windows of one weight row
  for (int r = 0; r < pl.nsf; ++r) {
    cuuint64_t dims[3] = {(cuuint64_t)N * pl.CSi, (cuuint64_t)((N - r + pl.nsf - 1) / pl.nsf), (cuuint64_t)Cout};
    cuuint64_t str[2] = {(cuuint64_t)pl.nsf * pl.CSi * 2, (cuuint64_t)pl.LFp * 2};
    cuuint32_t box[3] = {TC_KC, (cuuint32_t)(fJT / pl.nsf), (cuuint32_t)fCT};
    if (tc_encode(&pl.fB.h[r], pl.Wfh + (size_t)r * Cout * pl.LFp, 3, dims, str, box, TC_KSWZ) ||
        tc_encode(&pl.fB.l[r], pl.Wfl + (size_t)r * Cout * pl.LFp, 3, dims, str, box, TC_KSWZ)) return -1;
  }
  for (int r = 0; r < pl.nsd; ++r) {
    cuuint64_t dims[3] = {(cuuint64_t)N * pl.CSo, (cuuint64_t)((N - r + pl.nsd - 1) / pl.nsd), (cuuint64_t)Cin};
    cuuint64_t str[2] = {(cuuint64_t)pl.nsd * pl.CSo * 2, (cuuint64_t)pl.LDp * 2};
    cuuint32_t box[3] = {TC_KC, (cuuint32_t)(dJT / pl.nsd), (cuuint32_t)dCT};
    if (tc_encode(&pl.dB.h[r], pl.Wdh + (size_t)r * Cin * pl.LDp, 3, dims, str, box, TC_KSWZ) ||
        tc_encode(&pl.dB.l[r], pl.Wdl + (size_t)r * Cin * pl.LDp, 3, dims, str, box, TC_KSWZ)) return -1;
  }
  pl.ready = 1;
  return 0;
}
static void tc_plan_destroy(ToepPlan& pl) {
  if (pl.Wfh) cudaFree(pl.Wfh); if (pl.Wfl) cudaFree(pl.Wfl); if (pl.Wdh) cudaFree(pl.Wdh); if (pl.Wdl) cudaFree(pl.Wdl);
  memset(&pl, 0, sizeof pl);
}
static int tc_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "%s launch: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}
static int tc_plan_stage(ToepPlan& pl, const float* w, int Ctot, int coff, cudaStream_t st) {
  if (!pl.ready) { snprintf(g_tc_err, sizeof g_tc_err, "plan not initialised"); return -1; }
  long long total = (long long)pl.nsf * pl.Cout * pl.LFp + (long long)pl.nsd * pl.Cin * pl.LDp;
  tc_stage_weights_k<<<cdiv(total, 256), 256, 0, st>>>(w, pl.Wfh, pl.Wfl, pl.Wdh, pl.Wdl, pl.N, Ctot, coff, pl.Cin, pl.Cout, pl.CSi, pl.CSo,
                                                      pl.nsf, pl.nsd, pl.LFp, pl.LDp);
  return tc_check_launch("tc_stage_weights_k");
}
static int tc_split(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long rows, int C, int CS, cudaStream_t st) {
  tc_split_planes_k<<<cdiv(rows * C, 256), 256, 0, st>>>(src, hi, lo, rows, C, CS);
  return tc_check_launch("tc_split_planes_k");
}
static const size_t TC_WGRAD_SMEM = TC_WSTAGES * (2 * TC_WKC * TC_BM * 2 + 2 * TC_WKC * TC_WN * 2) + 1024;

// out[rows, N*Cout] (+)= in[rows, N*CSi planes] . Toeplitz(w)
static int tc_plan_fwd(ToepPlan& pl, const __nv_bfloat16* inh, const __nv_bfloat16* inl, float* out, long long rows, long long rows_alloc,
                       int accumulate, cudaStream_t st) {
  CUtensorMap ah, al;
  if (tc_encode_rows(&ah, inh, (long long)pl.N * pl.CSi, rows_alloc, TC_KC, TC_BM, TC_KSWZ) || tc_encode_rows(&al, inl, (long long)pl.N * pl.CSi, rows_alloc, TC_KC, TC_BM, TC_KSWZ)) return -1;
  ToepArgs a; a.out = out; a.rows = rows; a.N = pl.N; a.Cout = pl.Cout; a.CS = pl.CSi; a.CT = pl.fCT; a.JT = pl.fJT;
  a.n_ctiles = pl.Cout / pl.fCT; a.n_jtiles = (pl.N + pl.fJT - 1) / pl.fJT; a.pad_rows = pl.N - 1 - (pl.N - 1) / 2; a.KA = pl.N * pl.CSi;
  a.accumulate = accumulate; a.nsh = pl.nsf;
  unsigned grid = (unsigned)(((rows + TC_BM - 1) / TC_BM) * a.n_ctiles * a.n_jtiles);
  if (g_tc_2cta < 0) g_tc_2cta = getenv("SNDVAE_TC_2CTA") ? atoi(getenv("SNDVAE_TC_2CTA")) : 1;   // CTA pairs by default
  if (g_tc_2cta && a.CT * a.JT == 240 && a.nsh >= 2) {
    unsigned grid2 = (unsigned)(2 * ((rows + 2 * TC_BM - 1) / (2 * TC_BM)) * a.n_ctiles * a.n_jtiles);
    toep_gemm2_k<240><<<grid2, TC_THREADS, tc_toep2_smem(240), st>>>(ah, al, pl.fB, a);
  } else if (a.CT * a.JT == 240) toep_gemm_k<240, 2><<<grid, TC_THREADS, tc_toep_smem(240, 2), st>>>(ah, al, pl.fB, a);
  else toep_gemm_k<160, 3><<<grid, TC_THREADS, tc_toep_smem(160, 3), st>>>(ah, al, pl.fB, a);
  return tc_check_launch("toep_gemm_k(fwd)");
}
// din[rows, N*Cin] (+)= dout[rows, N*CSo planes] . Toeplitz(w)^T
static int tc_plan_dgrad(ToepPlan& pl, const __nv_bfloat16* doh, const __nv_bfloat16* dol, float* din, long long rows, long long rows_alloc,
                         int accumulate, cudaStream_t st) {
  CUtensorMap ah, al;
  if (tc_encode_rows(&ah, doh, (long long)pl.N * pl.CSo, rows_alloc, TC_KC, TC_BM, TC_KSWZ) || tc_encode_rows(&al, dol, (long long)pl.N * pl.CSo, rows_alloc, TC_KC, TC_BM, TC_KSWZ)) return -1;
  ToepArgs a; a.out = din; a.rows = rows; a.N = pl.N; a.Cout = pl.Cin; a.CS = pl.CSo; a.CT = pl.dCT; a.JT = pl.dJT;
  a.n_ctiles = pl.Cin / pl.dCT; a.n_jtiles = (pl.N + pl.dJT - 1) / pl.dJT; a.pad_rows = (pl.N - 1) / 2; a.KA = pl.N * pl.CSo;
  a.accumulate = accumulate; a.nsh = pl.nsd;
  unsigned grid = (unsigned)(((rows + TC_BM - 1) / TC_BM) * a.n_ctiles * a.n_jtiles);
  if (g_tc_2cta < 0) g_tc_2cta = getenv("SNDVAE_TC_2CTA") ? atoi(getenv("SNDVAE_TC_2CTA")) : 1;   // CTA pairs by default
  if (g_tc_2cta && a.CT * a.JT == 240 && a.nsh >= 2) {
    unsigned grid2 = (unsigned)(2 * ((rows + 2 * TC_BM - 1) / (2 * TC_BM)) * a.n_ctiles * a.n_jtiles);
    toep_gemm2_k<240><<<grid2, TC_THREADS, tc_toep2_smem(240), st>>>(ah, al, pl.dB, a);
  } else if (a.CT * a.JT == 240) toep_gemm_k<240, 2><<<grid, TC_THREADS, tc_toep_smem(240, 2), st>>>(ah, al, pl.dB, a);
  else toep_gemm_k<160, 3><<<grid, TC_THREADS, tc_toep_smem(160, 3), st>>>(ah, al, pl.dB, a);
  return tc_check_launch("toep_gemm_k(dgrad)");
}
// dw[t, coff+c, q] += sum_rows sum_j in[row, j+t-p, c] dout[row, j, q]
static int tc_plan_wgrad(ToepPlan& pl, const __nv_bfloat16* inh, const __nv_bfloat16* inl, const __nv_bfloat16* doh, const __nv_bfloat16* dol,
                         float* dw, int Ctot, int coff, long long rows, cudaStream_t st) {
  CUtensorMap ah, al, bh, bl;     // bounded by the exact row count: TMA zero-fills the K tail
  if (tc_encode_rows(&ah, inh, (long long)pl.N * pl.CSi, rows, 64, TC_WKC) || tc_encode_rows(&al, inl, (long long)pl.N * pl.CSi, rows, 64, TC_WKC) ||
      tc_encode_rows(&bh, doh, (long long)pl.N * pl.CSo, rows, 64, TC_WKC) || tc_encode_rows(&bl, dol, (long long)pl.N * pl.CSo, rows, 64, TC_WKC)) return -1;
  WgradArgs a; a.dw = dw; a.rows = rows; a.N = pl.N; a.CSi = pl.CSi; a.Cin = pl.Cin; a.CSo = pl.CSo; a.Cout = pl.Cout; a.Ctot = Ctot; a.coff = coff;
  a.n_ntiles = (pl.N * pl.CSo + TC_WN - 1) / TC_WN;
  int n_mtiles = (pl.N * pl.CSi + TC_BM - 1) / TC_BM;
  long long tiles = (long long)n_mtiles * a.n_ntiles;
  long long kchunks = (rows + TC_WKC - 1) / TC_WKC;
  int ks = 1;
  while (tiles * ks < 2 * 148 && ks * 8 < kchunks) ks *= 2;     // fill the 148 SMs when N is small
  a.ksplit = ks; a.plain = 0;
  wgrad_gemm_k<<<(unsigned)(tiles * ks), TC_THREADS, TC_WGRAD_SMEM, st>>>(ah, al, bh, bl, a);
  return tc_check_launch("wgrad_gemm_k");
}

static int tc_init(TcState& s, int N, int Chv, long long B, cudaStream_t st) {
  memset(&s, 0, sizeof s);
  if (tc_global_init()) return -1;
  const int compact = getenv("SNDVAE_TC_PADDED") ? 0 : 1;      // compact channel strides cut the padded-K MMA work
  // SNDVAE_TC_BN160: 128x160 tiles with a 3-stage ring instead of 128x240 with 2 stages
  const int bn160 = getenv("SNDVAE_TC_BN160") ? 1 : 0;
  if (tc_plan_init(s.l1, N, TC_C1, TC_C2, 20, bn160 ? 8 : 12, 10, bn160 ? 16 : 24, compact, st)) return -1;
  if (tc_plan_init(s.l0a, N, Chv, TC_C1, 10, 24, 10, 24, 0, st)) return -1;
  if (tc_plan_init(s.l0c, N, Chv, TC_C1, 10, 24, 10, 24, 0, st)) return -1;
  size_t ni = (size_t)B * N * s.l0a.CSi, no = (size_t)B * N * s.l0a.CSo;
  __nv_bfloat16** ps[8] = {&s.ah, &s.al, &s.ch, &s.cl, &s.dsh, &s.dsl, &s.drh, &s.drl};
  for (int i = 0; i < 8; ++i) {
    size_t n = i < 4 ? ni : no;
    if (cudaMalloc(ps[i], n * 2)) { snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of layer-0 planes failed"); return -1; }
    cudaMemsetAsync(*ps[i], 0, n * 2, st);
  }
  s.ready = 1;
  return 0;
}
static void tc_destroy(TcState& s) {
  tc_plan_destroy(s.l1); tc_plan_destroy(s.l0a); tc_plan_destroy(s.l0c);
  __nv_bfloat16* ps[8] = {s.ah, s.al, s.ch, s.cl, s.dsh, s.dsl, s.drh, s.drl};
  for (int i = 0; i < 8; ++i) if (ps[i]) cudaFree(ps[i]);
  memset(&s, 0, sizeof s);
}

// ------------------------------------------------------------------------------------------
// Layer-0 dense contractions on the tensor cores (the K = 2H products with the [rows, N*C1] tensors):
//   da[(b,i), :]     = dE1[(b,i), (j,o)] . WSa[(j,o), :]   (and dc with dE1^T / WSc)
//   dWSa[(j,o), :]  += dE1^T . a                           (and dWSc with dE1^T^T / c)
// They reuse the Toeplitz kernels with trivial window maps (one position, or position stride = a whole weight row).
// ------------------------------------------------------------------------------------------
struct L0Dense {
  int ready, N, Ch, CSk, C1, CSe;      // Ch = 2H (K of the forward product), CSk its plane stride; CSe = channel stride of dE1 planes
  __nv_bfloat16 *wf_h[2], *wf_l[2];    // forward B planes  [jr][o][CSk]      (0 = WSa, 1 = WSc)
  __nv_bfloat16 *wb_h[2], *wb_l[2];    // backward B planes [48][N*CSe]       rows = ch
  TMapSet mf[2], mb[2];
};
// WS fp32 [N][C1][Ch] -> forward planes (position reversed) and backward planes (channel-major rows over k = (j, o))
__global__ void tc_stage_ws_k(const float* __restrict__ WS, __nv_bfloat16* __restrict__ fh, __nv_bfloat16* __restrict__ fl,
                              __nv_bfloat16* __restrict__ bh, __nv_bfloat16* __restrict__ bl, int N, int C1, int Ch, int CSk, int CSe) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * C1 * Ch) return;
  const int ch = idx % Ch; const int o = (idx / Ch) % C1; const int j = idx / ((long long)Ch * C1);
  const float v = WS[idx];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  const long long f = ((long long)(N - 1 - j) * C1 + o) * CSk + ch;
  fh[f] = hi; fl[f] = lo;
  const long long b = (long long)ch * N * CSe + (long long)j * CSe + o;
  bh[b] = hi; bl[b] = lo;
}
static int l0d_init(L0Dense& d, int N, int Ch, int C1, cudaStream_t st) {
  memset(&d, 0, sizeof d);
  d.N = N; d.Ch = Ch; d.C1 = C1; d.CSk = tc_pad16(Ch); d.CSe = tc_stride(N, C1, 4, 1);
  for (int w = 0; w < 2; ++w) {
    size_t nf = (size_t)N * C1 * d.CSk + 2 * TC_KC, nb = (size_t)48 * N * d.CSe + 2 * TC_KC;
    if (cudaMalloc(&d.wf_h[w], nf * 2) || cudaMalloc(&d.wf_l[w], nf * 2) || cudaMalloc(&d.wb_h[w], nb * 2) || cudaMalloc(&d.wb_l[w], nb * 2)) {
      snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of layer-0 dense planes failed"); return -1;
    }
    cudaMemsetAsync(d.wf_h[w], 0, nf * 2, st); cudaMemsetAsync(d.wf_l[w], 0, nf * 2, st);
    cudaMemsetAsync(d.wb_h[w], 0, nb * 2, st); cudaMemsetAsync(d.wb_l[w], 0, nb * 2, st);
    {   // forward: dims (x = Ch, jr = N positions, o = C1 channels); row (o, jr) = planes[jr][o][:]
      cuuint64_t dims[3] = {(cuuint64_t)Ch, (cuuint64_t)N, (cuuint64_t)C1};
      cuuint64_t str[2] = {(cuuint64_t)C1 * d.CSk * 2, (cuuint64_t)d.CSk * 2};
      cuuint32_t box[3] = {TC_KC, 24, 10};
      if (tc_encode(&d.mf[w].h[0], d.wf_h[w], 3, dims, str, box, TC_KSWZ) || tc_encode(&d.mf[w].l[0], d.wf_l[w], 3, dims, str, box, TC_KSWZ)) return -1;
    }
    {   // backward: dims (x = N*CSe, one position, ch = Ch channels); box of 48 channel rows (rows >= Ch are zero-filled)
      cuuint64_t dims[3] = {(cuuint64_t)N * d.CSe, 1, (cuuint64_t)Ch};
      cuuint64_t str[2] = {(cuuint64_t)N * d.CSe * 2, (cuuint64_t)N * d.CSe * 2};
      cuuint32_t box[3] = {TC_KC, 1, 48};
      if (tc_encode(&d.mb[w].h[0], d.wb_h[w], 3, dims, str, box, TC_KSWZ) || tc_encode(&d.mb[w].l[0], d.wb_l[w], 3, dims, str, box, TC_KSWZ)) return -1;
    }
  }
  d.ready = 1;
  return 0;
}
static void l0d_destroy(L0Dense& d) {
  for (int w = 0; w < 2; ++w) { if (d.wf_h[w]) cudaFree(d.wf_h[w]); if (d.wf_l[w]) cudaFree(d.wf_l[w]); if (d.wb_h[w]) cudaFree(d.wb_h[w]); if (d.wb_l[w]) cudaFree(d.wb_l[w]); }
  memset(&d, 0, sizeof d);
}
static int l0d_stage(L0Dense& d, int w, const float* WS, cudaStream_t st) {
  long long n = (long long)d.N * d.C1 * d.Ch;
  tc_stage_ws_k<<<cdiv(n, 256), 256, 0, st>>>(WS, d.wf_h[w], d.wf_l[w], d.wb_h[w], d.wb_l[w], d.N, d.C1, d.Ch, d.CSk, d.CSe);
  return tc_check_launch("tc_stage_ws_k");
}
// dact[rows, Ch] = dE1 planes[rows, N*CSe] . WS
static int l0d_bwd_act(L0Dense& d, int w, const __nv_bfloat16* eh_, const __nv_bfloat16* el_, float* dact, long long rows, long long rows_alloc,
                       cudaStream_t st) {
  CUtensorMap ah, al;
  if (tc_encode_rows(&ah, eh_, (long long)d.N * d.CSe, rows_alloc, TC_KC, TC_BM, TC_KSWZ) || tc_encode_rows(&al, el_, (long long)d.N * d.CSe, rows_alloc, TC_KC, TC_BM, TC_KSWZ)) return -1;
  ToepArgs a; a.out = dact; a.rows = rows; a.N = 1; a.Cout = d.Ch; a.CS = d.N * d.CSe; a.CT = 48; a.JT = 1; a.n_ctiles = 1; a.n_jtiles = 1;
  a.pad_rows = 0; a.KA = d.N * d.CSe; a.accumulate = 0; a.nsh = 1;
  unsigned grid = (unsigned)((rows + TC_BM - 1) / TC_BM);
  toep_gemm_k<48, 4><<<grid, TC_THREADS, tc_toep_smem(48, 4), st>>>(ah, al, d.mb[w], a);
  return tc_check_launch("toep_gemm_k(l0 dense bwd)");
}
// dWS[(j,o), ch] += sum_rows dE1[row, (j,o)] act[row, ch]
static int l0d_bwd_w(L0Dense& d, const __nv_bfloat16* eh_, const __nv_bfloat16* el_, const __nv_bfloat16* ah_, const __nv_bfloat16* al_,
                     float* dWS, long long rows, cudaStream_t st) {
  CUtensorMap ah, al, bh, bl;
  if (tc_encode_rows(&ah, eh_, (long long)d.N * d.CSe, rows, 64, TC_WKC) || tc_encode_rows(&al, el_, (long long)d.N * d.CSe, rows, 64, TC_WKC) ||
      tc_encode_rows(&bh, ah_, d.CSk, rows, 64, TC_WKC) || tc_encode_rows(&bl, al_, d.CSk, rows, 64, TC_WKC)) return -1;
  WgradArgs a; a.dw = dWS; a.rows = rows; a.N = d.N; a.CSi = d.CSe; a.Cin = d.C1; a.CSo = d.CSk; a.Cout = d.Ch; a.Ctot = 0; a.coff = 0;
  a.n_ntiles = 1; a.plain = 1;
  int n_mtiles = (d.N * d.CSe + TC_BM - 1) / TC_BM;
  long long kchunks = (rows + TC_WKC - 1) / TC_WKC;
  int ks = 1;
  while ((long long)n_mtiles * ks < 2 * 148 && ks * 8 < kchunks) ks *= 2;
  a.ksplit = ks;
  wgrad_gemm_k<<<(unsigned)(n_mtiles * ks), TC_THREADS, TC_WGRAD_SMEM, st>>>(ah, al, bh, bl, a);
  return tc_check_launch("wgrad_gemm_k(l0 dense)");
}
