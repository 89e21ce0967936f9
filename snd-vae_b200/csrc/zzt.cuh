// zzt.cuh -- InnerProductDecoder (layers.py:400-410): logits[b] = Z_b Z_b^T for a batch of node embeddings.
// NOT part of the reference's models (model.py / model_joint.py decode edges with the e2e layers); the layer exists in
// layers.py and BASELINE config 4 names it, so it is offered as a standalone operator.
//
// One tcgen05 GEMM per 128 x 128 output tile: both operands are K-major row tiles of the SAME bf16 planes of Z (hi / lo
// split, bf16x3: hi.hi + hi.lo + lo.hi), fetched by TMA through a 3-D map (k, node, graph) whose out-of-range rows and
// columns read as zero, so ragged N and h need no padding pass.  The fp32 accumulator (128 TMEM columns) is drained by four
// warps through a swizzled shared-memory transpose, so that every store instruction writes 512 contiguous bytes of one
// output row.  The operator is bound by the N^2 fp32 logits it writes.
#pragma once
#include "e2e_tc.cuh"

#define ZZT_STAGES 1      /* h <= 64 is one K chunk; 65 KB of shared memory lets three CTAs share an SM and overlap their load, MMA and store phases */
struct ZztArgs { float* out; int N; int nt; int nk; };

__global__ void __launch_bounds__(192, 3) zzt_gemm_k(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmL, ZztArgs P) {
  constexpr int T_BYTES = 128 * 64 * 2;            // one 128-row, 64-wide bf16 operand tile
  constexpr int STAGE_BYTES = 4 * T_BYTES;         // A hi, A lo, B hi, B lo
  extern __shared__ __align__(1024) uint8_t zzt_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)zzt_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[ZZT_STAGES], empty_bar[ZZT_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile = blockIdx.x;
  const int tpg = P.nt * P.nt;
  const int b = (int)(tile / tpg), rem = (int)(tile - (long long)b * tpg);
  const int mt = rem / P.nt, ct = rem - mt * P.nt;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ZZT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < P.nk; ++it) {
        const int s = it % ZZT_STAGES; const uint32_t ph = (it / ZZT_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_3d(st, &tmH, &full_bar[s], it * 64, mt * 128, b);
        tma_load_3d(st + T_BYTES, &tmL, &full_bar[s], it * 64, mt * 128, b);
        tma_load_3d(st + 2 * T_BYTES, &tmH, &full_bar[s], it * 64, ct * 128, b);
        tma_load_3d(st + 3 * T_BYTES, &tmL, &full_bar[s], it * 64, ct * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, 128, 0, 0);
      for (int it = 0; it < P.nk; ++it) {
        const int s = it % ZZT_STAGES; const uint32_t ph = (it / ZZT_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ah = umma_desc(sa + k * 32, 16, 1024, 2ull), al = umma_desc(sa + T_BYTES + k * 32, 16, 1024, 2ull);
          const uint64_t bh = umma_desc(sa + 2 * T_BYTES + k * 32, 16, 1024, 2ull), bl = umma_desc(sa + 3 * T_BYTES + k * 32, 16, 1024, 2ull);
          umma_bf16(tmem_d, ah, bh, idesc, (it | k) ? 1u : 0u);
          umma_bf16(tmem_d, ah, bl, idesc, 1u);
          umma_bf16(tmem_d, al, bh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_full);
    }
  } else {
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    mbar_wait(&acc_full, 0);                         // every MMA has retired: the operand tiles are dead, reuse them as staging
    tc_fence_after();
    // warp-private 32 x 128 fp32 staging tile, float4 index (row, c4) -> row * 32 + (c4 ^ row): conflict-free for the
    // row-per-lane writes out of TMEM and for the row-per-instruction reads that feed fully coalesced 512-byte stores
    float4* stg = reinterpret_cast<float4*>(smem + (size_t)(warp - 2) * 32 * 128 * 4);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        stg[lane * 32 + (((c0 >> 2) + u) ^ lane)] = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
    }
    __syncwarp();
    const int j = ct * 128 + 4 * lane;
    const bool v4 = (P.N & 3) == 0;
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int i = mt * 128 + q * 32 + rr;
      if (i >= P.N) break;
      const float4 v = stg[rr * 32 + (lane ^ rr)];
      float* o = P.out + ((long long)b * P.N + i) * P.N + j;
      if (v4) { if (j < P.N) *reinterpret_cast<float4*>(o) = v; }
      else {
        if (j < P.N) o[0] = v.x;
        if (j + 1 < P.N) o[1] = v.y;
        if (j + 2 < P.N) o[2] = v.z;
        if (j + 3 < P.N) o[3] = v.w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 128); }
}

// z [B, N, h] fp32 -> out [B, N, N] fp32, on `st`.  `planes` / `cap` is a grow-only scratch buffer owned by the caller (bytes).
// Returns 0, or -1 with tc_last_error() set.
static int zzt_run(const float* z, long long B, int N, int h, float* out, void** planes, size_t* cap, cudaStream_t st) {
  if (tc_global_init()) return -1;
  static bool attr = false;
  const size_t smem = (size_t)ZZT_STAGES * 4 * 128 * 64 * 2 + 1024;
  if (!attr) { cudaFuncSetAttribute(zzt_gemm_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int Kp = (h + 7) / 8 * 8;                    // 16-byte plane rows
  const long long rows = B * N;
  const size_t plane = ((size_t)rows * Kp * 2 + 255) & ~(size_t)255;
  if (*cap < 2 * plane) {
    if (*planes) { cudaStreamSynchronize(st); cudaFree(*planes); *planes = nullptr; *cap = 0; }
    if (cudaMalloc(planes, 2 * plane) != cudaSuccess) {
      snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc of the embedding planes (%zu bytes) failed", 2 * plane); cudaGetLastError(); return -1;
    }
    *cap = 2 * plane;
  }
  __nv_bfloat16* hi = (__nv_bfloat16*)*planes; __nv_bfloat16* lo = (__nv_bfloat16*)((uint8_t*)*planes + plane);
  int rc = 0;
  if (Kp != h) { cudaMemsetAsync(hi, 0, (size_t)rows * Kp * 2, st); cudaMemsetAsync(lo, 0, (size_t)rows * Kp * 2, st); }
  if (tc_split(z, hi, lo, rows, h, Kp, st)) rc = -1;
  CUtensorMap mh, ml;
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)N * Kp * 2};
  cuuint32_t box[3] = {64, 128, 1};
  if (!rc && (tc_encode(&mh, hi, 3, dims, str, box) || tc_encode(&ml, lo, 3, dims, str, box))) rc = -1;
  if (!rc) {
    ZztArgs P; P.out = out; P.N = N; P.nt = (N + 127) / 128; P.nk = (Kp + 63) / 64;
    const long long tiles = B * P.nt * P.nt;
    if (tiles > 0x7fffffffLL) { snprintf(g_tc_err, sizeof g_tc_err, "too many output tiles"); rc = -1; }
    else { zzt_gemm_k<<<(unsigned)tiles, 192, smem, st>>>(mh, ml, P); rc = tc_check_launch("zzt_gemm_k"); }
  }
  return rc;
}
