// mma_probe.cu -- what does one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) cost as a function of N and of how the
// accumulators are chained?  The hot kernels of this repository that sit at "~200 cycles per MMA" (spec_wgrad_k, y_producer_tc_k,
// tsgemm_k) issue short-N products; this probe separates instruction latency, dependent-chain latency and throughput.
// One CTA per SM, one issuing thread, operands = zeroed SWIZZLE_128B K-major tiles in shared memory (contents do not matter),
// `iters` MMAs followed by one tcgen05.commit; cycles = clock64 from the first issue to the commit's arrival.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe tools/mma_probe.cu -lcuda && ./mma_probe
#include "../snd-vae_b200/csrc/e2e_tc.cuh"
#include <cstdio>
#include <vector>

// the same instruction with the election inside the asm statement: the surrounding C++ stays warp-converged, so the compiler can keep
// the (warp-uniform) descriptors in uniform registers instead of moving them there (R2UR) under an ELECT / BRA.U.ANY loop
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// nacc independent accumulators used round-robin; nab distinct A / B tile pairs used round-robin (operand re-use or not)
template <bool CONVERGED>
__global__ void __launch_bounds__(128, 1) mma_probe_k(int N, int nacc, int nab, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 4 * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  if (CONVERGED ? threadIdx.x < 32 : threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 4 * 16384;      // 4 A tiles of 128 x 64 bf16, then 4 B tiles of 256 x 64 bf16
    // nacc, nab are powers of two; descriptors of the 4 K steps x 4 tile pairs are built once (nothing but the MMA in the timed loop)
    uint64_t ad[4][4], bd[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) { ad[o][kk] = umma_desc(sa + o * 16384 + kk * 32, 16, 1024, 2ull); bd[o][kk] = umma_desc(sb + o * 32768 + kk * 32, 16, 1024, 2ull); }
    const uint32_t am = nacc - 1, om = nab - 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a = (uint32_t)(it + kk) & am, o = ((uint32_t)(it + kk) >> 2) & om;
        const uint64_t adesc = o == 0 ? ad[0][kk] : o == 1 ? ad[1][kk] : o == 2 ? ad[2][kk] : ad[3][kk];
        const uint64_t bdesc = o == 0 ? bd[0][kk] : o == 1 ? bd[1][kk] : o == 2 ? bd[2][kk] : bd[3][kk];
        if (CONVERGED) umma_bf16_elect(tmem_d + a * N, adesc, bdesc, idesc, (it + kk) >= nacc ? 1u : 0u);
        else umma_bf16(tmem_d + a * N, adesc, bdesc, idesc, (it + kk) >= nacc ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 4 * (16384 + 32768) + 1024;
  cudaFuncSetAttribute(mma_probe_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_probe_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("tcgen05.mma kind::f16 M=128 K=16, %d MMAs per CTA, 148 CTAs; cycles per MMA (issue loop | until commit arrives)\n", iters);
  printf("%6s %6s %6s %10s %10s | %10s %10s %12s\n", "N", "accs", "tiles", "lane0:issue", "complete", "conv:issue", "complete", "dense-rate");
  const int Ns[] = {48, 64, 96, 128, 256};
  for (int N : Ns)
    for (int nacc : {1, 2, 4})
      for (int nab : {1, 4}) {
        if (nacc * N > 512) continue;
        double res[2][2];
        for (int mode = 0; mode < 2; ++mode) {
          for (int rep = 0; rep < 2; ++rep) {      // the first launch warms up
            if (mode) mma_probe_k<true><<<148, 128, smem>>>(N, nacc, nab, iters, d); else mma_probe_k<false><<<148, 128, smem>>>(N, nacc, nab, iters, d);
          }
          long long h[2] = {0, 0};
          cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          res[mode][0] = (double)h[0] / iters; res[mode][1] = (double)h[1] / iters;
        }
        // dense bf16 rate of one SM: 128 x N x 16 MACs at 4096 MACs / clock
        printf("%6d %6d %6d %10.1f %10.1f | %10.1f %10.1f %12.1f\n", N, nacc, nab, res[0][0], res[0][1], res[1][0], res[1][1], 128.0 * N * 16 / 4096.0);
      }
  cudaFree(d);
  return 0;
}
