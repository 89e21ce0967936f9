// common.cuh -- shared helpers for the SND-VAE sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define BN_RS 0.99950037468777316f /* 1/sqrt(1 + 1e-3): Keras BN inference, moving var 1, eps 1e-3 */

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float lrelu_f(float x) { return fmaxf(x, 0.2f * x); }
// TF MaximumGrad sends the gradient to x when x >= 0.2x, i.e. x >= 0 (layers.py:112-113)
__device__ __forceinline__ float lrelu_g(float x) { return x >= 0.f ? 1.f : 0.2f; }
__device__ __forceinline__ float act_f(int act, float x) {
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_LRELU) return lrelu_f(x);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}
__device__ __forceinline__ float act_g(int act, float x) {
  if (act == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (act == ACT_LRELU) return lrelu_g(x);
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of `v`; result valid in thread 0.  `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = l < nw ? red[l] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

static inline unsigned int cdiv(long long a, long long b) { return (unsigned int)((a + b - 1) / b); }
