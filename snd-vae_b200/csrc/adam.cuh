// adam.cuh -- TF1 ApplyAdam (tf.train.AdamOptimizer, optimizer.py:125,197) over the flat
// parameter arena.  HBM-bound: 28 B/param (read p,g,m,v; write p,m,v), 128-bit accesses.
//   alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)      (b1^t, b2^t fp32 running products, host side)
//   m += (g - m)(1 - b1);  v += (g*g - v)(1 - b2);  p -= (m * alpha) / (sqrt(v) + eps)
// NOTE eps sits outside the bias correction: this is not torch.optim.Adam.
#pragma once
#include "common.cuh"

__global__ void __launch_bounds__(256) tf_adam_k(float4* __restrict__ p, const float4* __restrict__ g,
                                                 float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                 const float* __restrict__ alpha_p, float omb1, float omb2, float eps) {
  // alpha (the bias-corrected step size of this iteration) is read from device memory so that the launch arguments do not
  // change from step to step: the whole step can then be replayed as one CUDA graph
  const float alpha = __ldg(alpha_p);
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
#define ADAM1(c)                                   \
    M.c += (G.c - M.c) * omb1;                     \
    V.c += (G.c * G.c - V.c) * omb2;               \
    P.c -= (M.c * alpha) / (sqrtf(V.c) + eps);
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    p[i] = P; m[i] = M; v[i] = V;
  }
}
