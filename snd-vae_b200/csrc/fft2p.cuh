// fft2p.cuh -- two-pass transforms of length L = 48 M (M = 8: L = 384 for N = 193..256; M = 4: L = 192 for N = 97..128)
// for the spectral form of e2e layer 1 (spectral.cuh; layers.py:431-450 at model.py:202).
//
// The three-pass plans (6 x 8 x 8) move every point of a line through shared memory six times plus the channel separation;
// shared-memory wavefronts, issue slots and DRAM each sat near 50 % with four block barriers per line.  Here a line makes
// two passes and what surrounds the transform is folded into them:
//   forward  L = (3M) x 16:  pass 1 = radix 3M straight from the staged fp32 line (BN + relu on the fly).  Only the first
//            2M of its 3M inputs can be non-zero (positions >= 32 M >= N are padding), so the radix-3 stage is pruned.
//            pass 2 = radix 16; a thread takes the butterflies k and 3M - k together: their outputs f = k + 3M r and
//            L - f = (3M - k) + 3M (15 - r) are exactly the pairs the separation of the two packed real channels needs,
//            so the spectrum goes from registers to the bf16 hi / lo planes without a second buffer or another barrier.
//   inverse  L = 16 x (3M):  pass 1 = radix 16 from the staged spectrum rows; the butterflies j and 3M - j read mirrored
//            inputs (Z[L - p] from the same row as Z[p]), so one thread loads each row element once and builds both.
//            pass 2 = radix 3M with only the first 2M outputs computed (positions < 32 M; the rest is never stored).
// The per-thread pieces are __host__ __device__ so that tests/fft2p_host.cu can run them, thread by thread and phase by phase,
// against a double-precision DFT on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef FP_HD
#define FP_HD __host__ __device__ __forceinline__
#endif

FP_HD float2 fp_mk(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
FP_HD float2 fp_add(float2 a, float2 b) { return fp_mk(a.x + b.x, a.y + b.y); }
FP_HD float2 fp_sub(float2 a, float2 b) { return fp_mk(a.x - b.x, a.y - b.y); }
FP_HD float2 fp_mul(float2 a, float2 b) { return fp_mk(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
FP_HD float2 fp_mulnegi(float2 a) { return fp_mk(a.y, -a.x); }      // a * (-i)

// e^{-2 pi i m / 48}, m a compile-time constant once the calling loop is unrolled (the switch folds away)
FP_HD float2 fp_tw48(int m) {
  switch (m) {
    case 0: return fp_mk(1.0f, 0.0f);
    case 1: return fp_mk(0.9914448613738104f, -0.13052619222005157f);
    case 2: return fp_mk(0.9659258262890683f, -0.25881904510252074f);
    case 3: return fp_mk(0.9238795325112867f, -0.3826834323650898f);
    case 4: return fp_mk(0.8660254037844387f, -0.5f);
    case 5: return fp_mk(0.7933533402912352f, -0.6087614290087207f);
    case 6: return fp_mk(0.7071067811865476f, -0.7071067811865476f);
    case 7: return fp_mk(0.6087614290087207f, -0.7933533402912352f);
    case 8: return fp_mk(0.5f, -0.8660254037844387f);
    case 9: return fp_mk(0.3826834323650898f, -0.9238795325112867f);
    case 10: return fp_mk(0.25881904510252074f, -0.9659258262890683f);
    case 11: return fp_mk(0.13052619222005157f, -0.9914448613738104f);
    default: return fp_mk(0.0f, -1.0f);
  }
}
// a * e^{-2 pi i m / 48}: quarter turns are sign / component swaps, the rest one complex product with a literal
FP_HD float2 fp_mul48(float2 a, int m) {
  m %= 48;
  const int q = m / 12, r = m - 12 * q;
  if (r != 0) a = fp_mul(a, fp_tw48(r));
  if (q == 1) return fp_mk(a.y, -a.x);
  if (q == 2) return fp_mk(-a.x, -a.y);
  if (q == 3) return fp_mk(-a.y, a.x);
  return a;
}

// ---- small forward DFTs on registers (natural order in and out, e^{-2 pi i / R}) -----------------------------------
FP_HD void fp_dft4(float2* v) {
  const float2 t0 = fp_add(v[0], v[2]), t1 = fp_sub(v[0], v[2]), t2 = fp_add(v[1], v[3]), t3 = fp_mulnegi(fp_sub(v[1], v[3]));
  v[0] = fp_add(t0, t2); v[1] = fp_add(t1, t3); v[2] = fp_sub(t0, t2); v[3] = fp_sub(t1, t3);
}
FP_HD void fp_dft8(float2* v) {
  float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
  fp_dft4(e); fp_dft4(o);
  const float s = 0.70710678118654752f;
  const float2 o1 = fp_mk(s * (o[1].x + o[1].y), s * (o[1].y - o[1].x));        // * (s, -s)
  const float2 o2 = fp_mulnegi(o[2]);
  const float2 o3 = fp_mk(s * (o[3].y - o[3].x), -s * (o[3].x + o[3].y));       // * (-s, -s)
  v[0] = fp_add(e[0], o[0]); v[4] = fp_sub(e[0], o[0]);
  v[1] = fp_add(e[1], o1);   v[5] = fp_sub(e[1], o1);
  v[2] = fp_add(e[2], o2);   v[6] = fp_sub(e[2], o2);
  v[3] = fp_add(e[3], o3);   v[7] = fp_sub(e[3], o3);
}
template <int M> FP_HD void fp_dftM(float2* v) { if (M == 8) fp_dft8(v); else fp_dft4(v); }

// 16 = 4 x 4:  n = n1 + 4 n2, k = k2 + 4 k1:  w16^{nk} = w16^{n1 k2} w4^{n1 k1} w4^{n2 k2}
FP_HD void fp_dft16(float2* v) {
  float2 y[4][4];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    float2 t[4] = {v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]};
    fp_dft4(t);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) y[k2][n1] = fp_mul48(t[k2], 3 * n1 * k2);
  }
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    fp_dft4(y[k2]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) v[k2 + 4 * k1] = y[k2][k1];
  }
}
// 3M-point DFT whose inputs v[2M .. 3M) are zero (not read); all 3M outputs, natural order, into v[0 .. 3M).
//   n = n1 + M n2, k = k2 + 3 k1:  w^{nk} = w_{3M}^{n1 k2} w_M^{n1 k1} w_3^{n2 k2}
template <int M> FP_HD void fp_dft3M_in2M(float2* v) {
  const float h = 0.86602540378443865f;
  float2 e[3][M];
#pragma unroll
  for (int n1 = 0; n1 < M; ++n1) {
    const float2 a = v[n1], b = v[n1 + M];
    const float2 m = fp_mk(fmaf(-0.5f, b.x, a.x), fmaf(-0.5f, b.y, a.y));
    e[0][n1] = fp_add(a, b);
    e[1][n1] = fp_mul48(fp_mk(fmaf(h, b.y, m.x), fmaf(-h, b.x, m.y)), n1 * (16 / M));          // a + b w3
    e[2][n1] = fp_mul48(fp_mk(fmaf(-h, b.y, m.x), fmaf(h, b.x, m.y)), 2 * n1 * (16 / M));      // a + b w3^2
  }
  fp_dftM<M>(e[0]); fp_dftM<M>(e[1]); fp_dftM<M>(e[2]);
#pragma unroll
  for (int k1 = 0; k1 < M; ++k1) { v[3 * k1] = e[0][k1]; v[3 * k1 + 1] = e[1][k1]; v[3 * k1 + 2] = e[2][k1]; }
}
// 3M-point DFT of v[0 .. 3M) of which only the outputs k < 2M are produced (into v[0 .. 2M)).
//   n = 3 n1 + n2, k = k1 + M k2:  w^{nk} = w_M^{n1 k1} w_{3M}^{n2 k1} w_3^{n2 k2},  k2 = 0, 1
template <int M> FP_HD void fp_dft3M_out2M(float2* v) {
  const float h = 0.86602540378443865f;
  float2 t[3][M];
#pragma unroll
  for (int n1 = 0; n1 < M; ++n1) { t[0][n1] = v[3 * n1]; t[1][n1] = v[3 * n1 + 1]; t[2][n1] = v[3 * n1 + 2]; }
  fp_dftM<M>(t[0]); fp_dftM<M>(t[1]); fp_dftM<M>(t[2]);
#pragma unroll
  for (int k1 = 0; k1 < M; ++k1) {
    const float2 y0 = t[0][k1], y1 = fp_mul48(t[1][k1], k1 * (16 / M)), y2 = fp_mul48(t[2][k1], 2 * k1 * (16 / M));
    const float2 s = fp_add(y1, y2), d = fp_sub(y1, y2);
    v[k1] = fp_add(y0, s);
    v[k1 + M] = fp_mk(fmaf(h, d.y, fmaf(-0.5f, s.x, y0.x)), fmaf(-h, d.x, fmaf(-0.5f, s.y, y0.y)));   // y0 + y1 w3 + y2 w3^2
  }
}

// twiddles w[r] = tw[r k], r < R, of one butterfly: either R - 1 table reads or one read and a product chain
// (w^2 = w w, w^3 = w^2 w, w^4 = w^2 w^2, ...: at most log2 R + 1 products deep)
template <int R, bool TABLE> FP_HD void fp_twiddle(float2* v, const float2* tw, int k) {
  if (TABLE) {
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = fp_mul(v[r], tw[r * k]);
  } else {
    float2 w[R];
    w[1] = tw[k];
#pragma unroll
    for (int r = 2; r < R; ++r) w[r] = (r & 1) ? fp_mul(w[r - 1], w[1]) : fp_mul(w[r / 2], w[r / 2]);
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = fp_mul(v[r], w[r]);
  }
}

// ======================================================================================================================
// forward, L = 48 M = (3M) x 16.  Thread (jb, cp): cp = channel pair (two real channels packed as re / im), jb < 16.
// Buffers: stage = the fp32 line [pos][2G] seen as float2 [pos][G]; bufA = float2 [L][G].
// ======================================================================================================================
struct FpBn { bool on; float gx, gy, bx, by; };          // x <- relu(x * g + b) for the thread's two channels

// pass 1: butterfly jb (< T0) of the radix-3M pass (Ns = 1): inputs at positions jb + T0 r (r < 2M), outputs at 3M jb + r (r < 3M).
// T0 = L / (3M): 16 in the two-pass plans, 64 in the three-pass plan of L = 1536.
template <int M, int G, int T0> FP_HD void fp_fwd_pass1_t(const float2* stage, float2* bufA, int jb, int cp, int N, const FpBn& bn) {
  float2 v[3 * M];
  const float2* s = stage + jb * G + cp;
#pragma unroll
  for (int r = 0; r < 2 * M; ++r) {
    float2 x = fp_mk(0.f, 0.f);
    if (jb + T0 * r < N) {
      x = s[T0 * r * G];
      if (bn.on) { x.x = fmaxf(fmaf(x.x, bn.gx, bn.bx), 0.f); x.y = fmaxf(fmaf(x.y, bn.gy, bn.by), 0.f); }
    }
    v[r] = x;
  }
  fp_dft3M_in2M<M>(v);
  float2* o = bufA + (3 * M * jb) * G + cp;
#pragma unroll
  for (int r = 0; r < 3 * M; ++r) o[r * G] = v[r];
}
template <int M, int G> FP_HD void fp_fwd_pass1(const float2* stage, float2* bufA, int jb, int cp, int N, const FpBn& bn) {
  fp_fwd_pass1_t<M, G, 16>(stage, bufA, jb, cp, N, bn);
}
template <int R> FP_HD void fp_dftR(float2* v) { if (R == 16) fp_dft16(v); else fp_dft8(v); }
// one radix-R butterfly k (< NS) of the last pass (Ns = NS, L = R NS): inputs buf[k + NS r] * tw[r k], outputs (left in v) are
// the spectrum points f = k + NS r
template <int R, int NS, int G, bool TABLE> FP_HD void fp_fwd_bfly_t(const float2* buf, const float2* tw, int k, int cp, float2* v) {
  const float2* s = buf + k * G + cp;
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = s[NS * r * G];
  fp_twiddle<R, TABLE>(v, tw, k);
  fp_dftR<R>(v);
}
// last pass of unit u (u <= NS / 2): butterflies kA = u and kB = NS - u, channel separation of the pairs (f, L - f), and
// emit(f, X1, X2) for each of the unit's frequencies f <= L / 2, where X1 / X2 are the spectra of the pair's two real channels
//   X1 = (Z[f] + conj Z[L - f]) / 2,   X2 = (Z[f] - conj Z[L - f]) / (2 i)
// (R = 16, NS = 3M in the two-pass plans; R = 8, NS = 192 in the three-pass plan of L = 1536)
template <int R, int NS, int G, bool TABLE, class Emit> FP_HD void fp_fwd_last_t(const float2* buf, const float2* tw, int u, int cp, Emit& emit) {
  const bool pair = (u != 0) && (2 * u != NS);
  float2 a[R], b[R];
  fp_fwd_bfly_t<R, NS, G, TABLE>(buf, tw, u, cp, a);
  if (pair) fp_fwd_bfly_t<R, NS, G, TABLE>(buf, tw, NS - u, cp, b);
  else {
    // self-paired butterflies: u = NS / 2 mirrors onto itself (b = a); u = 0 onto itself shifted by one output (L - NS r = NS (R - r))
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = (u == 0) ? a[(r + 1) & (R - 1)] : a[r];
  }
#pragma unroll
  for (int r = 0; r < R / 2; ++r) {                   // f = u + NS r  <->  L - f = (NS - u) + NS (R - 1 - r)
    const float2 z1 = a[r], z2 = b[R - 1 - r];
    emit(u + NS * r, fp_mk(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y)), fp_mk(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x)));
  }
#pragma unroll
  for (int r = 0; r < R / 2; ++r) {                   // f = (NS - u) + NS r  <->  L - f = u + NS (R - 1 - r)
    // self-paired units would repeat the frequencies above, except f = L / 2 (u = 0, r = R / 2 - 1)
    if (pair || (u == 0 && r == R / 2 - 1)) {
      const float2 z1 = b[r], z2 = a[R - 1 - r];
      emit(NS - u + NS * r, fp_mk(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y)), fp_mk(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x)));
    }
  }
}
template <int M, int G, bool TABLE, class Emit> FP_HD void fp_fwd_pass2(const float2* bufA, const float2* tw, int u, int cp, Emit& emit) {
  fp_fwd_last_t<16, 3 * M, G, TABLE>(bufA, tw, u, cp, emit);
}
// middle pass of the three-pass plans: radix-8 butterfly j (< T = L / 8) of a pass with Ns = NS; TS = L / (8 NS)
//   v[r] = in[j + T r] * tw[r k TS],  k = j mod NS;  out[(j - k) 8 + k + NS r] = DFT_8(v)[r]
template <int G, int NS, int T, int TS> FP_HD void fp_mid8(const float2* in, float2* out, const float2* tw, int j, int cp) {
  const int k = j % NS;
  float2 v[8];
  const float2* s = in + j * G + cp;
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = s[T * r * G];
#pragma unroll
  for (int r = 1; r < 8; ++r) v[r] = fp_mul(v[r], tw[r * k * TS]);
  fp_dft8(v);
  float2* o = out + ((j - k) * 8 + k) * G + cp;
#pragma unroll
  for (int r = 0; r < 8; ++r) o[NS * r * G] = v[r];
}

// ======================================================================================================================
// inverse, L = 48 M = 16 x (3M).  Staged spectrum rows: stage[f][4G] floats = G x [re c, re c+1, im c, im c+1], f <= L / 2.
// Z = X1 + i X2 over the full circle (X[L - f] = conj X[f]); inverse DFT = swap . forward DFT . swap, so the transforms
// below run on the swapped values and the final store swaps back (and scales by 1 / L).
// ======================================================================================================================
// swapped Z[f] and swapped Z[L - f] from row f of the staged spectrum (pair cp: one 16-byte read): a = (re X1, re X2), b = (im X1, im X2)
template <int LH, int G> FP_HD void fp_inv_load(const float* stage, int f, int cp, float2& direct, float2& mirror) {
  const float4 ab = *reinterpret_cast<const float4*>(stage + f * (4 * G) + 4 * cp);
  const float2 a = fp_mk(ab.x, ab.y);
  float2 b = fp_mk(ab.z, ab.w);
  if (f == 0 || f == LH) b = fp_mk(0.f, 0.f);            // the imaginary parts of the self-conjugate frequencies (0, L / 2) do not enter
  direct = fp_mk(b.x + a.y, a.x - b.y);
  mirror = fp_mk(a.y - b.x, a.x + b.y);
}
// first pass of unit u (u <= T / 2): radix-R butterflies jA = u and jB = T - u (Ns = 1, no twiddles; L = R T).  Butterfly j reads
// the positions j + T r; L - (u + T r) = (T - u) + T (R - 1 - r), so the rows f = u + T r and f = (T - u) + T r (r < R / 2, all
// <= L / 2) give every input of both.  Outputs at R j + r.  (R = 16, T = 3M two-pass; R = 8, T = 192 three-pass L = 1536)
template <int R, int T, int G> FP_HD void fp_inv_first_t(const float* stage, float2* bufA, int u, int cp) {
  const bool pair = (u != 0) && (2 * u != T);
  float2 a[R], b[R];
#pragma unroll
  for (int r = 0; r < R / 2; ++r) fp_inv_load<R * T / 2, G>(stage, u + T * r, cp, a[r], b[R - 1 - r]);
#pragma unroll
  for (int r = 0; r < R / 2; ++r) fp_inv_load<R * T / 2, G>(stage, T - u + T * r, cp, b[r], a[R - 1 - r]);
  // (u = T / 2: both loops read the same rows and b = a; u = 0: the second loop reads the rows T (r + 1), b is a shifted by one)
  fp_dftR<R>(a);
  float2* o = bufA + (R * u) * G + cp;
#pragma unroll
  for (int r = 0; r < R; ++r) o[r * G] = a[r];
  if (pair) {
    fp_dftR<R>(b);
    float2* o2 = bufA + (R * (T - u)) * G + cp;
#pragma unroll
    for (int r = 0; r < R; ++r) o2[r * G] = b[r];
  }
}
template <int M, int G> FP_HD void fp_inv_pass1(const float* stage, float2* bufA, int u, int cp) { fp_inv_first_t<16, 3 * M, G>(stage, bufA, u, cp); }
// last pass: butterfly k (< T1) of the radix-3M pass (Ns = T1 = L / (3M)): inputs buf[k + T1 r] * tw[r k] (r < 3M), outputs at the
// positions k + T1 r (r < 2M; the others are >= 2M T1 >= N): emit(pos, value) with the value still swapped and unscaled
template <int M, int G, int T1, bool TABLE, class Emit> FP_HD void fp_inv_last_t(const float2* buf, const float2* tw, int k, int cp, Emit& emit) {
  float2 v[3 * M];
  const float2* s = buf + k * G + cp;
#pragma unroll
  for (int r = 0; r < 3 * M; ++r) v[r] = s[T1 * r * G];
  fp_twiddle<3 * M, TABLE>(v, tw, k);
  fp_dft3M_out2M<M>(v);
#pragma unroll
  for (int r = 0; r < 2 * M; ++r) emit(k + T1 * r, v[r]);
}
template <int M, int G, bool TABLE, class Emit> FP_HD void fp_inv_pass2(const float2* bufA, const float2* tw, int k, int cp, Emit& emit) {
  fp_inv_last_t<M, G, 16, TABLE>(bufA, tw, k, cp, emit);
}
