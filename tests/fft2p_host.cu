// Host check of the two-pass transforms (snd-vae_b200/csrc/fft2p.cuh): the per-thread pieces are __host__ __device__, so the
// kernels' phases are run here thread by thread (a block barrier = the end of a loop over the threads) against a double-precision
// DFT.  Built and run by tests/test_host.py::test_fft2p_host (nvcc, CPU only).
#include "../snd-vae_b200/csrc/fft2p.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>

typedef std::complex<double> cd;
static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

template <int R, class F> static double check_small(F f, int nin, int nout) {
  float2 v[R]; cd x[R];
  for (int i = 0; i < R; ++i) { const double a = i < nin ? frand() : 0.0, b = i < nin ? frand() : 0.0; x[i] = cd(a, b); v[i] = fp_mk((float)a, (float)b); }
  f(v);
  double err = 0.0;
  for (int k = 0; k < nout; ++k) {
    cd s = 0.0;
    for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, -2.0 * M_PI * n * k / R);
    err = fmax(err, std::abs(s - cd(v[k].x, v[k].y)));
  }
  return err;
}

struct FwdEmit {
  std::vector<cd>* X1; std::vector<cd>* X2; std::vector<int>* cnt; int cp, G;
  void operator()(int f, float2 x1, float2 x2) { (*X1)[f * G + cp] = cd(x1.x, x1.y); (*X2)[f * G + cp] = cd(x2.x, x2.y); (*cnt)[f * G + cp]++; }
};
struct InvEmit {
  std::vector<cd>* out; std::vector<int>* cnt; int cp, G; double scale;
  void operator()(int pos, float2 v) { (*out)[pos * G + cp] = cd(v.y * scale, v.x * scale); (*cnt)[pos * G + cp]++; }   // swap back + scale
};

template <int M, int G, bool TABLE> static int run(int N, bool bn) {
  constexpr int L = 48 * M, F = L / 2 + 1, R0 = 3 * M, NT = 16 * G;
  std::vector<float2> tw(L);
  for (int k = 0; k < L; ++k) tw[k] = fp_mk((float)cos(-2.0 * M_PI * k / L), (float)sin(-2.0 * M_PI * k / L));
  // ---- forward: line [N][2G] -> spectra of the 2G real channels
  std::vector<float> line((size_t)N * 2 * G);
  for (auto& x : line) x = (float)frand();
  std::vector<float> g(2 * G), b(2 * G);
  for (int c = 0; c < 2 * G; ++c) { g[c] = (float)(0.5 + frand()); b[c] = (float)(0.3 * frand()); }
  std::vector<float2> bufA((size_t)L * G);
  for (int tid = 0; tid < NT; ++tid) {
    const int jb = tid / G, cp = tid % G;
    FpBn p; p.on = bn; p.gx = g[2 * cp]; p.gy = g[2 * cp + 1]; p.bx = b[2 * cp]; p.by = b[2 * cp + 1];
    fp_fwd_pass1<M, G>(reinterpret_cast<const float2*>(line.data()), bufA.data(), jb, cp, N, p);
  }
  std::vector<cd> X1((size_t)F * G), X2((size_t)F * G); std::vector<int> cnt((size_t)F * G, 0);
  for (int tid = 0; tid < NT; ++tid) {
    const int jb = tid / G, cp = tid % G;
    if (jb > R0 / 2) continue;
    FwdEmit e{&X1, &X2, &cnt, cp, G};
    fp_fwd_pass2<M, G, TABLE>(bufA.data(), tw.data(), jb, cp, e);
  }
  double ferr = 0.0, fmaxv = 0.0; int bad = 0;
  for (int c = 0; c < 2 * G; ++c)
    for (int f = 0; f < F; ++f) {
      cd s = 0.0;
      for (int n = 0; n < N; ++n) {
        double x = line[(size_t)n * 2 * G + c];
        if (bn) x = fmax((double)fmaf((float)x, g[c], b[c]), 0.0);
        s += x * std::polar(1.0, -2.0 * M_PI * n * f / L);
      }
      const cd got = (c & 1) ? X2[f * G + c / 2] : X1[f * G + c / 2];
      ferr = fmax(ferr, std::abs(s - got)); fmaxv = fmax(fmaxv, std::abs(s));
    }
  for (int i = 0; i < F * G; ++i) if (cnt[i] < 1) ++bad;
  // ---- inverse: spectrum rows [F][4G] of 2G real signals -> first N positions
  std::vector<float> spec((size_t)F * 4 * G);
  std::vector<double> sig((size_t)L * 2 * G);
  for (auto& x : sig) x = frand();
  for (int c = 0; c < 2 * G; ++c)
    for (int f = 0; f < F; ++f) {
      cd s = 0.0;
      for (int n = 0; n < L; ++n) s += sig[(size_t)n * 2 * G + c] * std::polar(1.0, -2.0 * M_PI * n * f / L);
      spec[(size_t)f * 4 * G + 4 * (c / 2) + (c & 1)] = (float)s.real(); spec[(size_t)f * 4 * G + 4 * (c / 2) + 2 + (c & 1)] = (float)s.imag();
    }
  for (int tid = 0; tid < NT; ++tid) {
    const int jb = tid / G, cp = tid % G;
    if (jb > R0 / 2) continue;
    fp_inv_pass1<M, G>(spec.data(), bufA.data(), jb, cp);
  }
  std::vector<cd> out((size_t)L * G); std::vector<int> ocnt((size_t)L * G, 0);
  for (int tid = 0; tid < NT; ++tid) {
    const int jb = tid / G, cp = tid % G;
    InvEmit e{&out, &ocnt, cp, G, 1.0 / L};
    fp_inv_pass2<M, G, TABLE>(bufA.data(), tw.data(), jb, cp, e);
  }
  double ierr = 0.0;
  for (int pos = 0; pos < 32 * M; ++pos)
    for (int cp = 0; cp < G; ++cp) {
      if (ocnt[pos * G + cp] != 1) ++bad;
      const cd want(sig[(size_t)pos * 2 * G + 2 * cp], sig[(size_t)pos * 2 * G + 2 * cp + 1]);
      ierr = fmax(ierr, std::abs(want - out[pos * G + cp]));
    }
  printf("M=%d G=%d table=%d N=%d bn=%d: fwd err %.3g (max |X| %.3g)  inv err %.3g  uncovered %d\n", M, G, (int)TABLE, N, (int)bn, ferr, fmaxv, ierr, bad);
  return (ferr < 2e-4 * fmaxv && ierr < 2e-5 && bad == 0) ? 0 : 1;
}


// three-pass plan of L = 1536 (N <= 1024): forward 24 (pruned) x 8 x 8 (paired), inverse 8 (paired) x 8 x 24 (pruned)
template <int G> static int run3(int N, bool bn) {
  constexpr int L = 1536, F = L / 2 + 1;
  std::vector<float2> tw(L);
  for (int k = 0; k < L; ++k) tw[k] = fp_mk((float)cos(-2.0 * M_PI * k / L), (float)sin(-2.0 * M_PI * k / L));
  std::vector<float> line((size_t)N * 2 * G);
  for (auto& x : line) x = (float)frand();
  std::vector<float> g(2 * G), b(2 * G);
  for (int c = 0; c < 2 * G; ++c) { g[c] = (float)(0.5 + frand()); b[c] = (float)(0.3 * frand()); }
  std::vector<float2> bufA((size_t)L * G), bufB((size_t)L * G);
  for (int jb = 0; jb < 64; ++jb) for (int cp = 0; cp < G; ++cp) {
    FpBn p; p.on = bn; p.gx = g[2 * cp]; p.gy = g[2 * cp + 1]; p.bx = b[2 * cp]; p.by = b[2 * cp + 1];
    fp_fwd_pass1_t<8, G, 64>(reinterpret_cast<const float2*>(line.data()), bufA.data(), jb, cp, N, p);
  }
  for (int j = 0; j < 192; ++j) for (int cp = 0; cp < G; ++cp) fp_mid8<G, 24, 192, 8>(bufA.data(), bufB.data(), tw.data(), j, cp);
  std::vector<cd> X1((size_t)F * G), X2((size_t)F * G); std::vector<int> cnt((size_t)F * G, 0);
  for (int u = 0; u <= 96; ++u) for (int cp = 0; cp < G; ++cp) { FwdEmit e{&X1, &X2, &cnt, cp, G}; fp_fwd_last_t<8, 192, G, true>(bufB.data(), tw.data(), u, cp, e); }
  double ferr = 0.0, fmaxv = 0.0; int bad = 0;
  for (int c = 0; c < 2 * G; ++c)
    for (int f = 0; f < F; f += 7) {                     // every 7th frequency (and the last): the direct DFT is O(N) each
      cd s = 0.0;
      for (int n = 0; n < N; ++n) {
        double x = line[(size_t)n * 2 * G + c];
        if (bn) x = fmax((double)fmaf((float)x, g[c], b[c]), 0.0);
        s += x * std::polar(1.0, -2.0 * M_PI * (double)((long long)n * f % L) / L);
      }
      const cd got = (c & 1) ? X2[f * G + c / 2] : X1[f * G + c / 2];
      ferr = fmax(ferr, std::abs(s - got)); fmaxv = fmax(fmaxv, std::abs(s));
    }
  for (int i = 0; i < F * G; ++i) if (cnt[i] < 1) ++bad;
  // inverse: random spectrum rows (Hermitian by construction of the loader) against the direct inverse sum at sampled positions
  std::vector<float> spec((size_t)F * 4 * G);
  for (auto& x : spec) x = (float)frand();
  for (int u = 0; u <= 96; ++u) for (int cp = 0; cp < G; ++cp) fp_inv_first_t<8, 192, G>(spec.data(), bufA.data(), u, cp);
  for (int j = 0; j < 192; ++j) for (int cp = 0; cp < G; ++cp) fp_mid8<G, 8, 192, 24>(bufA.data(), bufB.data(), tw.data(), j, cp);
  std::vector<cd> out((size_t)L * G); std::vector<int> ocnt((size_t)L * G, 0);
  for (int k = 0; k < 64; ++k) for (int cp = 0; cp < G; ++cp) { InvEmit e{&out, &ocnt, cp, G, 1.0 / L}; fp_inv_last_t<8, G, 64, false>(bufB.data(), tw.data(), k, cp, e); }
  double ierr = 0.0, imax = 0.0;
  for (int pos = 0; pos < 1024; ++pos) for (int cp = 0; cp < G; ++cp) if (ocnt[pos * G + cp] != 1) ++bad;
  for (int pos = 0; pos < 1024; pos += 13)
    for (int c = 0; c < 2 * G; ++c) {
      // x[pos] = (1/L) sum_f kappa_f Re(X[f] e^{+2 pi i f pos / L}), X[f] = re + i im (im of f = 0, L/2 ignored)
      double acc = 0.0;
      for (int f = 0; f < F; ++f) {
        const double re = spec[(size_t)f * 4 * G + 4 * (c / 2) + (c & 1)];
        const double im = (f == 0 || f == L / 2) ? 0.0 : spec[(size_t)f * 4 * G + 4 * (c / 2) + 2 + (c & 1)];
        const double th = 2.0 * M_PI * (double)((long long)f * pos % L) / L;
        acc += ((f == 0 || f == L / 2) ? 1.0 : 2.0) * (re * cos(th) - im * sin(th));
      }
      acc /= L;
      const cd got = out[pos * G + c / 2];
      const double gv = (c & 1) ? got.imag() : got.real();
      ierr = fmax(ierr, fabs(acc - gv)); imax = fmax(imax, fabs(acc));
    }
  printf("L=1536 G=%d N=%d bn=%d: fwd err %.3g (max |X| %.3g)  inv err %.3g (max %.3g)  uncovered %d\n", G, N, (int)bn, ferr, fmaxv, ierr, imax, bad);
  return (ferr < 2e-4 * fmaxv && ierr < 2e-4 * imax && bad == 0) ? 0 : 1;
}

int main() {
  srand(7);
  int fails = 0;
  const double e16 = check_small<16>([](float2* v) { fp_dft16(v); }, 16, 16);
  const double e24i = check_small<24>([](float2* v) { fp_dft3M_in2M<8>(v); }, 16, 24);
  const double e24o = check_small<24>([](float2* v) { fp_dft3M_out2M<8>(v); }, 24, 16);
  const double e12i = check_small<12>([](float2* v) { fp_dft3M_in2M<4>(v); }, 8, 12);
  const double e12o = check_small<12>([](float2* v) { fp_dft3M_out2M<4>(v); }, 12, 8);
  printf("dft16 %.3g  dft24(in 16) %.3g  dft24(out 16) %.3g  dft12(in 8) %.3g  dft12(out 8) %.3g\n", e16, e24i, e24o, e12i, e12o);
  if (e16 > 1e-5 || e24i > 1e-5 || e24o > 1e-5 || e12i > 1e-5 || e12o > 1e-5) ++fails;
  fails += run<8, 25, false>(256, true);
  fails += run<8, 25, true>(256, false);
  fails += run<8, 10, false>(200, false);
  fails += run<8, 10, true>(193, true);
  fails += run<4, 25, false>(100, true);
  fails += run<4, 10, true>(128, false);
  fails += run3<5>(1024, true);
  fails += run3<5>(700, false);
  printf(fails ? "FAILED\n" : "OK\n");
  return fails ? 1 : 0;
}
