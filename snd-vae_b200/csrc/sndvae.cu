// sndvae.cu -- handle, parameter table, orchestration and the C ABI (include/sndvae.h)
// of the B200-native SND-VAE train / generate step.
//
// Path replaced (reference file:line): model.py:98-222 (encoder/get_z/decoder),
// model_joint.py:72-182, layers.py:112-125,143-198,431-450,566-576,
// optimizer.py:126-164,192-204 (ELBO + Adam), driven as main.py:331.
#include "../../include/sndvae.h"
#include "common.cuh"
#include "layers.cuh"
#include "sgc.cuh"
#include "sgc3.cuh"
#include "edge.cuh"
#include "adam.cuh"
#include "e2e_tc.cuh"
#include "spectral.cuh"
#include "synth.cuh"
#include "zzt.cuh"
#include "tsgemm.cuh"
#include <string>
#include <vector>
#include <cstring>
#include <cstdarg>
#include <cmath>

#define KS 5   /* conv1d kernel size (main.py:184,201,206) */

struct PT {   // parameter offsets (floats) into the flat arenas; -1 = absent
  long gg_w[2], gg_bng[2], gg_bnb[2], encg_g, encg_b, g_lin[3][2];
  long gs_k[3], gs_b[3], gs_bng[3], gs_bnb[3], encs_g, encs_b, s_lin[3][2];
  long sg3_M[2][4], sg3_b[2][4];     // 3-hop layers: Matrix0..3 / bias0..3
  long sg_M1[2], sg_b1[2], sg_M2[2], sg_b2[2], sg_M3[2], sg_b3[2], sg_bng[2], sg_bnb[2], encsg_g, encsg_b, sg_lin[3][2];
  long d_sg_lin1[2], d_s_lin1[2], d_g_lin1[2];
  long n_k[2], n_b[2], n_bng[2], n_bnb[2], decnode_g, decnode_b, d_n_lin2[2];
  long e_bng[2], e_bnb[2], e_w[2], e_b[2], decadj_g, decadj_b, d_e_lin2[2];
  long s_k[3], s_b[3], s_bng[3], s_bnb[3], d_s_lin2[2];
};

struct EvPair { cudaEvent_t a, b; double flops; };
struct StageTimer { std::vector<cudaEvent_t> ev; std::vector<const char*> name; size_t used; int on; int print;
                    std::vector<std::pair<std::string, double>> total; long long steps; };

struct sndvae_handle {
  sndvae_config cfg;
  cudaStream_t stream;
  std::string err;
  std::vector<sndvae_param_info> table;
  PT pt;
  long long nparam;          // padded arena length
  float *P, *G, *M, *V;      // params, grads, Adam slots
  float b1p, b2p;            // running beta powers (fp32, as TF's beta1_power / beta2_power)
  std::vector<void*> allocs;
  long long launches;
  // derived sizes
  int dis, N, F, D, S, H, Chv, C1, C2, Bc, SC;
  long long B, BS, Rn;
  // buffers (see alloc_buffers)
  float *t0, *c0, *g1, *t1, *c1, *g2, *fg, *hg, *mu_g, *ls_g;
  float *h1p, *h1, *h2p, *h2, *h3p, *h3, *fs, *hs, *mu_s, *ls_s;
  float *fsg, *hsg, *mu_sg, *ls_sg, *dfsg;
  float *x1, *x2, *dxa, *dxb;          // SGC chunk activations / grads
  SgcEdges E; SgcScratch S0, S1;
  int hops3; float *y3[2], *ws3; long long ws3_stride, SC3;   // SpatialGraphConvolution_3D: pre-BN layer outputs, per-sample workspace
  int sgc_keep;                        // forward activations of the joint encoder kept for every sample (no recompute in backward)
  float *z_s, *z_g, *z_sg, *zbar, *dz_s, *dz_g, *dzbar;
  float *dmu, *dls, *dh;               // head backward temporaries (sized for BS rows)
  float *n_sg, *n_s, *n_g, *dn_sg, *dn_s, *dn_g;
  float *v, *sp0, *q1p, *q1, *q2p, *q2, *q3, *xpre, *xhat, *dxpre;
  float *s1p, *s1, *s2p, *s2, *s3p, *s3, *ppre, *phat, *dppre;
  float *a, *c, *Rc, *Sa, *WSa, *WSc, *da, *dc, *dRc, *dSa, *dWSa, *dWSc, *dv, *dsp0;
  float *gA, *gB, *gC;                 // generic [Rn, 64] backward temporaries
  float *colbuf;                       // [Rn, 5 * 50] im2col staging for conv1d weight gradients
  float *E1, *E1T, *O12, *dY12, *Yf, *dOf;   // chunk buffers (fp32); E1T: transposed copy, spectral path only
  __nv_bfloat16 *Yhi, *Ylo, *dOhi, *dOlo;
  TcState tc; L0Dense l0d;
  SpecState sp; YtcState ytc; int spec;              // use_tensor_cores == 2: e2e layer 1 in the frequency domain (spectral.cuh)
  float* loss;                         // device [8]: ce, node, spatial, kl_s, kl_g, kl_sg, dip
  long long global_iter;               // main.py:329 (only the capacity loss reads it)
  float *dipS, *dipG[3], *dipm[3], *dipv;   // DIP regulariser: mu^T mu scratch, d reg / d cov, batch means, m G
  void* zz_planes; size_t zz_cap;           // bf16 hi / lo planes of the last sndvae_inner_product_decode call (grow-only)
  float *tcp, *tcL[3], *tcd[3]; double* tcJ[3];   // total correlation: exp(-2 ls), per-latent / joint log-sum-exps, (dz, dmu, dls) scratch
  int* errflag;
  float* pinned_loss;
  // host-feed staging (sndvae_train_step_host)
  float *hf_features, *hf_adj, *hf_rel, *hf_adj_truth, *hf_feature_truth, *hf_spatial_truth, *hf_eps_s, *hf_eps_sg, *hf_eps_g;
  long long* hf_gen_adj;
  // compact host feeds: packed staging (bits / per-graph tensors) expanded on the device into the dense staging above
  float *hc_features, *hc_rel; uint32_t *hc_adj_bits, *hc_adjt_bits, *hc_gen_bits; int hc_ready;
  // gemm timing
  std::vector<EvPair> ev; size_t ev_used;
  StageTimer stt;
  // per-graph views: every buffer indexed by graph (or by sample) registers its per-graph byte size; view_shift moves all of
  // them at once so that the forward functions can run on a contiguous piece of the batch (pipelined host step)
  struct Shift { char** p; long long bytes; };
  std::vector<Shift> shifts;
  float* adam_alpha;                   // device [1]: this iteration's Adam step size (see adam.cuh)
  // small-problem path: the device-resident train step (memsets, ~300 kernels, Adam, loss read-back) captured once and replayed as
  // a CUDA graph while the caller's buffers stay the same (SNDVAE_GRAPH=0 disables; default: problems of < 2^21 edge cells)
  cudaGraphExec_t graph_exec; cudaStream_t gs; unsigned long long graph_key; int graph_mode; int capturing; long long graph_launches, graph_replays;
  float* gemm_ws; size_t gemm_ws_floats;   // workspace of the deterministic split-K products (tsgemm.cuh)
  int hf_ready;                        // every host-feed staging buffer is allocated
  int max_c;                           // widest node-level channel count of the config (sizes gA / gB / gC / colbuf)
  int poisoned;                        // a step failed half-way: arenas / losses are undefined until the next successful run
  // data-parallel communicator (sndvae_comm_init): NCCL resolved at run time from the process (torch's libnccl.so.2)
  void* comm; int rank, world;
  cudaStream_t cs, ds;                 // host-step copy streams (H2D feeds, D2H generated_adj)
  std::vector<cudaEvent_t> pev;        // per-piece events: 2 per piece (feeds landed, piece computed)
  cudaEvent_t ev_start;
};
static void view_shift(sndvae_handle* h, long long g) { for (auto& s : h->shifts) *s.p += g * s.bytes; }
template <typename T> static void reg_shift(sndvae_handle* h, T** p, long long elems_per_graph) {
  h->shifts.push_back({reinterpret_cast<char**>(p), elems_per_graph * (long long)sizeof(T)});
}
// optional per-stage timing (env SNDVAE_STAGE_TIMING=1): mark(h, "name") closes the previous stage
static void mark(sndvae_handle* h, const char* name) {
  StageTimer& t = h->stt;
  if (!t.on) return;
  if (t.used == t.ev.size()) { cudaEvent_t e; cudaEventCreate(&e); t.ev.push_back(e); t.name.push_back(name); }
  t.name[t.used] = name;
  cudaEventRecord(t.ev[t.used++], h->stream);
}
static void report_stages(sndvae_handle* h) {
  StageTimer& t = h->stt;
  if (!t.on || t.used < 2) { t.used = 0; return; }
  std::vector<std::pair<std::string, double>> acc;
  for (size_t i = 0; i + 1 < t.used; ++i) {
    float ms = 0; cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]);
    bool found = false;
    for (auto& a : acc) if (a.first == t.name[i]) { a.second += ms; found = true; break; }
    if (!found) acc.push_back({t.name[i], ms});
  }
  for (auto& a : acc) {          // running totals for sndvae_stage_times
    bool found = false;
    for (auto& b : t.total) if (b.first == a.first) { b.second += a.second; found = true; break; }
    if (!found) t.total.push_back(a);
  }
  t.steps++;
  if (t.print) {
    fprintf(stderr, "[sndvae stages]");
    for (auto& a : acc) fprintf(stderr, " %s=%.2fms", a.first.c_str(), a.second);
    fprintf(stderr, "\n");
  }
  t.used = 0;
}

static int fail(sndvae_t* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf;
  return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CKB(call) do { cudaError_t s_ = (call); if (s_ != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(s_), __FILE__, __LINE__); } while (0)
#define LAUNCH(k, grid, block, smem, ...) do { k<<<(grid), (block), (smem), h->stream>>>(__VA_ARGS__); h->launches++; } while (0)
#define LEW(k, n, ...) LAUNCH(k, cdiv((n), 256), 256, 0, __VA_ARGS__)   /* elementwise launch */

// ------------------------------------------------------------------------------------------
// parameter table: tf.trainable_variables() in creation order
// ------------------------------------------------------------------------------------------
static long add_param(sndvae_t* h, const char* name, int rank, int s0, int s1 = 1, int s2 = 1, int s3 = 1) {
  sndvae_param_info pi; memset(&pi, 0, sizeof pi);
  snprintf(pi.name, sizeof pi.name, "%s", name);
  pi.rank = rank; pi.shape[0] = s0; pi.shape[1] = s1; pi.shape[2] = s2; pi.shape[3] = s3;
  pi.size = (int64_t)s0 * s1 * s2 * s3;
  pi.offset = h->nparam;
  h->nparam += (pi.size + 3) / 4 * 4;     // 16-byte aligned entries
  h->table.push_back(pi);
  return (long)pi.offset;
}
static void add_bn(sndvae_t* h, const char* name, int c, long* g, long* b) {
  char buf[96];
  snprintf(buf, sizeof buf, "%s/gamma", name); *g = add_param(h, buf, 1, c);
  snprintf(buf, sizeof buf, "%s/beta", name);  *b = add_param(h, buf, 1, c);
}
static void add_lin(sndvae_t* h, const char* name, int i, int o, long* mb) {
  char buf[96];
  snprintf(buf, sizeof buf, "%s/Matrix", name); mb[0] = add_param(h, buf, 2, i, o);
  snprintf(buf, sizeof buf, "%s/bias", name);   mb[1] = add_param(h, buf, 1, o);
}
static void add_conv(sndvae_t* h, const char* name, int ci, int co, long* k, long* b) {
  char buf[96];
  snprintf(buf, sizeof buf, "%s/kernel", name); *k = add_param(h, buf, 3, KS, ci, co);
  snprintf(buf, sizeof buf, "%s/bias", name);   *b = add_param(h, buf, 1, co);
}

static void build_table(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  PT& p = h->pt;
  memset(&p, 0xff, sizeof p);
  h->nparam = 0;
  const int N = c.num_nodes, F = c.num_feature, D = c.spatial_dim, H = c.node_h_size;
  char nm[96];
  if (h->dis) {
    int ci = F;
    for (int i = 0; i < 2; ++i) {
      snprintf(nm, sizeof nm, "encoder/g_g%d_conv/w", i); p.gg_w[i] = add_param(h, nm, 2, ci, c.g_conv_hidden[i]);
      snprintf(nm, sizeof nm, "encoder/g_bn_g%d", i); add_bn(h, nm, c.g_conv_hidden[i], &p.gg_bng[i], &p.gg_bnb[i]);
      ci = c.g_conv_hidden[i] + F;
    }
    add_bn(h, "encoder/encoder_g", ci, &p.encg_g, &p.encg_b);
    add_lin(h, "encoder/g_g1_lin", N * ci, c.g_hidden_size, p.g_lin[0]);
    add_lin(h, "encoder/g_g2_lin", c.g_hidden_size, c.g_latent_size, p.g_lin[1]);
    add_lin(h, "encoder/g_g3_lin", c.g_hidden_size, c.g_latent_size, p.g_lin[2]);
    ci = D;
    for (int i = 0; i < 3; ++i) {
      snprintf(nm, sizeof nm, "encoder/g_s%d_conv", i + 1); add_conv(h, nm, ci, c.s_channel[i], &p.gs_k[i], &p.gs_b[i]);
      snprintf(nm, sizeof nm, "encoder/g_bn_s%d", i); add_bn(h, nm, c.s_channel[i], &p.gs_bng[i], &p.gs_bnb[i]);
      ci = c.s_channel[i];
    }
    add_bn(h, "encoder/encoder_s", ci, &p.encs_g, &p.encs_b);
    add_lin(h, "encoder/g_s1_lin", N * ci, c.s_hidden_size, p.s_lin[0]);
    add_lin(h, "encoder/g_s2_lin", c.s_hidden_size, c.s_latent_size, p.s_lin[1]);
    add_lin(h, "encoder/g_s3_lin", c.s_hidden_size, c.s_latent_size, p.s_lin[2]);
  }
  int ci = F;
  for (int i = 0; i < 2 && h->hops3; ++i) {          // layers.py:210-225
    const int* hs = c.sg_conv_hidden3[i];
    const int rows[4] = {4 * ci + 5, 3 * ci + 3 + hs[0], 2 * ci + 1 + hs[1], ci + hs[2]};
    for (int m = 0; m < 4; ++m) {
      snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/Matrix%d", i, m); p.sg3_M[i][m] = add_param(h, nm, 2, rows[m], hs[m]);
      snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/bias%d", i, m);   p.sg3_b[i][m] = add_param(h, nm, 1, hs[m]);
    }
    snprintf(nm, sizeof nm, "encoder/g_bn_sg%d", i); add_bn(h, nm, hs[3], &p.sg_bng[i], &p.sg_bnb[i]);
    ci = hs[3];
  }
  for (int i = 0; i < 2 && !h->hops3; ++i) {
    const int* hs = c.sg_conv_hidden[i];
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/Matrix1", i); p.sg_M1[i] = add_param(h, nm, 2, 3 * ci + 3, hs[0]);
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/bias1", i);   p.sg_b1[i] = add_param(h, nm, 1, hs[0]);
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/Matrix2", i); p.sg_M2[i] = add_param(h, nm, 2, 2 * ci + hs[0] + 1, hs[1]);
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/bias2", i);   p.sg_b2[i] = add_param(h, nm, 1, hs[1]);
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/Matrix3", i); p.sg_M3[i] = add_param(h, nm, 2, ci + hs[1], hs[2]);
    snprintf(nm, sizeof nm, "encoder/g_sg%d_conv/bias3", i);   p.sg_b3[i] = add_param(h, nm, 1, hs[2]);
    snprintf(nm, sizeof nm, "encoder/g_bn_sg%d", i); add_bn(h, nm, hs[2], &p.sg_bng[i], &p.sg_bnb[i]);
    ci = hs[2];
  }
  if (h->dis) add_bn(h, "encoder/encoder_sg", ci, &p.encsg_g, &p.encsg_b);
  add_lin(h, "encoder/g_sg1_lin", N * ci, c.sg_hidden_size, p.sg_lin[0]);
  add_lin(h, "encoder/g_sg2_lin", c.sg_hidden_size, c.sg_latent_size, p.sg_lin[1]);
  add_lin(h, "encoder/g_sg3_lin", c.sg_hidden_size, c.sg_latent_size, p.sg_lin[2]);
  add_lin(h, "decoder/d_sg_lin1", c.sg_latent_size, N * H, p.d_sg_lin1);
  if (h->dis) {
    add_lin(h, "decoder/d_s_lin1", c.s_latent_size, N * H, p.d_s_lin1);
    add_lin(h, "decoder/d_g_lin1", c.g_latent_size, N * H, p.d_g_lin1);
  }
  const int cin0 = h->dis ? 2 * H : H;
  auto spatial_dec = [&]() {
    int cc = cin0;
    for (int i = 0; i < 3; ++i) {
      snprintf(nm, sizeof nm, "decoder/s%d_deconv", i + 1); add_conv(h, nm, cc, c.s_d_channel[i], &p.s_k[i], &p.s_b[i]);
      snprintf(nm, sizeof nm, "decoder/d_bn_s%d", i); add_bn(h, nm, c.s_d_channel[i], &p.s_bng[i], &p.s_bnb[i]);
      cc = c.s_d_channel[i];
    }
    add_lin(h, "decoder/d_s_lin2", cc, D, p.d_s_lin2);
  };
  if (!h->dis) spatial_dec();       // model_joint.py builds the spatial decoder first (113-121)
  int cc = cin0;
  for (int i = 0; i < 2; ++i) {
    snprintf(nm, sizeof nm, "decoder/n%d_deconv", i); add_conv(h, nm, cc, c.n_d_channel[i], &p.n_k[i], &p.n_b[i]);
    snprintf(nm, sizeof nm, "decoder/d_bn_n%d", i); add_bn(h, nm, c.n_d_channel[i], &p.n_bng[i], &p.n_bnb[i]);
    cc = c.n_d_channel[i];
  }
  if (h->dis) add_bn(h, "decoder/decoder_node", cc, &p.decnode_g, &p.decnode_b);
  add_lin(h, "decoder/d_n_lin2", cc, F, p.d_n_lin2);
  cc = 2 * cin0;
  for (int i = 0; i < 2; ++i) {
    snprintf(nm, sizeof nm, "decoder/d_bn_e%d", i); add_bn(h, nm, cc, &p.e_bng[i], &p.e_bnb[i]);
    snprintf(nm, sizeof nm, "decoder/e%d_deconv/w1", i); p.e_w[i] = add_param(h, nm, 4, 1, N, cc, c.e_d_hidden[i]);
    snprintf(nm, sizeof nm, "decoder/e%d_deconv/biases1", i); p.e_b[i] = add_param(h, nm, 1, c.e_d_hidden[i]);
    cc = c.e_d_hidden[i];
  }
  if (h->dis) add_bn(h, "decoder/decoder_adj", cc, &p.decadj_g, &p.decadj_b);
  add_lin(h, "decoder/d_e_lin2", cc, 2, p.d_e_lin2);
  if (h->dis) spatial_dec();
}

// ------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------
template <typename T>
static int dalloc(sndvae_t* h, T** p, long long n) {
  if (n <= 0) n = 1;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, (size_t)n * sizeof(T));
  if (e != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "cudaMalloc(%lld bytes): %s", (long long)(n * sizeof(T)), cudaGetErrorString(e));
  h->allocs.push_back(q);
  e = cudaMemsetAsync(q, 0, (size_t)n * sizeof(T), h->stream);
  if (e != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "cudaMemsetAsync(%lld bytes): %s", (long long)(n * sizeof(T)), cudaGetErrorString(e));
  *p = (T*)q;
  return 0;
}
#define DA(ptr, n) do { int r_ = dalloc(h, &(ptr), (long long)(n)); if (r_) return r_; } while (0)

// forward activations: for `fwd_samples` samples (chunk-sized, or -- sgc_keep -- for the whole batch: then they are per-graph
// buffers registered for view_shift, `per_graph_samples` = S); backward temporaries: always for one chunk of `bwd_samples`
static int alloc_scratch(sndvae_t* h, SgcScratch& s, int C, const int* hs, long long fwd_samples, long long bwd_samples, int N, int per_graph_samples) {
  const long long rf = fwd_samples * N, rows = bwd_samples * N; const int KQ = 2 * C + 2, K2 = 2 * C + 2 + hs[0], K3 = C + hs[1] + 1;
  DA(s.xphi, rf * C); DA(s.coefQ, rf * KQ); DA(s.P, rf * hs[0]); DA(s.Qc, rf * hs[0]); DA(s.coef2, rf * K2);
  DA(s.m2s, rf * hs[1]); DA(s.coef3, rf * K3); DA(s.y, rf * hs[2]);
  if (per_graph_samples > 0) {
    const long long pg = (long long)per_graph_samples * N;
    reg_shift(h, &s.xphi, pg * C); reg_shift(h, &s.coefQ, pg * KQ); reg_shift(h, &s.P, pg * hs[0]); reg_shift(h, &s.Qc, pg * hs[0]);
    reg_shift(h, &s.coef2, pg * K2); reg_shift(h, &s.m2s, pg * hs[1]); reg_shift(h, &s.coef3, pg * K3); reg_shift(h, &s.y, pg * hs[2]);
  }
  DA(s.dcoef3, rows * K3); DA(s.dm2s, rows * hs[1]); DA(s.dcoef2, rows * K2); DA(s.dP, rows * hs[0]); DA(s.dQc, rows * hs[0]);
  DA(s.dxphi, rows * C); DA(s.dcoefQ, rows * KQ); DA(s.dpx, rows * C);
  DA(s.WQ, KQ * hs[0]); DA(s.W2, K2 * hs[1]); DA(s.W3, K3 * hs[2]); DA(s.dWQ, KQ * hs[0]); DA(s.w46, 2 * hs[0]);
  DA(s.dW2, K2 * hs[1]); DA(s.dW3, K3 * hs[2]);
  return 0;
}
// the scratch of layer l as seen by the chunk of samples starting at s0 (kernels index it with chunk-local sample numbers)
static SgcScratch sgc_view(sndvae_t* h, int l, long long s0) {
  SgcScratch S = l == 0 ? h->S0 : h->S1;
  if (h->sgc_keep) {
    const sndvae_config& c = h->cfg;
    const int C = l == 0 ? c.num_feature : c.sg_conv_hidden[0][2]; const int* hs = c.sg_conv_hidden[l];
    const int KQ = 2 * C + 2, K2 = 2 * C + 2 + hs[0], K3 = C + hs[1] + 1; const long long r0 = s0 * h->N;
    S.xphi += r0 * C; S.coefQ += r0 * KQ; S.P += r0 * hs[0]; S.Qc += r0 * hs[0]; S.coef2 += r0 * K2; S.m2s += r0 * hs[1];
    S.coef3 += r0 * K3; S.y += r0 * hs[2];
  }
  return S;
}

// DG: a buffer with `pg` elements per graph of the batch (registered for view_shift)
#define DG(ptr, pg) do { DA(ptr, B * (long long)(pg)); reg_shift(h, &(ptr), (long long)(pg)); } while (0)
static int alloc_buffers(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  const long long B = h->B;
  const int N = h->N, F = h->F, D = h->D, H = h->H, Chv = h->Chv, C1 = h->C1, C2 = h->C2, S = h->S;
  DA(h->P, h->nparam); DA(h->G, h->nparam); DA(h->M, h->nparam); DA(h->V, h->nparam);
  if (h->dis) {
    int g0 = c.g_conv_hidden[0], g1c = c.g_conv_hidden[1];
    DG(h->t0, N * g0); DG(h->c0, N * g0); DG(h->g1, N * (g0 + F)); DG(h->t1, N * g1c); DG(h->c1, N * g1c);
    DG(h->g2, N * (g1c + F)); DG(h->fg, N * (g1c + F));
    DG(h->hg, c.g_hidden_size); DG(h->mu_g, c.g_latent_size); DG(h->ls_g, c.g_latent_size);
    DG(h->h1p, N * c.s_channel[0]); DG(h->h1, N * c.s_channel[0]); DG(h->h2p, N * c.s_channel[1]); DG(h->h2, N * c.s_channel[1]);
    DG(h->h3p, N * c.s_channel[2]); DG(h->h3, N * c.s_channel[2]); DG(h->fs, N * c.s_channel[2]);
    DG(h->hs, c.s_hidden_size); DG(h->mu_s, c.s_latent_size); DG(h->ls_s, c.s_latent_size);
    DG(h->z_s, c.s_latent_size); DG(h->z_g, c.g_latent_size); DG(h->dz_s, c.s_latent_size); DG(h->dz_g, c.g_latent_size);
    DG(h->n_s, N * H); DG(h->n_g, N * H); DG(h->dn_s, N * H); DG(h->dn_g, N * H);
  }
  const int hl = c.sg_conv_hidden[1][2];
  DG(h->fsg, (long long)S * N * hl); DG(h->dfsg, (long long)S * N * hl);
  DG(h->hsg, S * c.sg_hidden_size); DG(h->mu_sg, S * c.sg_latent_size); DG(h->ls_sg, S * c.sg_latent_size);
  // edge storage for all samples, SGC activations for one chunk of SC samples
  SgcEdges& E = h->E; E.cap = c.edge_capacity;
  DG(E.rowstart, S * N); DG(E.rowcnt, S * N); DG(E.erow, (long long)S * E.cap); DG(E.ecol, (long long)S * E.cap);
  DG(E.ea, (long long)S * E.cap); DG(E.epr, (long long)S * E.cap); DG(E.eG, (long long)S * E.cap); DG(E.deg, S * N); DG(E.ssum, S * N); DG(E.nedges, S);
  const long long SC = h->SC;
  int r;
  if (h->hops3) {
    // SpatialGraphConvolution_3D: layer outputs before BN for every sample (per-graph views), one workspace slice per sample of a chunk
    const int h30 = c.sg_conv_hidden3[0][3], h31 = c.sg_conv_hidden3[1][3];
    Sgc3Dims d0 = {F, c.sg_conv_hidden3[0][0], c.sg_conv_hidden3[0][1], c.sg_conv_hidden3[0][2], h30};
    Sgc3Dims d1 = {h30, c.sg_conv_hidden3[1][0], c.sg_conv_hidden3[1][1], c.sg_conv_hidden3[1][2], h31};
    const long long w0 = sgc3_ws_floats(N, d0), w1 = sgc3_ws_floats(N, d1);
    h->ws3_stride = (w0 > w1 ? w0 : w1);
    long long sc3 = (2LL << 30) / (h->ws3_stride * 4); if (sc3 < 1) sc3 = 1; if (sc3 > h->BS) sc3 = h->BS; if (sc3 > 148 * 8) sc3 = 148 * 8;
    h->SC3 = sc3;
    DA(h->ws3, sc3 * h->ws3_stride);
    DG(h->y3[0], (long long)S * N * h30); DG(h->y3[1], (long long)S * N * h31);
    DG(h->x1, (long long)S * N * h30); DG(h->x2, (long long)S * N * h31);
    DA(h->dxa, h->BS * N * (h31 > h30 ? h31 : h30)); DA(h->dxb, h->BS * N * (h31 > h30 ? h31 : h30));
    h->sgc_keep = 1;
  } else
  {
    // keep the joint encoder's forward activations for every sample when they fit (626 floats per node at the synthetic2 sizes:
    // 26 GB at N=256, B=4096, S=10) -- the backward pass then re-uses them instead of recomputing each chunk's forward
    const int* h0s = c.sg_conv_hidden[0]; const int* h1s = c.sg_conv_hidden[1]; const int C1s = h0s[2];
    const long long fl = (F + (2 * F + 2) + 2 * h0s[0] + (2 * F + 2 + h0s[0]) + h0s[1] + (F + h0s[1] + 1) + h0s[2]) +
                         (C1s + (2 * C1s + 2) + 2 * h1s[0] + (2 * C1s + 2 + h1s[0]) + h1s[1] + (C1s + h1s[1] + 1) + h1s[2]) + h0s[2] + hl;
    const long long bytes = h->BS * N * fl * 4;
    h->sgc_keep = (bytes <= (32LL << 30)) && !(getenv("SNDVAE_SGC_KEEP") && atoi(getenv("SNDVAE_SGC_KEEP")) == 0);
  }
  if (!h->hops3) {
    const long long SF = h->sgc_keep ? h->BS : SC; const int pgs = h->sgc_keep ? S : 0;
    if ((r = alloc_scratch(h, h->S0, F, c.sg_conv_hidden[0], SF, SC, N, pgs))) return r;
    if ((r = alloc_scratch(h, h->S1, c.sg_conv_hidden[0][2], c.sg_conv_hidden[1], SF, SC, N, pgs))) return r;
    DA(h->x1, SF * N * c.sg_conv_hidden[0][2]); DA(h->x2, SF * N * hl); DA(h->dxa, SC * N * hl); DA(h->dxb, SC * N * hl);
    if (h->sgc_keep) { reg_shift(h, &h->x1, (long long)S * N * c.sg_conv_hidden[0][2]); reg_shift(h, &h->x2, (long long)S * N * hl); }
  }
  DG(h->z_sg, S * c.sg_latent_size); DG(h->zbar, c.sg_latent_size); DG(h->dzbar, c.sg_latent_size);
  long long maxL = c.sg_latent_size > c.sg_hidden_size ? c.sg_latent_size : c.sg_hidden_size;
  if (h->dis) { int m2 = c.s_latent_size > c.g_latent_size ? c.s_latent_size : c.g_latent_size; if (m2 > maxL) maxL = m2;
                if (c.s_hidden_size > maxL) maxL = c.s_hidden_size; if (c.g_hidden_size > maxL) maxL = c.g_hidden_size; }
  DG(h->dmu, S * maxL); DG(h->dls, S * maxL); DG(h->dh, S * maxL);
  DG(h->n_sg, N * H); DG(h->dn_sg, N * H);
  DG(h->v, N * Chv);
  if (h->dis) DG(h->sp0, N * Chv); else { h->sp0 = h->v; reg_shift(h, &h->sp0, (long long)N * Chv); }
  DG(h->q1p, N * c.n_d_channel[0]); DG(h->q1, N * c.n_d_channel[0]); DG(h->q2p, N * c.n_d_channel[1]); DG(h->q2, N * c.n_d_channel[1]);
  if (h->dis) DG(h->q3, N * c.n_d_channel[1]); else { h->q3 = h->q2; reg_shift(h, &h->q3, (long long)N * c.n_d_channel[1]); }
  DG(h->xpre, N * F); DG(h->xhat, N * F); DG(h->dxpre, N * F);
  DG(h->s1p, N * c.s_d_channel[0]); DG(h->s1, N * c.s_d_channel[0]); DG(h->s2p, N * c.s_d_channel[1]); DG(h->s2, N * c.s_d_channel[1]);
  DG(h->s3p, N * c.s_d_channel[2]); DG(h->s3, N * c.s_d_channel[2]); DG(h->ppre, N * D); DG(h->phat, N * D); DG(h->dppre, N * D);
  DG(h->a, N * Chv); DG(h->c, N * Chv); DG(h->Rc, N * C1); DG(h->Sa, N * C1);
  DA(h->WSa, (long long)N * C1 * Chv); DA(h->WSc, (long long)N * C1 * Chv);
  DG(h->da, N * Chv); DG(h->dc, N * Chv); DG(h->dRc, N * C1); DG(h->dSa, N * C1);
  DA(h->dWSa, (long long)N * C1 * Chv); DA(h->dWSc, (long long)N * C1 * Chv);
  DG(h->dv, N * Chv); DG(h->dsp0, N * Chv);
  // generic backward temporaries: [N, widest channel count] per graph, and at least one hidden vector per graph (the
  // latent heads park [B, hidden] gradients in gC)
  { long long pg = (long long)N * h->max_c;
    const int hid[3] = {c.s_hidden_size, c.g_hidden_size, c.sg_hidden_size};
    for (int i = 0; i < 3; ++i) if (hid[i] > pg) pg = hid[i];
    DG(h->gA, pg); DG(h->gB, pg); DG(h->gC, pg); DG(h->colbuf, (long long)N * KS * h->max_c); }
  const long long cells = (long long)h->Bc * N * N;
  if (c.use_tensor_cores == 2) {     // graph-tiled layouts: whole tiles of 128 graphs
    const long long tcells = (long long)((h->Bc + 127) / 128) * 128 * N * N;
    DA(h->E1, tcells * C1); DA(h->E1T, tcells * C1);
  } else { DA(h->E1, cells * C1); h->E1T = nullptr; }
  DA(h->O12, 2 * cells * C2); DA(h->dY12, 2 * cells * C1);
  if (c.use_tensor_cores == 2) {
    // spectral path: the bf16 planes only carry dE1 (layer-0 backward operands); dO stays fp32
    DA(h->Yhi, 2 * cells * TC_CP); DA(h->Ylo, 2 * cells * TC_CP); DA(h->dOf, 2 * cells * C2);
    h->Yf = nullptr; h->dOhi = h->dOlo = nullptr;
  } else if (c.use_tensor_cores) {
    DA(h->Yhi, 2 * cells * TC_CP); DA(h->Ylo, 2 * cells * TC_CP); DA(h->dOhi, 2 * cells * TC_OP); DA(h->dOlo, 2 * cells * TC_OP);
    h->Yf = nullptr; h->dOf = nullptr;
  } else {
    DA(h->Yf, 2 * cells * C1); DA(h->dOf, 2 * cells * C2);
    h->Yhi = h->Ylo = h->dOhi = h->dOlo = nullptr;
  }
  DA(h->loss, 8); DA(h->errflag, 1); DA(h->adam_alpha, 4);
  h->gemm_ws_floats = (size_t)16 << 20; DA(h->gemm_ws, h->gemm_ws_floats);      // 64 MB
  if (c.loss_variant == SNDVAE_LOSS_DIP) {
    const int Ls[3] = {c.s_latent_size, c.g_latent_size, c.sg_latent_size};
    int Lm = 0; for (int i = 0; i < 3; ++i) if (Ls[i] > Lm) Lm = Ls[i];
    DA(h->dipS, (long long)Lm * Lm); DA(h->dipv, Lm);
    for (int i = 0; i < 3; ++i) { DA(h->dipG[i], (long long)Ls[i] * Ls[i]); DA(h->dipm[i], Ls[i]); }
  }
  if (c.loss_variant == SNDVAE_LOSS_TC) {
    const int Ls[3] = {c.s_latent_size, c.g_latent_size, c.sg_latent_size};
    const long long rows[3] = {h->B, h->B, h->BS};
    long long mx = 0; for (int i = 0; i < 3; ++i) if (rows[i] * Ls[i] > mx) mx = rows[i] * Ls[i];
    DA(h->tcp, mx); for (int i = 0; i < 3; ++i) { DA(h->tcJ[i], rows[i]); DA(h->tcL[i], rows[i] * Ls[i]); DA(h->tcd[i], mx); }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// small host helpers
// ------------------------------------------------------------------------------------------
__global__ void bias_rows_k(float* __restrict__ C, const float* __restrict__ bias, long long rows, int cols) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * cols) C[idx] = bias[idx % cols];
}
// row-major C[M,N] = alpha op(A) op(B) + beta C (+ bias[N] in the epilogue): the hand-written tcgen05 kernels of tsgemm.cuh
// (fp32 operands split into three bf16 planes by the loaders, six products, fp32 accumulation in TMEM)
static cudaError_t gemm_rm(sndvae_t* h, bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda,
                           const float* B, int ldb, float beta, float* C, int ldc, const float* bias = nullptr) {
  return tsgemm(h->stream, tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, &h->launches, h->gemm_ws, h->gemm_ws_floats);
}
// Y[rows, o] = X[rows, i] W[i, o] + bias   (layers.py:566-576 on flattened features)
static int lin_fwd(sndvae_t* h, const float* X, const long* mb, float* Y, long long rows, int i, int o) {
  CKB(gemm_rm(h, false, false, (int)rows, o, i, 1.f, X, i, h->P + mb[0], o, 0.f, Y, o, h->P + mb[1]));
  return 0;
}
// dW += X^T dY; db += colsum(dY); dX = dY W^T (optional)
static int lin_bwd(sndvae_t* h, const float* X, const long* mb, const float* dY, float* dX, long long rows, int i, int o) {
  CKB(gemm_rm(h, true, false, i, o, (int)rows, 1.f, X, i, dY, o, 1.f, h->G + mb[0], o));
  LAUNCH(colsum_k, dim3(slab_grid(rows, XTDY_SLAB), cdiv(o, 128)), 128, 0, dY, o, h->G + mb[1], rows, o);
  if (dX) CKB(gemm_rm(h, false, true, (int)rows, i, o, 1.f, dY, o, h->P + mb[0], o, 0.f, dX, i));
  return 0;
}
static void bn_fwd(sndvae_t* h, const float* in, int ldi, long g, long b, float* out, int ldo, long long rows, int C, int act, int order) {
  LEW(bn_act_fwd_k, rows * C, in, ldi, g >= 0 ? h->P + g : nullptr, b >= 0 ? h->P + b : nullptr, out, ldo, rows, C, act, order);
}
static void bn_bwd(sndvae_t* h, const float* dout, int ldd, const float* in, int ldi, long g, long b, float* din, int ldn,
                   long long rows, int C, int act, int order) {
  LAUNCH(bn_act_bwd_k, slab_grid(rows, BN_SLAB), 64, 0, dout, ldd, in, ldi, g >= 0 ? h->P + g : nullptr, b >= 0 ? h->P + b : nullptr,
         din, ldn, g >= 0 ? h->G + g : nullptr, b >= 0 ? h->G + b : nullptr, rows, C, act, order);
}
// conv1d k5 SAME over the node axis (model.py:122,191,216) as im2col + one tall library GEMM
static int conv_fwd(sndvae_t* h, const float* in, long k, long b, float* out, long long rows, int Ci, int Co) {
  if (rows * Ci < 4096) { LEW(conv1d_fwd_k, rows * Co, in, h->P + k, h->P + b, out, rows, h->N, Ci, Co, KS); return 0; }
  LEW(im2col_k, rows * KS * Ci, in, h->colbuf, rows, h->N, Ci, KS);
  CKB(gemm_rm(h, false, false, (int)rows, Co, KS * Ci, 1.f, h->colbuf, KS * Ci, h->P + k, Co, 0.f, out, Co, h->P + b));
  return 0;
}
// weight/bias grads + optional input grad of a conv1d layer
static int conv_bwd(sndvae_t* h, const float* in, long k, long b, const float* dout, float* din, long long rows, int Ci, int Co) {
  LEW(im2col_k, rows * KS * Ci, in, h->colbuf, rows, h->N, Ci, KS);
  CKB(gemm_rm(h, true, false, KS * Ci, Co, (int)rows, 1.f, h->colbuf, KS * Ci, dout, Co, 1.f, h->G + k, Co));
  LAUNCH(colsum_k, dim3(slab_grid(rows, XTDY_SLAB), cdiv(Co, 128)), 128, 0, dout, Co, h->G + b, rows, Co);
  if (din) {      // dcol = dout . K^T (re-using the im2col buffer), then the transposed gather
    CKB(gemm_rm(h, false, true, (int)rows, KS * Ci, Co, 1.f, dout, Co, h->P + k, Co, 0.f, h->colbuf, KS * Ci));
    LEW(col2im_k, rows * Ci, h->colbuf, din, rows, h->N, Ci, KS);
  }
  return 0;
}
// dW[K, hcols] += coef^T grad: the gradient of the packed block [M; b] (the last coefficient column multiplies the bias row);
// scattered into the arena once per step by sgc_unpack_w23_k
static int coef_grad(sndvae_t* h, const float* coef, int K, const float* grad, int hcols, long long rows, float* dW) {
  CKB(gemm_rm(h, true, false, K, hcols, (int)rows, 1.f, coef, K, grad, hcols, 1.f, dW, hcols));
  return 0;
}
static SgcDims sgc_dims(sndvae_t* h, int l) {
  const sndvae_config& c = h->cfg;
  SgcDims d; d.C = l == 0 ? c.num_feature : c.sg_conv_hidden[0][2];
  d.h0 = c.sg_conv_hidden[l][0]; d.h1 = c.sg_conv_hidden[l][1]; d.h2 = c.sg_conv_hidden[l][2];
  return d;
}
// assemble [M1b; M1c; w5; b1], [M2; b2], [M3; b3] for both layers (once per step)
static void sgc_pack(sndvae_t* h) {
  const PT& p = h->pt;
  for (int l = 0; l < 2; ++l) {
    SgcDims d = sgc_dims(h, l); SgcScratch& S = l == 0 ? h->S0 : h->S1;
    int n = (2 * d.C + 2) * d.h0 + (2 * d.C + 2 + d.h0) * d.h1 + (d.C + d.h1 + 1) * d.h2;
    LEW(sgc_pack_weights_k, n, h->P + p.sg_M1[l], h->P + p.sg_b1[l], h->P + p.sg_M2[l], h->P + p.sg_b2[l], h->P + p.sg_M3[l],
        h->P + p.sg_b3[l], S.WQ, S.W2, S.W3, d);
  }
}
// forward of SGC layer l for `ns` samples starting at global sample s0; x: [ns*N, C]; result in S.y
static int sgc_layer_fwd(sndvae_t* h, int l, const float* x, long long s0, long long ns) {
  const PT& p = h->pt; SgcDims d = sgc_dims(h, l); const SgcScratch S = sgc_view(h, l, s0);
  const int N = h->N, C = d.C, KQ = 2 * C + 2, K2 = 2 * C + 2 + d.h0, K3 = C + d.h1 + 1; const int rows = (int)(ns * N);
  const float* M1 = h->P + p.sg_M1[l];
  LAUNCH(sgc_prep_k, (unsigned)ns, 256, 0, x, h->E, d, S, N, s0);
  CKB(gemm_rm(h, false, false, rows, d.h0, C, 1.f, S.xphi, C, M1, d.h0, 0.f, S.P, d.h0));
  CKB(gemm_rm(h, false, false, rows, d.h0, KQ, 1.f, S.coefQ, KQ, S.WQ, d.h0, 0.f, S.Qc, d.h0));
  LAUNCH(sgc_edge_fwd_k, (unsigned)ns, 256, 0, h->E, d, S, M1 + (size_t)(3 * C) * d.h0, M1 + (size_t)(3 * C + 2) * d.h0, N, s0);
  CKB(gemm_rm(h, false, false, rows, d.h1, K2, 1.f, S.coef2, K2, S.W2, d.h1, 0.f, S.m2s, d.h1));
  LEW(sgc_cat_k, (long long)rows * K3, S.xphi, S.m2s, S.coef3, (long long)rows, C, d.h1);
  CKB(gemm_rm(h, false, false, rows, d.h2, K3, 1.f, S.coef3, K3, S.W3, d.h2, 0.f, S.y, d.h2));
  return 0;
}
// backward of SGC layer l (activations of the chunk must be in S): dy [ns*N, h2] -> parameter gradients and,
// when dx != NULL, dx [ns*N, C]
static int sgc_layer_bwd(sndvae_t* h, int l, const float* x, const float* dy, float* dx, long long s0, long long ns) {
  const PT& p = h->pt; SgcDims d = sgc_dims(h, l); const SgcScratch S = sgc_view(h, l, s0);
  const int N = h->N, C = d.C, KQ = 2 * C + 2, K2 = 2 * C + 2 + d.h0, K3 = C + d.h1 + 1; const int rows = (int)(ns * N);
  const float* M1 = h->P + p.sg_M1[l];
  int r;
  CKB(gemm_rm(h, false, true, rows, K3, d.h2, 1.f, dy, d.h2, S.W3, d.h2, 0.f, S.dcoef3, K3));
  LEW(sgc_cat_bwd_k, (long long)rows * (C + d.h1), S.dcoef3, x, S.m2s, dx, S.dm2s, (long long)rows, C, d.h1);
  CKB(gemm_rm(h, false, true, rows, K2, d.h1, 1.f, S.dm2s, d.h1, S.W2, d.h1, 0.f, S.dcoef2, K2));
  CK(cudaMemsetAsync(S.dQc, 0, sizeof(float) * (size_t)rows * d.h0, h->stream));
  LAUNCH(sgc_edge_bwd_k, (unsigned)ns, 256, sizeof(float) * 2 * d.h0, h->E, d, S, M1 + (size_t)(3 * C) * d.h0, M1 + (size_t)(3 * C + 2) * d.h0, N, s0);
  // parameter gradients: coefficient rows (transposed) times the gradient rows
  if ((r = coef_grad(h, S.coef3, K3, dy, d.h2, rows, S.dW3))) return r;
  if ((r = coef_grad(h, S.coef2, K2, S.dm2s, d.h1, rows, S.dW2))) return r;
  CKB(gemm_rm(h, true, false, C, d.h0, rows, 1.f, S.xphi, C, S.dP, d.h0, 1.f, h->G + p.sg_M1[l], d.h0));       // dM1a
  CKB(gemm_rm(h, true, false, KQ, d.h0, rows, 1.f, S.coefQ, KQ, S.dQc, d.h0, 1.f, S.dWQ, d.h0));                // d[M1b; M1c; w5; b1]
  if (dx) {
    CKB(gemm_rm(h, false, true, rows, C, d.h0, 1.f, S.dP, d.h0, M1, d.h0, 0.f, S.dxphi, C));
    CKB(gemm_rm(h, false, true, rows, KQ, d.h0, 1.f, S.dQc, d.h0, S.WQ, d.h0, 0.f, S.dcoefQ, KQ));
    LAUNCH(sgc_node_bwd_k, (unsigned)ns, 256, 0, x, dx, h->E, d, S, N, s0);
  }
  return 0;
}
static void ev_begin(sndvae_t* h, double flops) {
  if (h->capturing) return;
  if (h->ev_used < h->ev.size()) { h->ev[h->ev_used].flops = flops; cudaEventRecord(h->ev[h->ev_used].a, h->stream); }
}
static void ev_end(sndvae_t* h) {
  if (h->capturing) return;
  if (h->ev_used < h->ev.size()) { cudaEventRecord(h->ev[h->ev_used].b, h->stream); h->ev_used++; }
}

// ------------------------------------------------------------------------------------------
// encoder  (model.py:98-151 / model_joint.py:72-85)
// ------------------------------------------------------------------------------------------
// SGC forward for samples [s0, s0+ns); writes fsg rows s0.. ; leaves the chunk's activations in scratch
static int sgc_chunk_fwd(sndvae_t* h, const sndvae_inputs* in, long long s0, long long ns) {
  const sndvae_config& c = h->cfg; const PT& p = h->pt; const int N = h->N, F = h->F;
  const int h02 = c.sg_conv_hidden[0][2], h12 = c.sg_conv_hidden[1][2];
  const float* x0 = in->features + s0 * N * F;
  int r;
  float* x1 = h->x1 + (h->sgc_keep ? s0 * N * h02 : 0); float* x2 = h->x2 + (h->sgc_keep ? s0 * N * h12 : 0);
  if ((r = sgc_layer_fwd(h, 0, x0, s0, ns))) return r;
  bn_fwd(h, sgc_view(h, 0, s0).y, h02, p.sg_bng[0], p.sg_bnb[0], x1, h02, ns * N, h02, ACT_LRELU, 0);
  if ((r = sgc_layer_fwd(h, 1, x1, s0, ns))) return r;
  bn_fwd(h, sgc_view(h, 1, s0).y, h12, p.sg_bng[1], p.sg_bnb[1], x2, h12, ns * N, h12, ACT_LRELU, 0);
  // encoder_sg BN (model.py:148; absent in model_joint.py:83)
  bn_fwd(h, x2, h12, h->dis ? p.encsg_g : -1, h->dis ? p.encsg_b : -1, h->fsg + s0 * N * h12, h12, ns * N, h12, ACT_NONE, 0);
  return 0;
}

// ---- SpatialGraphConvolution_3D (sgc3.cuh) ---------------------------------------------------------------------------------
static Sgc3Dims sgc3_dims(sndvae_t* h, int l) {
  const sndvae_config& c = h->cfg; const int* hs = c.sg_conv_hidden3[l];
  Sgc3Dims d = {l == 0 ? c.num_feature : c.sg_conv_hidden3[0][3], hs[0], hs[1], hs[2], hs[3]};
  return d;
}
static Sgc3Params sgc3_params(sndvae_t* h, int l, float* arena) {
  const PT& p = h->pt;
  Sgc3Params w = {arena + p.sg3_M[l][0], arena + p.sg3_b[l][0], arena + p.sg3_M[l][1], arena + p.sg3_b[l][1],
                  arena + p.sg3_M[l][2], arena + p.sg3_b[l][2], arena + p.sg3_M[l][3], arena + p.sg3_b[l][3]};
  return w;
}
// layer l forward for every sample of the current view: x [BS, N, C] -> y3[l] (before BN), in chunks of SC3 samples
static int sgc3_layer_fwd(sndvae_t* h, int l, const float* x, const sndvae_inputs* in) {
  const Sgc3Dims d = sgc3_dims(h, l); const int N = h->N;
  for (long long s0 = 0; s0 < h->BS; s0 += h->SC3) {
    const long long ns = h->BS - s0 < h->SC3 ? h->BS - s0 : h->SC3;
    LAUNCH(sgc3_k<false>, (unsigned)ns, 256, 0, x + s0 * N * d.C, in->adj + s0 * N * N, in->rel + s0 * N * N, sgc3_params(h, l, h->P),
           sgc3_params(h, l, h->G), d, N, h->y3[l] + s0 * N * d.h3, (const float*)nullptr, (float*)nullptr, h->ws3, h->ws3_stride);
  }
  CK(cudaGetLastError());
  return 0;
}
static int sgc3_layer_bwd(sndvae_t* h, int l, const float* x, const float* dy, float* dx, const sndvae_inputs* in) {
  const Sgc3Dims d = sgc3_dims(h, l); const int N = h->N;
  for (long long s0 = 0; s0 < h->BS; s0 += h->SC3) {
    const long long ns = h->BS - s0 < h->SC3 ? h->BS - s0 : h->SC3;
    LAUNCH(sgc3_k<true>, (unsigned)ns, 256, 0, x + s0 * N * d.C, in->adj + s0 * N * N, in->rel + s0 * N * N, sgc3_params(h, l, h->P),
           sgc3_params(h, l, h->G), d, N, (float*)nullptr, dy + s0 * N * d.h3, dx ? dx + s0 * N * d.C : nullptr, h->ws3, h->ws3_stride);
  }
  CK(cudaGetLastError());
  return 0;
}
static int sgc3_encoder_fwd(sndvae_t* h, const sndvae_inputs* in) {
  const PT& p = h->pt; const int N = h->N; const long long BS = h->BS;
  const int h30 = h->cfg.sg_conv_hidden3[0][3], h31 = h->cfg.sg_conv_hidden3[1][3];
  int r;
  if ((r = sgc3_layer_fwd(h, 0, in->features, in))) return r;
  bn_fwd(h, h->y3[0], h30, p.sg_bng[0], p.sg_bnb[0], h->x1, h30, BS * N, h30, ACT_LRELU, 0);
  if ((r = sgc3_layer_fwd(h, 1, h->x1, in))) return r;
  bn_fwd(h, h->y3[1], h31, p.sg_bng[1], p.sg_bnb[1], h->x2, h31, BS * N, h31, ACT_LRELU, 0);
  bn_fwd(h, h->x2, h31, h->dis ? p.encsg_g : -1, h->dis ? p.encsg_b : -1, h->fsg, h31, BS * N, h31, ACT_NONE, 0);
  return 0;
}
// dfsg -> parameter gradients of both layers (the feeds need no gradient)
static int sgc3_encoder_bwd(sndvae_t* h, const sndvae_inputs* in) {
  const PT& p = h->pt; const int N = h->N; const long long BS = h->BS;
  const int h30 = h->cfg.sg_conv_hidden3[0][3], h31 = h->cfg.sg_conv_hidden3[1][3];
  int r;
  bn_bwd(h, h->dfsg, h31, h->x2, h31, h->dis ? p.encsg_g : -1, h->dis ? p.encsg_b : -1, h->dxa, h31, BS * N, h31, ACT_NONE, 0);
  bn_bwd(h, h->dxa, h31, h->y3[1], h31, p.sg_bng[1], p.sg_bnb[1], h->dxa, h31, BS * N, h31, ACT_LRELU, 0);      // dy1
  if ((r = sgc3_layer_bwd(h, 1, h->x1, h->dxa, h->dxb, in))) return r;                                           // dx1
  bn_bwd(h, h->dxb, h30, h->y3[0], h30, p.sg_bng[0], p.sg_bnb[0], h->dxb, h30, BS * N, h30, ACT_LRELU, 0);      // dy0
  return sgc3_layer_bwd(h, 0, in->features, h->dxb, nullptr, in);
}

static int encoder_fwd(sndvae_t* h, const sndvae_inputs* in) {
  const sndvae_config& c = h->cfg; const PT& p = h->pt;
  const int N = h->N, F = h->F, D = h->D; const long long B = h->B, BS = h->BS, Rn = h->Rn;
  if (h->dis) {
    // graph encoder (model.py:104-115): g <- BN(lrelu(A (g w))); g <- [g || X]
    int g0 = c.g_conv_hidden[0], g1c = c.g_conv_hidden[1];
    LEW(rowlin_fwd_k, Rn * g0, in->feature_truth, F, h->P + p.gg_w[0], (const float*)nullptr, h->t0, g0, Rn, F, g0, ACT_NONE);
    LAUNCH(graph_prop_fwd_k<32>, cdiv(Rn * 32, 256), 256, 0, in->adj_truth, h->t0, h->c0, Rn, N, g0);
    bn_fwd(h, h->c0, g0, p.gg_bng[0], p.gg_bnb[0], h->g1, g0 + F, Rn, g0, ACT_LRELU, 1);
    LEW(copy_cols_k, Rn * F, in->feature_truth, F, 0, h->g1, g0 + F, g0, Rn, F, 0);
    LEW(rowlin_fwd_k, Rn * g1c, h->g1, g0 + F, h->P + p.gg_w[1], (const float*)nullptr, h->t1, g1c, Rn, g0 + F, g1c, ACT_NONE);
    LAUNCH(graph_prop_fwd_k<32>, cdiv(Rn * 32, 256), 256, 0, in->adj_truth, h->t1, h->c1, Rn, N, g1c);
    bn_fwd(h, h->c1, g1c, p.gg_bng[1], p.gg_bnb[1], h->g2, g1c + F, Rn, g1c, ACT_LRELU, 1);
    LEW(copy_cols_k, Rn * F, in->feature_truth, F, 0, h->g2, g1c + F, g1c, Rn, F, 0);
    bn_fwd(h, h->g2, g1c + F, p.encg_g, p.encg_b, h->fg, g1c + F, Rn, g1c + F, ACT_NONE, 0);
    int r;
    if ((r = lin_fwd(h, h->fg, p.g_lin[0], h->hg, B, N * (g1c + F), c.g_hidden_size))) return r;
    if ((r = lin_fwd(h, h->hg, p.g_lin[1], h->mu_g, B, c.g_hidden_size, c.g_latent_size))) return r;
    if ((r = lin_fwd(h, h->hg, p.g_lin[2], h->ls_g, B, c.g_hidden_size, c.g_latent_size))) return r;
    // spatial encoder (model.py:119-129): relu(BN(conv1d k5 SAME)) x3
    const int* sc = c.s_channel;
    if ((r = conv_fwd(h, in->spatial_truth, p.gs_k[0], p.gs_b[0], h->h1p, Rn, D, sc[0]))) return r;
    bn_fwd(h, h->h1p, sc[0], p.gs_bng[0], p.gs_bnb[0], h->h1, sc[0], Rn, sc[0], ACT_RELU, 0);
    if ((r = conv_fwd(h, h->h1, p.gs_k[1], p.gs_b[1], h->h2p, Rn, sc[0], sc[1]))) return r;
    bn_fwd(h, h->h2p, sc[1], p.gs_bng[1], p.gs_bnb[1], h->h2, sc[1], Rn, sc[1], ACT_RELU, 0);
    if ((r = conv_fwd(h, h->h2, p.gs_k[2], p.gs_b[2], h->h3p, Rn, sc[1], sc[2]))) return r;
    bn_fwd(h, h->h3p, sc[2], p.gs_bng[2], p.gs_bnb[2], h->h3, sc[2], Rn, sc[2], ACT_RELU, 0);
    bn_fwd(h, h->h3, sc[2], p.encs_g, p.encs_b, h->fs, sc[2], Rn, sc[2], ACT_NONE, 0);
    if ((r = lin_fwd(h, h->fs, p.s_lin[0], h->hs, B, N * sc[2], c.s_hidden_size))) return r;
    if ((r = lin_fwd(h, h->hs, p.s_lin[1], h->mu_s, B, c.s_hidden_size, c.s_latent_size))) return r;
    if ((r = lin_fwd(h, h->hs, p.s_lin[2], h->ls_s, B, c.s_hidden_size, c.s_latent_size))) return r;
  }
  if (h->hops3) {      // model.py:139-140: SpatialGraphConvolution_3D x2 on the dense sampled adjacencies
    mark(h, "sgc_fwd");
    int r3 = sgc3_encoder_fwd(h, in); if (r3) return r3;
  } else {
  // joint encoder (model.py:134-151): edge lists once per step, SGC x2 in sample chunks
  mark(h, "sgc_edges");
  LAUNCH(sgc_build_edges_k, (unsigned)BS, 256, 0, in->adj, in->rel, h->E, N, h->errflag);
  mark(h, "sgc_fwd");
  sgc_pack(h);
  for (long long s0 = 0; s0 < BS; s0 += h->SC) {
    long long ns = BS - s0 < h->SC ? BS - s0 : h->SC;
    int r = sgc_chunk_fwd(h, in, s0, ns); if (r) return r;
  }
  }
  mark(h, "enc_heads");
  const int h12 = c.sg_conv_hidden[1][2];
  int r;
  if ((r = lin_fwd(h, h->fsg, p.sg_lin[0], h->hsg, BS, N * h12, c.sg_hidden_size))) return r;
  if ((r = lin_fwd(h, h->hsg, p.sg_lin[1], h->mu_sg, BS, c.sg_hidden_size, c.sg_latent_size))) return r;
  if ((r = lin_fwd(h, h->hsg, p.sg_lin[2], h->ls_sg, BS, c.sg_hidden_size, c.sg_latent_size))) return r;
  return 0;
}

// get_z (model.py:153-161) + KL sums (optimizer.py:160-162)
static int reparam_fwd(sndvae_t* h, const sndvae_noise* nz) {
  const sndvae_config& c = h->cfg;
  if (h->dis) {
    LEW(reparam_kl_k, h->B * c.s_latent_size, h->mu_s, h->ls_s, nz->eps_s, h->z_s, h->loss + 3, h->B * c.s_latent_size);
    LEW(reparam_kl_k, h->B * c.g_latent_size, h->mu_g, h->ls_g, nz->eps_g, h->z_g, h->loss + 4, h->B * c.g_latent_size);
  }
  LEW(reparam_kl_k, h->BS * c.sg_latent_size, h->mu_sg, h->ls_sg, nz->eps_sg, h->z_sg, h->loss + 5, h->BS * c.sg_latent_size);
  return 0;
}

// ------------------------------------------------------------------------------------------
// decoder  (model.py:172-222 / model_joint.py:94-182)
// ------------------------------------------------------------------------------------------
static int decoder_fwd(sndvae_t* h, const sndvae_inputs* in, sndvae_outputs* out, bool backward, float gB) {
  const sndvae_config& c = h->cfg; const PT& p = h->pt;
  const int N = h->N, F = h->F, D = h->D, H = h->H, Chv = h->Chv, C1 = h->C1, C2 = h->C2, S = h->S;
  const long long B = h->B, Rn = h->Rn;
  const int dact = h->dis ? ACT_NONE : ACT_LRELU;      // model_joint.py:116,139 lrelu after BN
  int r;
  mark(h, "dec_nodes");
  // latent -> node features; the S-mean (model.py:180) is hoisted before the linear map
  LEW(smean_k, B * c.sg_latent_size, h->z_sg, h->zbar, B, S, c.sg_latent_size);
  if ((r = lin_fwd(h, h->zbar, p.d_sg_lin1, h->n_sg, B, c.sg_latent_size, N * H))) return r;
  if (h->dis) {
    if ((r = lin_fwd(h, h->z_s, p.d_s_lin1, h->n_s, B, c.s_latent_size, N * H))) return r;
    if ((r = lin_fwd(h, h->z_g, p.d_g_lin1, h->n_g, B, c.g_latent_size, N * H))) return r;
    LEW(copy_cols_k, Rn * H, h->n_sg, H, 0, h->v, Chv, 0, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->n_g, H, 0, h->v, Chv, H, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->n_sg, H, 0, h->sp0, Chv, 0, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->n_s, H, 0, h->sp0, Chv, H, Rn, H, 0);
  } else {
    LEW(copy_cols_k, Rn * H, h->n_sg, H, 0, h->v, Chv, 0, Rn, H, 0);
  }
  // node-feature decoder (model.py:186-194)
  const int* nc = c.n_d_channel;
  if ((r = conv_fwd(h, h->v, p.n_k[0], p.n_b[0], h->q1p, Rn, Chv, nc[0]))) return r;
  bn_fwd(h, h->q1p, nc[0], p.n_bng[0], p.n_bnb[0], h->q1, nc[0], Rn, nc[0], dact, 0);
  if ((r = conv_fwd(h, h->q1, p.n_k[1], p.n_b[1], h->q2p, Rn, nc[0], nc[1]))) return r;
  bn_fwd(h, h->q2p, nc[1], p.n_bng[1], p.n_bnb[1], h->q2, nc[1], Rn, nc[1], dact, 0);
  if (h->dis) bn_fwd(h, h->q2, nc[1], p.decnode_g, p.decnode_b, h->q3, nc[1], Rn, nc[1], ACT_NONE, 0);
  LEW(rowlin_fwd_k, Rn * F, h->q3, nc[1], h->P + p.d_n_lin2[0], h->P + p.d_n_lin2[1], h->xpre, F, Rn, nc[1], F, ACT_NONE);
  LEW(sigmoid_mse_k, Rn * F, h->xpre, in ? in->feature_truth : nullptr, h->xhat, backward ? h->dxpre : nullptr, h->loss + 1,
         Rn * F, 1.f / (gB * N * F));
  // spatial decoder (model.py:213-219)
  const int* sc = c.s_d_channel;
  if ((r = conv_fwd(h, h->sp0, p.s_k[0], p.s_b[0], h->s1p, Rn, Chv, sc[0]))) return r;
  bn_fwd(h, h->s1p, sc[0], p.s_bng[0], p.s_bnb[0], h->s1, sc[0], Rn, sc[0], dact, 0);
  if ((r = conv_fwd(h, h->s1, p.s_k[1], p.s_b[1], h->s2p, Rn, sc[0], sc[1]))) return r;
  bn_fwd(h, h->s2p, sc[1], p.s_bng[1], p.s_bnb[1], h->s2, sc[1], Rn, sc[1], dact, 0);
  if ((r = conv_fwd(h, h->s2, p.s_k[2], p.s_b[2], h->s3p, Rn, sc[1], sc[2]))) return r;
  bn_fwd(h, h->s3p, sc[2], p.s_bng[2], p.s_bnb[2], h->s3, sc[2], Rn, sc[2], dact, 0);
  LEW(rowlin_fwd_k, Rn * D, h->s3, sc[2], h->P + p.d_s_lin2[0], h->P + p.d_s_lin2[1], h->ppre, D, Rn, sc[2], D, ACT_NONE);
  LEW(sigmoid_mse_k, Rn * D, h->ppre, in ? in->spatial_truth : nullptr, h->phat, backward ? h->dppre : nullptr, h->loss + 2,
         Rn * D, 1.f / (gB * N * D));
  if (out && out->generated_node_feat) CK(cudaMemcpyAsync(out->generated_node_feat, h->xhat, sizeof(float) * Rn * F, cudaMemcpyDeviceToDevice, h->stream));
  if (out && out->generated_spatial) CK(cudaMemcpyAsync(out->generated_spatial, h->phat, sizeof(float) * Rn * D, cudaMemcpyDeviceToDevice, h->stream));

  // ---- edge decoder (model.py:196-208) -------------------------------------------------
  mark(h, "l0_vec");
  const int Ctot = 2 * Chv;
  const float* w0 = h->P + p.e_w[0];
  bn_fwd(h, h->v, Chv, p.e_bng[0], p.e_bnb[0], h->a, Chv, Rn, Chv, ACT_RELU, 0);
  bn_fwd(h, h->v, Chv, p.e_bng[0] + Chv, p.e_bnb[0] + Chv, h->c, Chv, Rn, Chv, ACT_RELU, 0);
  const bool tc = c.use_tensor_cores != 0;
  if (tc) {
    // Rc = c . Toeplitz(w0[:, Ch:, :]),  Sa = a . Toeplitz(w0[:, :Ch, :])  on the tensor cores (split-bf16)
    TcState& T = h->tc;
    if (tc_split(h->a, T.ah, T.al, Rn, Chv, T.l0a.CSi, h->stream) || tc_split(h->c, T.ch, T.cl, Rn, Chv, T.l0c.CSi, h->stream) ||
        tc_plan_fwd(T.l0c, T.ch, T.cl, h->Rc, B, B, 0, h->stream) || tc_plan_fwd(T.l0a, T.ah, T.al, h->Sa, B, B, 0, h->stream))
      return fail(h, SNDVAE_E_CUDA, "tensor-core layer-0 products: %s", tc_last_error());
    h->launches += 4;
  } else {
    LEW(toep_vec_fwd_k, Rn * C1, h->c, w0, h->Rc, B, N, Ctot, Chv, Chv, C1);
    LEW(toep_vec_fwd_k, Rn * C1, h->a, w0, h->Sa, B, N, Ctot, 0, Chv, C1);
  }
  const double f1 = 2.0 * 2.0 * N * ((double)N * N - (double)((N - 1) / 2) * ((N - 1) / 2 + 1) / 2.0 -
                                     (double)(N - 1 - (N - 1) / 2) * (N - (N - 1) / 2) / 2.0) * C1 * C2;
  for (long long b0 = 0; b0 < B; b0 += h->Bc) {
    const int bc = (int)(B - b0 < h->Bc ? B - b0 : h->Bc);
    const long long rows = (long long)bc * N, cells = rows * N;
    mark(h, "y_producer");
    YOut Y; Y.E1 = h->E1; Y.Yf = h->Yf; Y.Yhi = h->Yhi; Y.Ylo = h->Ylo; Y.CP = tc ? h->tc.l1.CSi : C1; Y.bf16 = tc && !h->spec;
    if (h->spec) {
      // batch-major tcgen05 form (spectral.cuh): only E1 is written; Y = relu(BN(E1)) is applied inside the forward FFT
      TcState& T = h->tc; const long long po = b0 * N * h->ytc.CS;
      // E1 (rows (b,i) sweep j) and its transpose E1T (rows (b,j) sweep i): the same kernel with the roles of a / c swapped
      if (ytc_run(h->ytc, T.ah + po, T.al + po, T.ch + po, T.cl + po, 0, h->Rc + b0 * N * C1, h->Sa + b0 * N * C1, h->P + p.e_b[0], h->E1, bc, h->stream) ||
          ytc_run(h->ytc, T.ch + po, T.cl + po, T.ah + po, T.al + po, 1, h->Sa + b0 * N * C1, h->Rc + b0 * N * C1, h->P + p.e_b[0], h->E1T, bc, h->stream))
        return fail(h, SNDVAE_E_CUDA, "y_producer_tc: %s", tc_last_error());
      h->launches += 2;
    } else {
      // E1 / Y on the fp32 pipes (register-resident WS rows).  Routing the two K = 2H products through the tensor-core
      // kernel was measured slower (44 vs 31 ms per 512 graphs): with a single K chunk its epilogue dominates.
      dim3 yg(cdiv(N, YP_TJ), N);
      if (Chv == 40) LAUNCH(y_producer_k<40>, yg, YP_THREADS, 0, h->a + b0 * N * Chv, h->c + b0 * N * Chv, h->WSa, h->WSc, h->Rc + b0 * N * C1,
                            h->Sa + b0 * N * C1, h->P + p.e_b[0], h->P + p.e_bng[1], h->P + p.e_bnb[1], Y, bc, N, C1);
      else if (Chv == 20) LAUNCH(y_producer_k<20>, yg, YP_THREADS, 0, h->a + b0 * N * Chv, h->c + b0 * N * Chv, h->WSa, h->WSc, h->Rc + b0 * N * C1,
                  h->Sa + b0 * N * C1, h->P + p.e_b[0], h->P + p.e_bng[1], h->P + p.e_bnb[1], Y, bc, N, C1);
      else LEW(y_producer_generic_k, cells * C1, h->a + b0 * N * Chv, h->c + b0 * N * Chv, h->WSa, h->WSc, h->Rc + b0 * N * C1,
               h->Sa + b0 * N * C1, h->P + p.e_b[0], h->P + p.e_bng[1], h->P + p.e_bnb[1], Y, bc, N, C1, Chv);
    }
    mark(h, "gemm_fwd");
    ev_begin(h, f1 * bc);
    if (h->spec) { if (spec_forward(h->sp, h->E1, h->E1T, h->P + p.e_bng[1], h->P + p.e_bnb[1], h->Sa + b0 * N * C1, h->Rc + b0 * N * C1, h->P + p.e_b[0], rows, h->O12, h->stream)) return fail(h, SNDVAE_E_CUDA, "spectral fwd: %s", tc_last_error()); h->launches += 3; }
    else if (tc) { if ((r = tc_plan_fwd(h->tc.l1, h->Yhi, h->Ylo, h->O12, 2 * rows, 2LL * h->Bc * N, 0, h->stream))) return fail(h, SNDVAE_E_CUDA, "tc fwd: %s", tc_last_error()); h->launches++; }
    else LEW(e2e_l1_simt_fwd_k, 2 * cells * C2, h->Yf, h->P + p.e_w[1], h->O12, 2 * rows, N, C1, C2);
    ev_end(h);
    mark(h, "epilogue");
    EpiParams ep; memset(&ep, 0, sizeof ep);
    ep.O12 = h->O12; ep.b1 = h->P + p.e_b[1];
    ep.gd = h->dis ? h->P + p.decadj_g : nullptr; ep.bd = h->dis ? h->P + p.decadj_b : nullptr;
    ep.Me = h->P + p.d_e_lin2[0]; ep.be = h->P + p.d_e_lin2[1];
    ep.At = in ? in->adj_truth + b0 * N * N : nullptr;
    ep.gen_adj = (out && out->generated_adj) ? (long long*)out->generated_adj + b0 * N * N : nullptr;
    ep.logits = (out && out->generated_adj_prob) ? out->generated_adj_prob + b0 * N * N * 2 : nullptr;
    ep.dOf = h->dOf; ep.dOhi = h->dOhi; ep.dOlo = h->dOlo; ep.OP = (tc && !h->spec) ? h->tc.l1.CSo : C2; ep.bf16 = tc && !h->spec; ep.backward = backward;
    ep.loss_sum = h->loss + 0;
    ep.g_b1 = h->G + p.e_b[1]; ep.g_gd = h->dis ? h->G + p.decadj_g : nullptr; ep.g_bd = h->dis ? h->G + p.decadj_b : nullptr;
    ep.g_Me = h->G + p.d_e_lin2[0]; ep.g_be = h->G + p.d_e_lin2[1];
    ep.gscale = 1.f / (gB * N * N);
    { long long ntile = (long long)bc * cdiv(N, EPI_T) * cdiv(N, EPI_T);
      unsigned grid = (unsigned)(ntile < 148 * 3 ? ntile : 148 * 3);
      if (h->spec) LAUNCH(edge_epilogue_ew_k, 148 * 6, EPI_EW_THREADS, 0, ep, bc, N);
      else LAUNCH(edge_epilogue_k, grid, 256, EPI_SMEM_BYTES, ep, bc, N); }
    if (!backward) continue;
    // backward of e2e layer 1: dgrad + wgrad
    mark(h, "gemm_dgrad");
    ev_begin(h, (h->spec ? 2.0 : 1.0) * f1 * bc);
    if (h->spec) { if (spec_backward(h->sp, h->dOf, rows, h->dY12, h->stream)) return fail(h, SNDVAE_E_CUDA, "spectral bwd: %s", tc_last_error()); h->launches += 4; }
    else if (tc) { if ((r = tc_plan_dgrad(h->tc.l1, h->dOhi, h->dOlo, h->dY12, 2 * rows, 2LL * h->Bc * N, 0, h->stream))) return fail(h, SNDVAE_E_CUDA, "tc dgrad: %s", tc_last_error()); h->launches++; }
    else LEW(e2e_l1_simt_dgrad_k, 2 * cells * C1, h->dOf, h->P + p.e_w[1], h->dY12, 2 * rows, N, C1, C2);
    ev_end(h);
    mark(h, "gemm_wgrad");
    if (!h->spec) ev_begin(h, f1 * bc);
    if (h->spec) {}
    else if (tc) { if ((r = tc_plan_wgrad(h->tc.l1, h->Yhi, h->Ylo, h->dOhi, h->dOlo, h->G + p.e_w[1], C1, 0, 2 * rows, h->stream))) return fail(h, SNDVAE_E_CUDA, "tc wgrad: %s", tc_last_error()); h->launches++; }
    else { dim3 g(cdiv((long long)N * C1 * C2, 256), cdiv(2 * rows, WGRAD_RG));
           LAUNCH(e2e_l1_simt_wgrad_k, g, 256, 0, h->Yf, h->dOf, h->G + p.e_w[1], 2 * rows, N, C1, C2); }
    if (!h->spec) ev_end(h);
    mark(h, "combine");
    // back through relu/BN_e1 to dE1 (both layouts), then the layer-0 contractions
    if (tc) {
      TcState& T = h->tc; L0Dense& Ld = h->l0d;
      const int CSe = Ld.CSe; const long long poff = cells * CSe;        // direction-1 planes
      if (h->spec && C1 % 2 == 0 && CSe % 2 == 0) {          // two channels per thread: 4-byte plane stores, 8-byte loads
        const int nt = 5 * C1 / 2 * 2;                         // 10 positions x C1/2 channel pairs
        LAUNCH(l0_combine_planes2_k, (unsigned)rows, nt, sizeof(float) * 6 * nt, h->dY12, h->E1, h->P + p.e_bng[1], h->P + p.e_bnb[1],
               h->G + p.e_bng[1], h->G + p.e_bnb[1], h->G + p.e_b[0], h->Yhi, h->Ylo, h->dSa + b0 * N * C1, bc, N, C1, CSe, h->Sa + b0 * N * C1, h->P + p.e_b[0]);
        LAUNCH(rowsum_planes2_k, (unsigned)rows, nt, sizeof(float) * 2 * nt, h->Yhi + poff, h->Ylo + poff, h->dRc + b0 * N * C1, N, C1, CSe);
      } else {
        LAUNCH(l0_combine_planes_k, (unsigned)rows, 5 * C1, sizeof(float) * 15 * C1, h->dY12, h->E1, h->P + p.e_bng[1], h->P + p.e_bnb[1],
             h->G + p.e_bng[1], h->G + p.e_bnb[1], h->G + p.e_b[0], h->Yhi, h->Ylo, h->dSa + b0 * N * C1, bc, N, C1, CSe, h->spec, h->Sa + b0 * N * C1, h->P + p.e_b[0]);
        LAUNCH(rowsum_planes_k, (unsigned)rows, 5 * C1, sizeof(float) * 5 * C1, h->Yhi + poff, h->Ylo + poff, h->dRc + b0 * N * C1, N, C1, CSe);
      }
      mark(h, "l0_gemms");
      if (l0d_bwd_act(Ld, 0, h->Yhi, h->Ylo, h->da + b0 * N * Chv, rows, rows, h->stream) ||
          l0d_bwd_act(Ld, 1, h->Yhi + poff, h->Ylo + poff, h->dc + b0 * N * Chv, rows, rows, h->stream) ||
          l0d_bwd_w(Ld, h->Yhi, h->Ylo, T.ah + b0 * N * Ld.CSk, T.al + b0 * N * Ld.CSk, h->dWSa, rows, h->stream) ||
          l0d_bwd_w(Ld, h->Yhi + poff, h->Ylo + poff, T.ch + b0 * N * Ld.CSk, T.cl + b0 * N * Ld.CSk, h->dWSc, rows, h->stream))
        return fail(h, SNDVAE_E_CUDA, "layer-0 dense backward: %s", tc_last_error());
      h->launches += 4;
    } else {
      LAUNCH(l0_combine_k, (unsigned)rows, 5 * C1, sizeof(float) * 15 * C1, h->dY12, h->E1, h->P + p.e_bng[1], h->P + p.e_bnb[1],
             h->G + p.e_bng[1], h->G + p.e_bnb[1], h->G + p.e_b[0], bc, N, C1);
      const float* dE1 = h->dY12; const float* dE1t = h->dY12 + cells * C1;
      LAUNCH(rowsum_k, (unsigned)rows, 5 * C1, sizeof(float) * 5 * C1, dE1, h->dSa + b0 * N * C1, N, C1);
      LAUNCH(rowsum_k, (unsigned)rows, 5 * C1, sizeof(float) * 5 * C1, dE1t, h->dRc + b0 * N * C1, N, C1);
      mark(h, "l0_gemms");
      // da[i,:] = sum_{(j,o)} dE1[i,(j,o)] WSa[(j,o),:];  dc[j,:] = sum_{(i,o)} dE1t[j,(i,o)] WSc[(i,o),:]
      CKB(gemm_rm(h, false, false, (int)rows, Chv, N * C1, 1.f, dE1, N * C1, h->WSa, Chv, 0.f, h->da + b0 * N * Chv, Chv));
      CKB(gemm_rm(h, false, false, (int)rows, Chv, N * C1, 1.f, dE1t, N * C1, h->WSc, Chv, 0.f, h->dc + b0 * N * Chv, Chv));
      // dWSa[(j,o),:] += sum_rows dE1[row,(j,o)] a[row,:]
      CKB(gemm_rm(h, true, false, N * C1, Chv, (int)rows, 1.f, dE1, N * C1, h->a + b0 * N * Chv, Chv, 1.f, h->dWSa, Chv));
      CKB(gemm_rm(h, true, false, N * C1, Chv, (int)rows, 1.f, dE1t, N * C1, h->c + b0 * N * Chv, Chv, 1.f, h->dWSc, Chv));
    }
  }
  return 0;
}

// per-step staging shared by every piece of the batch: layer-0 weight sums, tensor-core operand copies of the e2e weights,
// accumulator resets  (decoder_finish folds the frequency-domain weight gradient into dw1 once all pieces are done)
static int decoder_prepare(sndvae_t* h, bool backward) {
  const sndvae_config& c = h->cfg; const PT& p = h->pt;
  const int N = h->N, Chv = h->Chv, C1 = h->C1; const int Ctot = 2 * Chv;
  const float* w0 = h->P + p.e_w[0];
  LEW(e2e_l0_prep_k, (long long)N * C1 * Chv, w0, h->WSa, N, Ctot, 0, Chv, C1);
  LEW(e2e_l0_prep_k, (long long)N * C1 * Chv, w0, h->WSc, N, Ctot, Chv, Chv, C1);
  if (c.use_tensor_cores) {
    TcState& T = h->tc;
    if ((!h->spec && tc_plan_stage(T.l1, h->P + p.e_w[1], C1, 0, h->stream)) || tc_plan_stage(T.l0a, w0, Ctot, 0, h->stream) ||
        tc_plan_stage(T.l0c, w0, Ctot, Chv, h->stream))
      return fail(h, SNDVAE_E_CUDA, "tensor-core weight staging: %s", tc_last_error());
    if (l0d_stage(h->l0d, 0, h->WSa, h->stream) || l0d_stage(h->l0d, 1, h->WSc, h->stream))
      return fail(h, SNDVAE_E_CUDA, "layer-0 dense staging: %s", tc_last_error());
    if (h->spec && (spec_stage_weights(h->sp, h->P + p.e_w[1], h->stream) || ytc_stage(h->ytc, h->WSa, h->WSc, h->stream)))
      return fail(h, SNDVAE_E_CUDA, "spectral weight staging: %s", tc_last_error());
    h->launches += 5;
  }
  if (backward && h->spec && spec_zero_wgrad(h->sp, h->stream)) return fail(h, SNDVAE_E_CUDA, "spectral wgrad reset failed");
  if (backward) {
    CK(cudaMemsetAsync(h->dWSa, 0, sizeof(float) * N * C1 * Chv, h->stream));
    CK(cudaMemsetAsync(h->dWSc, 0, sizeof(float) * N * C1 * Chv, h->stream));
  }
  return 0;
}
static int decoder_finish(sndvae_t* h, bool backward) {
  if (backward && h->spec) {     // fold the accumulated frequency-domain weight gradient into dw1 (once per step)
    if (spec_finalize_wgrad(h->sp, h->G + h->pt.e_w[1], h->stream)) return fail(h, SNDVAE_E_CUDA, "spectral wgrad: %s", tc_last_error());
    h->launches++;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// data-parallel communicator: NCCL, resolved at run time (dlopen of the libnccl.so.2 already in the process -- PyTorch
// loads it -- or on the loader path), so that libsndvae.so itself has no link-time dependency on it
// ------------------------------------------------------------------------------------------
#include <dlfcn.h>
typedef struct { char internal[128]; } nccl_unique_id;          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
  int (*GetUniqueId)(nccl_unique_id*);
  int (*CommInitRank)(void**, int, nccl_unique_id, int);
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
  int (*CommDestroy)(void*);
  const char* (*GetErrorString)(int);
  int ok;
};
static NcclApi g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
static const char* nccl_load() {
  if (g_nccl.ok) return nullptr;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return "libnccl.so.2 is neither loaded in this process nor on the loader path";
#define SYM_(f) do { *(void**)(&g_nccl.f) = dlsym(lib, "nccl" #f); if (!g_nccl.f) return "nccl" #f " not found in libnccl"; } while (0)
  SYM_(GetUniqueId); SYM_(CommInitRank); SYM_(AllReduce); SYM_(AllGather); SYM_(CommDestroy); SYM_(GetErrorString);
#undef SYM_
  g_nccl.ok = 1;
  return nullptr;
}
#define NCCL_FLOAT 7   /* ncclFloat32 */
#define NCCL_SUM 0     /* ncclSum */
#define CKN(call) do { int s_ = (call); if (s_ != 0) return fail(h, SNDVAE_E_CUDA, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(s_), __FILE__, __LINE__); } while (0)
// sum of the gradient arena (and of the 8 loss sums) over the ranks of the communicator; no-op without one
static int allreduce_arena(sndvae_t* h, bool with_losses) {
  if (!h->comm || h->world <= 1) return 0;
  CKN(g_nccl.AllReduce(h->G, h->G, (size_t)h->nparam, NCCL_FLOAT, NCCL_SUM, h->comm, h->stream));
  if (with_losses) CKN(g_nccl.AllReduce(h->loss, h->loss, 8, NCCL_FLOAT, NCCL_SUM, h->comm, h->stream));
  h->launches += with_losses ? 2 : 1;
  return 0;
}

// ---- loss variants on the latents (optimizer.py:166-183) -------------------------------------------------------------
static float capacity_C(const sndvae_t* h) {     // optimizer.py:172
  const sndvae_config& c = h->cfg;
  const long long step = (long long)c.C_step > 0 ? (long long)c.C_step : 1;
  float C = c.C_max * c.C_step / c.C_stop_iter * (float)(h->global_iter / step);
  return C < 0.f ? 0.f : (C > c.C_max ? c.C_max : C);
}
// DIP forward for the three posterior means: batch covariance -> regulariser value (loss[6]) and d reg / d cov (kept for backward).
// With a communicator the covariance is the GLOBAL batch's, as in the reference where one process sees the whole batch
// (optimizer.py:7-21): the second-moment and mean sums are all-reduced (L^2 + L floats per latent) before dip_cov_k.
static int dip_forward(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  const float* mus[3] = {h->mu_s, h->mu_g, h->mu_sg};
  const long long rows[3] = {h->B, h->B, h->BS};
  const int Ls[3] = {c.s_latent_size, c.g_latent_size, c.sg_latent_size};
  const int world = (h->comm && h->world > 1) ? h->world : 1;
  for (int i = 0; i < 3; ++i) {
    const int L = Ls[i];
    CK(cudaMemsetAsync(h->dipv, 0, sizeof(float) * L, h->stream));
    LAUNCH(colsum_k, dim3(slab_grid(rows[i], XTDY_SLAB), cdiv(L, 128)), 128, 0, mus[i], L, h->dipv, rows[i], L);
    CKB(gemm_rm(h, true, false, L, L, (int)rows[i], 1.f, mus[i], L, mus[i], L, 0.f, h->dipS, L));
    if (world > 1) {
      CKN(g_nccl.AllReduce(h->dipS, h->dipS, (size_t)L * L, NCCL_FLOAT, NCCL_SUM, h->comm, h->stream));
      CKN(g_nccl.AllReduce(h->dipv, h->dipv, (size_t)L, NCCL_FLOAT, NCCL_SUM, h->comm, h->stream));
      h->launches += 2;
    }
    LAUNCH(dip_cov_k, cdiv((long long)L * L, 256), 256, 0, h->dipS, h->dipv, h->dipG[i], h->dipm[i], h->loss + 6, L, 1.f / (float)(rows[i] * world),
           c.dip_lambda_od, c.dip_lambda_d);
  }
  return 0;
}
// dmu += beta (graphs under the statistic / global batch) (2 / rows of the statistic) (mu - m) G
static int dip_backward(sndvae_t* h, int which, const float* mu, float* dmu, long long rows, int L, float gB) {
  const int world = (h->comm && h->world > 1) ? h->world : 1;
  const float alpha = h->cfg.beta * 2.f / (float)(rows * world) * ((float)(h->B * world) / gB);
  CKB(gemm_rm(h, false, false, (int)rows, L, L, alpha, mu, L, h->dipG[which], L, 1.f, dmu, L));
  CKB(gemm_rm(h, false, false, 1, L, L, 1.f, h->dipm[which], L, h->dipG[which], L, 0.f, h->dipv, L));
  LEW(sub_row_k, rows * L, dmu, h->dipv, rows, L, alpha);
  return 0;
}

// Total correlation of the three latent groups (optimizer.py:190): loss[6] += sum TC, log-sum-exps kept for backward.
// O(rows^2 L) pairwise Gaussian log-densities, recomputed in each pass instead of stored ([rows, rows, L] in the reference).
static int tc_forward(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  const float* zs[3] = {h->z_s, h->z_g, h->z_sg}; const float* mus[3] = {h->mu_s, h->mu_g, h->mu_sg}; const float* lss[3] = {h->ls_s, h->ls_g, h->ls_sg};
  const long long rows[3] = {h->B, h->B, h->BS};
  const int Ls[3] = {c.s_latent_size, c.g_latent_size, c.sg_latent_size};
  for (int i = 0; i < 3; ++i) {
    LEW(tc_prec_k, rows[i] * Ls[i], lss[i], h->tcp, rows[i] * Ls[i]);
    LAUNCH(tc_fwd_k, (unsigned)rows[i], TCOR_THREADS, 0, zs[i], mus[i], lss[i], h->tcp, h->tcJ[i], h->tcL[i], h->loss + 6, rows[i], Ls[i], 1.f / (float)rows[i]);
  }
  return 0;
}
// dmu, dls += 10 (B / global batch) d TC / d (mu, ls), including the path through z = mu + eps exp(ls)
static int tc_backward(sndvae_t* h, int which, const float* z, const float* mu, const float* ls, const float* eps, float* dmu, float* dls,
                       long long rows, int L, float gB) {
  const float scale = 10.f / (float)rows * ((float)h->B / gB);
  LEW(tc_prec_k, rows * L, ls, h->tcp, rows * L);
  LAUNCH(tc_bwd_k<false>, (unsigned)rows, TCOR_THREADS, 0, z, mu, ls, h->tcp, h->tcJ[which], h->tcL[which], h->tcd[0], (float*)nullptr, rows, L, scale);
  LAUNCH(tc_bwd_k<true>, (unsigned)rows, TCOR_THREADS, 0, z, mu, ls, h->tcp, h->tcJ[which], h->tcL[which], h->tcd[1], h->tcd[2], rows, L, scale);
  LEW(tc_apply_k, rows * L, dmu, dls, h->tcd[0], h->tcd[1], h->tcd[2], eps, ls, rows * L);
  return 0;
}

// backward of everything except the per-chunk N^2 stages (already done inside decoder_fwd)
static int backward_rest(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, float gB) {
  const sndvae_config& c = h->cfg; const PT& p = h->pt;
  const int N = h->N, F = h->F, D = h->D, H = h->H, Chv = h->Chv, C1 = h->C1, S = h->S;
  const long long B = h->B, BS = h->BS, Rn = h->Rn;
  const int Ctot = 2 * Chv;
  const int dact = h->dis ? ACT_NONE : ACT_LRELU;
  const float* w0 = h->P + p.e_w[0];
  int r;
  mark(h, "l0_vec_bwd");
  // ---- e2e layer 0 vector terms and weight sums ----
  if (c.use_tensor_cores) {
    TcState& T = h->tc;
    if (tc_split(h->dRc, T.drh, T.drl, Rn, C1, T.l0c.CSo, h->stream) || tc_split(h->dSa, T.dsh, T.dsl, Rn, C1, T.l0a.CSo, h->stream) ||
        tc_plan_dgrad(T.l0c, T.drh, T.drl, h->dc, B, B, 1, h->stream) || tc_plan_dgrad(T.l0a, T.dsh, T.dsl, h->da, B, B, 1, h->stream) ||
        tc_plan_wgrad(T.l0c, T.ch, T.cl, T.drh, T.drl, h->G + p.e_w[0], Ctot, Chv, B, h->stream) ||
        tc_plan_wgrad(T.l0a, T.ah, T.al, T.dsh, T.dsl, h->G + p.e_w[0], Ctot, 0, B, h->stream))
      return fail(h, SNDVAE_E_CUDA, "tensor-core layer-0 backward products: %s", tc_last_error());
    h->launches += 6;
  } else {
    LEW(toep_vec_bwd_in_k, Rn * Chv, h->dRc, w0, h->dc, B, N, Ctot, Chv, Chv, C1);
    LEW(toep_vec_bwd_in_k, Rn * Chv, h->dSa, w0, h->da, B, N, Ctot, 0, Chv, C1);
    dim3 g(cdiv((long long)N * Chv * C1, 256), cdiv(B, TOEP_BG));
    LAUNCH(toep_vec_bwd_w_k, g, 256, 0, h->c, h->dRc, h->G + p.e_w[0], B, N, Ctot, Chv, Chv, C1);
    LAUNCH(toep_vec_bwd_w_k, g, 256, 0, h->a, h->dSa, h->G + p.e_w[0], B, N, Ctot, 0, Chv, C1);
  }
  LEW(e2e_l0_prep_bwd_k, (long long)N * Chv * C1, h->dWSa, h->G + p.e_w[0], N, Ctot, 0, Chv, C1);
  LEW(e2e_l0_prep_bwd_k, (long long)N * Chv * C1, h->dWSc, h->G + p.e_w[0], N, Ctot, Chv, Chv, C1);
  mark(h, "dec_nodes_bwd");
  // a = relu(BN_e0[:Ch](v)), c = relu(BN_e0[Ch:](v))
  bn_bwd(h, h->da, Chv, h->v, Chv, p.e_bng[0], p.e_bnb[0], h->dv, Chv, Rn, Chv, ACT_RELU, 0);
  bn_bwd(h, h->dc, Chv, h->v, Chv, p.e_bng[0] + Chv, p.e_bnb[0] + Chv, h->gA, Chv, Rn, Chv, ACT_RELU, 0);
  LEW(add_inplace_k, Rn * Chv, h->dv, h->gA, Rn * Chv);
  // ---- node-feature decoder ----
  const int* nc = c.n_d_channel;
  LAUNCH(xtdy_k, slab_grid(Rn, XTDY_SLAB), 64, 0, h->q3, nc[1], h->dxpre, F, h->G + p.d_n_lin2[0], Rn, N, nc[1], F, 1);
  LAUNCH(colsum_k, dim3(slab_grid(Rn, XTDY_SLAB), cdiv(F, 128)), 128, 0, h->dxpre, F, h->G + p.d_n_lin2[1], Rn, F);
  LEW(rowlin_bwd_in_k, Rn * nc[1], h->dxpre, F, h->P + p.d_n_lin2[0], h->gA, nc[1], Rn, nc[1], F, 0);      // dq3
  if (h->dis) bn_bwd(h, h->gA, nc[1], h->q2, nc[1], p.decnode_g, p.decnode_b, h->gA, nc[1], Rn, nc[1], ACT_NONE, 0);  // dq2
  bn_bwd(h, h->gA, nc[1], h->q2p, nc[1], p.n_bng[1], p.n_bnb[1], h->gA, nc[1], Rn, nc[1], dact, 0);                   // dq2p
  if ((r = conv_bwd(h, h->q1, p.n_k[1], p.n_b[1], h->gA, h->gB, Rn, nc[0], nc[1]))) return r;                                            // dq1
  bn_bwd(h, h->gB, nc[0], h->q1p, nc[0], p.n_bng[0], p.n_bnb[0], h->gB, nc[0], Rn, nc[0], dact, 0);                   // dq1p
  if ((r = conv_bwd(h, h->v, p.n_k[0], p.n_b[0], h->gB, h->gA, Rn, Chv, nc[0]))) return r;                                               // dv (node)
  LEW(add_inplace_k, Rn * Chv, h->dv, h->gA, Rn * Chv);
  // ---- spatial decoder ----
  const int* sc = c.s_d_channel;
  LAUNCH(xtdy_k, slab_grid(Rn, XTDY_SLAB), 64, 0, h->s3, sc[2], h->dppre, D, h->G + p.d_s_lin2[0], Rn, N, sc[2], D, 1);
  LAUNCH(colsum_k, dim3(slab_grid(Rn, XTDY_SLAB), cdiv(D, 128)), 128, 0, h->dppre, D, h->G + p.d_s_lin2[1], Rn, D);
  LEW(rowlin_bwd_in_k, Rn * sc[2], h->dppre, D, h->P + p.d_s_lin2[0], h->gA, sc[2], Rn, sc[2], D, 0);      // ds3
  bn_bwd(h, h->gA, sc[2], h->s3p, sc[2], p.s_bng[2], p.s_bnb[2], h->gA, sc[2], Rn, sc[2], dact, 0);                   // ds3p
  if ((r = conv_bwd(h, h->s2, p.s_k[2], p.s_b[2], h->gA, h->gB, Rn, sc[1], sc[2]))) return r;                                            // ds2
  bn_bwd(h, h->gB, sc[1], h->s2p, sc[1], p.s_bng[1], p.s_bnb[1], h->gB, sc[1], Rn, sc[1], dact, 0);
  if ((r = conv_bwd(h, h->s1, p.s_k[1], p.s_b[1], h->gB, h->gA, Rn, sc[0], sc[1]))) return r;                                            // ds1
  bn_bwd(h, h->gA, sc[0], h->s1p, sc[0], p.s_bng[0], p.s_bnb[0], h->gA, sc[0], Rn, sc[0], dact, 0);
  if ((r = conv_bwd(h, h->sp0, p.s_k[0], p.s_b[0], h->gA, h->dsp0, Rn, Chv, sc[0]))) return r;                                           // dsp0
  // ---- split back into n_sg / n_s / n_g and through the z -> [N,H] linears ----
  if (h->dis) {
    LEW(copy_cols_k, Rn * H, h->dv, Chv, 0, h->dn_sg, H, 0, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->dsp0, Chv, 0, h->dn_sg, H, 0, Rn, H, 1);
    LEW(copy_cols_k, Rn * H, h->dv, Chv, H, h->dn_g, H, 0, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->dsp0, Chv, H, h->dn_s, H, 0, Rn, H, 0);
    if ((r = lin_bwd(h, h->z_s, p.d_s_lin1, h->dn_s, h->dz_s, B, c.s_latent_size, N * H))) return r;
    if ((r = lin_bwd(h, h->z_g, p.d_g_lin1, h->dn_g, h->dz_g, B, c.g_latent_size, N * H))) return r;
  } else {
    LEW(copy_cols_k, Rn * H, h->dv, Chv, 0, h->dn_sg, H, 0, Rn, H, 0);
    LEW(copy_cols_k, Rn * H, h->dsp0, Chv, 0, h->dn_sg, H, 0, Rn, H, 1);
  }
  if ((r = lin_bwd(h, h->zbar, p.d_sg_lin1, h->dn_sg, h->dzbar, B, c.sg_latent_size, N * H))) return r;
  mark(h, "enc_bwd_small");
  // ---- reparameterisation + KL, heads, encoders ----
  // KL weights of the loss branch: ELBO beta each; capacity loss 1, 1, gamma * 1[kl_sg > C]; DIP 1 each (optimizer.py:164,173,182)
  const int lv = c.loss_variant;
  const float beta = (lv == SNDVAE_LOSS_ELBO || lv == SNDVAE_LOSS_TC) ? c.beta : 1.f;
  const float beta_sg = lv == SNDVAE_LOSS_CAPACITY ? c.gamma : beta;
  if (h->dis) {
    // graph head
    int L = c.g_latent_size, Hh = c.g_hidden_size;
    int g0 = c.g_conv_hidden[0], g1c = c.g_conv_hidden[1];
    LEW(reparam_kl_bwd_k, B * L, h->mu_g, h->ls_g, nz->eps_g, h->dz_g, h->dmu, h->dls, B, L, 1, beta / (gB * L));
    if (lv == SNDVAE_LOSS_DIP && (r = dip_backward(h, 1, h->mu_g, h->dmu, B, L, gB))) return r;
    if (lv == SNDVAE_LOSS_TC && (r = tc_backward(h, 1, h->z_g, h->mu_g, h->ls_g, nz->eps_g, h->dmu, h->dls, B, L, gB))) return r;
    if ((r = lin_bwd(h, h->hg, p.g_lin[1], h->dmu, h->dh, B, Hh, L))) return r;
    if ((r = lin_bwd(h, h->hg, p.g_lin[2], h->dls, h->gC, B, Hh, L))) return r;
    LEW(add_inplace_k, B * Hh, h->dh, h->gC, B * Hh);
    if ((r = lin_bwd(h, h->fg, p.g_lin[0], h->dh, h->gA, B, N * (g1c + F), Hh))) return r;                          // dfg [Rn, g1c+F]
    bn_bwd(h, h->gA, g1c + F, h->g2, g1c + F, p.encg_g, p.encg_b, h->gA, g1c + F, Rn, g1c + F, ACT_NONE, 0);        // dg2
    bn_bwd(h, h->gA, g1c + F, h->c1, g1c, p.gg_bng[1], p.gg_bnb[1], h->gB, g1c, Rn, g1c, ACT_LRELU, 1);             // dc1
    CK(cudaMemsetAsync(h->gC, 0, sizeof(float) * Rn * g1c, h->stream));
    LAUNCH(graph_prop_bwd_k, cdiv(Rn * 32, 256), 256, 0, in->adj_truth, h->gB, h->gC, Rn, N, g1c);                  // dt1
    LAUNCH(xtdy_k, slab_grid(Rn, XTDY_SLAB), 256, 0, h->g1, g0 + F, h->gC, g1c, h->G + p.gg_w[1], Rn, N, g0 + F, g1c, 1);
    LEW(rowlin_bwd_in_k, Rn * (g0 + F), h->gC, g1c, h->P + p.gg_w[1], h->gA, g0 + F, Rn, g0 + F, g1c, 0);  // dg1
    bn_bwd(h, h->gA, g0 + F, h->c0, g0, p.gg_bng[0], p.gg_bnb[0], h->gB, g0, Rn, g0, ACT_LRELU, 1);                 // dc0
    CK(cudaMemsetAsync(h->gC, 0, sizeof(float) * Rn * g0, h->stream));
    LAUNCH(graph_prop_bwd_k, cdiv(Rn * 32, 256), 256, 0, in->adj_truth, h->gB, h->gC, Rn, N, g0);                   // dt0
    LAUNCH(xtdy_k, slab_grid(Rn, XTDY_SLAB), 64, 0, in->feature_truth, F, h->gC, g0, h->G + p.gg_w[0], Rn, N, F, g0, 1);
    // spatial head
    L = c.s_latent_size; Hh = c.s_hidden_size;
    const int* ec = c.s_channel;
    LEW(reparam_kl_bwd_k, B * L, h->mu_s, h->ls_s, nz->eps_s, h->dz_s, h->dmu, h->dls, B, L, 1, beta / (gB * L));
    if (lv == SNDVAE_LOSS_DIP && (r = dip_backward(h, 0, h->mu_s, h->dmu, B, L, gB))) return r;
    if (lv == SNDVAE_LOSS_TC && (r = tc_backward(h, 0, h->z_s, h->mu_s, h->ls_s, nz->eps_s, h->dmu, h->dls, B, L, gB))) return r;
    if ((r = lin_bwd(h, h->hs, p.s_lin[1], h->dmu, h->dh, B, Hh, L))) return r;
    if ((r = lin_bwd(h, h->hs, p.s_lin[2], h->dls, h->gC, B, Hh, L))) return r;
    LEW(add_inplace_k, B * Hh, h->dh, h->gC, B * Hh);
    if ((r = lin_bwd(h, h->fs, p.s_lin[0], h->dh, h->gA, B, N * ec[2], Hh))) return r;                              // dfs
    bn_bwd(h, h->gA, ec[2], h->h3, ec[2], p.encs_g, p.encs_b, h->gA, ec[2], Rn, ec[2], ACT_NONE, 0);                // dh3
    bn_bwd(h, h->gA, ec[2], h->h3p, ec[2], p.gs_bng[2], p.gs_bnb[2], h->gA, ec[2], Rn, ec[2], ACT_RELU, 0);         // dh3p
    if ((r = conv_bwd(h, h->h2, p.gs_k[2], p.gs_b[2], h->gA, h->gB, Rn, ec[1], ec[2]))) return r;
    bn_bwd(h, h->gB, ec[1], h->h2p, ec[1], p.gs_bng[1], p.gs_bnb[1], h->gB, ec[1], Rn, ec[1], ACT_RELU, 0);
    if ((r = conv_bwd(h, h->h1, p.gs_k[1], p.gs_b[1], h->gB, h->gA, Rn, ec[0], ec[1]))) return r;
    bn_bwd(h, h->gA, ec[0], h->h1p, ec[0], p.gs_bng[0], p.gs_bnb[0], h->gA, ec[0], Rn, ec[0], ACT_RELU, 0);
    if ((r = conv_bwd(h, in->spatial_truth, p.gs_k[0], p.gs_b[0], h->gA, nullptr, Rn, D, ec[0]))) return r;
  }
  mark(h, "enc_bwd_sgc");
  // joint head (z_sg rows are graph-major: row b*S+s; the S-mean gives dz = dzbar[b]/S)
  {
    int L = c.sg_latent_size, Hh = c.sg_hidden_size;
    const int h02 = c.sg_conv_hidden[0][2], h12 = c.sg_conv_hidden[1][2];
    if (lv == SNDVAE_LOSS_CAPACITY)
      LEW(reparam_kl_bwd_k, BS * L, h->mu_sg, h->ls_sg, nz->eps_sg, h->dzbar, h->dmu, h->dls, BS, L, S, beta_sg / (gB * S * L), h->loss + 5,
          1.f / ((float)BS * L), capacity_C(h));
    else LEW(reparam_kl_bwd_k, BS * L, h->mu_sg, h->ls_sg, nz->eps_sg, h->dzbar, h->dmu, h->dls, BS, L, S, beta_sg / (gB * S * L));
    if (lv == SNDVAE_LOSS_DIP && (r = dip_backward(h, 2, h->mu_sg, h->dmu, BS, L, gB))) return r;
    if (lv == SNDVAE_LOSS_TC && (r = tc_backward(h, 2, h->z_sg, h->mu_sg, h->ls_sg, nz->eps_sg, h->dmu, h->dls, BS, L, gB))) return r;
    if ((r = lin_bwd(h, h->hsg, p.sg_lin[1], h->dmu, h->dh, BS, Hh, L))) return r;
    float* tmp = h->dmu;   // reuse: dmu is consumed
    if ((r = lin_bwd(h, h->hsg, p.sg_lin[2], h->dls, tmp, BS, Hh, L))) return r;
    LEW(add_inplace_k, BS * Hh, h->dh, tmp, BS * Hh);
    if ((r = lin_bwd(h, h->fsg, p.sg_lin[0], h->dh, h->dfsg, BS, N * h12, Hh))) return r;
    if (h->hops3) return sgc3_encoder_bwd(h, in);
    for (int l = 0; l < 2; ++l) {
      SgcDims d = sgc_dims(h, l); SgcScratch& S = l == 0 ? h->S0 : h->S1;
      CK(cudaMemsetAsync(S.dWQ, 0, sizeof(float) * (2 * d.C + 2) * d.h0, h->stream));
      CK(cudaMemsetAsync(S.w46, 0, sizeof(float) * 2 * d.h0, h->stream));
      CK(cudaMemsetAsync(S.dW2, 0, sizeof(float) * (2 * d.C + 2 + d.h0) * d.h1, h->stream));
      CK(cudaMemsetAsync(S.dW3, 0, sizeof(float) * (d.C + d.h1 + 1) * d.h2, h->stream));
    }
    for (long long s0 = 0; s0 < BS; s0 += h->SC) {
      long long ns = BS - s0 < h->SC ? BS - s0 : h->SC;
      mark(h, "sgc_refwd");
      // chunk-sized scratch: recompute the chunk's activations; sgc_keep: they are still there from the forward pass
      if (!h->sgc_keep && (r = sgc_chunk_fwd(h, in, s0, ns))) return r;
      const float* x1 = h->x1 + (h->sgc_keep ? s0 * N * h02 : 0); const float* x2 = h->x2 + (h->sgc_keep ? s0 * N * h12 : 0);
      mark(h, "sgc_bwd_act");
      const float* x0 = in->features + s0 * N * F;
      // fsg = BN_encsg(x2); x2 = lrelu(BN_sg1(y1))
      bn_bwd(h, h->dfsg + s0 * N * h12, h12, x2, h12, h->dis ? p.encsg_g : -1, h->dis ? p.encsg_b : -1, h->dxa, h12, ns * N, h12, ACT_NONE, 0);
      bn_bwd(h, h->dxa, h12, sgc_view(h, 1, s0).y, h12, p.sg_bng[1], p.sg_bnb[1], h->dxa, h12, ns * N, h12, ACT_LRELU, 0);      // dy1
      if ((r = sgc_layer_bwd(h, 1, x1, h->dxa, h->dxb, s0, ns))) return r;                                       // dx1
      mark(h, "sgc_bwd_act");
      bn_bwd(h, h->dxb, h02, sgc_view(h, 0, s0).y, h02, p.sg_bng[0], p.sg_bnb[0], h->dxb, h02, ns * N, h02, ACT_LRELU, 0);      // dy0
      if ((r = sgc_layer_bwd(h, 0, x0, h->dxb, nullptr, s0, ns))) return r;
    }
    for (int l = 0; l < 2; ++l) {
      SgcDims d = sgc_dims(h, l); SgcScratch& S = l == 0 ? h->S0 : h->S1;
      LEW(sgc_unpack_grads_k, (2 * d.C + 4) * d.h0, S.dWQ, S.w46, h->G + p.sg_M1[l], h->G + p.sg_b1[l], d);
      LEW(sgc_unpack_w23_k, (2 * d.C + 2 + d.h0) * d.h1 + (d.C + d.h1 + 1) * d.h2, S.dW2, S.dW3, h->G + p.sg_M2[l], h->G + p.sg_b2[l],
          h->G + p.sg_M3[l], h->G + p.sg_b3[l], d);
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// step drivers
// ------------------------------------------------------------------------------------------
static int check_inputs(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz) {
  if (!in || !nz) return fail(h, SNDVAE_E_ARG, "inputs / noise struct is NULL");
  if (!in->features || !in->adj || !in->rel || !in->adj_truth || !in->feature_truth || !in->spatial_truth)
    return fail(h, SNDVAE_E_ARG, "a required feed (features, adj, rel, adj_truth, feature_truth, spatial_truth) is NULL");
  if (!nz->eps_sg || (h->dis && (!nz->eps_s || !nz->eps_g))) return fail(h, SNDVAE_E_ARG, "a required noise tensor is NULL");
  return 0;
}

static int copy_latents(sndvae_t* h, sndvae_outputs* out) {
  if (!out) return 0;
  const sndvae_config& c = h->cfg;
#define CP_(dst, src, n) if (out->dst && (src)) CK(cudaMemcpyAsync(out->dst, (src), sizeof(float) * (n), cudaMemcpyDeviceToDevice, h->stream))
  if (h->dis) {
    CP_(z_mean_s, h->mu_s, h->B * c.s_latent_size); CP_(z_std_s, h->ls_s, h->B * c.s_latent_size); CP_(z_s, h->z_s, h->B * c.s_latent_size);
    CP_(z_mean_g, h->mu_g, h->B * c.g_latent_size); CP_(z_std_g, h->ls_g, h->B * c.g_latent_size); CP_(z_g, h->z_g, h->B * c.g_latent_size);
  }
  CP_(z_mean_sg, h->mu_sg, h->BS * c.sg_latent_size); CP_(z_std_sg, h->ls_sg, h->BS * c.sg_latent_size); CP_(z_sg, h->z_sg, h->BS * c.sg_latent_size);
#undef CP_
  return 0;
}

// losses_host <- optimizer.overall_loss (optimizer.py:200-203)
// the read-back of the loss sums and of the error flag: enqueue (capturable) / wait and evaluate
static int fetch_losses_enqueue(sndvae_t* h) {
  CK(cudaMemcpyAsync(h->pinned_loss, h->loss, sizeof(float) * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->pinned_loss + 8, h->errflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
static int fetch_losses_finish(sndvae_t* h, float* losses_host, int ranks_summed);
static int fetch_losses(sndvae_t* h, float* losses_host, int ranks_summed = 1) {
  int r = fetch_losses_enqueue(h); if (r) return r;
  return fetch_losses_finish(h, losses_host, ranks_summed);
}
static int fetch_losses_finish(sndvae_t* h, float* losses_host, int ranks_summed) {
  CK(cudaStreamSynchronize(h->stream));
  const int ef = *reinterpret_cast<const int*>(h->pinned_loss + 8);
  if (ef) { cudaMemsetAsync(h->errflag, 0, sizeof(int), h->stream);
            return fail(h, SNDVAE_E_DENSE, "a sampled adjacency has more than edge_capacity=%d non-zeros; the joint encoder expects spanning-forest samples (input_data.py:18-38)", h->cfg.edge_capacity); }
  if (!losses_host) return 0;
  const sndvae_config& c = h->cfg; const float* L = h->pinned_loss;
  const double B = (double)h->B * ranks_summed, N = h->N;
  float adj = (float)(L[0] / (B * N * N)), node = (float)(L[1] / (B * N * h->F)), sp = (float)(L[2] / (B * N * h->D));
  float kl_sg = (float)(L[5] / ((double)h->BS * ranks_summed * c.sg_latent_size));
  if (ranks_summed > 1) h->pinned_loss[6] /= (float)ranks_summed;     // DIP / TC regularisers are per-rank statistics: report their mean
  if (h->dis) {
    float kl_s = (float)(L[3] / (B * c.s_latent_size)), kl_g = (float)(L[4] / (B * c.g_latent_size));
    if (c.loss_variant == SNDVAE_LOSS_CAPACITY) { const float ex = kl_sg - capacity_C(h); losses_host[0] = adj + node + sp + c.gamma * (ex > 0.f ? ex : 0.f) + kl_s + kl_g; }
    else if (c.loss_variant == SNDVAE_LOSS_DIP) losses_host[0] = adj + node + sp + (kl_sg + kl_s + kl_g) + c.beta * L[6];
    else if (c.loss_variant == SNDVAE_LOSS_TC) losses_host[0] = adj + node + sp + c.beta * (kl_sg + kl_s + kl_g) + 10.f * L[6];
    else losses_host[0] = adj + node + sp + c.beta * (kl_sg + kl_s + kl_g);
    losses_host[1] = sp; losses_host[2] = adj; losses_host[3] = node; losses_host[4] = kl_g; losses_host[5] = kl_s; losses_host[6] = kl_sg;
  } else {
    losses_host[0] = adj + node + sp + c.beta * kl_sg;
    losses_host[1] = sp; losses_host[2] = adj; losses_host[3] = node; losses_host[4] = kl_sg;
  }
  return 0;
}

// views of the feed / noise / output structs for the graphs [g0, g0 + gn)
static void slice_io(sndvae_t* h, long long g0, const sndvae_inputs* in, const sndvae_noise* nz, const sndvae_outputs* out,
                     sndvae_inputs* vin, sndvae_noise* vnz, sndvae_outputs* vout) {
  const sndvae_config& c = h->cfg; const long long N = h->N, S = h->S, F = h->F, D = h->D;
#define OFF_(ptr, n) ((ptr) ? (ptr) + g0 * (n) : nullptr)
  if (in) {
    vin->features = OFF_(in->features, S * N * F); vin->spatial = OFF_(in->spatial, S * N * D); vin->adj = OFF_(in->adj, S * N * N);
    vin->rel = OFF_(in->rel, S * N * N); vin->adj_truth = OFF_(in->adj_truth, N * N); vin->feature_truth = OFF_(in->feature_truth, N * F);
    vin->spatial_truth = OFF_(in->spatial_truth, N * D); vin->rel_truth = OFF_(in->rel_truth, N * N);
  }
  if (nz) { vnz->eps_s = OFF_(nz->eps_s, c.s_latent_size); vnz->eps_sg = OFF_(nz->eps_sg, S * c.sg_latent_size); vnz->eps_g = OFF_(nz->eps_g, c.g_latent_size); }
  if (out) {
    vout->z_mean_s = OFF_(out->z_mean_s, c.s_latent_size); vout->z_std_s = OFF_(out->z_std_s, c.s_latent_size); vout->z_s = OFF_(out->z_s, c.s_latent_size);
    vout->z_mean_g = OFF_(out->z_mean_g, c.g_latent_size); vout->z_std_g = OFF_(out->z_std_g, c.g_latent_size); vout->z_g = OFF_(out->z_g, c.g_latent_size);
    vout->z_mean_sg = OFF_(out->z_mean_sg, S * c.sg_latent_size); vout->z_std_sg = OFF_(out->z_std_sg, S * c.sg_latent_size);
    vout->z_sg = OFF_(out->z_sg, S * c.sg_latent_size);
    vout->generated_adj = OFF_(out->generated_adj, N * N); vout->generated_adj_prob = OFF_(out->generated_adj_prob, N * N * 2);
    vout->generated_spatial = OFF_(out->generated_spatial, N * D); vout->generated_node_feat = OFF_(out->generated_node_feat, N * F);
  }
#undef OFF_
}

// ---- compact host feeds (sndvae_inputs_compact): 0/1 adjacencies as bit rows, per-graph tensors once -----------------
// dense[row, j] = bit j of bits[row, :]   (row = (sample or graph, i); W = ceil(N / 32) words per row)
__global__ void expand_bits_k(const uint32_t* __restrict__ bits, float* __restrict__ dense, long long rows, int N, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * N) return;
  const long long r = idx / N; const int j = (int)(idx - r * N);
  dense[idx] = (float)((__ldg(bits + r * W + (j >> 5)) >> (j & 31)) & 1u);
}
// out[(g * S + s), :] = in[g, :]   (per-graph tensor repeated for its S samples: `rel`, `features`)
__global__ void repeat_rows_k(const float* __restrict__ in, float* __restrict__ out, long long graphs, int S, long long per) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= graphs * S * per) return;
  const long long gs = idx / per, e = idx - gs * per;
  out[idx] = __ldg(in + (gs / S) * per + e);
}
// bits[row, w] = the 32 adjacency entries j = 32 w .. 32 w + 31 of generated_adj[row, :]  (one warp per word via ballot)
__global__ void pack_adj_bits_k(const long long* __restrict__ adj, uint32_t* __restrict__ bits, long long rows, int N, int W) {
  const long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (word >= rows * W) return;
  const long long r = word / W; const int j = (int)(word - r * W) * 32 + lane;
  const unsigned m = __ballot_sync(0xffffffffu, j < N && adj[r * N + j] != 0);
  if (lane == 0) bits[word] = m;
}

// Host feeds of one piece -> device staging, on the copy stream (sndvae_train_step_host)
struct HostFeeds { const sndvae_inputs* in; const sndvae_noise* nz; int64_t* gen_adj;
                   const sndvae_inputs_compact* cin; uint32_t* gen_bits; };

// The step.  Every graph's forward is independent of the rest of the batch (frozen-affine BN, SURVEY finding 3), so the
// encoder + decoder + N^2 backward run piece by piece over contiguous graph ranges; with host feeds the H2D copy of piece k+1
// and the D2H copy of piece k-1's generated_adj overlap piece k's kernels.  The node-level backward runs once at the end.
enum { RUN_ACCUMULATE = 1, RUN_ALLREDUCE = 2, RUN_NOFETCH = 4 };
static int run_body(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host,
                    bool backward, long long global_batch, const HostFeeds* hf, int flags);
// a step that fails after work was enqueued leaves the gradient arena / loss sums half-written: drain the streams (the caller may
// free its host buffers as soon as we return) and mark the handle so that the next call starts from a clean slate
static int run(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host,
               bool backward, long long global_batch, const HostFeeds* hf = nullptr, int flags = 0) {
  if (h->poisoned && (flags & RUN_ACCUMULATE))
    return fail(h, SNDVAE_E_STATE, "the previous step failed half-way: gradients cannot be accumulated onto it; call sndvae_zero_grads first");
  const int r = run_body(h, in, nz, out, losses_host, backward, global_batch, hf, flags);
  if (r) {
    std::string keep = h->err;
    if (h->cs) cudaStreamSynchronize(h->cs);
    if (h->ds) cudaStreamSynchronize(h->ds);
    cudaStreamSynchronize(h->stream);
    cudaGetLastError();
    h->B = h->cfg.batch_size; h->BS = h->B * h->S; h->Rn = h->B * h->N;
    h->poisoned = backward ? 1 : 0; h->stt.used = 0;
    h->err = keep;
  } else if (backward) h->poisoned = 0;
  return r;
}
static int run_body(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host,
                    bool backward, long long global_batch, const HostFeeds* hf, int flags) {
  int r = check_inputs(h, in, nz); if (r) return r;
  const sndvae_config& c = h->cfg;
  const long long Bfull = h->B; const int N = h->N, S = h->S;
  const float gB = (float)(global_batch > 0 ? global_batch : Bfull);
  CK(cudaMemsetAsync(h->loss, 0, sizeof(float) * 8, h->stream));
  if (backward && !(flags & RUN_ACCUMULATE)) CK(cudaMemsetAsync(h->G, 0, sizeof(float) * h->nparam, h->stream));
  if ((r = decoder_prepare(h, backward))) return r;
  // piece schedule: host mode walks the batch in chunk-sized pieces, with a half-sized first piece so that less of the
  // first H2D copy is exposed before any kernel can start
  const long long piece = hf ? (h->Bc < Bfull ? h->Bc : Bfull) : Bfull;
  std::vector<long long> pstart;
  { long long g = 0; if (hf && piece >= 2 && Bfull > piece / 2) { pstart.push_back(0); g = piece / 2; }
    for (; g < Bfull; g += piece) pstart.push_back(g); }
  pstart.push_back(Bfull);
  const long long npieces = (long long)pstart.size() - 1;
  if (hf) {
    while ((long long)h->pev.size() < 2 * npieces) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->pev.push_back(e); }
    CK(cudaEventRecord(h->ev_start, h->stream));             // staging buffers are free once earlier work on the stream is done
    CK(cudaStreamWaitEvent(h->cs, h->ev_start, 0));
    // enqueue every piece's feeds now: the copy engine runs ahead of the kernels
    for (long long pi = 0; pi < npieces; ++pi) {
      const long long g0 = pstart[pi], gn = pstart[pi + 1] - g0;
      sndvae_inputs hs, ds; sndvae_noise hn, dn; memset(&hs, 0, sizeof hs); memset(&ds, 0, sizeof ds); memset(&hn, 0, sizeof hn); memset(&dn, 0, sizeof dn);
      slice_io(h, g0, hf->in, hf->nz, nullptr, &hs, &hn, nullptr);
      slice_io(h, g0, in, nz, nullptr, &ds, &dn, nullptr);
#define H2D_(f, n) CK(cudaMemcpyAsync((void*)ds.f, hs.f, sizeof(float) * (size_t)(gn * (n)), cudaMemcpyHostToDevice, h->cs))
      if (hf->cin) {      // packed bytes over the bus; the dense staging is filled by the expand kernels of the piece loop
        const sndvae_inputs_compact* ci = hf->cin; const long long W = (N + 31) / 32;
#define H2C_(dst, src, per, T) CK(cudaMemcpyAsync((void*)((dst) + g0 * (per)), (src) + g0 * (per), sizeof(T) * (size_t)(gn * (per)), cudaMemcpyHostToDevice, h->cs))
        H2C_(h->hc_features, ci->features, (long long)N * h->F, float); H2C_(h->hc_rel, ci->rel, (long long)N * N, float);
        H2C_(h->hc_adj_bits, ci->adj_bits, (long long)S * N * W, uint32_t); H2C_(h->hc_adjt_bits, ci->adj_truth_bits, (long long)N * W, uint32_t);
#undef H2C_
        CK(cudaMemcpyAsync((void*)ds.feature_truth, ci->feature_truth + g0 * N * h->F, sizeof(float) * (size_t)(gn * N * h->F), cudaMemcpyHostToDevice, h->cs));
        CK(cudaMemcpyAsync((void*)ds.spatial_truth, ci->spatial_truth + g0 * N * h->D, sizeof(float) * (size_t)(gn * N * h->D), cudaMemcpyHostToDevice, h->cs));
      } else {
      H2D_(features, (long long)S * N * h->F); H2D_(adj, (long long)S * N * N); H2D_(rel, (long long)S * N * N); H2D_(adj_truth, (long long)N * N);
      H2D_(feature_truth, (long long)N * h->F); H2D_(spatial_truth, (long long)N * h->D);
      }
#undef H2D_
      CK(cudaMemcpyAsync((void*)dn.eps_sg, hn.eps_sg, sizeof(float) * (size_t)(gn * S * c.sg_latent_size), cudaMemcpyHostToDevice, h->cs));
      if (h->dis) {
        CK(cudaMemcpyAsync((void*)dn.eps_s, hn.eps_s, sizeof(float) * (size_t)(gn * c.s_latent_size), cudaMemcpyHostToDevice, h->cs));
        CK(cudaMemcpyAsync((void*)dn.eps_g, hn.eps_g, sizeof(float) * (size_t)(gn * c.g_latent_size), cudaMemcpyHostToDevice, h->cs));
      }
      CK(cudaEventRecord(h->pev[2 * pi], h->cs));
    }
  }
  for (long long pi = 0; pi < npieces; ++pi) {
    const long long g0 = pstart[pi], gn = pstart[pi + 1] - g0;
    sndvae_inputs vin; sndvae_noise vnz; sndvae_outputs vout; memset(&vin, 0, sizeof vin); memset(&vnz, 0, sizeof vnz); memset(&vout, 0, sizeof vout);
    slice_io(h, g0, in, nz, out, &vin, &vnz, &vout);
    if (hf) CK(cudaStreamWaitEvent(h->stream, h->pev[2 * pi], 0));
    if (hf && hf->cin) {    // unpack this piece's feeds into the dense staging the kernels read
      const long long W = (N + 31) / 32;
      LEW(expand_bits_k, gn * S * N * N, h->hc_adj_bits + g0 * S * N * W, (float*)vin.adj, gn * S * N, N, (int)W);
      LEW(expand_bits_k, gn * N * N, h->hc_adjt_bits + g0 * N * W, (float*)vin.adj_truth, gn * N, N, (int)W);
      LEW(repeat_rows_k, gn * S * N * N, h->hc_rel + g0 * N * N, (float*)vin.rel, gn, S, (long long)N * N);
      LEW(repeat_rows_k, gn * S * N * h->F, h->hc_features + g0 * N * h->F, (float*)vin.features, gn, S, (long long)N * h->F);
    }
    view_shift(h, g0); h->B = gn; h->BS = gn * S; h->Rn = gn * N;
    mark(h, "encoder");
    r = encoder_fwd(h, &vin);
    if (!r) r = reparam_fwd(h, &vnz);
    if (!r) r = copy_latents(h, out ? &vout : nullptr);
    if (!r) r = decoder_fwd(h, &vin, out ? &vout : nullptr, backward, gB);
    view_shift(h, -g0); h->B = Bfull; h->BS = Bfull * S; h->Rn = Bfull * N;
    if (r) return r;
    if (hf && hf->gen_bits && out && out->generated_adj) {
      const long long W = (N + 31) / 32;
      LAUNCH(pack_adj_bits_k, cdiv(gn * N * W * 32, 256), 256, 0, (const long long*)out->generated_adj + g0 * N * N, h->hc_gen_bits + g0 * N * W, gn * N, N, (int)W);
      CK(cudaEventRecord(h->pev[2 * pi + 1], h->stream));
      CK(cudaStreamWaitEvent(h->ds, h->pev[2 * pi + 1], 0));
      CK(cudaMemcpyAsync(hf->gen_bits + g0 * N * W, h->hc_gen_bits + g0 * N * W, sizeof(uint32_t) * (size_t)(gn * N * W), cudaMemcpyDeviceToHost, h->ds));
    }
    if (hf && hf->gen_adj && out && out->generated_adj) {
      CK(cudaEventRecord(h->pev[2 * pi + 1], h->stream));
      CK(cudaStreamWaitEvent(h->ds, h->pev[2 * pi + 1], 0));
      CK(cudaMemcpyAsync(hf->gen_adj + g0 * N * N, out->generated_adj + g0 * N * N, sizeof(int64_t) * (size_t)(gn * N * N), cudaMemcpyDeviceToHost, h->ds));
    }
  }
  if ((r = decoder_finish(h, backward))) return r;
  if (c.loss_variant == SNDVAE_LOSS_DIP && (r = dip_forward(h))) return r;      // needs every piece's posterior means
  if (c.loss_variant == SNDVAE_LOSS_TC && (r = tc_forward(h))) return r;        // ... and samples
  if (backward && (r = backward_rest(h, in, nz, gB))) return r;
  mark(h, "allreduce");
  CK(cudaGetLastError());
  const bool reduce = backward && (flags & RUN_ALLREDUCE) && h->comm && h->world > 1;
  if (reduce && (r = allreduce_arena(h, true))) return r;
  mark(h, "end");
  if (flags & RUN_NOFETCH) return fetch_losses_enqueue(h);      // graph capture: the caller waits and evaluates after the replay
  r = fetch_losses(h, losses_host, reduce ? h->world : 1);
  report_stages(h);
  return r;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int sndvae_default_config(sndvae_config* c) {
  if (!c) return SNDVAE_E_ARG;
  memset(c, 0, sizeof *c);
  c->model_type = SNDVAE_MODEL_DISENTANGLED; c->num_nodes = 25; c->num_feature = 1; c->spatial_dim = 2; c->sampling_num = 10;
  c->node_h_size = 20;
  c->s_channel[0] = 10; c->s_channel[1] = 10; c->s_channel[2] = 20; c->s_hidden_size = 100; c->s_latent_size = 100;
  c->g_conv_hidden[0] = 10; c->g_conv_hidden[1] = 20; c->g_hidden_size = 100; c->g_latent_size = 100;
  for (int k = 0; k < 3; ++k) { c->sg_conv_hidden[0][k] = 20; c->sg_conv_hidden[1][k] = 50; }
  c->sg_hidden_size = 100; c->sg_latent_size = 100;
  c->s_d_channel[0] = 50; c->s_d_channel[1] = 20; c->s_d_channel[2] = 10;
  c->n_d_channel[0] = 50; c->n_d_channel[1] = 20; c->e_d_hidden[0] = 50; c->e_d_hidden[1] = 20;
  c->batch_size = 10; c->chunk_graphs = 0; c->edge_capacity = 0; c->use_tensor_cores = 2;
  c->learning_rate = 0.0008f; c->beta = 1.f; c->adam_beta1 = 0.9f; c->adam_beta2 = 0.999f; c->adam_eps = 1e-8f;
  c->loss_variant = SNDVAE_LOSS_ELBO; c->gamma = 100.f; c->C_max = 100.f; c->C_stop_iter = 100.f; c->C_step = 20.f;
  c->dip_lambda_od = 10.f; c->dip_lambda_d = 100.f;
  return 0;
}

int sndvae_create(const sndvae_config* cfg, void* stream, sndvae_t** out) {
  if (!cfg || !out) return SNDVAE_E_ARG;
  *out = nullptr;
  sndvae_t* h = new sndvae_handle();
  *out = h;    // returned even on failure so that sndvae_last_error works; caller destroys
  h->cfg = *cfg; h->stream = (cudaStream_t)stream; h->launches = 0; h->pinned_loss = nullptr; h->ev_used = 0;
  h->graph_exec = nullptr; h->gs = nullptr; h->graph_key = 0; h->capturing = 0; h->graph_launches = 0; h->graph_replays = 0; h->graph_mode = 0;
  h->hf_ready = 0; h->hc_ready = 0; h->hc_features = nullptr; h->poisoned = 0; h->comm = nullptr; h->rank = 0; h->world = 1;
  h->hf_features = nullptr; h->cs = nullptr; h->ds = nullptr; h->ev_start = nullptr; h->zz_planes = nullptr; h->zz_cap = 0;
  cudaFuncSetAttribute(edge_epilogue_k, cudaFuncAttributeMaxDynamicSharedMemorySize, EPI_SMEM_BYTES);
  h->stt.used = 0; h->stt.on = h->stt.print = getenv("SNDVAE_STAGE_TIMING") != nullptr; h->stt.steps = 0;
  sndvae_config& c = h->cfg;
  if (c.num_nodes < 2 || c.batch_size < 1 || c.num_feature < 1 || c.spatial_dim < 1 || c.node_h_size < 1)
    return fail(h, SNDVAE_E_ARG, "bad config: num_nodes=%d batch_size=%d", c.num_nodes, c.batch_size);
  if (c.model_type != SNDVAE_MODEL_DISENTANGLED && c.model_type != SNDVAE_MODEL_BASE) return fail(h, SNDVAE_E_ARG, "bad model_type %d", c.model_type);
  h->dis = c.model_type == SNDVAE_MODEL_DISENTANGLED;
  h->global_iter = 0;
  if (c.loss_variant < 0 || c.loss_variant > SNDVAE_LOSS_TC || (c.loss_variant != SNDVAE_LOSS_ELBO && c.model_type != SNDVAE_MODEL_DISENTANGLED))
    return fail(h, SNDVAE_E_ARG, "loss_variant %d needs the disentangled model (optimizer.py:166-190)", c.loss_variant);
  if (c.loss_variant == SNDVAE_LOSS_TC && (c.s_latent_size > 32 * TCOR_LK || c.g_latent_size > 32 * TCOR_LK || c.sg_latent_size > 32 * TCOR_LK))
    return fail(h, SNDVAE_E_ARG, "the total-correlation kernels take latent sizes <= %d", 32 * TCOR_LK);
  if (c.node_h_size != 20 && c.use_tensor_cores) {
    // the tensor-core tiles of the edge decoder are built for node_h_size = 20 (synthetic2, main.py:209); other sizes
    // (synthetic1: 50, main.py:164; protein: 5, main.py:230) run the e2e layers on the fp32 SIMT kernels
    fprintf(stderr, "[sndvae] node_h_size = %d: edge decoder on the fp32 SIMT kernels (tensor-core tiles need node_h_size = 20)\n", c.node_h_size);
    c.use_tensor_cores = 0;
  }
  h->spec = c.use_tensor_cores == 2; memset(&h->sp, 0, sizeof h->sp); memset(&h->ytc, 0, sizeof h->ytc);
  if (c.sg_hops != 0 && c.sg_hops != 2 && c.sg_hops != 3) return fail(h, SNDVAE_E_ARG, "sg_hops must be 2 (SpatialGraphConvolution) or 3 (SpatialGraphConvolution_3D)");
  h->hops3 = c.sg_hops == 3; h->ws3 = nullptr; h->y3[0] = h->y3[1] = nullptr;
  if (h->hops3) {
    for (int l = 0; l < 2; ++l) {
      for (int m = 0; m < 4; ++m) if (c.sg_conv_hidden3[l][m] < 1) return fail(h, SNDVAE_E_ARG, "sg_conv_hidden3[%d][%d] = %d", l, m, c.sg_conv_hidden3[l][m]);
      c.sg_conv_hidden[l][2] = c.sg_conv_hidden3[l][3];      // the layer's output width, where the rest of the code reads it
    }
  }
  if (!h->dis) c.sampling_num = 1;     // model_joint.py is coherent only with one sample per graph
  if (c.sampling_num < 1) return fail(h, SNDVAE_E_ARG, "sampling_num must be >= 1");
  if (c.e_d_hidden[1] != EPI_C2) return fail(h, SNDVAE_E_ARG, "e_d_hidden[1] must be %d in this build", EPI_C2);
  if (c.e_d_hidden[0] > 52) return fail(h, SNDVAE_E_ARG, "this build supports e_d_hidden[0] <= 52 (main.py:209)");
  if (c.g_conv_hidden[0] > 32 || c.g_conv_hidden[1] > 32) return fail(h, SNDVAE_E_ARG, "g_conv_hidden must be <= 32");
  h->N = c.num_nodes; h->F = c.num_feature; h->D = c.spatial_dim; h->S = c.sampling_num; h->H = c.node_h_size;
  h->Chv = h->dis ? 2 * h->H : h->H; h->C1 = c.e_d_hidden[0]; h->C2 = c.e_d_hidden[1];
  h->B = c.batch_size; h->BS = h->B * h->S; h->Rn = h->B * h->N;
  if (c.edge_capacity <= 0) c.edge_capacity = 4 * h->N;
  { // widest per-node channel count any node-level layer reads or writes (conv1d inputs / outputs, concatenations)
    int m = 64;
    const int cand[] = {h->F, h->D, h->Chv, c.s_channel[0], c.s_channel[1], c.s_channel[2], c.g_conv_hidden[0] + h->F,
                        c.g_conv_hidden[1] + h->F, c.n_d_channel[0], c.n_d_channel[1], c.s_d_channel[0], c.s_d_channel[1],
                        c.s_d_channel[2], c.sg_conv_hidden[0][2], c.sg_conv_hidden[1][2]};
    for (int v : cand) { if (v < 1) return fail(h, SNDVAE_E_ARG, "bad config: a layer width is %d", v); if (v > m) m = v; }
    h->max_c = m; }
  if (c.use_tensor_cores && (h->C1 != TC_C1 || h->C2 != TC_C2))
    return fail(h, SNDVAE_E_ARG, "tensor-core e2e path requires e_d_hidden = (%d, %d)", TC_C1, TC_C2);
  if (c.chunk_graphs <= 0) {
    // bound the N^2 staging buffers to ~24 GB: E1 200 + O12 160 + dY12 400 + Y planes 448 + dO planes 192 B per (i,j) cell
    long long per_graph = (long long)h->N * h->N * 1400;     // (+ 200 for the transposed E1 copy of the spectral path, inside its budget)
    long long budget = 24LL << 30;
    if (c.use_tensor_cores == 2) {   // + spectra: F frequencies x 2N lines x (Y^ planes 416 + O^ 160 + dO^ planes 160 + dY^ 400) B
      per_graph += (long long)(spec_pick_L(h->N) / 2 + 1) * 2 * h->N * 1136; budget = 56LL << 30; }
    long long bc = budget / per_graph; if (bc < 1) bc = 1; if (bc > h->B) bc = h->B; if (bc > 512) bc = 512;
    if (c.use_tensor_cores == 2 && bc >= 128) bc = bc / 128 * 128;     // whole 128-graph tiles for the batch-major layer-0 GEMM
    if ((long long)2 * bc * h->N * h->N * h->C1 > 2000000000LL) bc = 2000000000LL / ((long long)2 * h->N * h->N * h->C1);
    c.chunk_graphs = (int)bc;
  }
  if (c.chunk_graphs > h->B) c.chunk_graphs = (int)h->B;
  h->Bc = c.chunk_graphs;
  { long long per_sample = (long long)h->N * 1200 * 4;     // SGC scratch: ~820 + ~210 floats per node for the two layers
    long long sc = (4LL << 30) / per_sample; if (sc < 1) sc = 1; if (sc > h->BS) sc = h->BS; h->SC = (int)sc; }
  if ((long long)2 * h->Bc * h->N * h->N * h->C1 > 2000000000LL) return fail(h, SNDVAE_E_ARG, "chunk too large for 32-bit GEMM dims");
  { const char* g = getenv("SNDVAE_GRAPH");
    h->graph_mode = g ? atoi(g) != 0 : ((long long)h->B * h->N * h->N < (1LL << 21)); }
  build_table(h);      // host-only: the table is valid even when no device is present (checked by the CPU tests)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, SNDVAE_E_CUDA, "no CUDA device: the SND-VAE hot path has no CPU fallback");
  cudaDeviceProp prop; int dev = 0; cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
  if (prop.major < 10) return fail(h, SNDVAE_E_CUDA, "device sm_%d%d is not sm_100: this library is built for B200 only", prop.major, prop.minor);
  int r = alloc_buffers(h); if (r) return r;
  if (cudaMallocHost((void**)&h->pinned_loss, 128) != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "cudaMallocHost failed");
  h->b1p = c.adam_beta1; h->b2p = c.adam_beta2;
  h->ev.resize(4096);
  for (auto& e : h->ev) { cudaEventCreate(&e.a); cudaEventCreate(&e.b); e.flops = 0; }
  if (c.use_tensor_cores) {
    if ((r = tc_init(h->tc, h->N, h->Chv, h->B, h->stream))) return fail(h, SNDVAE_E_CUDA, "tc_init: %s", tc_last_error());
    if ((r = l0d_init(h->l0d, h->N, h->Chv, h->C1, h->stream))) return fail(h, SNDVAE_E_CUDA, "l0d_init: %s", tc_last_error());
    { TcState& T = h->tc; const long long ni = (long long)h->N * T.l0a.CSi, no = (long long)h->N * T.l0a.CSo;
      reg_shift(h, &T.ah, ni); reg_shift(h, &T.al, ni); reg_shift(h, &T.ch, ni); reg_shift(h, &T.cl, ni);
      reg_shift(h, &T.dsh, no); reg_shift(h, &T.dsl, no); reg_shift(h, &T.drh, no); reg_shift(h, &T.drl, no); }
    if (h->spec && (r = spec_init(h->sp, h->N, 2LL * h->Bc * h->N, h->stream))) return fail(h, SNDVAE_E_CUDA, "spec_init: %s", tc_last_error());
    if (h->spec && (r = ytc_init(h->ytc, h->N, h->Chv, h->C1, h->stream))) return fail(h, SNDVAE_E_CUDA, "ytc_init: %s", tc_last_error());
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int sndvae_destroy(sndvae_t* h) {
  if (!h) return 0;
  cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->gs) cudaStreamDestroy(h->gs);
  tc_destroy(h->tc); if (h->cfg.use_tensor_cores) l0d_destroy(h->l0d);
  if (h->spec) { spec_destroy(h->sp); ytc_destroy(h->ytc); }
  for (void* p : h->allocs) cudaFree(p);
  if (h->zz_planes) cudaFree(h->zz_planes);
  for (auto& e : h->ev) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& e : h->pev) cudaEventDestroy(e);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->cs) cudaStreamDestroy(h->cs);
  if (h->ds) cudaStreamDestroy(h->ds);
  if (h->pinned_loss) cudaFreeHost(h->pinned_loss);
  delete h;
  return 0;
}

int sndvae_inner_product_decode(sndvae_t* h, const float* z, int64_t batch, int32_t num_nodes, int32_t dim, float* logits) {
  if (!h || !z || !logits) return SNDVAE_E_ARG;
  if (batch < 1 || num_nodes < 1 || dim < 1) return fail(h, SNDVAE_E_ARG, "inner_product_decode: batch=%lld num_nodes=%d dim=%d", (long long)batch, num_nodes, dim);
  if (zzt_run(z, batch, num_nodes, dim, logits, &h->zz_planes, &h->zz_cap, h->stream)) return fail(h, SNDVAE_E_CUDA, "inner_product_decode: %s", tc_last_error());
  h->launches += 2;
  return 0;
}

const char* sndvae_last_error(const sndvae_t* h) { return h ? h->err.c_str() : "null handle"; }
int64_t sndvae_param_count(const sndvae_t* h) { return h ? h->nparam : 0; }
int32_t sndvae_num_params(const sndvae_t* h) { return h ? (int32_t)h->table.size() : 0; }
int sndvae_param_table(const sndvae_t* h, sndvae_param_info* t, int32_t cap) {
  if (!h || !t) return SNDVAE_E_ARG;
  if (cap < (int32_t)h->table.size()) return SNDVAE_E_ARG;
  memcpy(t, h->table.data(), sizeof(sndvae_param_info) * h->table.size());
  return 0;
}
int sndvae_get_params(sndvae_t* h, float* dst) {
  if (!h || !dst) return SNDVAE_E_ARG;
  CK(cudaMemcpyAsync(dst, h->P, sizeof(float) * h->nparam, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream)); return 0;
}
int sndvae_set_params(sndvae_t* h, const float* src) {
  if (!h || !src) return SNDVAE_E_ARG;
  CK(cudaMemcpyAsync(h->P, src, sizeof(float) * h->nparam, cudaMemcpyHostToDevice, h->stream)); CK(cudaStreamSynchronize(h->stream)); return 0;
}
int sndvae_get_adam(sndvae_t* h, float* m, float* v, float* bp) {
  if (!h) return SNDVAE_E_ARG;
  if (m) CK(cudaMemcpyAsync(m, h->M, sizeof(float) * h->nparam, cudaMemcpyDeviceToHost, h->stream));
  if (v) CK(cudaMemcpyAsync(v, h->V, sizeof(float) * h->nparam, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (bp) { bp[0] = h->b1p; bp[1] = h->b2p; }
  return 0;
}
int sndvae_set_adam(sndvae_t* h, const float* m, const float* v, const float* bp) {
  if (!h) return SNDVAE_E_ARG;
  if (m) CK(cudaMemcpyAsync(h->M, m, sizeof(float) * h->nparam, cudaMemcpyHostToDevice, h->stream));
  if (v) CK(cudaMemcpyAsync(h->V, v, sizeof(float) * h->nparam, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (bp) { h->b1p = bp[0]; h->b2p = bp[1]; }
  return 0;
}
float* sndvae_params_device(sndvae_t* h) { return h ? h->P : nullptr; }
float* sndvae_grads_device(sndvae_t* h) { return h ? h->G : nullptr; }

int sndvae_forward(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host) {
  if (!h) return SNDVAE_E_ARG;
  return run(h, in, nz, out, losses_host, false, 0);
}
int sndvae_grads(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host, int64_t gb) {
  if (!h) return SNDVAE_E_ARG;
  return run(h, in, nz, out, losses_host, true, gb);
}
// this iteration's step size -> device scalar (outside any graph: it changes every step), then the running beta powers advance
static int adam_set_alpha(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) with fp32 running powers (TF ApplyAdam)
  h->pinned_loss[12] = c.learning_rate * sqrtf(1.f - h->b2p) / (1.f - h->b1p);
  CK(cudaMemcpyAsync(h->adam_alpha, h->pinned_loss + 12, sizeof(float), cudaMemcpyHostToDevice, h->stream));
  h->b1p *= c.adam_beta1; h->b2p *= c.adam_beta2;
  return 0;
}
static int adam_launch(sndvae_t* h) {
  const sndvae_config& c = h->cfg;
  long long n4 = h->nparam / 4;
  unsigned grid = cdiv(n4, 256); if (grid > 148 * 8) grid = 148 * 8;
  LAUNCH(tf_adam_k, grid, 256, 0, (float4*)h->P, (const float4*)h->G, (float4*)h->M, (float4*)h->V, n4, h->adam_alpha, 1.f - c.adam_beta1,
         1.f - c.adam_beta2, c.adam_eps);
  CK(cudaGetLastError());
  return 0;
}
int sndvae_apply_adam(sndvae_t* h) {
  if (!h) return SNDVAE_E_ARG;
  const sndvae_config& c = h->cfg;
  // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) with fp32 running powers (TF ApplyAdam)
  int r = adam_set_alpha(h); if (r) return r;
  return adam_launch(h);
}
int sndvae_zero_grads(sndvae_t* h) {
  if (!h) return SNDVAE_E_ARG;
  CK(cudaMemsetAsync(h->G, 0, sizeof(float) * h->nparam, h->stream));
  h->poisoned = 0;
  return 0;
}
int sndvae_grads_accumulate(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host, int64_t gb) {
  if (!h) return SNDVAE_E_ARG;
  return run(h, in, nz, out, losses_host, true, gb, nullptr, RUN_ACCUMULATE);
}
int sndvae_allreduce_grads(sndvae_t* h) {
  if (!h) return SNDVAE_E_ARG;
  if (!h->comm) return fail(h, SNDVAE_E_STATE, "sndvae_allreduce_grads: no communicator (call sndvae_comm_init first)");
  return allreduce_arena(h, false);
}
static unsigned long long graph_key_of(const sndvae_inputs* in, const sndvae_noise* nz, const sndvae_outputs* out) {
  unsigned long long k = 1469598103934665603ull;
  auto mix = [&k](const void* p) { k = (k ^ (unsigned long long)(uintptr_t)p) * 1099511628211ull; };
  const void* const* a = reinterpret_cast<const void* const*>(in);
  for (size_t i = 0; i < sizeof(*in) / sizeof(void*); ++i) mix(a[i]);
  a = reinterpret_cast<const void* const*>(nz);
  for (size_t i = 0; i < sizeof(*nz) / sizeof(void*); ++i) mix(a[i]);
  if (out) { a = reinterpret_cast<const void* const*>(out); for (size_t i = 0; i < sizeof(*out) / sizeof(void*); ++i) mix(a[i]); }
  return k | 1ull;
}
// The train step of a small problem is bound by launch latency (N = 25, B = 32: ~310 launches of a few microseconds each), so the
// second call with the same buffers captures the whole step on the handle's stream and later calls replay the graph.
static int train_step_graph(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host, bool* done) {
  *done = false;
  int r = check_inputs(h, in, nz); if (r) return r;
  const unsigned long long key = graph_key_of(in, nz, out);
  if (key != h->graph_key) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    if (h->graph_key == 0 || h->graph_key != (key ^ 2ull)) { h->graph_key = key ^ 2ull; return 0; }   // first sight of these buffers: run plainly (warms every lazy init)
    cudaGraph_t g = nullptr;
    // the legacy default stream (what a caller without a stream of its own passes) cannot be captured: the step is then
    // recorded on a stream of the handle's own -- nothing executes during capture -- and replayed on the caller's stream
    const cudaStream_t user = h->stream;
    if ((uintptr_t)user <= 2) {
      if (!h->gs && cudaStreamCreateWithFlags(&h->gs, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); h->graph_mode = 0; return 0; }
      h->stream = h->gs;
    }
    const cudaError_t eb = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (eb != cudaSuccess) {
      if (getenv("SNDVAE_GRAPH_DEBUG")) fprintf(stderr, "[sndvae graph] begin-capture: %s\n", cudaGetErrorString(eb));
      cudaGetLastError(); h->graph_mode = 0; h->stream = user; return 0;
    }
    h->capturing = 1;
    r = run_body(h, in, nz, out, nullptr, true, h->B, nullptr, RUN_NOFETCH);
    if (!r) r = adam_launch(h);
    h->capturing = 0;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->stream = user;
    cudaError_t ei = cudaSuccess;
    if (r || e != cudaSuccess || !g || (ei = cudaGraphInstantiate(&h->graph_exec, g, 0)) != cudaSuccess) {
      if (getenv("SNDVAE_GRAPH_DEBUG"))
        fprintf(stderr, "[sndvae graph] capture dropped: step rc %d (%s), end-capture %s, instantiate %s\n", r, h->err.c_str(), cudaGetErrorString(e), cudaGetErrorString(ei));
      if (g) cudaGraphDestroy(g);
      cudaGetLastError(); h->graph_exec = nullptr; h->graph_mode = 0;       // something in the step cannot be captured: stay on plain launches
      h->B = h->cfg.batch_size; h->BS = h->B * h->S; h->Rn = h->B * h->N;
      return 0;
    }
    cudaGraphDestroy(g);
    h->graph_key = key;
  }
  r = adam_set_alpha(h); if (r) return r;
  CK(cudaGraphLaunch(h->graph_exec, h->stream));
  h->launches += h->graph_launches; h->graph_replays++;
  *done = true;
  r = fetch_losses_finish(h, losses_host, 1);
  h->poisoned = r ? 1 : 0;
  return r;
}
int sndvae_train_step(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, sndvae_outputs* out, float* losses_host) {
  if (!h) return SNDVAE_E_ARG;
  if (h->graph_mode && !h->comm && !h->stt.on && h->cfg.loss_variant != SNDVAE_LOSS_CAPACITY && in && nz) {
    bool done = false;
    const long long l0 = h->launches;
    int r = train_step_graph(h, in, nz, out, losses_host, &done);
    if (r || done) return r;
    if (h->graph_exec == nullptr && h->graph_mode) {
      // plain run that the next call's capture will mirror: remember how many kernels one step launches
      r = run(h, in, nz, out, losses_host, true, (long long)h->world * h->B, nullptr, RUN_ALLREDUCE); if (r) return r;
      r = sndvae_apply_adam(h);
      h->graph_launches = h->launches - l0;
      return r;
    }
  }
  // with a communicator: local sums scaled by 1 / (world B), one all-reduce of the arena (+ the loss sums), then Adam
  int r = run(h, in, nz, out, losses_host, true, (long long)h->world * h->B, nullptr, RUN_ALLREDUCE); if (r) return r;
  return sndvae_apply_adam(h);
}
int sndvae_set_beta(sndvae_t* h, float beta) {
  if (!h) return SNDVAE_E_ARG;
  if (!(beta >= 0.f)) return fail(h, SNDVAE_E_ARG, "beta must be >= 0");
  h->cfg.beta = beta;
  return 0;
}

int sndvae_comm_unique_id(uint8_t* id_host) {
  if (!id_host) return SNDVAE_E_ARG;
  if (nccl_load()) return SNDVAE_E_CUDA;
  nccl_unique_id id;
  if (g_nccl.GetUniqueId(&id) != 0) return SNDVAE_E_CUDA;
  memcpy(id_host, id.internal, 128);
  return 0;
}
int sndvae_comm_init(sndvae_t* h, const uint8_t* id_host, int32_t rank, int32_t world) {
  if (!h || !id_host) return SNDVAE_E_ARG;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, SNDVAE_E_ARG, "sndvae_comm_init: rank %d of %d", rank, world);
  if (h->comm) return fail(h, SNDVAE_E_STATE, "sndvae_comm_init: the handle already has a communicator");
  if (const char* e = nccl_load()) return fail(h, SNDVAE_E_CUDA, "NCCL unavailable: %s", e);
  nccl_unique_id id; memcpy(id.internal, id_host, 128);
  CKN(g_nccl.CommInitRank(&h->comm, world, id, rank));
  h->rank = rank; h->world = world;
  return 0;
}
int sndvae_comm_info(const sndvae_t* h, int32_t* rank, int32_t* world) {
  if (!h) return SNDVAE_E_ARG;
  if (rank) *rank = h->rank;
  if (world) *world = h->comm ? h->world : 1;
  return 0;
}

int sndvae_generate(sndvae_t* h, const float* z_s, const float* z_sg, const float* z_g, sndvae_outputs* out) {
  if (!h || !z_sg || !out) return SNDVAE_E_ARG;
  if (h->dis && (!z_s || !z_g)) return fail(h, SNDVAE_E_ARG, "z_s / z_g required for the disentangled model");
  const sndvae_config& c = h->cfg;
  CK(cudaMemcpyAsync(h->z_sg, z_sg, sizeof(float) * h->BS * c.sg_latent_size, cudaMemcpyDeviceToDevice, h->stream));
  if (h->dis) {
    CK(cudaMemcpyAsync(h->z_s, z_s, sizeof(float) * h->B * c.s_latent_size, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->z_g, z_g, sizeof(float) * h->B * c.g_latent_size, cudaMemcpyDeviceToDevice, h->stream));
  }
  int r = decoder_prepare(h, false); if (r) return r;
  r = decoder_fwd(h, nullptr, out, false, (float)h->B); if (r) return r;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// host feeds -> device staging -> forward + backward (pipelined pieces); optional accumulate / all-reduce; no update
static int grads_host(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, int64_t* gen_adj_host, float* losses_host,
                      long long global_batch, int flags, const sndvae_inputs_compact* cin = nullptr, uint32_t* gen_bits_host = nullptr) {
  int r;
  if (cin) {
    if (!nz) return fail(h, SNDVAE_E_ARG, "noise struct is NULL");
    if (!cin->features || !cin->adj_bits || !cin->rel || !cin->adj_truth_bits || !cin->feature_truth || !cin->spatial_truth)
      return fail(h, SNDVAE_E_ARG, "a required compact feed (features, adj_bits, rel, adj_truth_bits, feature_truth, spatial_truth) is NULL");
    if (!nz->eps_sg || (h->dis && (!nz->eps_s || !nz->eps_g))) return fail(h, SNDVAE_E_ARG, "a required noise tensor is NULL");
  } else if ((r = check_inputs(h, in, nz))) return r;
  const sndvae_config& c = h->cfg; const int N = h->N; const long long B = h->B, BS = h->BS;
  if (!h->hf_ready) {
    if (h->hf_features) return fail(h, SNDVAE_E_STATE, "an earlier host-feed staging allocation failed; destroy the handle");
    DA(h->hf_features, BS * N * h->F); DA(h->hf_adj, BS * N * N); DA(h->hf_rel, BS * N * N); DA(h->hf_adj_truth, B * N * N);
    DA(h->hf_feature_truth, B * N * h->F); DA(h->hf_spatial_truth, B * N * h->D);
    DA(h->hf_eps_s, B * c.s_latent_size); DA(h->hf_eps_sg, BS * c.sg_latent_size); DA(h->hf_eps_g, B * c.g_latent_size);
    DA(h->hf_gen_adj, B * N * N);
    h->hf_ready = 1;
  }
  if (cin && !h->hc_ready) {
    if (h->hc_features) return fail(h, SNDVAE_E_STATE, "an earlier compact-feed staging allocation failed; destroy the handle");
    const long long W = (N + 31) / 32;
    DA(h->hc_features, B * N * h->F); DA(h->hc_rel, B * N * N); DA(h->hc_adj_bits, BS * N * W); DA(h->hc_adjt_bits, B * N * W);
    DA(h->hc_gen_bits, B * N * W);
    h->hc_ready = 1;
  }
  if (!h->cs) {
    CK(cudaStreamCreateWithFlags(&h->cs, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&h->ds, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  }
  sndvae_inputs din; memset(&din, 0, sizeof din);
  din.features = h->hf_features; din.adj = h->hf_adj; din.rel = h->hf_rel; din.adj_truth = h->hf_adj_truth;
  din.feature_truth = h->hf_feature_truth; din.spatial_truth = h->hf_spatial_truth;
  sndvae_noise dnz; dnz.eps_s = h->hf_eps_s; dnz.eps_sg = h->hf_eps_sg; dnz.eps_g = h->hf_eps_g;
  sndvae_outputs o; memset(&o, 0, sizeof o); o.generated_adj = (gen_adj_host || gen_bits_host) ? (int64_t*)h->hf_gen_adj : nullptr;
  HostFeeds hfd; hfd.in = in; hfd.nz = nz; hfd.gen_adj = gen_adj_host; hfd.cin = cin; hfd.gen_bits = gen_bits_host;
  return run(h, &din, &dnz, &o, losses_host, true, global_batch, &hfd, flags);
}
int sndvae_grads_host(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, int64_t* gen_adj_host, float* losses_host,
                      int64_t global_batch, int32_t accumulate) {
  if (!h) return SNDVAE_E_ARG;
  int r = grads_host(h, in, nz, gen_adj_host, losses_host, global_batch, accumulate ? RUN_ACCUMULATE : 0); if (r) return r;
  CK(cudaStreamSynchronize(h->ds));      // the caller's generated_adj buffer is complete on return
  return 0;
}
int sndvae_train_step_host(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* nz, int64_t* gen_adj_host, float* losses_host) {
  if (!h) return SNDVAE_E_ARG;
  int r = grads_host(h, in, nz, gen_adj_host, losses_host, (long long)h->world * h->B, RUN_ALLREDUCE); if (r) return r;
  r = sndvae_apply_adam(h);
  cudaError_t e1 = cudaStreamSynchronize(h->ds), e2 = cudaStreamSynchronize(h->stream);
  if (r) return r;
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "stream synchronize: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return 0;
}

int sndvae_train_step_host_compact(sndvae_t* h, const sndvae_inputs_compact* in, const sndvae_noise* nz, uint32_t* gen_bits_host,
                                   float* losses_host) {
  if (!h || !in) return SNDVAE_E_ARG;
  int r = grads_host(h, nullptr, nz, nullptr, losses_host, (long long)h->world * h->B, RUN_ALLREDUCE, in, gen_bits_host); if (r) return r;
  r = sndvae_apply_adam(h);
  cudaError_t e1 = cudaStreamSynchronize(h->ds), e2 = cudaStreamSynchronize(h->stream);
  if (r) return r;
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "stream synchronize: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return 0;
}

int sndvae_synth_inputs(sndvae_t* h, uint64_t seed, const sndvae_inputs* io) {
  if (!h || !io) return SNDVAE_E_ARG;
  const int N = h->N, F = h->F, D = h->D, S = h->S; const long long B = h->B, BS = h->BS;
  if (N > SY_MAXN) return fail(h, SNDVAE_E_ARG, "sndvae_synth_inputs supports num_nodes <= %d", SY_MAXN);
  if (!io->spatial_truth || !io->adj) return fail(h, SNDVAE_E_ARG, "sndvae_synth_inputs needs at least spatial_truth and adj buffers");
  const float r2 = (float)(6.0 / (3.14159265358979323846 * (double)N));       // radius of the random geometric graph: mean degree ~ 6
  LAUNCH(synth_graph_k, (unsigned)B, 256, sizeof(float) * N * D, (unsigned long long)seed, N, F, D, S, r2, (float*)io->spatial_truth,
         (float*)io->feature_truth, (float*)io->rel_truth, (float*)io->adj_truth, (float*)io->features, (float*)io->spatial);
  const size_t smem = (size_t)N * (8 + 4 + 4 * D + 1) + 16;
  LAUNCH(synth_sample_k, (unsigned)BS, SY_THREADS, smem, (unsigned long long)seed, N, D, S, r2, io->spatial_truth, (float*)io->adj, (float*)io->rel);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t sndvae_launch_count(const sndvae_t* h) { return h ? h->launches : 0; }
int64_t sndvae_graph_replays(const sndvae_t* h) { return h ? h->graph_replays : 0; }
int sndvae_set_global_iter(sndvae_t* h, int64_t it) { if (!h) return SNDVAE_E_ARG; h->global_iter = it; return 0; }

int sndvae_gemm_timing(sndvae_t* h, int reset, double* total_ms, int64_t* launches, double* flops) {
  if (!h) return SNDVAE_E_ARG;
  CK(cudaStreamSynchronize(h->stream));
  double ms = 0, fl = 0;
  for (size_t i = 0; i < h->ev_used; ++i) { float t = 0; cudaEventElapsedTime(&t, h->ev[i].a, h->ev[i].b); ms += t; fl += h->ev[i].flops; }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = (int64_t)h->ev_used;
  if (flops) *flops = fl;
  if (reset) h->ev_used = 0;
  return 0;
}

int sndvae_stage_times(sndvae_t* h, int32_t enable, char* names_host, double* ms_host, int32_t capacity, int64_t* steps_host) {
  if (!h) return SNDVAE_E_ARG;
  StageTimer& t = h->stt;
  int n = 0;
  if (names_host && ms_host) {
    CK(cudaStreamSynchronize(h->stream));
    for (auto& a : t.total) {
      if (n >= capacity) break;
      snprintf(names_host + (size_t)n * 32, 32, "%s", a.first.c_str()); ms_host[n] = a.second; ++n;
    }
  }
  if (steps_host) *steps_host = t.steps;
  t.total.clear(); t.steps = 0; t.used = 0;
  t.on = enable ? 1 : t.print;
  return n;
}

int sndvae_debug_gemm(sndvae_t* h, int32_t tA, int32_t tB, int64_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                      const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias) {
  if (!h || !A || !B || !C) return SNDVAE_E_ARG;
  cudaError_t e = tsgemm(h->stream, tA != 0, tB != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, &h->launches, h->gemm_ws, h->gemm_ws_floats);
  if (e != cudaSuccess) return fail(h, SNDVAE_E_CUDA, "tsgemm: %s", cudaGetErrorString(e));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int sndvae_threshold_logits(sndvae_t* h, const float* logits, int64_t n, int64_t* out) {
  if (!h || !logits || !out) return SNDVAE_E_ARG;
  LEW(threshold_logits_k, n, logits, (long long)n, (long long*)out);
  CK(cudaGetLastError()); CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t sndvae_debug_read(sndvae_t* h, const char* name, float* dst, int64_t cap) {
  if (!h || !name || !dst) return SNDVAE_E_ARG;
  const sndvae_config& c = h->cfg;
  struct { const char* n; const float* p; long long len; } tab[] = {
    {"fg", h->fg, h->dis ? h->Rn * (c.g_conv_hidden[1] + h->F) : 0}, {"fs", h->fs, h->dis ? h->Rn * c.s_channel[2] : 0},
    {"fsg", h->fsg, h->BS * h->N * c.sg_conv_hidden[1][2]}, {"n_sg", h->n_sg, h->Rn * h->H}, {"v", h->v, h->Rn * h->Chv},
    {"a", h->a, h->Rn * h->Chv}, {"c", h->c, h->Rn * h->Chv}, {"Rc", h->Rc, h->Rn * h->C1}, {"Sa", h->Sa, h->Rn * h->C1},
    {"E1", h->E1, (long long)h->Bc * h->N * h->N * h->C1}, {"O12", h->O12, 2LL * h->Bc * h->N * h->N * h->C2},
    {"dY12", h->dY12, 2LL * h->Bc * h->N * h->N * h->C1}, {"da", h->da, h->Rn * h->Chv}, {"dc", h->dc, h->Rn * h->Chv},
    {"dv", h->dv, h->Rn * h->Chv}, {"dfsg", h->dfsg, h->BS * h->N * c.sg_conv_hidden[1][2]}, {"loss", h->loss, 8},
  };
  for (auto& t : tab) if (!strcmp(t.n, name)) {
    long long n = t.len < cap ? t.len : cap;
    if (n <= 0 || !t.p) return 0;
    CK(cudaMemcpyAsync(dst, t.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
    return n;
  }
  return fail(h, SNDVAE_E_ARG, "unknown debug buffer '%s'", name);
}

}  // extern "C"
