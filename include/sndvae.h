/* sndvae.h -- C ABI of the B200-native SND-VAE train / generate step.
 *
 * The reference (xguo7/SND-VAE, TensorFlow 1.x) has no FFI: its only runtime
 * boundary is `sess.run(fetches, feed_dict)` on the graph built by
 * `SGCNModelVAE.__init__` (model.py:22, model_joint.py:14) and
 * `OptimizerVAE.__init__` (optimizer.py:124).  This header is the C surface a
 * drop-in for that boundary binds; each entry point
 * cites the reference call it replaces.
 *
 * Conventions: every function returns 0 on success or a negative SNDVAE_E_*
 * code; `sndvae_last_error(h)` gives the message.  No C++ exceptions cross the
 * ABI.  Pointers are DEVICE pointers unless the name ends in `_host`.  The
 * caller owns input/output buffers; the library owns parameters, gradients,
 * Adam state and workspace.  All work is issued on the stream given at create
 * time.  One handle per device, one host thread per handle (the reference has
 * one driver thread, main.py:310-349).  There is no CPU fallback: create fails
 * with SNDVAE_E_CUDA when no sm_100 device is present.
 */
#ifndef SNDVAE_H_
#define SNDVAE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNDVAE_OK          0
#define SNDVAE_E_ARG      -1   /* bad argument / shape (TF: InvalidArgumentError) */
#define SNDVAE_E_CUDA     -2   /* CUDA failure, or no usable device */
#define SNDVAE_E_DENSE    -3   /* a sampled adjacency row set exceeded the edge capacity */
#define SNDVAE_E_STATE    -4   /* call sequence error */

#define SNDVAE_MODEL_DISENTANGLED 0   /* model.py       (z_s, z_sg, z_g) */
#define SNDVAE_MODEL_BASE         1   /* model_joint.py (z_sg only, S = 1) */

/* Hyper-parameters = the reference's flags (main.py:42-103, synthetic2 block
 * main.py:181-209).  Layer counts are those of the reference (2 graph conv, 3
 * spatial conv, 2 SGC, 2 node deconv, 3 spatial deconv, 2 e2e); kernel size 5. */
typedef struct sndvae_config {
  int32_t model_type;        /* SNDVAE_MODEL_*                       main.py:72 */
  int32_t num_nodes;         /* N                                     main.py:243 */
  int32_t num_feature;       /* F  = FLAGS.num_feature                main.py:83 */
  int32_t spatial_dim;       /* D  = FLAGS.spatial_dim                main.py:84 */
  int32_t sampling_num;      /* S  = FLAGS.sampling_num (base: 1)     main.py:100 */
  int32_t node_h_size;       /* H                                     main.py:209 */
  int32_t s_channel[3];      /* spatial conv1d channels               main.py:183 */
  int32_t s_hidden_size, s_latent_size;
  int32_t g_conv_hidden[2];  /* graph conv channels                   main.py:190 */
  int32_t g_hidden_size, g_latent_size;
  int32_t sg_conv_hidden[2][3];                                    /* main.py:195 */
  int32_t sg_hidden_size, sg_latent_size;
  int32_t s_d_channel[3];    /* spatial deconv channels               main.py:200 */
  int32_t n_d_channel[2];    /* node deconv channels                  main.py:205 */
  int32_t e_d_hidden[2];     /* e2e channels                          main.py:209 */
  int32_t batch_size;        /* B graphs per step = FLAGS.batch_size  main.py:213 */
  int32_t chunk_graphs;      /* graphs per device micro-batch for the N^2 stages (0 = auto) */
  int32_t edge_capacity;     /* per-sample nnz capacity of `adj` (0 = 4N) */
  int32_t use_tensor_cores;  /* e2e layer 1 (layers.py:431-450): 2 (default): spectral -- Stockham FFT lines + per-frequency tcgen05
                              * bf16x3 channel-mix GEMMs; 1: block-Toeplitz tcgen05 bf16x3 GEMMs; 0: fp32 SIMT reference kernels */
  float   learning_rate;     /* FLAGS.learning_rate                   main.py:211 */
  float   beta;              /* KL weight                             main.py:515 */
  float   adam_beta1, adam_beta2, adam_eps;   /* tf.train.AdamOptimizer defaults */
  /* loss branch of OptimizerVAE (optimizer.py:160-190; disentangled model only):
   *   SNDVAE_LOSS_ELBO      'disentangled' / 'base':  mse + beta (kl_sg + kl_s + kl_g)
   *   SNDVAE_LOSS_CAPACITY  'disentangled_C':         mse + gamma relu(kl_sg - C) + kl_s + kl_g,
   *                          C = clip(C_max C_step / C_stop_iter (global_iter // C_step), 0, C_max)   (optimizer.py:170-172)
   *   SNDVAE_LOSS_DIP       'NED-VAE-IP':             mse + kl + beta sum_latents DIP(z_mean, lambda_od, lambda_d)  (optimizer.py:7-21,183)
   *   SNDVAE_LOSS_TC        'beta-TCVAE':             mse + beta kl + 10 sum_latents TC(z, z_mean, z_std): the minibatch
   *                          total-correlation estimate over all pairs of rows (optimizer.py:23-63,185-190); latent sizes <= 128
   * DIP: with a communicator (sndvae_comm_init) the second-moment sums of the posterior means are all-reduced, so the
   * covariance is the GLOBAL batch's (optimizer.py:7-21 couples every sample of the batch) and the 2-GPU step equals the
   * 1-GPU step on the concatenated batch (tests/test_multigpu.py).  TC is a statistic of the batch the handle sees (per rank
   * under data parallelism; its gradient is weighted batch_size / global_batch so that the all-reduced sum is the mean
   * over ranks). */
  int32_t loss_variant;
  float   gamma, C_max, C_stop_iter, C_step;  /* main.py:95-98 */
  float   dip_lambda_od, dip_lambda_d;        /* 10, 100 (optimizer.py:183) */
  /* Joint-encoder layer type.  2 (and 0): SpatialGraphConvolution (layers.py:143-198; FLAGS.dataset synthetic1/2/3,
   * model.py:137-138) with sg_conv_hidden[l][0..2] on per-sample edge lists (spanning-forest samples).
   * 3: SpatialGraphConvolution_3D (layers.py:200-277; FLAGS.dataset protein / mnist, model.py:139-140, main.py:225,241) with the
   * four hidden sizes sg_conv_hidden3[l][0..3] on dense adjacencies (any real A). */
  int32_t sg_hops;
  int32_t sg_conv_hidden3[2][4];
} sndvae_config;

#define SNDVAE_LOSS_ELBO      0
#define SNDVAE_LOSS_CAPACITY  1
#define SNDVAE_LOSS_DIP       2
#define SNDVAE_LOSS_TC        3

/* The eight feeds of construct_feed_dict_train (preprocessing.py:32-42) with
 * the static shapes of main.py:253-264.  fp32, C-contiguous.  `spatial` and
 * `rel_truth` are accepted and unused on this path. */
typedef struct sndvae_inputs {
  const float* features;       /* [B*S, N, F] */
  const float* spatial;        /* [B*S, N, D]   (unused; may be NULL) */
  const float* adj;            /* [B*S, N, N]   row b*S+s = sample s of graph b */
  const float* rel;            /* [B*S, N, N, 1] */
  const float* adj_truth;      /* [B, N, N] */
  const float* feature_truth;  /* [B, N, F] */
  const float* spatial_truth;  /* [B, N, D] */
  const float* rel_truth;      /* [B, N, N, 1]  (unused; may be NULL) */
} sndvae_inputs;

/* The tf.random.normal draws of get_z / get_random_z (model.py:153-169) made
 * explicit so that results are reproducible against the oracle. */
typedef struct sndvae_noise {
  const float* eps_s;          /* [B, Ls]     (NULL for base) */
  const float* eps_sg;         /* [B*S, Lsg] */
  const float* eps_g;          /* [B, Lg]     (NULL for base) */
} sndvae_noise;

/* Model attributes a caller fetches (model.py:78-80,114-151).  Any pointer may
 * be NULL = not fetched. */
typedef struct sndvae_outputs {
  float* z_mean_s;  float* z_std_s;   float* z_s;     /* [B, Ls]    */
  float* z_mean_g;  float* z_std_g;   float* z_g;     /* [B, Lg]    */
  float* z_mean_sg; float* z_std_sg;  float* z_sg;    /* [B*S, Lsg] */
  int64_t* generated_adj;         /* [B, N, N]    tf.argmax -> int64 (model.py:208) */
  float* generated_adj_prob;      /* [B, N, N, 2] masked logits   (model.py:207) */
  float* generated_spatial;       /* [B, N, D] */
  float* generated_node_feat;     /* [B, N, F] */
} sndvae_outputs;

typedef struct sndvae_handle sndvae_t;

/* One parameter of tf.trainable_variables() (model.py:92-94): TF variable
 * name, offset (in floats) into the flat arenas, shape. */
typedef struct sndvae_param_info {
  char    name[64];
  int64_t offset;
  int64_t size;
  int32_t rank;
  int32_t shape[4];
} sndvae_param_info;

/* Fill a config with the synthetic2 defaults (main.py:181-215). */
int sndvae_default_config(sndvae_config* cfg_host);

/* Replaces SGCNModelVAE(...) + OptimizerVAE(...) + tf.Session() +
 * global_variables_initializer (model.py:22, optimizer.py:124, main.py:301-302).
 * `stream` is a cudaStream_t (NULL = default stream).  Parameters start at
 * zero; load them with sndvae_set_params. */
int sndvae_create(const sndvae_config* cfg_host, void* stream, sndvae_t** out);
int sndvae_destroy(sndvae_t* h);
const char* sndvae_last_error(const sndvae_t* h);

/* tf.trainable_variables() in creation order. */
int64_t sndvae_param_count(const sndvae_t* h);        /* padded arena length in floats */
int32_t sndvae_num_params(const sndvae_t* h);
int sndvae_param_table(const sndvae_t* h, sndvae_param_info* table_host, int32_t capacity);

/* tf.train.Saver save/restore of variables and Adam slots (main.py:299,351-352).
 * Host buffers of sndvae_param_count floats. */
int sndvae_get_params(sndvae_t* h, float* params_host);
int sndvae_set_params(sndvae_t* h, const float* params_host);
int sndvae_get_adam(sndvae_t* h, float* m_host, float* v_host, float* beta_pows_host /*[2]*/);
int sndvae_set_adam(sndvae_t* h, const float* m_host, const float* v_host, const float* beta_pows_host);

/* Device pointers of the library-owned arenas (for NCCL all-reduce of the
 * gradient arena and zero-copy views from the host language). */
float* sndvae_params_device(sndvae_t* h);
float* sndvae_grads_device(sndvae_t* h);

/* sess.run([model.z_*, model.generated_*], feed) with FLAGS.type in
 * {'train','test_reconstruct'} (main.py:368-371): encoder, get_z, decoder.
 * losses_host (may be NULL) receives optimizer.overall_loss (optimizer.py:200-203):
 * [cost, spatial, adj, node, kl_g, kl_s, kl_sg] (base: 5 entries, last = kl_sg). */
int sndvae_forward(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* noise,
                   sndvae_outputs* out, float* losses_host);

/* optimizer.compute_gradients(cost) (optimizer.py:198): forward + backward
 * into the gradient arena, no update.  The arena holds LOCAL-shard sums scaled
 * by 1/(global batch): with `global_batch` = world * B the all-reduced arena is
 * the full-batch gradient.  global_batch <= 0 means B. */
int sndvae_grads(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* noise,
                 sndvae_outputs* out, float* losses_host, int64_t global_batch);

/* Gradient accumulation over device-resident micro-batches (BASELINE config 4: a global batch larger than what one
 * call holds).  The reference has no counterpart -- its batch is whatever one sess.run is fed (main.py:316-331) -- but
 * the ELBO is a mean over the batch (optimizer.py:144-162), so the gradient of a batch of k micro-batches is the sum
 * of k calls with global_batch = k * world * B:
 *     sndvae_zero_grads(h);  k x sndvae_grads_accumulate(h, ..., k * world * B);  [sndvae_allreduce_grads(h);]
 *     sndvae_apply_adam(h);
 * sndvae_grads == sndvae_zero_grads + one sndvae_grads_accumulate.  losses_host is the micro-batch's own mean. */
int sndvae_zero_grads(sndvae_t* h);
int sndvae_grads_accumulate(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* noise,
                            sndvae_outputs* out, float* losses_host, int64_t global_batch);

/* The apply half of AdamOptimizer.minimize (optimizer.py:197): TF1 ApplyAdam on
 * the whole arena using the current gradient arena. */
int sndvae_apply_adam(sndvae_t* h);

/* The KL weight `beta` of OptimizerVAE(..., beta=...) (optimizer.py:124,164; main.py:515), settable after create
 * because the reference passes it to the optimizer constructor, not to the model's. */
int sndvae_set_beta(sndvae_t* h, float beta);

/* Data parallelism (the reference is single-process).  One handle per rank; graphs are sharded, parameters
 * replicated.  Rank 0 calls sndvae_comm_unique_id and distributes the 128 bytes (any side channel: torch.distributed
 * broadcast, MPI, a file); every rank then calls sndvae_comm_init.  NCCL is resolved at run time from the process
 * (PyTorch's libnccl.so.2) -- the library has no link-time dependency on it.  With a communicator,
 * sndvae_train_step / sndvae_train_step_host run  grads (scaled 1 / (world B)) -> one ncclAllReduce(sum) of the
 * gradient arena and of the loss sums over NVLink on the handle's stream -> Adam, and losses_host is the mean over
 * the global batch.  sndvae_allreduce_grads is the bare collective for callers that drive sndvae_grads /
 * sndvae_grads_accumulate themselves. */
int sndvae_comm_unique_id(uint8_t* id_host /*[128]*/);
int sndvae_comm_init(sndvae_t* h, const uint8_t* id_host /*[128]*/, int32_t rank, int32_t world);
int sndvae_comm_info(const sndvae_t* h, int32_t* rank_host, int32_t* world_host);
int sndvae_allreduce_grads(sndvae_t* h);

/* sess.run([opt.opt_op, opt.overall_loss, model.generated_adj], feed)
 * (main.py:331): forward, backward and Adam on one device. */
int sndvae_train_step(sndvae_t* h, const sndvae_inputs* in, const sndvae_noise* noise,
                      sndvae_outputs* out, float* losses_host);

/* model.sample(z) with FLAGS.type == 'test_generation' (model.py:83-85,
 * 163-169,227-229): decoder only, from caller-provided latents. */
int sndvae_generate(sndvae_t* h, const float* z_s, const float* z_sg, const float* z_g,
                    sndvae_outputs* out);

/* Same call as sndvae_train_step but with HOST buffers in and out, like the
 * reference's feed_dict (numpy in, numpy out; main.py:327-331): copies the
 * feeds host->device, runs the step, copies losses and generated_adj back. */
int sndvae_train_step_host(sndvae_t* h, const sndvae_inputs* in_host, const sndvae_noise* noise_host,
                           int64_t* generated_adj_host, float* losses_host);

/* The gradient half of sndvae_train_step_host (host feeds in, generated_adj and the micro-batch's losses out, no
 * all-reduce, no update): with `accumulate` != 0 the gradient arena is added to, as sndvae_grads_accumulate does. */
int sndvae_grads_host(sndvae_t* h, const sndvae_inputs* in_host, const sndvae_noise* noise_host,
                      int64_t* generated_adj_host, float* losses_host, int64_t global_batch, int32_t accumulate);

/* Compact host feeds ("optional compact input formats").  The dense feed_dict of main.py:253-264 carries
 * S identical copies of `rel` / `features` (np.tile, main.py:307-309) and 0/1 adjacencies as fp32: 84 N^2 bytes per graph.
 * A loader that keeps what input_data.py:18-38,77-83 actually computes (one `rel` per graph, spanning forests as sets of
 * edges) can hand over 5.4 N^2 bytes instead:
 *   adjacency bit rows: W = ceil(N / 32) uint32 words per row, bit (j % 32) of word (j / 32) = (adj[i][j] != 0).
 * Only valid for 0/1 adjacencies (what the reference's loaders produce); the dense entry points stay the default. */
typedef struct sndvae_inputs_compact {
  const float*    features;        /* [B, N, F]     per graph (dense row b*S+s = graph b's, main.py:307) */
  const uint32_t* adj_bits;        /* [B*S, N, W]   sampled spanning forests */
  const float*    rel;             /* [B, N, N]     per graph */
  const uint32_t* adj_truth_bits;  /* [B, N, W] */
  const float*    feature_truth;   /* [B, N, F] */
  const float*    spatial_truth;   /* [B, N, D] */
} sndvae_inputs_compact;

/* sndvae_train_step_host on compact feeds: the packed bytes cross the bus, the dense tensors are rebuilt in HBM by
 * unpack kernels; generated_adj comes back as bit rows [B, N, W] (may be NULL).  Same step, same results. */
int sndvae_train_step_host_compact(sndvae_t* h, const sndvae_inputs_compact* in_host, const sndvae_noise* noise_host,
                                   uint32_t* generated_adj_bits_host, float* losses_host);

/* Number of library kernels launched since create (bench.py's gpu_launches). */
/* Device-side synthetic data (replaces input_data.py:18-38,54-96 for synthetic spatial graphs): fills the
 * caller's DEVICE feed buffers (shapes as in sndvae_inputs; any pointer except spatial_truth and adj may be NULL) with
 * random-geometric graphs in the unit square and `sampling_num` random spanning forests per graph, reproducibly from `seed`
 * (counter-based hash; oracle/synth.py gives the same arrays bit for bit).  The pointers are written despite the const. */
int sndvae_synth_inputs(sndvae_t* h, uint64_t seed, const sndvae_inputs* device_buffers);

int64_t sndvae_launch_count(const sndvae_t* h);

/* Small problems (batch_size * num_nodes^2 < 2^21 edge cells; SNDVAE_GRAPH=0/1 overrides) replay sndvae_train_step as one CUDA graph
 * once the same feed / output buffers are seen a second time: the N = 25 step is ~260 launches of a few microseconds each
 * (main.py's own configuration, BASELINE configs[0]).  Number of steps executed as a graph replay since create: */
int64_t sndvae_graph_replays(const sndvae_t* h);

/* The `global_iter` placeholder (main.py:262,329): only the 'disentangled_C' loss reads it (optimizer.py:172). */
int sndvae_set_global_iter(sndvae_t* h, int64_t global_iter);

/* Average duration (ms) and launch count of the e2e layer-1 GEMM launches
 * (forward + dgrad + wgrad) since the last reset, measured with CUDA events on
 * the handle's stream; used for bench.py's roofline block. */
int sndvae_gemm_timing(sndvae_t* h, int reset, double* total_ms_host, int64_t* launches_host,
                       double* flops_host);

/* Per-stage CUDA-event times of the step (the reference prints wall time per step only, main.py:348-350): with
 * `enable` != 0 every following step records an event at each stage boundary on the handle's stream.  A call returns
 * the totals accumulated since the previous call -- up to `capacity` stages, names in 32-byte slots of names_host, ms
 * in ms_host, the number of steps in *steps_host -- and resets them; the return value is the number of stages. */
int sndvae_stage_times(sndvae_t* h, int32_t enable, char* names_host, double* ms_host, int32_t capacity, int64_t* steps_host);

/* InnerProductDecoder._call (layers.py:400-410): logits[b] = z[b] z[b]^T, no activation (the layer returns the raw
 * product).  z: device [batch, num_nodes, dim] fp32, logits: device [batch, num_nodes, num_nodes] fp32; shapes are the
 * caller's, independent of the handle's config.  NOT used by the reference's models (model.py / model_joint.py decode edges
 * with the e2e layers); offered as a standalone operator because layers.py defines it.  tcgen05 bf16x3. */
int sndvae_inner_product_decode(sndvae_t* h, const float* z, int64_t batch, int32_t num_nodes, int32_t dim, float* logits);

/* The node-level contraction kernel on its own (tsgemm.cuh: the GEMM under `linear`, layers.py:566-576, tf.layers.conv1d and the
 * SpatialGraphConvolution coefficient products), exposed so that it can be checked against numpy at arbitrary shapes:
 * row-major C[M,N] = alpha op(A) op(B) + beta C (+ bias[N], may be NULL); tA: A is stored [K,M]; tB: B is stored [N,K].
 * Device pointers; synchronous. */
int sndvae_debug_gemm(sndvae_t* h, int32_t tA, int32_t tB, int64_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t lda,
                      const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias);

/* argmax(softmax([l0,l1])) of model.py:208 on caller logits [n,2] -> int64 [n];
 * exposed so the thresholding rule can be checked bit-exactly on its own. */
int sndvae_threshold_logits(sndvae_t* h, const float* logits, int64_t n, int64_t* out);

/* Read a named intermediate buffer of the last chunk (debug / parity tests).
 * Returns the number of floats written, or a negative error. */
int64_t sndvae_debug_read(sndvae_t* h, const char* name, float* dst_host, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* SNDVAE_H_ */
