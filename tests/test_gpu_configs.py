"""GPU parity tests at BASELINE.json's own configurations (run under gpurun with -m gpu): N = 256 (configs 2-3: the 3-latent
model and the base model) with EVERY gradient against the fp64 oracle, N = 1024 (config 4), gradient accumulation over
micro-batches (config 3's global batch), the Saver stand-in's save -> restore -> identical next step, the beta setter.
All calls go through ctypes -> the C ABI of libsndvae.so.  Tolerances as in test_gpu_parity.py: losses / outputs rtol 1e-4,
gradients 1e-3 of the tensor's largest magnitude, adjacency bit-exact given the logits."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch

from oracle import sndvae_oracle as O

pytestmark = pytest.mark.gpu


def _setup(N, B, S, model, dtype=torch.float64, perturb=0.05, seed_in=5, mesh=False, **kw):
    cfg = O.Config(num_nodes=N, model_type=model, sampling_num=S, **kw)
    P = O.init_params(cfg, 7, dtype)
    g = torch.Generator().manual_seed(1)
    for k in P:
        P[k] = P[k] + perturb * torch.randn(P[k].shape, generator=g, dtype=dtype)
    return cfg, P, O.synthetic_inputs(cfg, B, seed_in, dtype, mesh=mesh), O.synthetic_noise(cfg, B, 9, dtype)


def _relmax(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _check_against_oracle(built, cfg, P, inp, noise, mode, tc, chunk, model, grad_max_tol=1e-3):
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, mode)
    B = inp["adj_truth"].shape[0]
    eng = built.Engine(built.make_config(cfg.N, B, model, sampling_num=cfg.S, use_tensor_cores=tc, chunk_graphs=chunk))
    eng.set_params(P)
    res = eng.grads(inp, noise, fetch=("generated_adj_prob", "generated_adj", "generated_spatial", "generated_node_feat", "z_mean_sg"))
    np.testing.assert_allclose(res["overall_loss"], [x.item() for x in L["overall_loss"]], rtol=1e-4)
    ref = {**enc, **dec}
    for k in ("generated_adj_prob", "generated_spatial", "generated_node_feat", "z_mean_sg"):
        assert _relmax(res[k].cpu().numpy(), ref[k].detach().numpy()) < 1e-4, k
    lg = res["generated_adj_prob"].cpu()
    assert torch.equal(torch.argmax(torch.softmax(lg, -1), -1), res["generated_adj"].cpu())
    # cells whose logit gap exceeds what the (checked above: 1e-4 of the largest logit) device error can close must threshold alike
    gap = max(1e-5, 2e-4 * float(ref["generated_adj_prob"].abs().max()))
    margin = (ref["generated_adj_prob"][..., 1] - ref["generated_adj_prob"][..., 0]).abs() > gap
    assert torch.equal(res["generated_adj"].cpu()[margin], ref["generated_adj"][margin])
    gg = eng.get_grads()
    worst = max((_relmax(gg[k].numpy(), v.numpy()), k) for k, v in grads.items())
    worst2 = max((_rel_l2(gg[k].numpy(), v.numpy()), k) for k, v in grads.items())
    print(f"[parity N={cfg.N} {model}] worst gradient: max-norm {worst}, l2 {worst2}")
    assert worst[0] < grad_max_tol, worst
    assert worst2[0] < 1e-3, worst2
    eng.close()


@pytest.mark.parametrize("model,S", [("disentangled", 2), ("base", 1)])
def test_n256_all_gradients_vs_oracle(built, model, S):
    """BASELINE configs 2-3 at their N: two graphs (two micro-batches of one), every output and every parameter gradient of the
    spectral tensor-core path against the fp64 block-Toeplitz restatement (T is 0.5 GB in fp64)."""
    cfg, P, inp, noise = _setup(256, 2, S, model)
    _check_against_oracle(built, cfg, P, inp, noise, "factored", 2, 1, model)


def test_n1024_all_gradients_vs_oracle(built):
    """BASELINE config 4 (N = 1024, mesh-like inputs, D = 2 as the synthetic flags have it): transform length 1536 = 3 * 2^9
    (runtime-plan FFT kernels).  Checker: the oracle's torch.fft form (== the Toeplitz form to 1e-12, tests/test_oracle.py)."""
    cfg, P, inp, noise = _setup(1024, 2, 2, "disentangled", mesh=True)
    # Gradient tolerance.  With two graphs nothing averages out the relu-mask decisions of BN_e1(E1) > 0: the device's fp32 /
    # split-bf16 forward differs from the fp64 oracle by ~1e-5 relative, which flips the mask of ~1e-5 of the 52 M layer-0 cells of
    # a graph; every flip moves one row of da / dc (a cancelling sum of N*50 terms) by ~1/sqrt(N*50) of its size, and
    # d_*_lin1/Matrix sums only B = 2 such rows.  So the max-norm bound is 3e-3 here (measured 1.1e-3 on d_sg_lin1/Matrix) while
    # every tensor still agrees to 1e-3 in the l2 norm, which sparse flips do not move.
    _check_against_oracle(built, cfg, P, inp, noise, "fft", 2, 1, "disentangled", grad_max_tol=3e-3)


@pytest.mark.parametrize("N,B,S,tol", [(200, 2, 2, 2e-3), (201, 2, 2, 2e-3), (112, 3, 2, 1e-3), (97, 2, 3, 1e-3), (800, 1, 2, 3e-3)])
def test_ragged_sizes_of_the_compile_time_transforms_vs_oracle(built, N, B, S, tol):
    """Sizes that do not fill their transform length: N = 200 (L = 384: the two-pass kernels with 56 padding positions inside their
    pruned first pass), N = 201 (odd: lines are not whole 16-byte pieces, so the three-pass kernels take them), N = 112 and 97
    (L = 192), N = 800 (L = 1536: the three-pass (line, channel-group) kernels with 224 padding positions) -- every output and every
    gradient against the oracle's torch.fft form.  (Max-norm tolerance as in test_n1024_all_gradients_vs_oracle: relu-mask flips.)"""
    cfg, P, inp, noise = _setup(N, B, S, "disentangled")
    _check_against_oracle(built, cfg, P, inp, noise, "fft", 2, 1, "disentangled", grad_max_tol=tol)


def test_gradient_accumulation_equals_full_batch(built):
    """BASELINE config 3 (global batch larger than one call holds): zero_grads + k x grads_accumulate(global_batch = k B) on
    micro-batches == grads on the concatenated batch; then one Adam step gives the same parameters."""
    N, B, S, k = 12, 6, 3, 3
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", dtype=torch.float32)
    full = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S))
    full.set_params(P)
    rf = full.grads(inp, noise)
    g_full = full.get_grads()
    full.apply_adam(); P_full = full.get_params(); full.close()
    b = B // k
    eng = built.Engine(built.make_config(N, b, "disentangled", sampling_num=S))
    eng.set_params(P)
    eng.zero_grads()
    costs = []
    for m in range(k):
        sl = slice(m * b, (m + 1) * b); sls = slice(m * b * S, (m + 1) * b * S)
        si = {kk: (v[sls] if kk in ("adj", "features", "spatial", "rel") else v[sl]) for kk, v in inp.items()}
        sn = {"eps_s": noise["eps_s"][sl], "eps_g": noise["eps_g"][sl], "eps_sg": noise["eps_sg"][sls]}
        costs.append(eng.grads_accumulate(si, sn, global_batch=B)["overall_loss"])
    g_acc = eng.get_grads()
    for kk in g_full:
        assert _relmax(g_acc[kk].numpy(), g_full[kk].numpy()) < 2e-5, kk
    np.testing.assert_allclose(np.mean(costs, axis=0), rf["overall_loss"], rtol=2e-5)
    eng.apply_adam()
    P_acc = eng.get_params(); eng.close()
    for kk in P_full:
        np.testing.assert_allclose(P_acc[kk].numpy(), P_full[kk].numpy(), rtol=0, atol=2e-6, err_msg=kk)


def test_checkpoint_roundtrip_identical_next_step(built, tmp_path):
    """tf.train.Saver stand-in (main.py:299,351-352,376): save after two Adam steps, restore into a fresh model, and the third
    step (losses, adjacency, parameters, Adam slots) is identical to the uninterrupted run; the same path string round-trips."""
    flags = import_module("snd-vae_b200.flags"); model_m = import_module("snd-vae_b200.model")
    sess_m = import_module("snd-vae_b200.session")
    F = flags.FLAGS; F.reset(); F.apply_dataset("synthetic2")
    F.type = "train"; F.batch_size = 3; F.sampling_num = 2
    N = 10
    cfg = O.Config(num_nodes=N, sampling_num=2)
    inp = O.synthetic_inputs(cfg, 3, 5, torch.float32); noise = O.synthetic_noise(cfg, 3, 9, torch.float32)
    ph = sess_m.make_placeholders(F.batch_size, F.sampling_num, N, F.num_feature, F.spatial_dim)
    a = model_m.SGCNModelVAE(ph, F.num_feature, N)
    for _ in range(2):
        a.engine.train_step(inp, noise)
    path = str(tmp_path / "model_epoch_2.ckpt")              # no .npz suffix, as saver.save(sess, path) is called
    a.save(path)
    ra = a.engine.train_step(inp, noise)
    b = model_m.SGCNModelVAE(ph, F.num_feature, N, seed=99)  # different initial weights: everything must come from the file
    b.restore(path)
    rb = b.engine.train_step(inp, noise)
    np.testing.assert_allclose(ra["overall_loss"], rb["overall_loss"], rtol=1e-6)     # loss sums go through atomics
    assert torch.equal(ra["generated_adj"].cpu(), rb["generated_adj"].cpu())
    Pa, Pb = a.engine.get_params(), b.engine.get_params()
    (ma, va, bpa), (mb, vb, bpb) = a.engine.get_adam(), b.engine.get_adam()
    assert np.array_equal(bpa, bpb)
    for k in Pa:     # atomics order is not fixed: identical up to the last bits
        np.testing.assert_allclose(Pa[k].numpy(), Pb[k].numpy(), rtol=0, atol=1e-7, err_msg=k)
        np.testing.assert_allclose(ma[k].numpy(), mb[k].numpy(), rtol=1e-5, atol=1e-9, err_msg=k)
        np.testing.assert_allclose(va[k].numpy(), vb[k].numpy(), rtol=1e-5, atol=1e-12, err_msg=k)
    # the file is keyed by TF variable names
    z = np.load(path + ".npz")
    assert "decoder/e1_deconv/w1" in z and "adam_m/encoder/g_sg1_lin/Matrix" in z and z["adam_beta_pows"].shape == (2,)
    F.reset()


@pytest.mark.parametrize("variant,beta", [("disentangled", 4.0), ("NED-VAE-IP", 0.5)])
def test_beta_reaches_the_engine(built, variant, beta):
    """OptimizerVAE(..., beta=...) (optimizer.py:124): beta weights the KL terms (optimizer.py:164) or the DIP regulariser
    (optimizer.py:183); set after create through sndvae_set_beta."""
    N, B, S = 9, 4, 2
    cfg, P, inp, noise = _setup(N, B, S, "disentangled")
    cfg.loss_variant = variant; cfg.beta = beta
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, "factored")
    eng = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S, loss_variant=built._lib.LOSS_VARIANTS[variant]))
    eng.set_params(P); eng.set_beta(beta)
    res = eng.grads(inp, noise)
    np.testing.assert_allclose(res["overall_loss"], [x.item() for x in L["overall_loss"]], rtol=1e-4)
    gg = eng.get_grads()
    floor = 0.0
    if variant == "NED-VAE-IP":      # see test_loss_variants: bias-like DIP gradients are cancellation noise
        floor = 1e-2 * max(grads[k].abs().max().item() for k in grads if "_lin/" in k and k.startswith("encoder/"))
    for k, v in grads.items():
        err = np.abs(gg[k].double().numpy() - v.numpy()).max() / max(v.abs().max().item(), floor, 1e-30)
        assert err < 1e-3, (k, err)
    eng.close()


def test_session_accepts_the_initializer_fetch(built):
    """main.py:301-302: sess.run(tf.global_variables_initializer()) is a no-op here and must not raise."""
    sess_m = import_module("snd-vae_b200.session")
    with sess_m.Session() as sess:
        assert sess.run(sess_m.global_variables_initializer()) is None


def test_compact_host_feeds_equal_dense_host_feeds(built):
    """sndvae_train_step_host_compact (bit-row adjacencies, per-graph rel / features) against
    sndvae_train_step_host on the dense feed_dict arrays: same losses, same adjacency (bit rows unpack to the int64 tensor),
    same parameters after Adam; ragged pieces (B = 5, chunk 2) and N not a multiple of 32."""
    data = import_module("snd-vae_b200.data")
    N, B, S = 37, 5, 3
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", dtype=torch.float32)
    dense = {k: np.ascontiguousarray(inp[k].numpy()) for k in ("features", "adj", "rel", "adj_truth", "feature_truth", "spatial_truth")}
    nz = {k: noise[k].numpy() for k in noise}
    a = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S, chunk_graphs=2))
    a.set_params(P)
    gen = np.zeros((B, N, N), np.int64); la = np.zeros(8, np.float32)
    a.train_step_host(dense, nz, gen, la)
    Pa = a.get_params(); a.close()
    compact = data.pack_feeds(dense, S)
    assert sum(v.nbytes for v in compact.values()) * 4 < sum(v.nbytes for v in dense.values())
    b = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S, chunk_graphs=2))
    b.set_params(P)
    bits = np.zeros((B, N, (N + 31) // 32), np.uint32); lb = np.zeros(8, np.float32)
    b.train_step_host_compact(compact, nz, bits, lb)
    Pb = b.get_params(); b.close()
    np.testing.assert_allclose(lb, la, rtol=1e-6)
    assert np.array_equal(data.unpack_adj_bits(bits, N), gen)
    for k in Pa:
        np.testing.assert_allclose(Pb[k].numpy(), Pa[k].numpy(), rtol=0, atol=2e-6, err_msg=k)
    bad = dict(dense); bad["adj"] = dense["adj"] * 0.5
    with pytest.raises(ValueError, match="0/1"):
        data.pack_feeds(bad, S)


GEMM_SHAPES = [  # tA, tB, M, N, K, beta, bias
    (0, 0, 1000, 50, 92, 0.0, False),      # SGC coef2 . W2: rows x [2C+2+h0] (16-byte aligned rows: vector loads)
    (0, 0, 777, 20, 23, 0.0, False),       # layer-0 SGC coefficient rows (lda = 23: scalar loads), ragged M
    (0, 0, 300, 20, 1, 0.0, False),        # K = 1 (xphi . M1a with one input feature)
    (0, 0, 1500, 50, 250, 1.0, True),      # conv1d as im2col rows x kernel, bias in the epilogue, beta = 1
    (0, 0, 300, 100, 5000, 0.0, True),     # latent head: K = N * channels (two-level accumulation over 157 chunks)
    (0, 0, 64, 5120, 100, 1.0, False),     # d_*_lin1: z . Matrix with N * H columns (40 column tiles)
    (0, 0, 130, 130, 70, 0.0, False),      # a 2-column last tile (MMA N = 16)
    (0, 1, 1000, 92, 50, 0.0, False),      # input gradients dX = dY . W^T
    (0, 1, 500, 250, 50, 0.0, False),
    (1, 0, 92, 50, 100000, 1.0, False),    # weight gradients: reduction over rows split over CTAs, atomics, accumulate
    (1, 0, 250, 50, 40000, 0.0, False),    # ... beta = 0 (the launcher clears C)
    (1, 0, 1, 20, 30000, 1.0, False),      # bias-row gradient of the coefficient products (M = 1)
    (1, 0, 5000, 100, 4096, 1.0, False),   # head weight gradient [N * channels, hidden]
    (1, 0, 100, 100, 3000, 0.0, False),    # DIP covariance mu^T mu
    # M >= 148 tiles: the persistent kernel with op(B) resident in shared memory
    (0, 0, 20000, 50, 92, 0.0, False),     # SGC layer-1 coef2 . W2
    (0, 0, 19001, 20, 23, 0.0, False),     # unaligned rows (scalar loads), ragged last tile
    (0, 0, 25000, 50, 250, 1.0, True),     # conv1d rows x [5 Ci, Co], bias + beta, 8 resident chunks
    (0, 1, 19500, 92, 50, 0.0, False),     # dcoef2 = dm2s . W2^T: 128-wide N tile
    (0, 1, 20000, 250, 50, 0.0, False),    # conv1d input gradient rows x [Co, 5 Ci]^T: two N tiles
    (0, 0, 19000, 20, 1, 0.0, False),      # K = 1
    (0, 0, 40000, 71, 70, 1.0, False),     # odd widths both ways
]


@pytest.mark.parametrize("tA,tB,M,N,K,beta,use_bias", GEMM_SHAPES)
def test_node_level_gemm_kernel(built, tA, tB, M, N, K, beta, use_bias):
    """tsgemm.cuh (the tcgen05 split-bf16 GEMM under `linear`, conv1d and the SGC coefficient products) against numpy fp64 at
    every shape family the step uses, with ragged tiles, unaligned leading dimensions, K tails, transposed operands, bias,
    beta and split-K accumulation.  Bound: 4e-6 of sum |a||b| per output -- fp32-grade (a 3-pass bf16 split would sit at ~1e-5)."""
    eng = built.Engine(built.make_config(8, 2, "disentangled", sampling_num=2))
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g)
    B = torch.randn((N, K) if tB else (K, N), generator=g)
    C0 = torch.randn((M, N), generator=g) if beta != 0 else None
    bias = torch.randn(N, generator=g) if use_bias else None
    alpha = 0.75
    out = eng.debug_gemm(A, B, tA=bool(tA), tB=bool(tB), alpha=alpha, beta=beta, C0=C0, bias=bias).cpu().double().numpy()
    a = (A.T if tA else A).double().numpy(); b = (B.T if tB else B).double().numpy()
    want = alpha * (a @ b) + (beta * C0.double().numpy() if C0 is not None else 0.0) + (bias.double().numpy() if bias is not None else 0.0)
    bound = 4e-6 * (np.abs(a) @ np.abs(b)) + 1e-6 * np.abs(want) + 1e-6
    assert np.isfinite(out).all()
    assert (np.abs(out - want) <= bound).all(), float((np.abs(out - want) / bound).max())
    eng.close()


def test_drivers_train_reconstruct_generate(built, tmp_path):
    """The three loops of main.py (train 299-356, test_reconstruct 374-426, test_generation 428-469) through the shims: batching
    with the last partial batch dropped, accuracy / loss bookkeeping, checkpoint every `save_every` epochs, restore by path, the
    z_*.npy dumps (main.py:411-416), and quirk Q11 (generation returns the encoder's posterior means)."""
    flags = import_module("snd-vae_b200.flags"); model_m = import_module("snd-vae_b200.model"); opt_m = import_module("snd-vae_b200.optimizer")
    sess_m = import_module("snd-vae_b200.session"); drv = import_module("snd-vae_b200.drivers"); data = import_module("snd-vae_b200.data")
    F = flags.FLAGS; F.reset(); F.apply_dataset("synthetic2")
    N, B, S, G = 12, 3, 2, 7                                   # 7 graphs -> two batches of 3, one graph dropped
    F.batch_size = B; F.sampling_num = S; F.epochs = 2
    d = data.synthetic_graphs(N, G, S, seed=3)
    ph = sess_m.make_placeholders(B, S, N, F.num_feature, F.spatial_dim)

    def build(kind):
        F.type = kind
        m = model_m.SGCNModelVAE(ph, F.num_feature, N)
        o = opt_m.OptimizerVAE(preds_edge=m.generated_adj_prob, preds_node=m.generated_node_feat, preds_spatial=m.generated_spatial,
                               labels_edge=ph["adj_truth"], labels_node=ph["feature_truth"], labels_spatial=ph["spatial_truth"],
                               labels_rel=ph["rel_truth"], global_iter=ph["global_iter"], model=m, num_nodes=N, pos_weight=1.0, norm=1.0, beta=1)
        return m, o
    m, o = build("train")
    log = drv.LossesLogger(str(tmp_path / "train_losses.csv"))
    check, hist = drv.train(m, o, ph, d, ckpt_dir=str(tmp_path / "ckpt"), save_every=1, logger=log)
    assert check.shape == (2, B, N, N) and len(hist) == 2 and set(hist[0]) >= {"loss", "adj_acc", "graph_kl", "spatial_kl", "sg_kl"}
    assert all(np.isfinite(list(h.values())).all() for h in hist) and hist[1]["loss"] < hist[0]["loss"]
    ck = str(tmp_path / "ckpt" / "model_dgt_global_1.ckpt")
    assert os.path.exists(ck + ".npz") and len(log.rows) == 2 * len(hist[0])
    Ptrained = {k: v.clone() for k, v in m.engine.get_params().items()}
    mr, _ = build("test_reconstruct")
    rec = drv.reconstruct(mr, ph, d, restore=ck, out_dir=str(tmp_path / "qual"), vae_type="disentangled")
    for k, L in (("z_s", F.s_latent_size), ("z_sg", F.sg_latent_size), ("z_g", F.g_latent_size)):
        z = np.load(tmp_path / "qual" / f"disentangled_{k}.npy")
        assert z.shape == (2, B, L) and np.array_equal(z, rec[k])
    assert rec["generated_adj"].shape == (2 * B, N, N) and rec["generated_spatial"].shape == (2 * B, N, 2)
    # the restored model is the trained one: its posterior means are those of a direct forward pass on the first batch
    eng = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S))
    eng.set_params(Ptrained)
    first = {k: torch.from_numpy(np.ascontiguousarray(d[k][: B * S] if k in ("features", "spatial", "adj", "rel") else d[k][:B])) for k in d}
    f = eng.forward(first, O.synthetic_noise(O.Config(num_nodes=N, sampling_num=S), B, 9, torch.float32), fetch=("z_mean_s", "z_mean_sg"))
    np.testing.assert_allclose(rec["z_s"][0], f["z_mean_s"].cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rec["z_sg"][0], f["z_mean_sg"].cpu().numpy().reshape(B, S, -1).mean(1), rtol=1e-5, atol=1e-6)
    eng.close()
    mg, _ = build("test_generation")
    gen = drv.generate(mg, ph, d, restore=ck)
    np.testing.assert_allclose(gen["z_s"], rec["z_s"], rtol=1e-6, atol=1e-7)          # encoder means, whatever the decoder is fed (quirk Q11)
    idx = np.arange(N)
    assert gen["generated_adj"].shape == (2 * B, N, N) and (gen["generated_adj"][:, idx, idx] == 0).all()
    assert not np.array_equal(gen["generated_spatial"], rec["generated_spatial"])     # prior draws, not posterior samples
    with pytest.raises(ValueError):
        drv.generate(mr, ph, d)
    F.reset()


@pytest.mark.parametrize("node_h,hid", [(50, 500), (5, 50)])
def test_other_node_h_sizes_vs_oracle(built, node_h, hid):
    """The reference's other flag blocks: synthetic1 (main.py:128-172: node_h_size = 50, sg hidden / latent 500) and the decoder
    shape of `protein` (main.py:230: node_h_size = 5).  The tensor-core tiles are built for node_h_size = 20, so these run the
    edge decoder on the fp32 SIMT kernels (selected by sndvae_create); losses, outputs and every gradient against the oracle."""
    N, B, S = 25, 3, 2
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", node_h_size=node_h, sg_hidden_size=hid, sg_latent_size=hid)
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, "factored")
    eng = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S, node_h_size=node_h, sg_hidden_size=hid, sg_latent_size=hid, chunk_graphs=2))
    assert eng.cfg.use_tensor_cores == 2          # asked for the default; the handle itself switched to the SIMT edge decoder
    eng.set_params(P)
    res = eng.grads(inp, noise, fetch=("generated_adj_prob", "generated_spatial", "generated_node_feat"))
    np.testing.assert_allclose(res["overall_loss"], [x.item() for x in L["overall_loss"]], rtol=1e-4)
    for k in ("generated_adj_prob", "generated_spatial", "generated_node_feat"):
        assert _relmax(res[k].cpu().numpy(), dec[k].detach().numpy()) < 1e-4, k
    gg = eng.get_grads()
    worst = max((_relmax(gg[k].numpy(), v.numpy()), k) for k, v in grads.items())
    assert worst[0] < 1e-3, worst
    eng.close()


def test_synthetic1_flag_block_trains(built):
    """flags.apply_dataset('synthetic1') (main.py:128-172) -> SGCNModelVAE / OptimizerVAE / Session.run: constructs and trains."""
    flags = import_module("snd-vae_b200.flags"); model_m = import_module("snd-vae_b200.model"); opt_m = import_module("snd-vae_b200.optimizer")
    sess_m = import_module("snd-vae_b200.session"); prep = import_module("snd-vae_b200.preprocessing"); data = import_module("snd-vae_b200.data")
    F = flags.FLAGS; F.reset(); F.apply_dataset("synthetic1"); F.type = "train"; F.batch_size = 4; F.sampling_num = 3
    N = 25
    d = data.synthetic_graphs(N, F.batch_size, F.sampling_num, seed=11)
    ph = sess_m.make_placeholders(F.batch_size, F.sampling_num, N, F.num_feature, F.spatial_dim)
    m = model_m.SGCNModelVAE(ph, F.num_feature, N)
    o = opt_m.OptimizerVAE(preds_edge=m.generated_adj_prob, preds_node=m.generated_node_feat, preds_spatial=m.generated_spatial,
                           labels_edge=ph["adj_truth"], labels_node=ph["feature_truth"], labels_spatial=ph["spatial_truth"],
                           labels_rel=ph["rel_truth"], global_iter=ph["global_iter"], model=m, num_nodes=N, pos_weight=1.0, norm=1.0, beta=1)
    assert m.engine.cfg.node_h_size == 50 and m.engine.cfg.sg_latent_size == 500
    fd = prep.construct_feed_dict_train(d["features"], d["spatial"], d["adj"], d["rel"], d["adj_truth"], d["feature_truth"], d["spatial_truth"], d["rel_truth"], ph)
    costs = []
    with sess_m.Session() as sess:
        sess.run(sess_m.global_variables_initializer())
        for _ in range(4):
            costs.append(sess.run([o.opt_op, o.cost], feed_dict=fd)[1])
    assert np.isfinite(costs).all() and costs[-1] < costs[0]
    F.reset()


def test_cuda_graph_step_equals_plain_step(built):
    """Small problems replay the device-resident train step as one CUDA graph (captured on the second call with the same
    buffers).  Five steps with the graph against five plain steps: same losses, same adjacency, same parameters; new feed buffers
    half-way force a re-capture."""
    N, B, S = 25, 4, 3
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", dtype=torch.float32)
    out = {}
    for mode in ("1", "0"):
        os.environ["SNDVAE_GRAPH"] = mode
        try:
            eng = built.Engine(built.make_config(N, B, "disentangled", sampling_num=S))
        finally:
            os.environ.pop("SNDVAE_GRAPH", None)
        eng.set_params(P)
        feeds = {k: v.to("cuda") for k, v in inp.items()}; nz = {k: v.to("cuda") for k, v in noise.items()}
        ipk, npk, keep = eng._pack(feeds, nz)
        o, res = eng._outs(("generated_adj",))
        losses = np.zeros(8, np.float32)
        hist, l0 = [], eng.launch_count()
        for step in range(5):
            if step == 3:      # other buffers with the same contents: the graph must be rebuilt, not replayed on stale pointers
                feeds2 = {k: v.clone() for k, v in feeds.items()}
                ipk, npk, keep2 = eng._pack(feeds2, nz)
            eng.train_step_packed(ipk, npk, o, losses)
            hist.append(losses[:7].copy())
        assert eng.launch_count() - l0 > 5 * 100
        # steps 0 (plain) 1 (capture + replay) 2 (replay) 3 (new buffers: plain) 4 (capture + replay)
        assert eng.graph_replays() == (3 if mode == "1" else 0)
        out[mode] = (np.array(hist), res["generated_adj"].cpu().numpy().copy(), eng.get_params())
        eng.close()
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=2e-5, atol=1e-8)
    assert np.array_equal(out["1"][1], out["0"][1])
    for k in out["0"][2]:
        np.testing.assert_allclose(out["1"][2][k].numpy(), out["0"][2][k].numpy(), rtol=0, atol=3e-6, err_msg=k)


PROTEIN = dict(spatial_dim=3, node_h_size=5, sg_conv_hidden=((10, 10, 10, 10), (20, 20, 20, 20)), sg_hidden_size=50, sg_latent_size=50,
               s_hidden_size=5, s_latent_size=5, g_hidden_size=5, g_latent_size=5)     # main.py:218-236


def _protein_engine(built, N, B, S, **kw):
    return built.Engine(built.make_config(N, B, "disentangled", sampling_num=S, spatial_dim=3, node_h_size=5,
                                          sg_conv_hidden=PROTEIN["sg_conv_hidden"], sg_hidden_size=50, sg_latent_size=50, s_hidden_size=5,
                                          s_latent_size=5, g_hidden_size=5, g_latent_size=5, **kw))


def test_protein_golden_fixture(built):
    """tests/golden/protein_n6.npz (the reference's protein flag block, main.py:218-236: SpatialGraphConvolution_3D joint encoder,
    layers.py:200-277, D = 3, node_h_size = 5) against the CUDA path: losses, outputs, gradient sums, Adam trajectory."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "protein_n6.npz"))
    N, B, S = int(z["N"]), int(z["B"]), int(z["S"])
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", **PROTEIN)
    eng = _protein_engine(built, N, B, S)
    assert [n for n, _, _ in eng.table] == [n for n, _, _ in O.param_table(cfg)]
    eng.set_params(P)
    res = eng.grads(inp, noise, fetch=("generated_adj_prob", "generated_spatial", "generated_node_feat", "z_sg"))
    np.testing.assert_allclose(res["overall_loss"], z["overall_loss"], rtol=1e-4)
    for k in ("generated_adj_prob", "generated_spatial", "generated_node_feat", "z_sg"):
        assert _relmax(res[k].cpu().numpy(), z[k]) < 1e-4, k
    gg = eng.get_grads()
    for name_, _, _ in eng.table:
        if "grad/" + name_ in z:
            assert _relmax(gg[name_].numpy(), z["grad/" + name_]) < 1e-3, name_
        s = z["gradsum/" + name_]
        assert abs(gg[name_].double().sum().item() - s[0]) < 1e-3 * max(s[1], 1e-12), name_
    costs = [eng.train_step(inp, noise)["overall_loss"][0] for _ in range(len(z["adam_costs"]))]
    np.testing.assert_allclose(costs, z["adam_costs"], rtol=2e-4)
    eng.close()


@pytest.mark.parametrize("N,B,S,density", [(9, 3, 2, 0.45), (14, 2, 3, 0.25)])
def test_three_hop_sgc_dense_adjacency_vs_oracle(built, N, B, S, density):
    """SpatialGraphConvolution_3D on contact-map-like adjacencies (dense, symmetric 0/1 -- not forests) and on real-valued ones:
    every output and gradient against the oracle's factored form (== the literal N^4 form to 1e-12, tests/test_oracle.py)."""
    cfg, P, inp, noise = _setup(N, B, S, "disentangled", **PROTEIN)
    g = torch.Generator().manual_seed(4)
    A = (torch.rand((B * S, N, N), generator=g, dtype=torch.float64) < density).double()
    A = torch.triu(A, 1); A = A + A.transpose(1, 2)
    if density < 0.3:
        A = A * torch.rand((B * S, N, N), generator=g, dtype=torch.float64)        # real-valued, asymmetric weights
    inp = dict(inp); inp["adj"] = A
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, "factored")
    eng = _protein_engine(built, N, B, S, chunk_graphs=2)
    eng.set_params(P)
    res = eng.grads(inp, noise, fetch=("generated_adj_prob", "z_mean_sg", "z_std_sg"))
    np.testing.assert_allclose(res["overall_loss"], [x.item() for x in L["overall_loss"]], rtol=1e-4)
    for k in ("z_mean_sg", "z_std_sg"):
        assert _relmax(res[k].cpu().numpy(), enc[k].detach().numpy()) < 1e-4, k
    assert _relmax(res["generated_adj_prob"].cpu().numpy(), dec["generated_adj_prob"].detach().numpy()) < 1e-4
    gg = eng.get_grads()
    worst = max((_relmax(gg[k].numpy(), v.numpy()), k) for k, v in grads.items())
    assert worst[0] < 1e-3, worst
    eng.close()


def test_protein_flag_block_trains(built):
    """flags.apply_dataset('protein') (main.py:218-236) -> SGCNModelVAE / OptimizerVAE / Session.run: constructs and trains."""
    flags = import_module("snd-vae_b200.flags"); model_m = import_module("snd-vae_b200.model"); opt_m = import_module("snd-vae_b200.optimizer")
    sess_m = import_module("snd-vae_b200.session"); prep = import_module("snd-vae_b200.preprocessing"); data = import_module("snd-vae_b200.data")
    F = flags.FLAGS; F.reset(); F.apply_dataset("protein"); F.type = "train"; F.batch_size = 4; F.sampling_num = 2
    N = 16
    d = data.synthetic_graphs(N, F.batch_size, F.sampling_num, spatial_dim=3, seed=12)
    d["adj"] = np.repeat(d["adj_truth"], F.sampling_num, axis=0)          # the protein loader feeds the contact map itself
    ph = sess_m.make_placeholders(F.batch_size, F.sampling_num, N, F.num_feature, F.spatial_dim)
    m = model_m.SGCNModelVAE(ph, F.num_feature, N)
    o = opt_m.OptimizerVAE(preds_edge=m.generated_adj_prob, preds_node=m.generated_node_feat, preds_spatial=m.generated_spatial,
                           labels_edge=ph["adj_truth"], labels_node=ph["feature_truth"], labels_spatial=ph["spatial_truth"],
                           labels_rel=ph["rel_truth"], global_iter=ph["global_iter"], model=m, num_nodes=N, pos_weight=1.0, norm=1.0, beta=1)
    assert m.engine.cfg.sg_hops == 3 and "encoder/g_sg1_conv/Matrix0" in m.vars
    fd = prep.construct_feed_dict_train(d["features"], d["spatial"], d["adj"], d["rel"], d["adj_truth"], d["feature_truth"], d["spatial_truth"], d["rel_truth"], ph)
    costs = []
    with sess_m.Session() as sess:
        for _ in range(4):
            costs.append(sess.run([o.opt_op, o.cost], feed_dict=fd)[1])
    assert np.isfinite(costs).all() and costs[-1] < costs[0]
    F.reset()
