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


def _check_against_oracle(built, cfg, P, inp, noise, mode, tc, chunk, model):
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
    margin = (ref["generated_adj_prob"][..., 1] - ref["generated_adj_prob"][..., 0]).abs() > 1e-5
    assert torch.equal(res["generated_adj"].cpu()[margin], ref["generated_adj"][margin])
    gg = eng.get_grads()
    worst = max((_relmax(gg[k].numpy(), v.numpy()), k) for k, v in grads.items())
    assert worst[0] < 1e-3, worst
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
    _check_against_oracle(built, cfg, P, inp, noise, "fft", 2, 1, "disentangled")


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
    assert sum(v.nbytes for v in compact.values()) * 6 < sum(v.nbytes for v in dense.values())
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
