"""CPU tests that pin the oracle: literal == factored restatement,
autograd == finite differences, invariants, TF-Adam formula, golden vectors.
The oracle is a restatement of the TensorFlow reference, not TensorFlow itself
(parity unpinned at that boundary)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sndvae_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _setup(N, B, S, model, seed=1, perturb=0.05, dtype=torch.float64):
    cfg = O.Config(num_nodes=N, model_type=model, sampling_num=S)
    P = O.init_params(cfg, 7, dtype)
    g = torch.Generator().manual_seed(seed)
    for k in P:
        P[k] = P[k] + perturb * torch.randn(P[k].shape, generator=g, dtype=dtype)
    return cfg, P, O.synthetic_inputs(cfg, B, 5, dtype), O.synthetic_noise(cfg, B, 9, dtype)


@pytest.mark.parametrize("model", ["disentangled", "base"])
@pytest.mark.parametrize("N", [6, 7])
def test_literal_equals_factored(model, N):
    cfg, P, inp, noise = _setup(N, 2, 3, model)
    e1, z1, d1, L1, g1 = O.loss_and_grads(P, inp, noise, cfg, "literal")
    e2, z2, d2, L2, g2 = O.loss_and_grads(P, inp, noise, cfg, "factored")
    for k in e1:
        assert (e1[k] - e2[k]).abs().max() < 1e-12
    assert (d1["generated_adj_prob"] - d2["generated_adj_prob"]).abs().max() < 1e-12
    assert torch.equal(d1["generated_adj"], d2["generated_adj"])
    for k in g1:
        assert (g1[k] - g2[k]).abs().max() < 1e-12, k


def test_sgc_factorisation_general_adjacency():
    """Appendix C.1 holds for arbitrary real A and arbitrary-sign rel, not only 0/1 trees."""
    torch.manual_seed(0)
    Bn, N, C = 3, 6, 4
    P = {"s/Matrix1": torch.randn(3 * C + 3, 5, dtype=torch.float64), "s/bias1": torch.randn(5, dtype=torch.float64),
         "s/Matrix2": torch.randn(2 * C + 5 + 1, 7, dtype=torch.float64), "s/bias2": torch.randn(7, dtype=torch.float64),
         "s/Matrix3": torch.randn(C + 7, 3, dtype=torch.float64), "s/bias3": torch.randn(3, dtype=torch.float64)}
    A = torch.randn(Bn, N, N, dtype=torch.float64)
    x = torch.randn(Bn, N, C, dtype=torch.float64)
    rel = torch.randn(Bn, N, N, 1, dtype=torch.float64)
    assert (O.sgc_literal(A, x, rel, P, "s") - O.sgc_factored(A, x, rel, P, "s")).abs().max() < 1e-11


def test_e2e_forms_agree():
    torch.manual_seed(1)
    for N in (5, 8):
        x = torch.randn(2, N, N, 6, dtype=torch.float64)
        w = torch.randn(1, N, 6, 4, dtype=torch.float64)
        b = torch.randn(4, dtype=torch.float64)
        assert (O.e2e_literal(x, w, b) - O.e2e_toeplitz(x, w, b)).abs().max() < 1e-12
        # density of the block-Toeplitz matrix = V(N) / N^2
        p = (N - 1) // 2; q = N - 1 - p
        V = N * N - p * (p + 1) // 2 - q * (q + 1) // 2
        T = O.toeplitz_matrix(torch.ones(1, N, 1, 1, dtype=torch.float64))
        assert int(T.sum().item()) == V


def test_autograd_matches_finite_differences():
    cfg, P, inp, noise = _setup(6, 2, 2, "disentangled")
    _, _, _, L, g = O.loss_and_grads(P, inp, noise, cfg, "factored")
    rng = np.random.default_rng(0)
    eps = 1e-6
    for name in ["encoder/g_sg1_conv/Matrix1", "encoder/g_g1_conv/w", "decoder/e1_deconv/w1", "decoder/e0_deconv/w1",
                 "decoder/d_bn_e1/gamma", "encoder/g_s2_conv/kernel", "decoder/decoder_adj/beta", "decoder/d_sg_lin1/Matrix"]:
        flat = P[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=3, replace=False):
            old = flat[idx].item()
            flat[idx] = old + eps
            lp = O.forward(P, inp, noise, cfg)[3]["cost"].item()
            flat[idx] = old - eps
            lm = O.forward(P, inp, noise, cfg)[3]["cost"].item()
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            an = g[name].reshape(-1)[idx].item()
            assert abs(fd - an) < 1e-7 + 1e-5 * abs(an), (name, idx, fd, an)


def test_shard_sum_equals_full_batch_gradient():
    """Every graph's forward is independent of the rest of the batch (SURVEY finding 3), so the
    data-parallel sum of shard gradients scaled by 1/world equals the full-batch gradient."""
    cfg, P, inp, noise = _setup(6, 4, 2, "disentangled")
    _, _, _, L, g = O.loss_and_grads(P, inp, noise, cfg)
    S = cfg.S
    acc = {k: torch.zeros_like(v) for k, v in P.items()}
    for r in range(2):
        sl = slice(2 * r, 2 * r + 2); sls = slice(2 * r * S, (2 * r + 2) * S)
        si = {k: (v[sls] if k in ("adj", "features", "spatial", "rel") else v[sl]) for k, v in inp.items()}
        sn = {"eps_s": noise["eps_s"][sl], "eps_g": noise["eps_g"][sl], "eps_sg": noise["eps_sg"][sls]}
        gs = O.loss_and_grads(P, si, sn, cfg)[4]
        for k in acc:
            acc[k] += gs[k] / 2
    for k in g:
        assert (acc[k] - g[k]).abs().max() < 1e-13, k


def test_batch_permutation_and_diagonal():
    cfg, P, inp, noise = _setup(7, 3, 2, "disentangled")
    enc, z, dec, L = O.forward(P, inp, noise, cfg)
    perm = torch.tensor([2, 0, 1]); S = cfg.S
    perms = (perm[:, None] * S + torch.arange(S)[None]).reshape(-1)
    pi = {k: (v[perms] if k in ("adj", "features", "spatial", "rel") else v[perm]) for k, v in inp.items()}
    pn = {"eps_s": noise["eps_s"][perm], "eps_g": noise["eps_g"][perm], "eps_sg": noise["eps_sg"][perms]}
    enc2, z2, dec2, L2 = O.forward(P, pi, pn, cfg)
    assert (dec2["generated_adj_prob"] - dec["generated_adj_prob"][perm]).abs().max() < 1e-12
    assert abs(L2["cost"].item() - L["cost"].item()) < 1e-12
    N = cfg.N
    d = torch.arange(N)
    assert (dec["generated_adj"][:, d, d] == 0).all()                      # diag logits are (1, 0)
    assert (dec["generated_adj_prob"][:, d, d, 0] == 1).all() and (dec["generated_adj_prob"][:, d, d, 1] == 0).all()


def test_step0_anchors():
    """with the reference initialisers adj_cost ~ (ln2 (N-1) + ln(1+1/e))/N etc."""
    cfg = O.Config(num_nodes=25)
    P = O.init_params(cfg, 7, torch.float64)
    inp = O.synthetic_inputs(cfg, 4, 5, torch.float64)
    noise = O.synthetic_noise(cfg, 4, 9, torch.float64)
    L = O.forward(P, inp, noise, cfg)[3]
    N = 25
    assert abs(L["adj_cost"].item() - (math.log(2) * (N - 1) + math.log(1 + math.exp(-1))) / N) < 5e-3
    assert abs(L["node_cost"].item() - 1 / 12) < 2e-2 and abs(L["spatial_cost"].item() - 1 / 12) < 2e-2
    assert L["kl_sg"].item() < 1e-2


def test_tf_adam_formula():
    """TF1 ApplyAdam against a hand-rolled scalar recurrence; differs from torch Adam."""
    P = {"w": torch.tensor([0.5, -1.0], dtype=torch.float64)}
    opt = O.TFAdam(P, 0.01)
    m = v = np.zeros(2); th = np.array([0.5, -1.0]); b1p, b2p = 0.9, 0.999
    for t in range(1, 4):
        g = np.array([0.1 * t, -1e-7])
        opt.step(P, {"w": torch.tensor(g)})
        alpha = 0.01 * math.sqrt(1 - b2p) / (1 - b1p)
        m = m + (g - m) * 0.1; v = v + (g * g - v) * 0.001
        th = th - m * alpha / (np.sqrt(v) + 1e-8)
        b1p *= 0.9; b2p *= 0.999
    assert np.allclose(P["w"].numpy(), th, rtol=1e-6, atol=0)


def test_param_counts():
    n = lambda c: sum(int(np.prod(s)) for _, s, _ in O.param_table(c))
    assert n(O.Config(num_nodes=25)) == 611587
    assert n(O.Config(num_nodes=256)) == 5268547
    assert n(O.Config(num_nodes=256, model_type="base")) == 2619755


@pytest.mark.parametrize("name", ["dis_n8", "base_n8", "dis_n25"])
def test_golden_vectors(name):
    """The committed fixtures (tests/golden/make_golden.py) are reproduced by the oracle."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    model = "base" if name.startswith("base") else "disentangled"
    N, B, S = int(z["N"]), int(z["B"]), int(z["S"])
    cfg, P, inp, noise = _setup(N, B, S, model)
    enc, zz, dec, L, g = O.loss_and_grads(P, inp, noise, cfg)
    assert np.allclose(np.array([x.item() for x in L["overall_loss"]]), z["overall_loss"], rtol=1e-10)
    assert np.allclose(dec["generated_adj_prob"].detach().numpy(), z["generated_adj_prob"], rtol=0, atol=1e-10)
    assert np.array_equal(dec["generated_adj"].numpy(), z["generated_adj"])
    for k in g:
        d = z["gradsum/" + k]
        assert np.allclose([g[k].sum().item(), g[k].abs().sum().item()], d, rtol=1e-7, atol=1e-12), k


def test_loss_variants_of_the_oracle():
    """optimizer.py:166-190: the capacity schedule C(global_iter), the relu gate of 'disentangled_C', DIP() against a
    direct numpy evaluation of its definition (covariance of the posterior means over the batch), and total_correlation()
    against plain loops."""
    cfg0 = O.Config(num_nodes=6, sampling_num=2)
    P = O.init_params(cfg0, 7, torch.float64)
    g = torch.Generator().manual_seed(1)
    for k in P:
        P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg0, 3, 5, torch.float64); nz = O.synthetic_noise(cfg0, 3, 9, torch.float64)
    base = O.forward(P, inp, nz, cfg0)[3]
    mse = base["adj_cost"] + base["node_cost"] + base["spatial_cost"]
    for it, C in ((0, 0.0), (19, 0.0), (20, 20.0), (45, 40.0), (100, 100.0), (500, 100.0)):   # 100 * 20 / 100 * (it // 20), clipped
        cfg = O.Config(num_nodes=6, sampling_num=2, loss_variant="disentangled_C", global_iter=it)
        L = O.forward(P, inp, nz, cfg)[3]
        assert L["C"] == C
        want = mse + 100.0 * max(float(base["kl_sg"]) - C, 0.0) + base["kl_s"] + base["kl_g"]
        assert abs(float(L["cost"]) - float(want)) < 1e-12
    cfg = O.Config(num_nodes=6, sampling_num=2, loss_variant="NED-VAE-IP")
    enc, _, _, L = O.forward(P, inp, nz, cfg)
    tot = 0.0
    for k in ("z_mean_s", "z_mean_g", "z_mean_sg"):
        mu = enc[k].numpy()
        cov = np.cov(mu.T, bias=True)
        d = np.diag(cov)
        tot += 100.0 * ((d - 1) ** 2).sum() + 10.0 * ((cov - np.diag(d)) ** 2).sum()
    assert abs(float(L["dip"]) - tot) < 1e-8 * max(tot, 1.0)
    assert abs(float(L["cost"]) - float(mse + base["kl_sg"] + base["kl_s"] + base["kl_g"] + tot)) < 1e-8 * max(tot, 1.0)
    # 'beta-TCVAE' (optimizer.py:23-63,185-190): the minibatch total correlation against plain loops over its definition
    cfg = O.Config(num_nodes=6, sampling_num=2, loss_variant="beta-TCVAE")
    enc, z, _, L = O.forward(P, inp, nz, cfg)
    tot = 0.0
    for k in ("s", "g", "sg"):
        zz, mu, lv = z["z_" + k].numpy(), enc["z_mean_" + k].numpy(), 2.0 * enc["z_std_" + k].numpy()
        R, Ld = zz.shape
        acc = 0.0
        for j in range(R):
            lq = np.array([[-0.5 * ((zz[j, l] - mu[i, l]) ** 2 * np.exp(-lv[i, l]) + lv[i, l] + np.log(2 * np.pi)) for l in range(Ld)] for i in range(R)])
            acc += np.log(np.exp(lq.sum(1)).sum()) - np.log(np.exp(lq).sum(0)).sum()
        tot += acc / R
    assert abs(float(L["tc"]) - tot) < 1e-9 * max(abs(tot), 1.0)
    assert abs(float(L["cost"]) - float(mse + base["kl_sg"] + base["kl_s"] + base["kl_g"] + 10.0 * tot)) < 1e-8 * max(abs(tot), 1.0)


def test_synthetic_data_oracle_invariants():
    """oracle/synth.py (the checker of the device-side generator): coordinates / features in [0,1), rel = pairwise
    distance, symmetric zero-diagonal adjacency, and every sample a spanning forest of its graph's adjacency (|E| = N - #components,
    acyclic, subset of the truth edges) -- the properties of input_data.py:18-38,62-67,77-82."""
    from scipy.sparse.csgraph import connected_components
    from oracle import synth
    N, B, S = 25, 3, 4
    d = synth.synth_inputs(N, B, S, 1, 2, seed=99)
    P, A, As, rel = d["spatial_truth"], d["adj_truth"], d["adj"], d["rel_truth"][..., 0]
    assert P.min() >= 0 and P.max() < 1 and d["feature_truth"].min() >= 0 and d["feature_truth"].max() < 1
    np.testing.assert_allclose(rel, np.sqrt(((P[:, :, None] - P[:, None]) ** 2).sum(-1)), rtol=1e-6, atol=1e-7)
    assert (A == A.transpose(0, 2, 1)).all() and (A[:, np.arange(N), np.arange(N)] == 0).all()
    assert 2.0 < A.sum() / (B * N) < 9.0                       # mean degree ~ 6 (boundary effects lower it)
    for b in range(B):
        nc, _ = connected_components(A[b])
        for s in range(S):
            T = As[b * S + s]
            assert (T == T.T).all() and (T <= A[b]).all() and T.sum() / 2 == N - nc
            assert connected_components(T)[0] == nc            # N - nc edges and nc components: a forest spanning every component
    assert not (As[0] == As[1]).all()                          # different samples of one graph differ
    for k in ("features", "spatial", "rel"):
        assert d[k].shape[0] == B * S
    assert (d["rel"][S:2 * S, ..., 0] == rel[1]).all() and (d["features"][S:2 * S] == d["feature_truth"][1]).all()
    again = synth.synth_inputs(N, B, S, 1, 2, seed=99)
    assert all((again[k] == d[k]).all() for k in d)


def test_sgc3d_factored_equals_literal():
    """the 3-hop SpatialGraphConvolution_3D (layers.py:200-277).  The literal restatement materialises the
    [B,N,N,N,N,4C+5] tensor; the factored one removes the N^4 term exactly (the p-sum is linear).  fp64, real-valued
    adjacency (the identity does not need 0/1 entries), forests and dense graphs."""
    torch.manual_seed(3)
    Bn, N, C = 2, 5, 3
    h = (4, 5, 6, 7)
    name = "sg3"
    P = {name + "/Matrix0": 0.3 * torch.randn(4 * C + 5, h[0], dtype=torch.float64), name + "/bias0": 0.1 * torch.randn(h[0], dtype=torch.float64),
         name + "/Matrix1": 0.3 * torch.randn(3 * C + 3 + h[0], h[1], dtype=torch.float64), name + "/bias1": 0.1 * torch.randn(h[1], dtype=torch.float64),
         name + "/Matrix2": 0.3 * torch.randn(2 * C + 1 + h[1], h[2], dtype=torch.float64), name + "/bias2": 0.1 * torch.randn(h[2], dtype=torch.float64),
         name + "/Matrix3": 0.3 * torch.randn(C + h[2], h[3], dtype=torch.float64), name + "/bias3": 0.1 * torch.randn(h[3], dtype=torch.float64)}
    x = torch.randn(Bn, N, C, dtype=torch.float64)
    rel = torch.randn(Bn, N, N, 1, dtype=torch.float64)           # signed: exercises both branches of the leaky relu
    for kind in ("dense_real", "binary"):
        adj = torch.randn(Bn, N, N, dtype=torch.float64) if kind == "dense_real" else (torch.rand(Bn, N, N) < 0.4).double()
        lit = O.sgc3d_literal(adj, x, rel, P, name)
        fac = O.sgc3d_factored(adj, x, rel, P, name)
        assert lit.shape == (Bn, N, h[3])
        assert (lit - fac).abs().max().item() < 1e-11 * max(1.0, lit.abs().max().item()), kind
    # with no edges only the x -> Matrix3 path is left
    out = O.sgc3d_factored(torch.zeros(Bn, N, N, dtype=torch.float64), x, rel, P, name)
    ref = O.lrelu(torch.cat([x, torch.zeros(Bn, N, h[2], dtype=torch.float64)], dim=2)) @ P[name + "/Matrix3"] + P[name + "/bias3"]
    assert torch.allclose(out, ref, atol=1e-14)


def test_protein_branch_literal_equals_factored():
    """The `protein` configuration of the reference (main.py:218-236: 3-hop SpatialGraphConvolution_3D layers, spatial_dim 3,
    node_h_size 5, small latents) through the whole model: losses and every gradient of the literal restatement (N^4 tensors)
    equal those of the factored one in fp64.  Checker side only -- the CUDA engine does not build this branch yet."""
    cfg = O.Config(num_nodes=5, spatial_dim=3, node_h_size=5, sampling_num=2, sg_conv_hidden=((4, 4, 4, 4), (6, 6, 6, 6)),
                   sg_hidden_size=8, sg_latent_size=8, s_hidden_size=5, s_latent_size=5, g_hidden_size=5, g_latent_size=5)
    names = [n for n, _, _ in O.param_table(cfg)]
    assert "encoder/g_sg0_conv/Matrix0" in names and "encoder/g_sg1_conv/bias3" in names
    shapes = {n: sh for n, sh, _ in O.param_table(cfg)}
    assert shapes["encoder/g_sg0_conv/Matrix0"] == (4 * 1 + 5, 4) and shapes["encoder/g_sg1_conv/Matrix1"] == (3 * 4 + 3 + 6, 6)
    P = O.init_params(cfg, 7, torch.float64)
    g = torch.Generator().manual_seed(2)
    for k in P:
        P[k] = P[k] + 0.1 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg, 2, 5, torch.float64)
    noise = O.synthetic_noise(cfg, 2, 9, torch.float64)
    _, _, _, L1, g1 = O.loss_and_grads(P, inp, noise, cfg, "literal")
    _, _, _, L2, g2 = O.loss_and_grads(P, inp, noise, cfg, "factored")
    for a, b in zip(L1["overall_loss"], L2["overall_loss"]):
        assert abs(a.item() - b.item()) < 1e-12 * max(1.0, abs(a.item()))
    for k in g1:
        assert (g1[k] - g2[k]).abs().max().item() < 1e-11 * max(1.0, g1[k].abs().max().item()), k
    assert g1["encoder/g_sg0_conv/Matrix0"].abs().max().item() > 0


def test_golden_protein_fixture():
    """tests/golden/protein_n6.npz (the reference's protein hyper-parameters, main.py:218-236, at N=6) is reproduced by the
    factored oracle and, independently, by the literal restatement that materialises the N^4 tensors."""
    z = np.load(os.path.join(GOLD, "protein_n6.npz"))
    N, B, S = int(z["N"]), int(z["B"]), int(z["S"])
    cfg = O.Config(num_nodes=N, sampling_num=S, spatial_dim=3, node_h_size=5, sg_conv_hidden=((10, 10, 10, 10), (20, 20, 20, 20)),
                   sg_hidden_size=50, sg_latent_size=50, s_hidden_size=5, s_latent_size=5, g_hidden_size=5, g_latent_size=5)
    P = O.init_params(cfg, 7, torch.float64)
    g = torch.Generator().manual_seed(1)
    for k in P:
        P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg, B, 5, torch.float64)
    noise = O.synthetic_noise(cfg, B, 9, torch.float64)
    for mode, tol in (("factored", 1e-10), ("literal", 1e-9)):
        enc, zz, dec, L, gr = O.loss_and_grads(P, inp, noise, cfg, mode)
        assert np.allclose(np.array([x.item() for x in L["overall_loss"]]), z["overall_loss"], rtol=tol), mode
        assert np.allclose(dec["generated_adj_prob"].detach().numpy(), z["generated_adj_prob"], rtol=0, atol=tol), mode
        for k in gr:
            d = z["gradsum/" + k]
            assert np.allclose([gr[k].sum().item(), gr[k].abs().sum().item()], d, rtol=1e-6, atol=1e-11), (mode, k)


@pytest.mark.parametrize("model,N,B,S", [("disentangled", 8, 3, 2), ("disentangled", 25, 2, 3), ("base", 7, 2, 1)])
def test_fft_form_equals_factored(model, N, B, S):
    """Third form of the width-N correlations (torch.fft; what the N = 1024 GPU parity test uses as its checker) against the
    block-Toeplitz form: losses, logits and every gradient to fp64 round-off."""
    cfg = O.Config(num_nodes=N, model_type=model, sampling_num=S)
    P = O.init_params(cfg, 7, torch.float64)
    g = torch.Generator().manual_seed(1)
    for k in P:
        P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg, B, 5, torch.float64)
    nz = O.synthetic_noise(cfg, B, 9, torch.float64)
    a = O.loss_and_grads(P, inp, nz, cfg, "factored")
    b = O.loss_and_grads(P, inp, nz, cfg, "fft")
    assert abs(float(a[3]["cost"]) - float(b[3]["cost"])) < 1e-12
    assert (a[2]["generated_adj_prob"] - b[2]["generated_adj_prob"]).abs().max().item() < 1e-12
    for k in a[4]:
        assert (a[4][k] - b[4][k]).abs().max().item() < 1e-12, k
