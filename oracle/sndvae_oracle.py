"""CPU oracle for the SND-VAE train / generate step  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch-CPU tensors, fp64 or fp32) of the
reference algorithm in /root/reference (xguo7/SND-VAE, TensorFlow 1.x).  It is
*not* part of the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it, and there only
as the checker (or as the timed CPU baseline).  The product path
(`snd-vae_b200/`) never imports this module and fails loudly without its CUDA
extension.

PARITY UNPINNED at the TensorFlow boundary: TensorFlow is not installed in the
build image and not in /opt/wheelhouse, the reference ships no tests, golden
vectors, logs or checkpoints.  The pin is therefore
internal: (1) a *literal* restatement that materialises what TF materialises,
(2) an independent *factored* restatement, (1)==(2) to 1e-12 in fp64 (and a third form, mode "fft": the width-N
correlations through torch.fft, == (2) to 1e-12, which is what scales to N = 1024 on a CPU),
(3) autograd == central finite differences, (4) seeded golden vectors under
tests/golden/ made by tests/golden/make_golden.py.

Reference lines followed (file:line into /root/reference):
  lrelu                         layers.py:112-113
  GraphConvolution              layers.py:115-125
  SpatialGraphConvolution       layers.py:143-198
  e2e                           layers.py:431-450
  linear                        layers.py:566-576
  encoder / get_z / decoder     model.py:98-222      (3-latent "disentangled")
  base model                    model_joint.py:72-182 (single latent z_sg)
  ELBO                          optimizer.py:126-164,192-204
  Adam                          optimizer.py:125,197 (tf.train.AdamOptimizer, TF1)
  feeds / shapes                main.py:253-264, preprocessing.py:32-42
Semantics restated from TF knowledge (not visible in the repo): Keras
BatchNormalization called without `training=` in TF1 graph mode runs the
inference branch with moving_mean=0, moving_var=1, eps=1e-3 forever (SURVEY
finding 3); tf.layers.conv1d = channels-last cross-correlation + bias, glorot
uniform kernel; conv2d SAME pad_before=(k-1)//2; softmax_cross_entropy over the
last axis; tf.argmax -> int64, first max on ties; TF1 ApplyAdam formula.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as Fn

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon


# --------------------------------------------------------------------------
# configuration  (main.py:42-103 flags, synthetic2 block main.py:181-209)
# --------------------------------------------------------------------------
@dataclass
class Config:
    num_nodes: int = 25
    num_feature: int = 1            # main.py:83
    spatial_dim: int = 2            # main.py:84
    sampling_num: int = 10          # main.py:100
    node_h_size: int = 20           # main.py:209
    s_channel: Tuple[int, ...] = (10, 10, 20)
    s_hidden_size: int = 100
    s_latent_size: int = 100
    g_conv_hidden: Tuple[int, ...] = (10, 20)
    g_hidden_size: int = 100
    g_latent_size: int = 100
    sg_conv_hidden: Tuple[Tuple[int, ...], ...] = ((20, 20, 20), (50, 50, 50))   # 4-tuples select the 3-hop layer (protein / mnist)
    sg_hidden_size: int = 100
    sg_latent_size: int = 100
    s_d_channel: Tuple[int, ...] = (50, 20, 10)
    n_d_channel: Tuple[int, ...] = (50, 20)      # graph_deconv_layers=2 of [50,20,10]
    e_d_hidden: Tuple[int, ...] = (50, 20)       # graph_deconv_layers=2 of [50,20,10]
    kernel_size: int = 5
    model_type: str = "disentangled"             # or "base" (model_joint.py)
    learning_rate: float = 0.0008                # main.py:211
    beta: float = 1.0                            # main.py:515
    num_edge_feature: int = 2                    # model_joint.py:171 (undefined flag)
    # loss variants of optimizer.py:166-190 (FLAGS.model_type selects the branch; the network is model.py's in all of them)
    loss_variant: str = "elbo"                   # "elbo" | "disentangled_C" | "NED-VAE-IP" | "beta-TCVAE"
    gamma: float = 100.0                         # main.py:97
    C_max: float = 100.0                         # main.py:95
    C_stop_iter: float = 1e2                     # main.py:96
    C_step: float = 20.0                         # main.py:98
    global_iter: int = 0                         # placeholder fed per epoch (main.py:329)
    dip_lambda_od: float = 10.0                  # DIP(enc_mean, 10, 100), optimizer.py:183
    dip_lambda_d: float = 100.0

    @property
    def N(self): return self.num_nodes
    @property
    def S(self): return self.sampling_num if self.model_type != "base" else 1


# --------------------------------------------------------------------------
# parameter table  (TF variable creation order)
# --------------------------------------------------------------------------
def param_table(cfg: Config) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(tf_name, shape, init) in creation order.  init in
    {trunc02, normal02, glorot, zeros, ones}."""
    N, F, D, H = cfg.N, cfg.num_feature, cfg.spatial_dim, cfg.node_h_size
    k = cfg.kernel_size
    t: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(name, c):
        t.append((f"{name}/gamma", (c,), "ones"))
        t.append((f"{name}/beta", (c,), "zeros"))

    def lin(name, i, o):
        t.append((f"{name}/Matrix", (i, o), "normal02"))
        t.append((f"{name}/bias", (o,), "zeros"))

    def conv(name, ci, co):
        t.append((f"{name}/kernel", (k, ci, co), "glorot"))
        t.append((f"{name}/bias", (co,), "zeros"))

    def sgc(name, C, hs):
        R = 1
        if len(hs) == 4:        # SpatialGraphConvolution_3D (layers.py:210-225; protein / mnist branch, model.py:139-140)
            t.append((f"{name}/Matrix0", (4 * C + 3 * R + 2, hs[0]), "normal02"))
            t.append((f"{name}/bias0", (hs[0],), "zeros"))
            t.append((f"{name}/Matrix1", (3 * C + 2 * R + hs[0] + 1, hs[1]), "normal02"))
            t.append((f"{name}/bias1", (hs[1],), "zeros"))
            t.append((f"{name}/Matrix2", (2 * C + R + hs[1], hs[2]), "normal02"))
            t.append((f"{name}/bias2", (hs[2],), "zeros"))
            t.append((f"{name}/Matrix3", (C + hs[2], hs[3]), "normal02"))
            t.append((f"{name}/bias3", (hs[3],), "zeros"))
            return
        t.append((f"{name}/Matrix1", (3 * C + 2 * R + 1, hs[0]), "normal02"))
        t.append((f"{name}/bias1", (hs[0],), "zeros"))
        t.append((f"{name}/Matrix2", (2 * C + hs[0] + R, hs[1]), "normal02"))
        t.append((f"{name}/bias2", (hs[1],), "zeros"))
        t.append((f"{name}/Matrix3", (C + hs[1], hs[2]), "normal02"))
        t.append((f"{name}/bias3", (hs[2],), "zeros"))

    dis = cfg.model_type != "base"
    if dis:
        # graph encoder  model.py:104-115
        c = F
        for i, h in enumerate(cfg.g_conv_hidden):
            t.append((f"encoder/g_g{i}_conv/w", (c, h), "trunc02"))
            bn(f"encoder/g_bn_g{i}", h)
            c = h + F
        bn("encoder/encoder_g", c)
        lin("encoder/g_g1_lin", N * c, cfg.g_hidden_size)
        lin("encoder/g_g2_lin", cfg.g_hidden_size, cfg.g_latent_size)
        lin("encoder/g_g3_lin", cfg.g_hidden_size, cfg.g_latent_size)
        # spatial encoder  model.py:119-129
        c = D
        for i, h in enumerate(cfg.s_channel):
            conv(f"encoder/g_s{i+1}_conv", c, h)
            bn(f"encoder/g_bn_s{i}", h)
            c = h
        bn("encoder/encoder_s", c)
        lin("encoder/g_s1_lin", N * c, cfg.s_hidden_size)
        lin("encoder/g_s2_lin", cfg.s_hidden_size, cfg.s_latent_size)
        lin("encoder/g_s3_lin", cfg.s_hidden_size, cfg.s_latent_size)
    # joint encoder  model.py:134-151 / model_joint.py:76-85
    c = F
    for i, hs in enumerate(cfg.sg_conv_hidden):
        sgc(f"encoder/g_sg{i}_conv", c, hs)
        bn(f"encoder/g_bn_sg{i}", hs[-1])
        c = hs[-1]
    if dis:
        bn("encoder/encoder_sg", c)
    lin("encoder/g_sg1_lin", N * c, cfg.sg_hidden_size)
    lin("encoder/g_sg2_lin", cfg.sg_hidden_size, cfg.sg_latent_size)
    lin("encoder/g_sg3_lin", cfg.sg_hidden_size, cfg.sg_latent_size)
    # decoder  model.py:177-219 / model_joint.py:97-171
    lin("decoder/d_sg_lin1", cfg.sg_latent_size, N * H)
    if dis:
        lin("decoder/d_s_lin1", cfg.s_latent_size, N * H)
        lin("decoder/d_g_lin1", cfg.g_latent_size, N * H)
    cin0 = 2 * H if dis else H
    if not dis:
        # model_joint.py builds the spatial decoder first (113-121)
        c = cin0
        for i, h in enumerate(cfg.s_d_channel):
            conv(f"decoder/s{i+1}_deconv", c, h)
            bn(f"decoder/d_bn_s{i}", h)
            c = h
        lin("decoder/d_s_lin2", c, D)
    c = cin0
    for i, h in enumerate(cfg.n_d_channel):
        conv(f"decoder/n{i}_deconv", c, h)
        bn(f"decoder/d_bn_n{i}", h)
        c = h
    if dis:
        bn("decoder/decoder_node", c)
    lin("decoder/d_n_lin2", c, F)
    c = 2 * cin0
    for i, h in enumerate(cfg.e_d_hidden):
        bn(f"decoder/d_bn_e{i}", c)
        t.append((f"decoder/e{i}_deconv/w1", (1, N, c, h), "trunc02"))
        t.append((f"decoder/e{i}_deconv/biases1", (h,), "zeros"))
        c = h
    if dis:
        bn("decoder/decoder_adj", c)
    lin("decoder/d_e_lin2", c, 2)
    if dis:
        c = cin0
        for i, h in enumerate(cfg.s_d_channel):
            conv(f"decoder/s{i+1}_deconv", c, h)
            bn(f"decoder/d_bn_s{i}", h)
            c = h
        lin("decoder/d_s_lin2", c, D)
    return t


def init_params(cfg: Config, seed: int = 7, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference initialisers (layers.py:118-119,158-169,434-437,569-572; Keras
    glorot_uniform for tf.layers.conv1d), drawn from torch.Generator(seed)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, init in param_table(cfg):
        if init == "zeros":
            v = torch.zeros(shape, dtype=torch.float64)
        elif init == "ones":
            v = torch.ones(shape, dtype=torch.float64)
        elif init == "normal02":
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.02
        elif init == "trunc02":
            v = torch.randn(shape, generator=g, dtype=torch.float64)
            for _ in range(50):  # resample beyond 2 sigma (tf.truncated_normal)
                bad = v.abs() > 2.0
                if not bad.any():
                    break
                v = torch.where(bad, torch.randn(shape, generator=g, dtype=torch.float64), v)
            v = v * 0.02
        elif init == "glorot":
            kk, ci, co = shape
            lim = math.sqrt(6.0 / (kk * ci + kk * co))
            v = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        else:
            raise ValueError(init)
        out[name] = v.to(dtype)
    return out


# --------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------
def lrelu(x, leak=0.2):
    """layers.py:112-113"""
    return torch.maximum(x, leak * x)


def bn(x, P, name):
    """Keras BatchNormalization, inference branch, moving_mean=0, moving_var=1
    (model.py:41-71; SURVEY finding 3): y = x * gamma*rsqrt(1+eps) + beta."""
    gam, bet = P[name + "/gamma"], P[name + "/beta"]
    inv = gam * (1.0 / math.sqrt(1.0 + BN_EPS))
    return x * inv + bet


def linear(x, P, name):
    """layers.py:566-576"""
    return x @ P[name + "/Matrix"] + P[name + "/bias"]


def conv1d_same(x, P, name):
    """tf.layers.conv1d(..., padding='SAME'), stride 1 (model.py:122,191,216):
    channels-last cross-correlation over the node axis, plus bias."""
    K, b = P[name + "/kernel"], P[name + "/bias"]          # [k, Cin, Cout]
    k = K.shape[0]
    pb = (k - 1) // 2
    xp = Fn.pad(x, (0, 0, pb, k - 1 - pb))                 # pad node axis
    n = x.shape[1]
    out = b.expand(x.shape[0], n, K.shape[2]).clone()
    for t in range(k):
        out = out + xp[:, t:t + n, :] @ K[t]
    return out


def graph_convolution(adj, x, P, name):
    """layers.py:115-125: lrelu(adj @ (x @ w)); raw adjacency, no bias."""
    return lrelu(adj @ (x @ P[name + "/w"]))


def sgc_literal(adj, x, rel, P, name):
    """layers.py:143-198, materialising the [B,N,N,N,3C+3] tensor as TF does."""
    Bn, N, C = x.shape
    r = rel.reshape(Bn, N, N, 1)
    rel_ij = r.reshape(Bn, N, N, 1, 1).expand(Bn, N, N, N, 1)
    rel_jk = r.reshape(Bn, 1, N, N, 1).expand(Bn, N, N, N, 1)
    dis_ik = r.reshape(Bn, N, 1, N, 1).expand(Bn, N, N, N, 1)
    adj3 = adj.reshape(Bn, N, N, 1) * adj.reshape(Bn, 1, N, N)       # A_ij A_jk
    fx = x.reshape(Bn, N, 1, 1, C).expand(Bn, N, N, N, C)
    fy = x.reshape(Bn, 1, N, 1, C).expand(Bn, N, N, N, C)
    fz = x.reshape(Bn, 1, 1, N, C).expand(Bn, N, N, N, C)
    m3 = torch.cat([fx, fy, fz, rel_ij, rel_jk, dis_ik], dim=-1)
    m3 = lrelu(m3) @ P[name + "/Matrix1"] + P[name + "/bias1"]        # [B,N,N,N,h0]
    m3s = (m3 * adj3.unsqueeze(-1)).sum(dim=3)                        # sum over k
    fx2 = x.reshape(Bn, N, 1, C).expand(Bn, N, N, C)
    fy2 = x.reshape(Bn, 1, N, C).expand(Bn, N, N, C)
    m2 = torch.cat([fx2, fy2, r, m3s], dim=3)
    m2 = lrelu(m2) @ P[name + "/Matrix2"] + P[name + "/bias2"]        # [B,N,N,h1]
    m2s = (m2 * adj.unsqueeze(-1)).sum(dim=2)                         # sum over j
    m1 = torch.cat([x, m2s], dim=2)
    return lrelu(m1) @ P[name + "/Matrix3"] + P[name + "/bias3"]


def sgc_factored(adj, x, rel, P, name):
    """Exact factorisation of layers.py:171-196: no
    nonlinearity sits between Matrix1 and the adjacency-weighted k-sum."""
    Bn, N, C = x.shape
    M1, b1 = P[name + "/Matrix1"], P[name + "/bias1"]
    M2, b2 = P[name + "/Matrix2"], P[name + "/bias2"]
    M3, b3 = P[name + "/Matrix3"], P[name + "/bias3"]
    h0 = M1.shape[1]
    px = lrelu(x)
    pr = lrelu(rel.reshape(Bn, N, N))
    Pm, Qm, Rm = px @ M1[0:C], px @ M1[C:2 * C], px @ M1[2 * C:3 * C]
    w4, w5, w6 = M1[3 * C], M1[3 * C + 1], M1[3 * C + 2]
    deg = adj.sum(dim=2)                                   # deg_j = sum_k A_jk
    s = (adj * pr).sum(dim=2)                              # s_j = sum_k A_jk phi(r_jk)
    G = pr @ adj.transpose(1, 2)                           # G_ij = sum_k phi(r_ik) A_jk
    AR = adj @ Rm
    inner = (deg.reshape(Bn, 1, N, 1) * (Pm.reshape(Bn, N, 1, h0) + Qm.reshape(Bn, 1, N, h0)
                                         + pr.unsqueeze(-1) * w4 + b1)
             + AR.reshape(Bn, 1, N, h0) + s.reshape(Bn, 1, N, 1) * w5 + G.unsqueeze(-1) * w6)
    m3s = adj.unsqueeze(-1) * inner                        # [B,N,N,h0]
    M2a, M2b, M2c, M2d = M2[0:C], M2[C:2 * C], M2[2 * C], M2[2 * C + 1:]
    U, V = px @ M2a, px @ M2b
    T = (adj.unsqueeze(-1) * lrelu(m3s)).sum(dim=2)        # [B,N,h0]
    m2s = deg.unsqueeze(-1) * (U + b2) + adj @ V + s.unsqueeze(-1) * M2c + T @ M2d
    return lrelu(torch.cat([x, m2s], dim=2)) @ M3 + b3


def sgc3d_literal(adj, x, rel, P, name):
    """SpatialGraphConvolution_3D (layers.py:200-277; the `protein` / `mnist` branch, model.py:139-140), materialising the
    [B,N,N,N,N,4C+5] tensor as TF does.  Parameters name/Matrix0..3, name/bias0..3 (layers.py:210-225).  SURVEY 8(f) N4:
    restatement only -- no CUDA path yet; `tests/test_oracle.py::test_sgc3d_factored_equals_literal` pins the two forms
    against each other."""
    Bn, N, C = x.shape
    r = rel.reshape(Bn, N, N, 1)
    e = lambda t, shape: t.reshape(shape).expand(Bn, N, N, N, N, t.shape[-1])
    m4 = torch.cat([e(x, (Bn, N, 1, 1, 1, C)), e(x, (Bn, 1, N, 1, 1, C)), e(x, (Bn, 1, 1, N, 1, C)), e(x, (Bn, 1, 1, 1, N, C)),
                    e(r, (Bn, N, N, 1, 1, 1)), e(r, (Bn, 1, N, N, 1, 1)), e(r, (Bn, 1, 1, N, N, 1)),
                    e(r, (Bn, N, 1, N, 1, 1)), e(r, (Bn, N, 1, 1, N, 1))], dim=-1)         # i, j, k, p; r_ij, r_jk, r_kp, r_ik, r_ip
    m4 = lrelu(m4) @ P[name + "/Matrix0"] + P[name + "/bias0"]                            # [B,N,N,N,N,h0]
    adj4 = (adj.reshape(Bn, N, N, 1, 1) * adj.reshape(Bn, 1, N, N, 1) * adj.reshape(Bn, 1, 1, N, N))   # A_ij A_jk A_kp
    m4s = (m4 * adj4.unsqueeze(-1)).sum(dim=4)                                            # sum over p
    f = lambda t, shape: t.reshape(shape).expand(Bn, N, N, N, t.shape[-1])
    m3 = torch.cat([f(x, (Bn, N, 1, 1, C)), f(x, (Bn, 1, N, 1, C)), f(x, (Bn, 1, 1, N, C)),
                    f(r, (Bn, N, N, 1, 1)), f(r, (Bn, 1, N, N, 1)), f(r, (Bn, N, 1, N, 1)), m4s], dim=-1)
    m3 = lrelu(m3) @ P[name + "/Matrix1"] + P[name + "/bias1"]                            # [B,N,N,N,h1]
    adj3 = adj.reshape(Bn, N, N, 1) * adj.reshape(Bn, 1, N, N)                            # A_ij A_jk
    m3s = (m3 * adj3.unsqueeze(-1)).sum(dim=3)                                            # sum over k
    m2 = torch.cat([x.reshape(Bn, N, 1, C).expand(Bn, N, N, C), x.reshape(Bn, 1, N, C).expand(Bn, N, N, C), r, m3s], dim=3)
    m2 = lrelu(m2) @ P[name + "/Matrix2"] + P[name + "/bias2"]                            # [B,N,N,h2]
    m2s = (m2 * adj.unsqueeze(-1)).sum(dim=2)                                             # sum over j
    return lrelu(torch.cat([x, m2s], dim=2)) @ P[name + "/Matrix3"] + P[name + "/bias3"]


def sgc3d_factored(adj, x, rel, P, name):
    """Exact factorisation of layers.py:228-273: the p-sum is linear (no nonlinearity between Matrix0 and it), so the N^4
    tensor collapses to per-(i,j,k) closed forms; the k-sum keeps one N^3 h0 pointwise term (lrelu of m4_sum), the j-sum one
    N^2 h1 term.  With deg_k = sum_p A_kp, s_k = sum_p A_kp phi(r_kp), G_ik = sum_p A_kp phi(r_ip):
      m4s_ijk = A_ij A_jk [deg_k (P0_i + Q0_j + R0_k + phi(r_ij) a1 + phi(r_jk) a2 + phi(r_ik) a4 + b0) + (A S0)_k + s_k a3 + G_ik a5]
      m3s_ij  = A_ij [deg_j (P1_i + Q1_j + phi(r_ij) c1 + b1) + (A R1)_j + s_j c2 + G_ij c3 + (sum_k A_jk phi(m4s_ijk)) M1e]
      m2s_i   = deg_i (U_i + b2) + (A V)_i + s_i M2c + (sum_j A_ij phi(m3s_ij)) M2d"""
    Bn, N, C = x.shape
    M0, b0 = P[name + "/Matrix0"], P[name + "/bias0"]
    M1, b1 = P[name + "/Matrix1"], P[name + "/bias1"]
    M2, b2 = P[name + "/Matrix2"], P[name + "/bias2"]
    M3, b3 = P[name + "/Matrix3"], P[name + "/bias3"]
    h0, h1 = M0.shape[1], M1.shape[1]
    px = lrelu(x)
    pr = lrelu(rel.reshape(Bn, N, N))
    deg = adj.sum(dim=2)
    s = (adj * pr).sum(dim=2)
    G = pr @ adj.transpose(1, 2)
    # level 4 -> m4s [B,N,N,N,h0]
    P0, Q0, R0, S0 = px @ M0[0:C], px @ M0[C:2 * C], px @ M0[2 * C:3 * C], px @ M0[3 * C:4 * C]
    a1, a2, a3, a4, a5 = (M0[4 * C + t] for t in range(5))                                # r_ij, r_jk, r_kp, r_ik, r_ip
    AS = adj @ S0
    inner4 = (deg.reshape(Bn, 1, 1, N, 1) * (P0.reshape(Bn, N, 1, 1, h0) + Q0.reshape(Bn, 1, N, 1, h0) + R0.reshape(Bn, 1, 1, N, h0)
                                             + pr.reshape(Bn, N, N, 1, 1) * a1 + pr.reshape(Bn, 1, N, N, 1) * a2
                                             + pr.reshape(Bn, N, 1, N, 1) * a4 + b0)
              + AS.reshape(Bn, 1, 1, N, h0) + s.reshape(Bn, 1, 1, N, 1) * a3 + G.reshape(Bn, N, 1, N, 1) * a5)
    adj3 = adj.reshape(Bn, N, N, 1) * adj.reshape(Bn, 1, N, N)
    m4s = adj3.unsqueeze(-1) * inner4
    # level 3 -> m3s [B,N,N,h1]
    P1, Q1, R1 = px @ M1[0:C], px @ M1[C:2 * C], px @ M1[2 * C:3 * C]
    c1, c2, c3 = M1[3 * C], M1[3 * C + 1], M1[3 * C + 2]                                  # r_ij, r_jk, r_ik
    M1e = M1[3 * C + 3:]
    T3 = (adj.reshape(Bn, 1, N, N, 1) * lrelu(m4s)).sum(dim=3)                            # sum_k A_jk phi(m4s_ijk)
    AR = adj @ R1
    inner3 = (deg.reshape(Bn, 1, N, 1) * (P1.reshape(Bn, N, 1, h1) + Q1.reshape(Bn, 1, N, h1) + pr.unsqueeze(-1) * c1 + b1)
              + AR.reshape(Bn, 1, N, h1) + s.reshape(Bn, 1, N, 1) * c2 + G.unsqueeze(-1) * c3 + T3 @ M1e)
    m3s = adj.unsqueeze(-1) * inner3
    # level 2 -> m2s [B,N,h2]
    M2a, M2b, M2c, M2d = M2[0:C], M2[C:2 * C], M2[2 * C], M2[2 * C + 1:]
    U, V = px @ M2a, px @ M2b
    T2 = (adj.unsqueeze(-1) * lrelu(m3s)).sum(dim=2)
    m2s = deg.unsqueeze(-1) * (U + b2) + adj @ V + s.unsqueeze(-1) * M2c + T2 @ M2d
    return lrelu(torch.cat([x, m2s], dim=2)) @ M3 + b3


def e2e_literal(x, w1, bias):
    """layers.py:431-450 via conv2d on a materialised [B,N,N,C] tensor.
    w1: [1,N,C,O].  TF SAME padding: pad_before=(k-1)//2."""
    Bn, N, _, C = x.shape
    k = w1.shape[1]
    p = (k - 1) // 2
    xin = x.permute(0, 3, 1, 2)                            # NCHW, H=i, W=j
    wk = w1[0].permute(2, 1, 0)                            # [O,C,k]
    xw = Fn.pad(xin, (p, k - 1 - p, 0, 0))
    c1 = Fn.conv2d(xw, wk.unsqueeze(2))                    # slide along W (j)
    xh = Fn.pad(xin, (0, 0, p, k - 1 - p))
    c2 = Fn.conv2d(xh, wk.unsqueeze(3))                    # slide along H (i)
    out = c1 + c2 + 2.0 * bias.reshape(1, -1, 1, 1)        # bias added twice (438,446)
    return out.permute(0, 2, 3, 1)


def toeplitz_matrix(w1):
    """T[(j',c),(j,o)] = w1[0, j'-j+p, c, o]."""
    _, N, C, O = w1.shape
    p = (N - 1) // 2
    jp = torch.arange(N).reshape(N, 1)
    j = torch.arange(N).reshape(1, N)
    t = jp - j + p
    valid = (t >= 0) & (t < N)
    Tm = w1[0][t.clamp(0, N - 1)] * valid.reshape(N, N, 1, 1).to(w1.dtype)   # [j',j,C,O]
    return Tm.permute(0, 2, 1, 3).reshape(N * C, N * O)


def e2e_toeplitz(x, w1, bias):
    """e2e as two GEMMs against the block-Toeplitz matrix (Appendix C.3)."""
    Bn, N, _, C = x.shape
    O = w1.shape[3]
    Tm = toeplitz_matrix(w1)
    o1 = (x.reshape(Bn * N, N * C) @ Tm).reshape(Bn, N, N, O)
    o2 = (x.transpose(1, 2).reshape(Bn * N, N * C) @ Tm).reshape(Bn, N, N, O).transpose(1, 2)
    return o1 + o2 + 2.0 * bias


def corr_same_fft(x, w):
    """out[..., s, o] = sum_{s', c} x[..., s', c] w[s' - s + p, c, o]  (p = (N-1)//2, zero outside 0 <= s' - s + p < N): the
    width-N SAME cross-correlation of layers.py:436,443 along the second-to-last axis, evaluated with torch.fft (a third,
    independent form of the same sum, O(N log N) per line: what makes N = 1024 checkable on a CPU).
    x: [..., N, C], w: [N, C, O].  Circular convolution of the zero-padded line with g[m] = w[p - m], L >= 2N."""
    N, C, O = w.shape
    p = (N - 1) // 2
    L = 2 * N
    g = torch.zeros((L, C, O), dtype=w.dtype)
    m = torch.arange(p - N + 1, p + 1)                     # g[m] = w[p - m] for p-N+1 <= m <= p, indices mod L
    g[m % L] = w[p - m]
    Xf = torch.fft.rfft(x, n=L, dim=-2)                    # [..., L/2+1, C]
    Gf = torch.fft.rfft(g, dim=0)                          # [L/2+1, C, O]
    Of = torch.einsum("...fc,fco->...fo", Xf, Gf)
    return torch.fft.irfft(Of, n=L, dim=-2)[..., :N, :]


def e2e_fft(x, w1, bias):
    """e2e (layers.py:431-450) with both directions evaluated by corr_same_fft."""
    o1 = corr_same_fft(x, w1[0])                                               # slide along j
    o2 = corr_same_fft(x.transpose(1, 2), w1[0]).transpose(1, 2)               # slide along i
    return o1 + o2 + 2.0 * bias


def e2e_l0_factored(a, c, w1, bias, fft=False):
    """e2e on the never-materialised pair tensor [a_i || c_j] (Appendix C.2).
    a, c: [B,N,Ch] (already BN+relu'd halves)."""
    Bn, N, Ch = a.shape
    p = (N - 1) // 2
    wa, wc = w1[0][:, :Ch, :], w1[0][:, Ch:, :]            # [N(t),Ch,O]
    pos = torch.arange(N).reshape(N, 1)
    t = torch.arange(N).reshape(1, N)
    valid = ((pos + t - p >= 0) & (pos + t - p < N)).to(w1.dtype)     # [pos,t]
    WSa = torch.einsum("pt,tco->pco", valid, wa)           # [N(j),Ch,O]
    WSc = torch.einsum("pt,tco->pco", valid, wc)           # [N(i),Ch,O]
    O = w1.shape[3]
    if fft:
        Rc, Sa = corr_same_fft(c, wc), corr_same_fft(a, wa)
    else:
        Tc = toeplitz_matrix(wc.unsqueeze(0))              # [(j',c),(j,o)]
        Ta = toeplitz_matrix(wa.unsqueeze(0))
        Rc = (c.reshape(Bn, N * Ch) @ Tc).reshape(Bn, N, O)    # depends on j
        Sa = (a.reshape(Bn, N * Ch) @ Ta).reshape(Bn, N, O)    # depends on i
    t1 = torch.einsum("bic,jco->bijo", a, WSa)
    t3 = torch.einsum("bjc,ico->bijo", c, WSc)
    return t1 + t3 + Rc.unsqueeze(1) + Sa.unsqueeze(2) + 2.0 * bias


# --------------------------------------------------------------------------
# model forward
# --------------------------------------------------------------------------
def encoder(P, inp, cfg: Config, mode="factored"):
    """model.py:98-151 (disentangled) / model_joint.py:72-85 (base)."""
    N, S = cfg.N, cfg.S
    out = {}
    sgc = sgc_literal if mode == "literal" else sgc_factored
    dis = cfg.model_type != "base"
    if dis:
        X, Pt, At = inp["feature_truth"], inp["spatial_truth"], inp["adj_truth"]
        B = X.shape[0]
        g = X
        for i in range(len(cfg.g_conv_hidden)):
            g = bn(graph_convolution(At, g, P, f"encoder/g_g{i}_conv"), P, f"encoder/g_bn_g{i}")
            g = torch.cat([g, X], dim=-1)
        g = bn(g, P, "encoder/encoder_g")
        hg = linear(g.reshape(B, -1), P, "encoder/g_g1_lin")
        out["z_mean_g"] = linear(hg, P, "encoder/g_g2_lin")
        out["z_std_g"] = linear(hg, P, "encoder/g_g3_lin")
        h = Pt
        for i in range(len(cfg.s_channel)):
            h = torch.relu(bn(conv1d_same(h, P, f"encoder/g_s{i+1}_conv"), P, f"encoder/g_bn_s{i}"))
        h = bn(h, P, "encoder/encoder_s")
        hs = linear(h.reshape(B, -1), P, "encoder/g_s1_lin")
        out["z_mean_s"] = linear(hs, P, "encoder/g_s2_lin")
        out["z_std_s"] = linear(hs, P, "encoder/g_s3_lin")
    x, A, R = inp["features"], inp["adj"], inp["rel"]
    BS = x.shape[0]
    sgc3 = sgc3d_literal if mode == "literal" else sgc3d_factored
    for i in range(len(cfg.sg_conv_hidden)):
        # a 4-tuple of hidden sizes selects the 3-hop layer (FLAGS.dataset in {protein, mnist}: main.py:225,241, model.py:139-140)
        layer = sgc3 if len(cfg.sg_conv_hidden[i]) == 4 else sgc
        x = lrelu(bn(layer(A, x, R, P, f"encoder/g_sg{i}_conv"), P, f"encoder/g_bn_sg{i}"))
    if dis:
        x = bn(x, P, "encoder/encoder_sg")
    hsg = linear(x.reshape(BS, -1), P, "encoder/g_sg1_lin")
    out["z_mean_sg"] = linear(hsg, P, "encoder/g_sg2_lin")
    out["z_std_sg"] = linear(hsg, P, "encoder/g_sg3_lin")
    return out


def get_z(enc, noise, cfg: Config):
    """model.py:153-161 / model_joint.py:87-91: z = mu + eps * exp(logsigma)."""
    z = {"z_sg": enc["z_mean_sg"] + noise["eps_sg"] * torch.exp(enc["z_std_sg"])}
    if cfg.model_type != "base":
        z["z_s"] = enc["z_mean_s"] + noise["eps_s"] * torch.exp(enc["z_std_s"])
        z["z_g"] = enc["z_mean_g"] + noise["eps_g"] * torch.exp(enc["z_std_g"])
    return z


def mask_and_threshold(lg):
    """model.py:185,205-208: diagonal mask then argmax(softmax) -> int64."""
    Bn, N = lg.shape[0], lg.shape[1]
    m = (1.0 - torch.eye(N, dtype=lg.dtype)).reshape(1, N, N)
    l1 = m * lg[..., 1]
    l0 = m * lg[..., 0] + (1 - m)
    prob = torch.stack([l0, l1], dim=-1)
    return torch.argmax(torch.softmax(prob, dim=-1), dim=-1), prob


def threshold_rule_fp32(lg):
    """argmax(softmax([l0,l1])) evaluated with correctly rounded fp32 operations (numpy fp32
    add / divide are correctly rounded; exp is taken in fp64 and rounded once), first index on
    ties -- the arithmetic tf.nn.softmax + tf.argmax specify (model.py:208).  lg: [...,2] fp32."""
    a = np.asarray(lg, dtype=np.float32)
    p0, p1 = a[..., 0], a[..., 1]
    mx = np.maximum(p0, p1)
    e0 = np.exp((p0 - mx).astype(np.float64)).astype(np.float32)
    e1 = np.exp((p1 - mx).astype(np.float64)).astype(np.float32)
    s = e0 + e1
    return (e1 / s > e0 / s).astype(np.int64)


def decoder(P, z, cfg: Config, mode="factored"):
    """model.py:172-222 (disentangled) / model_joint.py:94-182 (base)."""
    N, H, S = cfg.N, cfg.node_h_size, cfg.S
    dis = cfg.model_type != "base"
    out = {}
    if dis:
        zsg = linear(z["z_sg"], P, "decoder/d_sg_lin1")
        B = zsg.shape[0] // S
        n_sg = zsg.reshape(B, S, N, H).mean(dim=1)          # model.py:177,180
        n_s = linear(z["z_s"], P, "decoder/d_s_lin1").reshape(B, N, H)
        n_g = linear(z["z_g"], P, "decoder/d_g_lin1").reshape(B, N, H)
        v = torch.cat([n_sg, n_g], dim=-1)
        q = v
        for i in range(len(cfg.n_d_channel)):               # no activation (model.py:192)
            q = bn(conv1d_same(q, P, f"decoder/n{i}_deconv"), P, f"decoder/d_bn_n{i}")
        q = bn(q, P, "decoder/decoder_node")
        out["generated_node_feat"] = torch.sigmoid(linear(q, P, "decoder/d_n_lin2"))
        sp = torch.cat([n_sg, n_s], dim=-1)
        for i in range(len(cfg.s_d_channel)):
            sp = bn(conv1d_same(sp, P, f"decoder/s{i+1}_deconv"), P, f"decoder/d_bn_s{i}")
        out["generated_spatial"] = torch.sigmoid(linear(sp, P, "decoder/d_s_lin2"))
    else:
        B = z["z_sg"].shape[0]
        v = linear(z["z_sg"], P, "decoder/d_sg_lin1").reshape(B, N, H)
        sp = v
        for i in range(len(cfg.s_d_channel)):               # lrelu after BN (model_joint.py:115-116)
            sp = lrelu(bn(conv1d_same(sp, P, f"decoder/s{i+1}_deconv"), P, f"decoder/d_bn_s{i}"))
        out["generated_spatial"] = torch.sigmoid(linear(sp, P, "decoder/d_s_lin2"))
        q = v
        for i in range(len(cfg.n_d_channel)):
            q = lrelu(bn(conv1d_same(q, P, f"decoder/n{i}_deconv"), P, f"decoder/d_bn_n{i}"))
        out["generated_node_feat"] = torch.sigmoid(linear(q, P, "decoder/d_n_lin2"))
    # edge decoder  model.py:196-208 / model_joint.py:164-179
    Ch = v.shape[-1]
    w0, b0 = P["decoder/e0_deconv/w1"], P["decoder/e0_deconv/biases1"]
    w1, b1 = P["decoder/e1_deconv/w1"], P["decoder/e1_deconv/biases1"]
    if mode == "literal":
        E = torch.cat([v.reshape(B, N, 1, Ch).expand(B, N, N, Ch),
                       v.reshape(B, 1, N, Ch).expand(B, N, N, Ch)], dim=-1)
        E = e2e_literal(torch.relu(bn(E, P, "decoder/d_bn_e0")), w0, b0)
        E = e2e_literal(torch.relu(bn(E, P, "decoder/d_bn_e1")), w1, b1)
    else:
        gam, bet = P["decoder/d_bn_e0/gamma"], P["decoder/d_bn_e0/beta"]
        inv = gam * (1.0 / math.sqrt(1.0 + BN_EPS))
        a = torch.relu(v * inv[:Ch] + bet[:Ch])
        c = torch.relu(v * inv[Ch:] + bet[Ch:])
        E = e2e_l0_factored(a, c, w0, b0, fft=(mode == "fft"))
        Y = torch.relu(bn(E, P, "decoder/d_bn_e1"))
        E = e2e_fft(Y, w1, b1) if mode == "fft" else e2e_toeplitz(Y, w1, b1)
    if dis:
        E = bn(E, P, "decoder/decoder_adj")
    lg = linear(torch.relu(E).reshape(B * N * N, -1), P, "decoder/d_e_lin2").reshape(B, N, N, 2)
    out["generated_adj"], out["generated_adj_prob"] = mask_and_threshold(lg)
    return out


def losses(P, inp, enc, dec, cfg: Config, z=None):
    """optimizer.py:126-164,192-204.  Returns dict + overall_loss list order."""
    At = inp["adj_truth"] if cfg.model_type != "base" else inp["adj_truth"]
    lab = torch.stack([1 - At, At], dim=-1)
    lg = dec["generated_adj_prob"]
    ce = -(lab * torch.log_softmax(lg, dim=-1)).sum(dim=-1)
    adj_cost = ce.mean()
    node_cost = ((inp["feature_truth"] - dec["generated_node_feat"]) ** 2).mean()
    spatial_cost = ((inp["spatial_truth"] - dec["generated_spatial"]) ** 2).mean()

    def kl(mu, ls):
        return -0.5 * (1 + 2 * ls - mu ** 2 - torch.exp(ls) ** 2).mean()

    kl_sg = kl(enc["z_mean_sg"], enc["z_std_sg"])
    L = {"adj_cost": adj_cost, "node_cost": node_cost, "spatial_cost": spatial_cost, "kl_sg": kl_sg}
    if cfg.model_type != "base":
        L["kl_s"] = kl(enc["z_mean_s"], enc["z_std_s"])
        L["kl_g"] = kl(enc["z_mean_g"], enc["z_std_g"])
        mse = adj_cost + node_cost + spatial_cost
        if cfg.loss_variant == "disentangled_C":
            # optimizer.py:166-174: capacity-controlled joint KL; C follows global_iter in steps of C_step
            C = min(max(cfg.C_max * cfg.C_step / cfg.C_stop_iter * (cfg.global_iter // int(cfg.C_step)), 0.0), cfg.C_max)
            L["C"] = C
            L["cost"] = mse + cfg.gamma * torch.relu(kl_sg - C) + L["kl_s"] + L["kl_g"]
        elif cfg.loss_variant == "NED-VAE-IP":
            # optimizer.py:176-183: ELBO (KL weight 1) + beta * DIP-VAE-I regulariser on the three posterior means
            L["dip"] = sum(dip_regulariser(enc[k], cfg.dip_lambda_od, cfg.dip_lambda_d) for k in ("z_mean_s", "z_mean_g", "z_mean_sg"))
            L["cost"] = mse + (kl_sg + L["kl_s"] + L["kl_g"]) + cfg.beta * L["dip"]
        elif cfg.loss_variant == "beta-TCVAE":
            # optimizer.py:185-190: ELBO + 10 * minibatch total-correlation estimate of each latent group
            L["tc"] = sum(total_correlation(z[k], enc["z_mean" + k[1:]], enc["z_std" + k[1:]]) for k in ("z_s", "z_g", "z_sg"))
            L["cost"] = mse + cfg.beta * (kl_sg + L["kl_s"] + L["kl_g"]) + 10.0 * L["tc"]
        else:
            L["cost"] = mse + cfg.beta * (kl_sg + L["kl_s"] + L["kl_g"])
        L["overall_loss"] = [L["cost"], spatial_cost, adj_cost, node_cost, L["kl_g"], L["kl_s"], kl_sg]
    else:
        L["cost"] = adj_cost + node_cost + spatial_cost + cfg.beta * kl_sg
        L["overall_loss"] = [L["cost"], spatial_cost, adj_cost, node_cost, kl_sg]
    return L


def dip_regulariser(mu, lambda_od, lambda_d):
    """DIP() of optimizer.py:7-21: covariance of the posterior means over the batch, pushed towards the identity."""
    m = mu.mean(dim=0)
    cov = (mu.unsqueeze(1) * mu.unsqueeze(2)).mean(dim=0) - m.unsqueeze(0) * m.unsqueeze(1)
    d = torch.diagonal(cov)
    off = cov - torch.diag(d)
    return lambda_d * ((d - 1) ** 2).sum() + lambda_od * (off ** 2).sum()


def total_correlation(z, z_mean, z_logstd):
    """total_correlation() of optimizer.py:30-63 with gaussian_log_density (optimizer.py:23-28): the minibatch estimate
    mean_j [ log sum_i prod_l q(z_jl | x_i)  -  sum_l log sum_i q(z_jl | x_i) ], constants dropped as the reference does.
    z, z_mean, z_logstd: [rows, L]."""
    logvar = torch.log(torch.exp(z_logstd) * torch.exp(z_logstd))                     # optimizer.py:43
    d = z.unsqueeze(1) - z_mean.unsqueeze(0)                                          # [j, i, l]
    lq = -0.5 * (d * d * torch.exp(-logvar).unsqueeze(0) + logvar.unsqueeze(0) + math.log(2.0 * math.pi))
    log_qz_product = torch.logsumexp(lq, dim=1).sum(dim=1)
    log_qz = torch.logsumexp(lq.sum(dim=2), dim=1)
    return (log_qz - log_qz_product).mean()


def forward(P, inp, noise, cfg: Config, mode="factored"):
    enc = encoder(P, inp, cfg, mode)
    z = get_z(enc, noise, cfg)
    dec = decoder(P, z, cfg, mode)
    L = losses(P, inp, enc, dec, cfg, z)
    return enc, z, dec, L


def loss_and_grads(P, inp, noise, cfg: Config, mode="factored"):
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    enc, z, dec, L = forward(Pg, inp, noise, cfg, mode)
    L["cost"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return enc, z, dec, L, grads


# --------------------------------------------------------------------------
# TF1 Adam  (optimizer.py:125,197)
# --------------------------------------------------------------------------
class TFAdam:
    """tf.train.AdamOptimizer(lr) defaults beta1=.9 beta2=.999 eps=1e-8, the
    ApplyAdam kernel formula with fp32 running beta powers."""

    def __init__(self, P, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}
        self.b1p = np.float32(beta1)
        self.b2p = np.float32(beta2)

    def step(self, P, grads):
        f32 = np.float32
        if next(iter(P.values())).dtype == torch.float32:
            alpha = float(f32(self.lr) * np.sqrt(f32(1) - self.b2p) / (f32(1) - self.b1p))
        else:
            alpha = self.lr * math.sqrt(1 - float(self.b2p)) / (1 - float(self.b1p))
        for k in P:
            g = grads[k]
            self.m[k] += (g - self.m[k]) * (1 - self.b1)
            self.v[k] += (g * g - self.v[k]) * (1 - self.b2)
            P[k] -= (self.m[k] * alpha) / (torch.sqrt(self.v[k]) + self.eps)
        self.b1p = f32(self.b1p * f32(self.b1))
        self.b2p = f32(self.b2p * f32(self.b2))


# --------------------------------------------------------------------------
# synthetic inputs  (input_data.py:18-38,54-96 semantics)
# --------------------------------------------------------------------------
def synthetic_inputs(cfg: Config, B: int, seed: int = 1234, dtype=torch.float32, mesh=False):
    """Random-geometric spatial graphs + S random spanning forests per graph.
    Row b*S+s of the sampled tensors belongs to graph b (graph-major)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree
    rng = np.random.default_rng(seed)
    N, F, D, S = cfg.N, cfg.num_feature, cfg.spatial_dim, cfg.S
    P = rng.random((B, N, D), dtype=np.float32)
    X = rng.random((B, N, F), dtype=np.float32)
    diff = P[:, :, None, :] - P[:, None, :, :]
    rel = np.sqrt((diff ** 2).sum(-1)).astype(np.float32)
    r = math.sqrt(6.0 / (math.pi * N))
    A = (rel < r).astype(np.float32)
    idx = np.arange(N)
    A[:, idx, idx] = 0.0
    As = np.zeros((B, S, N, N), dtype=np.float32)
    for b in range(B):
        x, y = np.where(A[b])
        for s in range(S):
            if len(x) == 0:
                continue
            cg = csr_matrix((rng.random(len(x)) + 1, (x, y)), shape=(N, N))
            tr, tc = minimum_spanning_tree(cg).nonzero()
            As[b, s, tr, tc] = 1.0
            As[b, s, tc, tr] = 1.0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    inp = {
        "adj_truth": t(A), "feature_truth": t(X), "spatial_truth": t(P),
        "rel_truth": t(rel[..., None]),
        "adj": t(As.reshape(B * S, N, N)),
        "features": t(np.repeat(X, S, axis=0)),
        "spatial": t(np.repeat(P, S, axis=0)),
        "rel": t(np.repeat(rel, S, axis=0)[..., None]),
    }
    return inp


def synthetic_noise(cfg: Config, B: int, seed: int = 4321, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    S = cfg.S
    n = {"eps_s": torch.randn(B, cfg.s_latent_size, generator=g, dtype=torch.float64),
         "eps_sg": torch.randn(B * S, cfg.sg_latent_size, generator=g, dtype=torch.float64),
         "eps_g": torch.randn(B, cfg.g_latent_size, generator=g, dtype=torch.float64)}
    return {k: v.to(dtype) for k, v in n.items()}


def cast(d, dtype):
    return {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}
