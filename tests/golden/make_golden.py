"""Generate the committed golden fixtures from the CPU oracle (fp64).

The reference is TensorFlow 1.x and cannot run here (no TensorFlow wheel, no
network), so these are outputs of the oracle restatement, NOT of TensorFlow:
parity unpinned at the TF boundary.  Seeds: params seed 7 + N(0,.05)
perturbation seed 1, inputs seed 5, noise seed 9.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sndvae_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


PROTEIN = dict(spatial_dim=3, node_h_size=5, sg_conv_hidden=((10, 10, 10, 10), (20, 20, 20, 20)), sg_hidden_size=50, sg_latent_size=50,
               s_hidden_size=5, s_latent_size=5, g_hidden_size=5, g_latent_size=5)     # main.py:218-236 (3-hop SGC layers)


def make(name, model, N, B, S, adam_steps=(1, 3), full=True, cfg_kw=None):
    cfg = O.Config(num_nodes=N, model_type=model, sampling_num=S, **(cfg_kw or {}))
    P = O.init_params(cfg, 7, torch.float64)
    g = torch.Generator().manual_seed(1)
    for k in P:
        P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg, B, 5, torch.float64)
    noise = O.synthetic_noise(cfg, B, 9, torch.float64)
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, "factored")
    out = {"N": N, "B": B, "S": cfg.S, "overall_loss": np.array([x.item() for x in L["overall_loss"]])}
    for k, v in {**enc, **z, **dec}.items():
        a = v.detach().numpy()
        out[k] = a.astype(np.float32) if (v.is_floating_point() and k != "generated_adj_prob") else a
    for k, v in grads.items():
        out["gradsum/" + k] = np.array([v.sum().item(), v.abs().sum().item()])
        if full:
            out["grad/" + k] = v.numpy().astype(np.float32)
    # fp32 TF-Adam trajectory
    P32 = {k: v.to(torch.float32).clone() for k, v in P.items()}
    adam = O.TFAdam(P32, cfg.learning_rate)
    i32, n32 = O.cast(inp, torch.float32), O.cast(noise, torch.float32)
    costs = []
    for st in range(1, max(adam_steps) + 1):
        _, _, _, Ls, g32 = O.loss_and_grads(P32, i32, n32, cfg, "factored")
        costs.append(Ls["cost"].item())
        adam.step(P32, g32)
    out["adam_costs"] = np.array(costs)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written; losses", out["overall_loss"])


if __name__ == "__main__":
    make("dis_n8", "disentangled", 8, 4, 3)
    make("base_n8", "base", 8, 4, 1)
    make("dis_n25", "disentangled", 25, 3, 10, full=False)
    # the reference's protein configuration (SpatialGraphConvolution_3D): oracle-only until the CUDA engine builds the branch
    make("protein_n6", "disentangled", 6, 3, 2, cfg_kw=PROTEIN)
