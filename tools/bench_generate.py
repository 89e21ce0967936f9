"""Generation throughput (BASELINE config 5): decoder-only sampling from the prior, as main.py:428-469 does
(get_random_z draws + sample()/decoder, model.py:163-169,227-229).  Prints one JSON line: generated graphs per second.

  python tools/bench_generate.py [--nodes 256] [--batch 4096] [--steps 5] [--warmup 3]
"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sndvae_b200 as sv
from importlib import import_module
params = import_module("snd-vae_b200.params")

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=256); ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--sampling", type=int, default=10); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--tc", type=int, default=2)
a = ap.parse_args()
torch.cuda.set_device(0)
cfg = sv.make_config(a.nodes, a.batch, "disentangled", sampling_num=a.sampling, use_tensor_cores=a.tc)
eng = sv.Engine(cfg)
eng.set_params({k: torch.from_numpy(v) for k, v in params.init_params(eng.table, seed=7).items()})
g = torch.Generator(device="cuda").manual_seed(11)
B, S = a.batch, a.sampling
z_s = torch.randn(B, cfg.s_latent_size, device="cuda", generator=g)
z_sg = torch.randn(B * S, cfg.sg_latent_size, device="cuda", generator=g)
z_g = torch.randn(B, cfg.g_latent_size, device="cuda", generator=g)
fetch = ("generated_adj", "generated_spatial", "generated_node_feat")
for _ in range(a.warmup):
    eng.generate(z_s, z_sg, z_g, fetch=fetch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = eng.generate(z_s, z_sg, z_g, fetch=fetch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
dens = float(out["generated_adj"].float().mean().item())
print(json.dumps({"metric": "generate_graphs_per_sec", "value": B / (ms * 1e-3), "unit": "graphs/s", "ms_per_batch": ms,
                  "config": {"workload": f"decoder-only generation from prior samples, N={a.nodes}, {B} graphs per call, model.py", "tc": a.tc},
                  "edge_density_of_samples": dens}))
