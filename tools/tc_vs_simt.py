"""Diagnostic: gradients of the tensor-core path vs the fp32 SIMT path at a given size."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sndvae_b200 as sv
from oracle import sndvae_oracle as O
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=256); ap.add_argument("--b", type=int, default=4)
ap.add_argument("--s", type=int, default=2); ap.add_argument("--chunk", type=int, default=3); ap.add_argument("--tc", type=int, default=1)
a = ap.parse_args()
cfg = O.Config(num_nodes=a.n, sampling_num=a.s)
P = O.init_params(cfg, 7, torch.float32)
g = torch.Generator().manual_seed(1)
for k in P: P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g)
inp = O.synthetic_inputs(cfg, a.b, 5, torch.float32); noise = O.synthetic_noise(cfg, a.b, 9, torch.float32)
out = {}
for name, tc, chunk in (("simt", 0, a.chunk), ("tc", a.tc, a.chunk)):
    eng = sv.Engine(sv.make_config(a.n, a.b, "disentangled", sampling_num=a.s, use_tensor_cores=tc, chunk_graphs=chunk))
    eng.set_params(P); r = eng.grads(inp, noise, fetch=("generated_adj_prob",))
    out[name] = (r["overall_loss"], eng.get_grads(), r["generated_adj_prob"].cpu())
    dbg = {k: eng.debug_read(k, 1 << 24) for k in ("Rc", "Sa", "E1", "O12", "dY12", "da", "dc", "dv")}
    out[name] += (dbg,)
    eng.close()
ref = out["simt"]
for name in ("tc",):
    o = out[name]
    print("==", name, "loss", o[0], "ref", ref[0])
    print("  logits relmax", float((o[2] - ref[2]).abs().max() / ref[2].abs().max()))
    for k in ("Rc", "Sa", "E1", "O12", "dY12", "da", "dc", "dv"):
        x, y = o[3][k].astype(np.float64), ref[3][k].astype(np.float64)
        print(f"  dbg {k:5s} relmax {np.abs(x - y).max() / max(np.abs(y).max(), 1e-30):.2e}  rel-rms {np.sqrt(((x-y)**2).mean()) / max(np.sqrt((y**2).mean()), 1e-30):.2e}  mean-signed-ratio {((x-y)*np.sign(y)).mean() / max(np.abs(y).mean(),1e-30):.2e}")
    worst = sorted(((float((o[1][k] - ref[1][k]).abs().max() / max(ref[1][k].abs().max(), 1e-30)), k) for k in ref[1]), reverse=True)[:8]
    for e, k in worst: print(f"  grad {k:36s} {e:.2e}")
