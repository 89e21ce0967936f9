"""Throughput of the standalone InnerProductDecoder operator (layers.py:400-410): z z^T per graph on the
tensor cores.  The operator is bound by the fp32 logits it writes (4 N^2 bytes per graph)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1024); ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dim", type=int, default=40); ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    sv = import_module("snd-vae_b200")
    eng = sv.Engine(sv.make_config(8, 2, "disentangled", sampling_num=2))
    lib, h = eng.lib, eng._h
    z = torch.randn((a.batch, a.nodes, a.dim), device="cuda")
    out = torch.empty((a.batch, a.nodes, a.nodes), device="cuda")
    for _ in range(10):
        eng._check(lib.sndvae_inner_product_decode(h, z.data_ptr(), a.batch, a.nodes, a.dim, out.data_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()                                   # the handle runs on the stream that was current at its creation
    for _ in range(a.steps):
        eng._check(lib.sndvae_inner_product_decode(h, z.data_ptr(), a.batch, a.nodes, a.dim, out.data_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    byts = 4.0 * a.batch * a.nodes * a.nodes + 4.0 * a.batch * a.nodes * a.dim
    print(json.dumps({"op": "inner_product_decode", "nodes": a.nodes, "batch": a.batch, "dim": a.dim, "ms": ms,
                      "graphs_per_s": a.batch / ms * 1e3, "GBps_algorithmic": byts / ms / 1e6}))


if __name__ == "__main__":
    main()
