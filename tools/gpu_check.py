"""Per-tensor parity report of the CUDA path against the CPU oracle (diagnostic).

Prints one line per output / gradient instead of stopping at the first failure.
Usage: python tools/gpu_check.py [--n 8] [--b 4] [--s 3] [--model disentangled] [--tc 0]
"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sndvae_b200 as sv
from oracle import sndvae_oracle as O


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / scale, np.abs(b).max()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8); ap.add_argument("--b", type=int, default=4)
    ap.add_argument("--s", type=int, default=3); ap.add_argument("--model", default="disentangled")
    ap.add_argument("--tc", type=int, default=0); ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2); ap.add_argument("--perturb", type=float, default=0.05)
    a = ap.parse_args()
    cfg = O.Config(num_nodes=a.n, model_type=a.model, sampling_num=a.s)
    P = O.init_params(cfg, 7, torch.float64)
    g = torch.Generator().manual_seed(1)
    for k in P:      # perturb so that biases / BN parameters matter
        P[k] = P[k] + a.perturb * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    inp = O.synthetic_inputs(cfg, a.b, 5, torch.float64)
    noise = O.synthetic_noise(cfg, a.b, 9, torch.float64)
    t0 = time.time()
    enc, z, dec, L, grads = O.loss_and_grads(P, inp, noise, cfg, "factored")
    print(f"oracle fp64 done in {time.time()-t0:.1f}s; losses", [round(float(x), 6) for x in L["overall_loss"]])
    ecfg = sv.make_config(a.n, a.b, a.model, sampling_num=a.s, use_tensor_cores=a.tc, chunk_graphs=a.chunk)
    eng = sv.Engine(ecfg)
    names_o = [n for n, _, _ in O.param_table(cfg)]
    names_e = [n for n, _, _ in eng.table]
    assert names_o == names_e, (names_o, names_e)
    eng.set_params(P)
    res = eng.grads(inp, noise, fetch=sv._lib.OUTPUT_FIELDS)
    bad = 0
    print("losses gpu  ", [round(float(x), 6) for x in res["overall_loss"]])
    for i, (x, y) in enumerate(zip(res["overall_loss"], L["overall_loss"])):
        e = abs(float(x) - float(y)) / max(abs(float(y)), 1e-12)
        if e > 1e-4: bad += 1; print(f"  LOSS[{i}] rel err {e:.2e}  !!")
    allref = {**enc, **z, **dec}
    for k in sv._lib.OUTPUT_FIELDS:
        if k not in res: continue
        ref = allref[k].detach().numpy(); got = res[k].cpu().numpy()
        if k == "generated_adj":
            nd = int((ref != got).sum()); print(f"  {k:24s} mismatches {nd} / {ref.size}")
            # bit-exact rule on the GPU's own logits
            own = O.mask_and_threshold(torch.zeros(1))[0] if False else None
            continue
        e, m = rel_err(got, ref)
        flag = "" if e < 1e-4 else "  !!"
        if flag: bad += 1
        print(f"  {k:24s} max|ref| {m:.3e}  rel-to-max err {e:.2e}{flag}")
    gg = eng.get_grads()
    for k in names_o:
        e, m = rel_err(gg[k].numpy(), grads[k].numpy())
        flag = "" if e < 1e-3 else "  !!"
        if flag: bad += 1
        print(f"  grad {k:36s} max|ref| {m:.3e}  rel-to-max err {e:.2e}{flag}")
    # Adam steps
    P32 = {k: v.to(torch.float32).clone() for k, v in P.items()}
    adam = O.TFAdam(P32, cfg.learning_rate)
    inp32, noise32 = O.cast(inp, torch.float32), O.cast(noise, torch.float32)
    for st in range(a.steps):
        _, _, _, Ls, g32 = O.loss_and_grads(P32, inp32, noise32, cfg, "factored")
        adam.step(P32, g32)
        r = eng.train_step(inp, noise)
        print(f"  step {st}: oracle cost {float(Ls['cost']):.6f} gpu cost {float(r['overall_loss'][0]):.6f}")
    pg = eng.get_params()
    worst = 0
    for k in names_o:
        d = (pg[k].double() - P32[k].double()).abs().max().item()
        worst = max(worst, d)
    print(f"  params after {a.steps} Adam steps: max abs diff {worst:.3e} (lr {cfg.learning_rate})")
    if worst > 0.2 * cfg.learning_rate: bad += 1
    print("RESULT", "FAIL" if bad else "OK", "bad =", bad, " launches =", eng.launch_count())
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
