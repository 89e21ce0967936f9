#!/usr/bin/env python
"""Pin the golden fixtures against the REAL reference graph (TensorFlow 1.14 / 1.15).

The build image has no TensorFlow, so tests/golden/*.npz are outputs of the CPU oracle (oracle/sndvae_oracle.py) and every
parity claim of this repository is "unpinned at the TensorFlow boundary".  This script closes that pin the day a TF-1.x
environment exists.  It has two halves:

  export   (runs HERE: numpy + torch + the oracle)  writes tests/golden/inputs_<case>.npz -- the seeded parameters (TF variable
           names), the eight feeds of main.py:253-264 and the three noise tensors eps_s / eps_sg / eps_g.  Committed.

  run      (runs in a TF-1.x environment: numpy + tensorflow only, plus a checkout of xguo7/SND-VAE)
           1. defines the flags of main.py:42-103 (values of the synthetic2 block, main.py:181-215) through tf.app.flags;
           2. replaces tf.random.normal by placeholders while the model is built, so that the draws of get_z
              (model.py:155-159, order s, sg, g) are fed with the fixture's eps_* instead of fresh noise;
           3. builds the reference's own SGCNModelVAE (model.py:22 / model_joint.py:14) and OptimizerVAE (optimizer.py:124) with
              the placeholders of main.py:253-264;
           4. assigns the seeded parameters to tf.trainable_variables() (matched by name, falling back to creation order +
              shape -- the fixture is in creation order), leaves the Keras BN moving statistics at 0 / 1;
           5. fetches overall_loss, z_*, generated_*, tf.gradients(cost, variables) and the cost of three opt_op steps;
           6. rewrites tests/golden/<case>.npz with the same keys tests/golden/make_golden.py writes, plus `source`.
           After that `python -m pytest tests -m "not gpu"` checks the oracle against TensorFlow's numbers and
           `pytest -m gpu` checks the CUDA path against them (tests/test_oracle.py, tests/test_gpu_parity.py::test_golden_fixtures).

  python tools/pin_against_tf.py export
  python tools/pin_against_tf.py run --reference /path/to/SND-VAE [--case dis_n8 base_n8]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {"dis_n8": ("disentangled", 8, 4, 3), "base_n8": ("base", 8, 4, 1)}       # model, N, B, S  (as make_golden.py)
FEEDS = ("features", "spatial", "adj", "rel", "adj_truth", "feature_truth", "spatial_truth", "rel_truth")


# ----------------------------------------------------------------------------------------------------------------------------
def export():
    import torch
    sys.path.insert(0, ROOT)
    from oracle import sndvae_oracle as O
    for name, (model, N, B, S) in CASES.items():
        cfg = O.Config(num_nodes=N, model_type=model, sampling_num=S)
        P = O.init_params(cfg, 7, torch.float64)
        g = torch.Generator().manual_seed(1)
        for k in P:
            P[k] = P[k] + 0.05 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
        inp = O.synthetic_inputs(cfg, B, 5, torch.float64)
        noise = O.synthetic_noise(cfg, B, 9, torch.float64)
        out = {"model": model, "N": N, "B": B, "S": cfg.S, "param_order": np.array([n for n, _, _ in O.param_table(cfg)])}
        out.update({"param/" + k: v.numpy().astype(np.float32) for k, v in P.items()})
        out.update({"feed/" + k: inp[k].numpy().astype(np.float32) for k in FEEDS})
        out.update({"noise/" + k: v.numpy().astype(np.float32) for k, v in noise.items()})
        path = os.path.join(GOLD, f"inputs_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, f"({os.path.getsize(path) / 1e3:.0f} KB)")


# ----------------------------------------------------------------------------------------------------------------------------
SYNTHETIC2_FLAGS = dict(      # main.py:42-103 with the synthetic2 overrides of main.py:181-215
    spatial_conv_layers=3, s_channel=[10, 10, 20], s_kernel_size=[5, 5, 5], s_strides=[1, 1, 1], s_hidden_size=100, s_latent_size=100,
    graph_conv_layers=2, g_conv_hidden=[10, 20], g_hidden_size=100, g_latent_size=100,
    spatial_graph_conv_layers=2, sg_conv_hidden=[[20, 20, 20], [50, 50, 50]], sg_hidden_size=100, sg_latent_size=100,
    spatial_deconv_layers=3, s_d_channel=[50, 20, 10], s_d_kernel_size=[5, 5, 5], s_d_strides=[1, 1, 1],
    graph_deconv_layers=2, n_d_channel=[50, 20], n_d_kernel_size=[5, 5], n_d_strides=[1, 1], e_d_hidden=[50, 20], node_h_size=20,
    learning_rate=0.0008, epochs=1, dropout=0.0, num_feature=1, spatial_dim=2, num_edge_feature=2,
    dataset="synthetic2", vae=1, C_max=100.0, C_stop_iter=100.0, C_step=20.0, gamma=100.0,
)


def define_flags(tf, model, B, S):
    flags = tf.app.flags
    vals = dict(SYNTHETIC2_FLAGS)
    vals.update(batch_size=B, sg_batch_size=B, decoder_batch_size=B, sg_decoder_batch_size=B, sampling_num=S,
                model_type="disentangled" if model == "disentangled" else "base", type="train")
    for k, v in vals.items():
        if k in flags.FLAGS:
            setattr(flags.FLAGS, k, v)
        elif isinstance(v, bool) or isinstance(v, int):
            flags.DEFINE_integer(k, v, k)
        elif isinstance(v, float):
            flags.DEFINE_float(k, v, k)
        elif isinstance(v, str):
            flags.DEFINE_string(k, v, k)
        else:
            flags.DEFINE_list(k, v, k)
    flags.FLAGS(sys.argv[:1])          # mark as parsed
    for k, v in vals.items():          # DEFINE_list stores strings when parsed from argv: force the Python values
        setattr(flags.FLAGS, k, v)
    return flags.FLAGS


def run_case(tf, reference, name):
    model_kind, N, B, S = CASES[name]
    fx = np.load(os.path.join(GOLD, f"inputs_{name}.npz"), allow_pickle=False)
    S = int(fx["S"])
    tf.reset_default_graph()
    FLAGS = define_flags(tf, model_kind, B, S)
    sys.path.insert(0, reference)
    for m in ("model", "model_joint", "optimizer", "layers"):
        sys.modules.pop(m, None)
    # the tf.random.normal draws of get_z become placeholders, in draw order (model.py:155-159: s, sg, g; model_joint.py: sg)
    eps_ph = []
    real_normal = tf.random.normal

    def fed_normal(shape, *a, **k):
        ph = tf.placeholder(tf.float32, shape=[int(x) for x in shape], name=f"eps_{len(eps_ph)}")
        eps_ph.append(ph)
        return ph
    tf.random.normal = fed_normal
    tf.random_normal = fed_normal
    try:
        F, D = int(FLAGS.num_feature), int(FLAGS.spatial_dim)
        ph = {      # main.py:253-264
            "features": tf.placeholder(tf.float32, [B * S, N, F]), "spatial": tf.placeholder(tf.float32, [B * S, N, D]),
            "adj": tf.placeholder(tf.float32, [B * S, N, N]), "adj_truth": tf.placeholder(tf.float32, [B, N, N]),
            "feature_truth": tf.placeholder(tf.float32, [B, N, F]), "spatial_truth": tf.placeholder(tf.float32, [B, N, D]),
            "rel_truth": tf.placeholder(tf.float32, [B, N, N, 1]), "rel": tf.placeholder(tf.float32, [B * S, N, N, 1]),
            "dropout": tf.placeholder_with_default(0., shape=()), "global_iter": tf.placeholder_with_default(0., shape=()),
        }
        mod = __import__("model" if model_kind == "disentangled" else "model_joint")
        optm = __import__("optimizer")
        model = mod.SGCNModelVAE(ph, F, N)
        with tf.name_scope("optimizer"):
            opt = optm.OptimizerVAE(preds_edge=model.generated_adj_prob, preds_node=model.generated_node_feat, preds_spatial=model.generated_spatial,
                                    labels_edge=ph["adj_truth"], labels_node=ph["feature_truth"], labels_spatial=ph["spatial_truth"],
                                    labels_rel=ph["rel_truth"], global_iter=ph["global_iter"], model=model, num_nodes=N,
                                    pos_weight=1.0, norm=1.0, beta=1)
    finally:
        tf.random.normal = real_normal
        tf.random_normal = real_normal
    order = [str(x) for x in fx["param_order"]]
    tvars = tf.trainable_variables()
    assert len(tvars) == len(order), f"{len(tvars)} trainable variables in the graph, {len(order)} in the fixture"
    assign, names = [], []
    for v, want in zip(tvars, order):                     # creation order; names must agree up to the ':0' suffix
        got = v.name.rsplit(":", 1)[0]
        val = fx["param/" + want]
        assert tuple(v.shape.as_list()) == val.shape, f"{got} has shape {v.shape}, fixture {want} has {val.shape}"
        if got != want:
            print(f"[pin] name differs (matched by creation order and shape): graph '{got}' <- fixture '{want}'")
        assign.append(v.assign(val)); names.append(want)
    noise_keys = ["eps_s", "eps_sg", "eps_g"] if model_kind == "disentangled" else ["eps_sg"]
    assert len(eps_ph) >= len(noise_keys), f"expected {len(noise_keys)} tf.random.normal draws in get_z, saw {len(eps_ph)}"
    feed = {ph[k]: fx["feed/" + k] for k in FEEDS}
    feed.update({p: fx["noise/" + k] for p, k in zip(eps_ph, noise_keys)})
    for p in eps_ph[len(noise_keys):]:                    # get_random_z draws of the test branches: unused when type == 'train'
        feed[p] = np.zeros([int(x) for x in p.shape], np.float32)
    grads = tf.gradients(opt.cost, tvars)
    fetch_names = ["z_mean_sg", "z_std_sg", "z_sg", "generated_adj", "generated_adj_prob", "generated_spatial", "generated_node_feat"]
    if model_kind == "disentangled":
        fetch_names += ["z_mean_s", "z_std_s", "z_s", "z_mean_g", "z_std_g", "z_g"]
    out = {"N": N, "B": B, "S": S, "source": f"tensorflow {tf.__version__}, reference checkout {os.path.abspath(reference)}"}
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        sess.run(assign)
        vals = sess.run([opt.overall_loss] + [getattr(model, k) for k in fetch_names] + [g for g in grads if g is not None], feed)
        out["overall_loss"] = np.asarray(vals[0], np.float64)
        for k, v in zip(fetch_names, vals[1:1 + len(fetch_names)]):
            out[k] = np.asarray(v)
        gi = iter(vals[1 + len(fetch_names):])
        for n, g in zip(names, grads):
            gv = np.zeros(fx["param/" + n].shape, np.float32) if g is None else np.asarray(next(gi))
            out["gradsum/" + n] = np.array([gv.astype(np.float64).sum(), np.abs(gv.astype(np.float64)).sum()])
            out["grad/" + n] = gv.astype(np.float32)
        costs = []
        for _ in range(3):                                # fp32 TF-Adam trajectory (optimizer.py:125,197)
            c, _ = sess.run([opt.cost, opt.opt_op], feed)
            costs.append(float(c))
        out["adam_costs"] = np.array(costs)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("[pin] rewrote", path, "from TensorFlow; losses", out["overall_loss"])


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("mode", choices=["export", "run"])
    ap.add_argument("--reference", default=os.environ.get("SNDVAE_REFERENCE", "/root/reference"))
    ap.add_argument("--case", nargs="*", default=list(CASES))
    a = ap.parse_args()
    if a.mode == "export":
        return export()
    try:
        import tensorflow as tf
    except ImportError:
        sys.exit("TensorFlow is not installed in this environment: `run` needs TF 1.14 / 1.15 (tensorflow.compat.v1 of TF 2 will not "
                 "do: layers.py uses tf.contrib).  The `export` half has no such requirement.")
    if not tf.__version__.startswith("1."):
        sys.exit(f"TensorFlow {tf.__version__} found; the reference needs 1.14 / 1.15 (tf.contrib, tf.app.flags)")
    for name in a.case:
        run_case(tf, a.reference, name)


if __name__ == "__main__":
    main()
