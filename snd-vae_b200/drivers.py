"""The three driver loops of the reference's `main(beta, type_model)` on top of the Session shim: training
(main.py:299-356), reconstruction (main.py:374-426) and generation (main.py:428-469), with its Saver protocol (save every
100 epochs, restore a named epoch: main.py:299,351-352,376,430) and the `z_*.npy` dumps the latent-traversal code loads
(main.py:411-416).  Hard-coded absolute paths of the reference become arguments; the unshipped
`utils.evaluation` metrics (quirk Q1) are not re-invented -- the loops return the arrays those functions were given.

Batching is the reference's: `int(len / (batch_size * sampling_num))` batches of `batch_size` graphs, the last partial batch
dropped (quirk Q14).  The per-sample tensors are aligned graph-major (row b*S+s belongs to graph b; quirk Q6 fixed, see
data.py): `features` / `spatial` / `rel` may be passed per graph ([G, ...]) and are repeated S times, or already per sample."""
from __future__ import annotations

import os
import time
from collections import defaultdict
from typing import Dict, Optional

import numpy as np

from .flags import FLAGS
from .preprocessing import construct_feed_dict_train
from .session import Session

_DIS_TYPES = ("disentangled", "disentangled_C", "NED-VAE-IP", "beta-TCVAE")


class LossesLogger:
    """Stand-in for utils.utils.LossesLogger (main.py:25,278-280,353): one CSV line per epoch and loss name."""

    def __init__(self, path: Optional[str] = None):
        self.path, self.rows = path, []
        if path:
            with open(path, "w") as f:
                f.write("Epoch,Loss,Value\n")

    def log(self, epoch, storer):
        for k, v in storer.items():
            m = float(np.mean(v))
            self.rows.append((epoch, k, m))
            if self.path:
                with open(self.path, "a") as f:
                    f.write(f"{epoch},{k},{m}\n")


def _per_sample(a, S, rows):
    a = np.asarray(a)
    return a if a.shape[0] == rows else np.repeat(a, S, axis=0)


def _batches(data: Dict[str, np.ndarray], B: int, S: int):
    """main.py:305-323: yields the eight feed arrays of every full batch."""
    adj = np.asarray(data["adj"])
    G = adj.shape[0] // S
    feature, spatial, rel = (_per_sample(data[k], S, G * S) for k in ("features", "spatial", "rel"))
    ft = np.asarray(data.get("feature_truth", np.asarray(data["features"])[::S] if np.asarray(data["features"]).shape[0] == G * S else data["features"]))
    st = np.asarray(data.get("spatial_truth", np.asarray(data["spatial"])[::S] if np.asarray(data["spatial"]).shape[0] == G * S else data["spatial"]))
    rt = data.get("rel_truth")
    if rt is None:
        r = np.asarray(data["rel"]); r = r[::S] if r.shape[0] == G * S else r
        rt = r.reshape(r.shape[:3] + (1,))
    rel = rel.reshape(rel.shape[:3] + (1,))
    for i in range(int(adj.shape[0] / (B * S))):
        s, g = slice(i * B * S, (i + 1) * B * S), slice(i * B, (i + 1) * B)
        yield i, (feature[s], spatial[s], adj[s], rel[s], np.asarray(data["adj_truth"])[g], ft[g], st[g], np.asarray(rt)[g])


def train(model, opt, placeholders, data, epochs=None, ckpt_dir=None, save_every=100, logger: Optional[LossesLogger] = None,
          verbose=False):
    """main.py:299-356.  Returns (generated_adj of the last epoch's batches, per-epoch mean losses)."""
    B, S, N = FLAGS.batch_size, model.engine.S, model.engine.N
    epochs = FLAGS.epochs if epochs is None else epochs
    history, check = [], []
    dis = FLAGS.model_type in _DIS_TYPES
    with Session() as sess:
        for epoch in range(epochs):
            storer, check, t_epoch = defaultdict(list), [], time.time()
            for i, arrs in _batches(data, B, S):
                t = time.time()
                feed_dict = construct_feed_dict_train(*arrs, placeholders)
                feed_dict.update({placeholders["dropout"]: FLAGS.dropout, placeholders["global_iter"]: epoch})
                outs = sess.run([opt.opt_op, opt.overall_loss, model.generated_adj], feed_dict=feed_dict)
                ol = outs[1]
                acc = float((outs[2] == arrs[4]).sum()) / (B * N * N)              # main.py:334
                for k, v in (("loss", ol[0]), ("spatial_loss", ol[1]), ("adj_loss", ol[2]), ("adj_acc", acc), ("node_loss", ol[3])):
                    storer[k].append(v)
                if dis:
                    storer["graph_kl"].append(ol[4]); storer["spatial_kl"].append(ol[5]); storer["sg_kl"].append(ol[6])
                else:
                    storer["sg_kl"].append(ol[4])
                check.append(outs[2])
                if verbose:
                    print("Epoch:", "%04d" % (epoch + 1), "loss=", "{:.5f}".format(ol[0]), "time=", "{:.5f}".format(time.time() - t))
            if verbose:
                print("epoch time=", "{:.5f}".format(time.time() - t_epoch))
            if ckpt_dir and epoch % save_every == 0:                                  # main.py:351-352
                os.makedirs(ckpt_dir, exist_ok=True)
                model.save(os.path.join(ckpt_dir, f"model_dgt_global_{epoch}.ckpt"))
            history.append({k: float(np.mean(v)) for k, v in storer.items()})
            if logger:
                logger.log(epoch, storer)
    return np.array(check), history


def _encode_decode(model, placeholders, data, restore=None):
    """The shared loop of main.py:374-409 / 428-461: `generate_new_train` on every batch."""
    if restore:
        model.restore(restore)                                                       # saver.restore (main.py:376,430)
    B, S = FLAGS.batch_size, model.engine.S
    dis = model.engine.dis
    out = defaultdict(list)
    with Session() as sess:
        for i, arrs in _batches(data, B, S):
            feed_dict = construct_feed_dict_train(*arrs, placeholders)
            feed_dict.update({placeholders["dropout"]: 1.0})                         # generate_new_train (main.py:358-362)
            if dis:
                z_s, z_sg, z_g, adj, spatial, node = sess.run([model.z_mean_s, model.z_mean_sg, model.z_mean_g, model.generated_adj,
                                                               model.generated_spatial, model.generated_node_feat], feed_dict=feed_dict)
                out["z_s"].append(z_s.reshape((B, -1))); out["z_g"].append(z_g.reshape((B, -1)))
            else:
                z_sg, adj, node, spatial = sess.run([model.z_mean_sg, model.generated_adj, model.generated_node_feat,
                                                     model.generated_spatial], feed_dict=feed_dict)
            out["z_sg"].append(z_sg.reshape((B, S, -1)).mean(axis=1))                # main.py:405
            out["generated_adj"].append(adj); out["generated_nodes"].append(node); out["generated_spatial"].append(spatial)
    N, F, D = model.engine.N, model.engine.F, model.engine.D
    res = {k: np.array(v) for k, v in out.items()}
    res["generated_adj"] = res["generated_adj"].reshape(-1, N, N)
    res["generated_nodes"] = res["generated_nodes"].reshape(-1, N, F)
    res["generated_spatial"] = res["generated_spatial"].reshape(-1, N, D)
    return res


def reconstruct(model, placeholders, data, restore=None, out_dir=None, vae_type="disentangled"):
    """FLAGS.type == 'test_reconstruct' (main.py:374-426): the model must have been built with that type, so that the decoder
    sees the posterior samples of get_z (model.py:81-82).  Writes `<vae_type>_z_{s,sg,g}.npy` (main.py:411-416) into out_dir."""
    if model.mode != "test_reconstruct":
        raise ValueError("reconstruct() needs a model built with FLAGS.type == 'test_reconstruct' (model.py:81-82)")
    res = _encode_decode(model, placeholders, data, restore)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        keys = ("z_s", "z_sg", "z_g") if model.engine.dis else ("z_sg",)
        for k in keys:
            np.save(os.path.join(out_dir, f"{vae_type}_{k}.npy"), res[k])
    return res


def generate(model, placeholders, data, restore=None):
    """FLAGS.type == 'test_generation' (main.py:428-469): the encoder still runs on the fed graphs (the returned z_* are its
    posterior means, quirk Q11) while the decoder is driven by prior draws (get_random_z, model.py:83-85,163-169)."""
    if model.mode != "test_generation":
        raise ValueError("generate() needs a model built with FLAGS.type == 'test_generation' (model.py:83-85)")
    return _encode_decode(model, placeholders, data, restore)
