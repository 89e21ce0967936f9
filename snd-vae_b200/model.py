"""`SGCNModelVAE` of model.py (the 3-latent "disentangled" SND-VAE) with the reference's
constructor signature (model.py:22) and attribute names (model.py:78-80,114-151), backed
by the sm_100a engine.  Attributes are fetch handles for `Session.run`."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .engine import Engine, SndvaeError, make_config
from .flags import FLAGS
from .params import init_params
from .session import Fetch, Placeholder

_LOSS_NAMES_DIS = ("cost", "spatial_cost", "adj_cost", "node_cost", "kl_g", "kl_s", "kl_sg")
_LOSS_NAMES_BASE = ("cost", "spatial_cost", "adj_cost", "node_cost", "kl_sg")
_MODEL_FETCHES = ("z_mean_s", "z_std_s", "z_mean_g", "z_std_g", "z_mean_sg", "z_std_sg", "z_s", "z_sg", "z_g",
                  "generated_adj", "generated_adj_prob", "generated_spatial", "generated_node_feat")


from ._lib import LOSS_VARIANTS


class _ModelBase:
    model_type = "disentangled"

    def __init__(self, placeholders: Dict[str, Placeholder], num_features, num_nodes, seed=7, **kwargs):
        self.placeholders = placeholders
        self.input_dim, self.n_samples = num_features, num_nodes
        F = FLAGS
        S = F.sampling_num if self.model_type != "base" else 1
        cfg = make_config(
            num_nodes, F.batch_size, self.model_type, num_feature=num_features, spatial_dim=F.spatial_dim, sampling_num=S,
            node_h_size=F.node_h_size, s_channel=F.s_channel[:3], s_hidden_size=F.s_hidden_size, s_latent_size=F.s_latent_size,
            g_conv_hidden=F.g_conv_hidden[:2], g_hidden_size=F.g_hidden_size, g_latent_size=F.g_latent_size,
            sg_conv_hidden=F.sg_conv_hidden, sg_hidden_size=F.sg_hidden_size, sg_latent_size=F.sg_latent_size,
            s_d_channel=F.s_d_channel[:3], n_d_channel=F.n_d_channel[:F.graph_deconv_layers],
            e_d_hidden=F.e_d_hidden[:F.graph_deconv_layers], learning_rate=F.learning_rate,
            use_tensor_cores=F.use_tensor_cores, chunk_graphs=F.chunk_graphs,
            # loss branch of OptimizerVAE selected by FLAGS.model_type (optimizer.py:160-190)
            loss_variant=LOSS_VARIANTS.get(F.model_type, 0), gamma=float(F.gamma), C_max=float(F.C_max),
            C_stop_iter=float(F.C_stop_iter), C_step=float(F.C_step))
        self.engine = Engine(cfg)
        self.engine.set_params({k: torch.from_numpy(v) for k, v in init_params(self.engine.table, seed).items()})
        self.mode = F.type                     # 'train' | 'test_reconstruct' | 'test_generation' (model.py:79-90)
        self._gen = torch.Generator(device="cpu").manual_seed(seed + 1)
        for name in _MODEL_FETCHES:
            setattr(self, name, Fetch(self, name))
        self.vars = [name for name, _, _ in self.engine.table]     # tf.trainable_variables() (model.py:92-94)
        self.optimizer = None

    # -- execution behind Session.run ---------------------------------------------------
    def _noise(self, feeds):
        c = self.engine.cfg
        B, S = self.engine.B, self.engine.S
        shapes = {"eps_s": (B, c.s_latent_size), "eps_sg": (B * S, c.sg_latent_size), "eps_g": (B, c.g_latent_size)}
        noise = {}
        for k, shp in shapes.items():          # tf.random.normal draw order: s, sg, g (model.py:155-159)
            noise[k] = feeds[k] if k in feeds else torch.randn(shp, generator=self._gen)
        return noise

    def _run(self, names, feed_dict):
        feeds = {}
        for ph, val in feed_dict.items():
            key = ph.name if isinstance(ph, Placeholder) else str(ph)
            feeds[key] = val
        noise = self._noise(feeds)
        if "global_iter" in feeds:             # main.py:329: only the capacity loss reads it
            self.engine.set_global_iter(int(np.asarray(feeds["global_iter"]).reshape(-1)[0]))
        loss_names = _LOSS_NAMES_DIS if self.engine.dis else _LOSS_NAMES_BASE
        want_opt = "opt_op" in names
        want_loss = any(n in loss_names or n == "overall_loss" for n in names)
        fetch = tuple(n for n in names if n in _MODEL_FETCHES)
        if (want_opt or want_loss) and self.optimizer is None:
            raise SndvaeError("loss / opt_op fetched before OptimizerVAE was constructed")
        if want_opt:
            if self.mode != "train":
                raise SndvaeError("opt_op exists only when FLAGS.type == 'train' (main.py:283)")
            res = self.engine.train_step(feeds, noise, fetch=fetch)
        elif self.mode == "test_generation":
            # encoder on the fed data for z_mean_*, decoder on prior draws (model.py:83-85,163-169; quirk Q11)
            enc_f = tuple(n for n in fetch if n.startswith("z_mean") or n.startswith("z_std"))
            res = self.engine.forward(feeds, noise, fetch=enc_f) if enc_f or want_loss else {}
            z = self._noise({})
            dec_f = tuple(n for n in fetch if n.startswith("generated"))
            if dec_f or any(n in ("z_s", "z_sg", "z_g") for n in fetch):
                res.update(self.engine.generate(z["eps_s"], z["eps_sg"], z["eps_g"], fetch=dec_f))
                res.update({"z_s": z["eps_s"], "z_sg": z["eps_sg"], "z_g": z["eps_g"]})
        else:
            res = self.engine.forward(feeds, noise, fetch=fetch)
        out = []
        for n in names:
            if n == "opt_op":
                out.append(None)
            elif n == "overall_loss":
                out.append([np.float32(x) for x in res["overall_loss"]])
            elif n in loss_names:
                out.append(np.float32(res["overall_loss"][loss_names.index(n)]))
            else:
                v = res[n]
                out.append(v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
        return out

    # -- tf.train.Saver stand-in (main.py:299,351-352): .npz keyed by TF variable name ----
    @staticmethod
    def _ckpt_path(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"      # saver.save(p) / saver.restore(p) take the same name

    def save(self, path):
        p = {k: v.numpy() for k, v in self.engine.get_params().items()}
        m, v, bp = self.engine.get_adam()
        with open(self._ckpt_path(path), "wb") as f:
            np.savez(f, **p, **{"adam_m/" + k: x.numpy() for k, x in m.items()},
                     **{"adam_v/" + k: x.numpy() for k, x in v.items()}, adam_beta_pows=bp)
        return self._ckpt_path(path)

    def restore(self, path):
        z = np.load(self._ckpt_path(path))
        names = [n for n, _, _ in self.engine.table]
        self.engine.set_params({k: torch.from_numpy(z[k]) for k in names})
        if "adam_beta_pows" in z:
            self.engine.set_adam({k: torch.from_numpy(z["adam_m/" + k]) for k in names},
                                 {k: torch.from_numpy(z["adam_v/" + k]) for k in names}, z["adam_beta_pows"])


class SGCNModelVAE(_ModelBase):
    """model.py:19-229.  Placeholders used: feature_truth, features, spatial_truth, adj_truth,
    adj, rel, dropout (model.py:24-33); `spatial` / `rel_truth` are accepted and unused."""
    model_type = "disentangled"

    def __init__(self, placeholders, num_features, num_nodes, group_type=None, dim=None, dim_a=None, dim_b=None, dim_c=None,
                 **kwargs):
        self.group_type, self.dim, self.dim_a, self.dim_b, self.dim_c = group_type, dim, dim_a, dim_b, dim_c
        super().__init__(placeholders, num_features, num_nodes, **kwargs)

    def sample(self, z_s, z_sg, z_g):
        """model.py:227-229: decode caller-provided latents (arrays, not fetch handles)."""
        r = self.engine.generate(z_s, z_sg, z_g)
        return r["generated_adj"], r["generated_adj_prob"], r["generated_spatial"], r["generated_node_feat"]
