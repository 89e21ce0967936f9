"""Placeholder / fetch / Session shim: the `sess.run(fetches, feed_dict)` seam of
main.py:331,361,368,371 on top of the Engine.  Feeds are host numpy arrays with
the static shapes of main.py:253-264; fetches come back as fresh numpy arrays.
Shape mismatches raise (TensorFlow: InvalidArgumentError), nothing is silent."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .engine import SndvaeError

PLACEHOLDER_KEYS = ("features", "spatial", "adj", "adj_truth", "feature_truth", "spatial_truth", "rel_truth", "rel",
                    "dropout", "global_iter")
NOISE_PLACEHOLDERS = ("eps_s", "eps_sg", "eps_g")


class Placeholder:
    """tf.placeholder(tf.float32, shape) stand-in: hashable key of a feed_dict."""

    def __init__(self, name, shape=None):
        self.name, self.shape = name, None if shape is None else tuple(shape)

    def __repr__(self):
        return f"<placeholder {self.name} {self.shape}>"


class Fetch:
    """A fetchable model / optimizer attribute (e.g. model.generated_adj, opt.opt_op)."""

    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __repr__(self):
        return f"<fetch {self.name}>"


def make_placeholders(batch_size, sampling_num, num_nodes, num_features, spatial_dim) -> Dict[str, Placeholder]:
    """The placeholder dict of main.py:253-264."""
    B, S, N, F, D = batch_size, sampling_num, num_nodes, num_features, spatial_dim
    ph = {
        "features": Placeholder("features", (B * S, N, F)), "spatial": Placeholder("spatial", (B * S, N, D)),
        "adj": Placeholder("adj", (B * S, N, N)), "adj_truth": Placeholder("adj_truth", (B, N, N)),
        "feature_truth": Placeholder("feature_truth", (B, N, F)), "spatial_truth": Placeholder("spatial_truth", (B, N, D)),
        "rel_truth": Placeholder("rel_truth", (B, N, N, 1)), "rel": Placeholder("rel", (B * S, N, N, 1)),
        "dropout": Placeholder("dropout", ()), "global_iter": Placeholder("global_iter", ()),
    }
    # explicit noise inputs: when fed they replace the tf.random.normal draws
    for k in NOISE_PLACEHOLDERS:
        ph[k] = Placeholder(k)
    return ph


class Session:
    """with Session() as sess: sess.run([opt.opt_op, opt.overall_loss, model.generated_adj], feed_dict)"""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        if fetches is None or isinstance(fetches, _NoOp):
            return None                           # sess.run(tf.global_variables_initializer()) (main.py:302)
        single = not isinstance(fetches, (list, tuple))
        flist: List = [fetches] if single else list(fetches)
        flat = []
        for f in flist:
            flat.extend(f if isinstance(f, (list, tuple)) else [f])
        if not any(isinstance(f, Fetch) for f in flat):
            return None if single else [None for _ in flist]
        models = {id(f.owner.model if hasattr(f.owner, "model") else f.owner): (f.owner.model if hasattr(f.owner, "model") else f.owner)
                  for f in flat if isinstance(f, Fetch)}
        if len(models) != 1:
            raise SndvaeError("Session.run needs fetches of exactly one model")
        model = next(iter(models.values()))
        values = model._run([f.name for f in flat if isinstance(f, Fetch)], feed_dict or {})
        it = iter(values)
        nxt = lambda f: next(it) if isinstance(f, Fetch) else None     # no-op fetches (initialisers) come back as None
        out = []
        for f in flist:
            if isinstance(f, (list, tuple)):
                out.append([nxt(x) for x in f])
            else:
                out.append(nxt(f))
        return out[0] if single else out


class _NoOp:
    """A fetch that does nothing and returns None (TF: an Operation, e.g. the variable initialiser)."""

    def __repr__(self):
        return "<no-op>"


def global_variables_initializer():
    """The reference calls sess.run(tf.global_variables_initializer()) (main.py:302); here the
    constructor already initialised the variables, so this is an accepted no-op fetch."""
    return _NoOp()
