"""Reference initialisers for tf.trainable_variables() (what
`sess.run(tf.global_variables_initializer())` does, main.py:302): truncated
normal(0.02) for GraphConvolution / e2e weights (layers.py:118-119,434-435),
normal(0.02) for `linear` and SGC matrices (layers.py:158-169,569-570), zeros
for biases, Keras glorot_uniform for tf.layers.conv1d kernels, gamma=1 / beta=0
for BatchNormalization.  Keyed by TF variable name; shapes come from the
library's parameter table."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np


def init_params(table: List[Tuple[str, int, Tuple[int, ...]]], seed: int = 7) -> Dict[str, "np.ndarray"]:
    rng = np.random.default_rng(seed)
    out = {}
    for name, _off, shape in table:
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("bias", "bias1", "bias2", "bias3", "biases1", "beta"):
            v = np.zeros(shape, np.float32)
        elif leaf == "gamma":
            v = np.ones(shape, np.float32)
        elif leaf in ("w", "w1"):                       # tf.truncated_normal_initializer(stddev=0.02)
            v = rng.standard_normal(shape)
            bad = np.abs(v) > 2.0
            while bad.any():
                v[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(v) > 2.0
            v = (v * 0.02).astype(np.float32)
        elif leaf == "kernel":                          # glorot_uniform, fan = k * channels
            k, ci, co = shape
            lim = math.sqrt(6.0 / (k * ci + k * co))
            v = rng.uniform(-lim, lim, shape).astype(np.float32)
        else:                                           # Matrix, Matrix1..3: random_normal(stddev=0.02)
            v = (rng.standard_normal(shape) * 0.02).astype(np.float32)
        out[name] = v
    return out
