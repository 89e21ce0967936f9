"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/sndvae.h declares, and its parameter table (host logic) matches the
oracle's tf.trainable_variables() order.  No compute calls: there is no GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import sndvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(built):
    lib = built._lib.load()
    hdr = open(os.path.join(ROOT, "include", "sndvae.h")).read()
    declared = set(re.findall(r"\b(sndvae_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"sndvae_config", "sndvae_inputs", "sndvae_noise", "sndvae_outputs", "sndvae_param_info", "sndvae_handle"}
    assert declared, "no declarations parsed"
    assert declared == set(built._lib.SYMBOLS), declared ^ set(built._lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_layout(built):
    lib = built._lib.load()
    cfg = built._lib.Config()
    assert lib.sndvae_default_config(C.byref(cfg)) == 0
    # synthetic2 defaults (main.py:181-215)
    assert (cfg.num_nodes, cfg.num_feature, cfg.spatial_dim, cfg.sampling_num, cfg.node_h_size) == (25, 1, 2, 10, 20)
    assert list(cfg.s_channel) == [10, 10, 20] and list(cfg.g_conv_hidden) == [10, 20]
    assert [list(r) for r in cfg.sg_conv_hidden] == [[20, 20, 20], [50, 50, 50]]
    assert list(cfg.s_d_channel) == [50, 20, 10] and list(cfg.n_d_channel) == [50, 20] and list(cfg.e_d_hidden) == [50, 20]
    assert (cfg.sg_hidden_size, cfg.sg_latent_size, cfg.batch_size) == (100, 100, 10)
    assert abs(cfg.learning_rate - 0.0008) < 1e-9 and cfg.beta == 1.0
    assert abs(cfg.adam_beta1 - 0.9) < 1e-7 and abs(cfg.adam_beta2 - 0.999) < 1e-7 and abs(cfg.adam_eps - 1e-8) < 1e-15
    # loss-branch tail of the struct (main.py:95-98, optimizer.py:183): a layout slip between header and ctypes mirror shows here
    assert (cfg.loss_variant, cfg.gamma, cfg.C_max, cfg.C_stop_iter, cfg.C_step, cfg.dip_lambda_od, cfg.dip_lambda_d) == (0, 100.0, 100.0, 100.0, 20.0, 10.0, 100.0)
    assert cfg.use_tensor_cores == 2


@pytest.mark.parametrize("model,N", [("disentangled", 25), ("base", 25), ("disentangled", 256)])
def test_param_table_matches_oracle(built, model, N):
    """sndvae_create builds the table on the host before touching the device, so it can be read
    back even when create fails for lack of a GPU."""
    lib = built._lib.load()
    cfg = built.make_config(N, 4, model)
    h = C.c_void_p()
    rc = lib.sndvae_create(C.byref(cfg), None, C.byref(h))
    assert h.value is not None
    if not torch.cuda.is_available():
        assert rc == -2 and b"no CUDA device" in lib.sndvae_last_error(h)      # fails loudly, no CPU fallback
    n = lib.sndvae_num_params(h)
    tab = (built._lib.ParamInfo * n)()
    assert lib.sndvae_param_table(h, tab, n) == 0
    ocfg = O.Config(num_nodes=N, model_type=model)
    ref = O.param_table(ocfg)
    assert n == len(ref)
    off = 0
    for t, (name, shape, _) in zip(tab, ref):
        assert t.name.decode() == name
        assert tuple(t.shape[i] for i in range(t.rank)) == tuple(shape)
        assert t.offset == off and t.offset % 4 == 0
        off += (int(np.prod(shape)) + 3) // 4 * 4
    assert lib.sndvae_param_count(h) == off
    lib.sndvae_destroy(h)


def test_bad_config_rejected(built):
    lib = built._lib.load()
    cfg = built.make_config(25, 4)
    cfg.num_nodes = 1
    h = C.c_void_p()
    assert lib.sndvae_create(C.byref(cfg), None, C.byref(h)) == -1
    assert b"bad config" in lib.sndvae_last_error(h)
    lib.sndvae_destroy(h)
    # the capacity / DIP losses exist for the 3-latent model only (optimizer.py:166-183 read z_mean_s / z_mean_g)
    cfg = built.make_config(25, 4, "base", loss_variant=1)
    h = C.c_void_p()
    assert lib.sndvae_create(C.byref(cfg), None, C.byref(h)) == -1
    assert b"loss_variant" in lib.sndvae_last_error(h)
    lib.sndvae_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_engine_fails_loudly_without_gpu(built):
    with pytest.raises(built.SndvaeError, match="no CPU fallback"):
        built.Engine(built.make_config(25, 4))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (tier rule 3)."""
    pkg = os.path.join(ROOT, "snd-vae_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dp, f)
