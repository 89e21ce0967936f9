"""ctypes binding of the C ABI in include/sndvae.h (libsndvae.so, built in-tree).

No CPU fallback: importing works without a GPU (so that the symbol table can be
checked), but every compute entry point fails loudly when the library or a B200
is missing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsndvae.so")

I32, I64, F32 = C.c_int32, C.c_int64, C.c_float
PF = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [
        ("model_type", I32), ("num_nodes", I32), ("num_feature", I32), ("spatial_dim", I32),
        ("sampling_num", I32), ("node_h_size", I32),
        ("s_channel", I32 * 3), ("s_hidden_size", I32), ("s_latent_size", I32),
        ("g_conv_hidden", I32 * 2), ("g_hidden_size", I32), ("g_latent_size", I32),
        ("sg_conv_hidden", (I32 * 3) * 2), ("sg_hidden_size", I32), ("sg_latent_size", I32),
        ("s_d_channel", I32 * 3), ("n_d_channel", I32 * 2), ("e_d_hidden", I32 * 2),
        ("batch_size", I32), ("chunk_graphs", I32), ("edge_capacity", I32), ("use_tensor_cores", I32),
        ("learning_rate", F32), ("beta", F32), ("adam_beta1", F32), ("adam_beta2", F32), ("adam_eps", F32),
        ("loss_variant", I32), ("gamma", F32), ("C_max", F32), ("C_stop_iter", F32), ("C_step", F32),
        ("dip_lambda_od", F32), ("dip_lambda_d", F32),
        ("sg_hops", I32), ("sg_conv_hidden3", (I32 * 4) * 2),
    ]


LOSS_VARIANTS = {"disentangled": 0, "base": 0, "disentangled_C": 1, "NED-VAE-IP": 2, "beta-TCVAE": 3}


class Inputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("features", "spatial", "adj", "rel", "adj_truth", "feature_truth",
                                            "spatial_truth", "rel_truth")]


class InputsCompact(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("features", "adj_bits", "rel", "adj_truth_bits", "feature_truth", "spatial_truth")]


class Noise(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("eps_s", "eps_sg", "eps_g")]


OUTPUT_FIELDS = ("z_mean_s", "z_std_s", "z_s", "z_mean_g", "z_std_g", "z_g", "z_mean_sg", "z_std_sg", "z_sg",
                 "generated_adj", "generated_adj_prob", "generated_spatial", "generated_node_feat")


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_FIELDS]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("offset", I64), ("size", I64), ("rank", I32), ("shape", I32 * 4)]


# every symbol include/sndvae.h declares: (restype, argtypes)
SYMBOLS = {
    "sndvae_default_config": (C.c_int, [C.POINTER(Config)]),
    "sndvae_create": (C.c_int, [C.POINTER(Config), C.c_void_p, C.POINTER(C.c_void_p)]),
    "sndvae_destroy": (C.c_int, [C.c_void_p]),
    "sndvae_last_error": (C.c_char_p, [C.c_void_p]),
    "sndvae_param_count": (I64, [C.c_void_p]),
    "sndvae_num_params": (I32, [C.c_void_p]),
    "sndvae_param_table": (C.c_int, [C.c_void_p, C.POINTER(ParamInfo), I32]),
    "sndvae_get_params": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sndvae_set_params": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sndvae_get_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sndvae_set_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sndvae_params_device": (C.c_void_p, [C.c_void_p]),
    "sndvae_grads_device": (C.c_void_p, [C.c_void_p]),
    "sndvae_forward": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.POINTER(Outputs), C.c_void_p]),
    "sndvae_grads": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.POINTER(Outputs), C.c_void_p, I64]),
    "sndvae_apply_adam": (C.c_int, [C.c_void_p]),
    "sndvae_zero_grads": (C.c_int, [C.c_void_p]),
    "sndvae_grads_accumulate": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.POINTER(Outputs), C.c_void_p, I64]),
    "sndvae_set_beta": (C.c_int, [C.c_void_p, F32]),
    "sndvae_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sndvae_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, I32, I32]),
    "sndvae_comm_info": (C.c_int, [C.c_void_p, C.POINTER(I32), C.POINTER(I32)]),
    "sndvae_allreduce_grads": (C.c_int, [C.c_void_p]),
    "sndvae_grads_host": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.c_void_p, C.c_void_p, I64, I32]),
    "sndvae_train_step": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.POINTER(Outputs), C.c_void_p]),
    "sndvae_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Outputs)]),
    "sndvae_train_step_host": (C.c_int, [C.c_void_p, C.POINTER(Inputs), C.POINTER(Noise), C.c_void_p, C.c_void_p]),
    "sndvae_train_step_host_compact": (C.c_int, [C.c_void_p, C.POINTER(InputsCompact), C.POINTER(Noise), C.c_void_p, C.c_void_p]),
    "sndvae_set_global_iter": (C.c_int, [C.c_void_p, I64]),
    "sndvae_synth_inputs": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(Inputs)]),
    "sndvae_launch_count": (I64, [C.c_void_p]),
    "sndvae_graph_replays": (I64, [C.c_void_p]),
    "sndvae_gemm_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(I64), C.POINTER(C.c_double)]),
    "sndvae_stage_times": (C.c_int, [C.c_void_p, I32, C.c_void_p, C.c_void_p, I32, C.POINTER(I64)]),
    "sndvae_debug_gemm": (C.c_int, [C.c_void_p, I32, I32, I64, I32, I32, F32, C.c_void_p, I64, C.c_void_p, I64, F32, C.c_void_p, I64, C.c_void_p]),
    "sndvae_threshold_logits": (C.c_int, [C.c_void_p, C.c_void_p, I64, C.c_void_p]),
    "sndvae_inner_product_decode": (C.c_int, [C.c_void_p, C.c_void_p, I64, I32, I32, C.c_void_p]),
    "sndvae_debug_read": (I64, [C.c_void_p, C.c_char_p, C.c_void_p, I64]),
}

_lib = None


def load():
    """dlopen libsndvae.so and bind every symbol; raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The SND-VAE hot path is CUDA-only (sm_100a); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
