"""The reference's flag system (main.py:39-103, `tf.app.flags`) with the same
names and defaults, plus the per-dataset overrides main.py applies at run time
(main.py:136-172 synthetic1, 181-217 synthetic2) and the flag model_joint.py
reads but main.py never defines (`num_edge_feature`, model_joint.py:171; SURVEY
quirk Q3).  Unlike the reference, the model never mutates the flags (quirk Q5)."""
from __future__ import annotations

import copy

_DEFAULTS = dict(
    spatial_conv_layers=3, s_channel=[10, 10, 20], s_kernel_size=[5, 5, 5], s_strides=[1, 1, 1],
    s_hidden_size=100, s_latent_size=100,
    graph_conv_layers=2, g_conv_hidden=[10, 20], g_hidden_size=100, g_latent_size=100,
    spatial_graph_conv_layers=2, sg_conv_hidden=[[20, 20, 20], [50, 50, 50]], sg_hidden_size=200, sg_latent_size=200,
    spatial_deconv_layers=3, s_d_channel=[50, 20, 10], s_d_kernel_size=[5, 5, 5], s_d_strides=[1, 1, 1],
    graph_deconv_layers=2, n_d_channel=[50, 20, 10], n_d_kernel_size=[5, 5, 5], n_d_strides=[1, 1, 1],
    d_hidden_size=20, e_d_hidden=[50, 20, 10],
    node_h_size=20, model_type="disentangled",
    learning_rate=0.001, epochs=2000, dropout=1.0, batch_size=2, decoder_batch_size=2, sg_batch_size=5,
    sg_decoder_batch_size=5, dataset_path="../dataset/", num_feature=1, spatial_dim=2, verbose=1, test_count=10,
    model="feedback", seeded=1, connected_split=0, type="test_reconstruct", if_traverse=1, visualize_length=5,
    dataset="synthetic2", C_max=100.0, C_stop_iter=1e2, gamma=100.0, C_step=20.0, sampling_num=10, dim=None,
    group_type=None,
    num_edge_feature=2,          # model_joint.py:171 (undefined in main.py)
    use_tensor_cores=2, chunk_graphs=0,   # B200-side knobs (not in the reference)
)

_SYNTHETIC1 = dict(sg_hidden_size=500, sg_latent_size=500, node_h_size=50, learning_rate=0.001, epochs=1000, dropout=1.0,
                   batch_size=10, decoder_batch_size=10, sg_batch_size=10, sg_decoder_batch_size=10)
_SYNTHETIC2 = dict(sg_hidden_size=100, sg_latent_size=100, node_h_size=20, learning_rate=0.0008, epochs=1000, dropout=1.0,
                   batch_size=10, decoder_batch_size=10, sg_batch_size=10, sg_decoder_batch_size=10)


class _Flags:
    def __init__(self):
        self.reset()

    def reset(self):
        self.__dict__.update(copy.deepcopy(_DEFAULTS))

    def apply_dataset(self, dataset: str):
        """The run-time overrides of main.py:128-217."""
        self.dataset = dataset
        base = dict(spatial_conv_layers=3, s_channel=[10, 10, 20], s_kernel_size=[5, 5, 5], s_strides=[1, 1, 1], s_hidden_size=100,
                    s_latent_size=100, graph_conv_layers=2, g_conv_hidden=[10, 20], g_hidden_size=100, g_latent_size=100,
                    spatial_graph_conv_layers=2, sg_conv_hidden=[[20, 20, 20], [50, 50, 50]], spatial_deconv_layers=3,
                    s_d_channel=[50, 20, 10], graph_deconv_layers=2, n_d_channel=[50, 20, 10], d_hidden_size=20, e_d_hidden=[50, 20, 10])
        if dataset == "synthetic1":
            self.__dict__.update(copy.deepcopy(base)); self.__dict__.update(_SYNTHETIC1)
        elif dataset == "synthetic2":
            self.__dict__.update(copy.deepcopy(base)); self.__dict__.update(_SYNTHETIC2)
        elif dataset == "protein":         # main.py:218-236: 3-hop joint encoder (model.py:139-140), 3-D coordinates, small heads
            self.__dict__.update(copy.deepcopy(base))
            self.__dict__.update(spatial_dim=3, sg_conv_hidden=[[10, 10, 10, 10], [20, 20, 20, 20]], sg_hidden_size=50, sg_latent_size=50,
                                 s_hidden_size=5, s_latent_size=5, g_hidden_size=5, g_latent_size=5, node_h_size=5,
                                 s_channel=[10, 10, 20], s_kernel_size=[5, 5, 5], batch_size=50, decoder_batch_size=50,
                                 sg_batch_size=50, sg_decoder_batch_size=50)
        elif dataset == "mnist":           # main.py:237-241
            self.__dict__.update(copy.deepcopy(base))
            self.__dict__.update(spatial_dim=3, sg_conv_hidden=[[20, 20, 20, 20], [50, 50, 50, 50]])
        else:
            raise ValueError(f"dataset '{dataset}' is outside the hot path (SURVEY section 2, rows 11-13)")

    def flag_values_dict(self):
        return dict(self.__dict__)


FLAGS = _Flags()
