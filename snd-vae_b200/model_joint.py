"""`SGCNModelVAE` of model_joint.py (the single-latent "base" model, one `z_sg`) with
the reference's constructor signature (model_joint.py:14).  Coherent only with one
sample per graph: S is forced to 1."""
from .model import _ModelBase


class SGCNModelVAE(_ModelBase):
    model_type = "base"

    def __init__(self, placeholders, num_features, num_nodes, **kwargs):
        super().__init__(placeholders, num_features, num_nodes, **kwargs)

    def sample(self, z_sg):
        r = self.engine.generate(None, z_sg, None)
        return r["generated_adj"], r["generated_adj_prob"], r["generated_spatial"], r["generated_node_feat"]
