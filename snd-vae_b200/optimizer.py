"""`OptimizerVAE` with the reference's constructor signature (optimizer.py:124) and
attribute names (optimizer.py:144-203).  The ELBO, its backward pass and the TF1 Adam
update run inside the engine; this object only exposes the fetch handles.  `pos_weight`,
`norm` and `labels_rel` are accepted and unused; `global_iter` is read by
the 'disentangled_C' branch only (optimizer.py:172) and reaches the engine through the feed
dict.  Branches built: 'disentangled', 'base', 'disentangled_C', 'NED-VAE-IP', 'beta-TCVAE'."""
from .flags import FLAGS
from .session import Fetch


class OptimizerVAE(object):
    def __init__(self, preds_edge, preds_node, preds_spatial, labels_edge, labels_node, labels_spatial, labels_rel, model,
                 num_nodes, pos_weight, norm, beta, global_iter):
        if FLAGS.model_type not in ("disentangled", "base", "disentangled_C", "NED-VAE-IP", "beta-TCVAE"):
            raise ValueError(f"model_type '{FLAGS.model_type}' is not built (geoGCN, posGCN: SURVEY section 2 row 13)")
        self.model = model
        model.optimizer = self
        # beta weights the KL terms ('disentangled', 'base', 'beta-TCVAE': optimizer.py:164,190) or the DIP regulariser
        # ('NED-VAE-IP': optimizer.py:183); the reference hands it to this constructor (main.py:285-296,515)
        model.engine.set_beta(float(beta))
        for name in ("opt_op", "cost", "adj_cost", "node_cost", "spatial_cost", "kl_sg", "kl_s", "kl_g"):
            setattr(self, name, Fetch(self, name))
        if model.engine.dis:
            self.overall_loss = [self.cost, self.spatial_cost, self.adj_cost, self.node_cost, self.kl_g, self.kl_s, self.kl_sg]
        else:
            self.overall_loss = [self.cost, self.spatial_cost, self.adj_cost, self.node_cost, self.kl_sg]

    @property
    def grads_vars(self):
        """optimizer.compute_gradients(cost) (optimizer.py:198): (gradient, variable-name) pairs of
        the last backward pass, as views of the flat gradient arena."""
        flat = self.model.engine.grads_tensor()          # zero-copy view of the device arena; entries are slices of it
        out = []
        for name, off, shape in self.model.engine.table:
            n = 1
            for d in shape:
                n *= d
            out.append((flat[off:off + n].view(shape), name))
        return out
