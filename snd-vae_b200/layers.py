"""Host mirror of the one layers.py class that is offered outside the fused step: `InnerProductDecoder`
(layers.py:400-410).  The reference's models never instantiate it (model.py / model_joint.py decode edges with the e2e
layers), so it is a standalone operator on an existing engine, not a part of the train step."""


class InnerProductDecoder(object):
    """Decoder model layer for link prediction: `outputs = inputs . inputs^T` per graph.  Constructor arguments as in
    layers.py:402 (`input_dim`, `dropout`, `act` are accepted; the reference's `_call` applies neither dropout nor `act`)."""

    def __init__(self, input_dim, engine, dropout=0., act=None, **kwargs):
        self.input_dim = input_dim
        self.engine = engine
        self.dropout = dropout
        self.act = act

    def __call__(self, inputs):
        if inputs.shape[-1] != self.input_dim:
            raise ValueError(f"InnerProductDecoder built for input_dim={self.input_dim}, got {inputs.shape[-1]}")
        return self.engine.inner_product_decode(inputs)
