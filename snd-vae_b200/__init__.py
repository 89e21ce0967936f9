"""snd-vae_b200: B200-native (sm_100a) SND-VAE train / generate step behind the
reference's Python entry points.  Import with
`importlib.import_module("snd-vae_b200")` or through the `sndvae_b200` alias
module at the repository root (a dash is not valid in an `import` statement).
"""
from . import _lib                      # noqa: F401
from .engine import Engine, SndvaeError, make_config   # noqa: F401
