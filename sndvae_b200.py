"""Importable alias of the `snd-vae_b200` package (the dash blocks `import`)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("snd-vae_b200")
