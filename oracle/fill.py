"""Synthetic input / weight generators: re-exported from the product package (they are torch-free and
reference-free; the golden fixtures were generated with exactly these streams)."""
from multimodal_transformer_b200.synthetic import *          # noqa: F401,F403
from multimodal_transformer_b200.synthetic import _rs          # noqa: F401
