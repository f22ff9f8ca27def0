"""B200-native (sm_100a) implementation of the fusion-model hot path of frankaging/Multimodal-Transformer.

Import as `multimodal_transformer_b200` (the importable alias of this directory).
"""
from . import functional
from .functional import fix_seed, get_compute_dtype, manual_seed, set_compute_dtype, set_parallel_stacks
from .multiTransformer import *          # noqa: F401,F403  (the reference's class names)
from .multiTransformer import fusion_layer
