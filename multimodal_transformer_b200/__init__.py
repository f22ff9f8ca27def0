"""Importable alias of the `multimodal-transformer_b200/` package directory (a hyphen is not a valid module name)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'multimodal-transformer_b200'))
from .functional import fix_seed, get_compute_dtype, manual_seed, set_compute_dtype, set_parallel_stacks          # noqa: E402,F401
from . import functional, multiTransformer                                                    # noqa: E402,F401
from .multiTransformer import *                                                              # noqa: E402,F401,F403
from .multiTransformer import fusion_layer                                                    # noqa: E402,F401
