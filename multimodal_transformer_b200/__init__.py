"""Importable alias of the `multimodal-transformer_b200/` package directory (a hyphen is not a valid module name)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'multimodal-transformer_b200'))
from .functional import (ccc_batched, fix_seed, get_compute_dtype, manual_seed, ragged_batch, set_compute_dtype,   # noqa: E402,F401
                         set_grouped_stacks, set_parallel_stacks)
from . import functional, multiTransformer                                                    # noqa: E402,F401
from .multiTransformer import *                                                              # noqa: E402,F401,F403
from .multiTransformer import fusion_layer                                                    # noqa: E402,F401
from . import models, evaluation, batching                                                              # noqa: E402,F401
from .models import *                                                                        # noqa: E402,F401,F403
from .evaluation import evaluate                                                              # noqa: E402,F401
from .batching import DeviceCorpus, generateTrainBatch                                        # noqa: E402,F401
