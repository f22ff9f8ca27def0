"""B200-native (sm_100a) implementation of the fusion-model hot path of frankaging/Multimodal-Transformer.

Import as `multimodal_transformer_b200` (the importable alias of this directory).
"""
from . import functional
from .functional import ccc_batched, fix_seed, get_compute_dtype, manual_seed, ragged_batch, set_compute_dtype, set_grouped_stacks, set_parallel_stacks
from .multiTransformer import *          # noqa: F401,F403  (the reference's class names)
from .multiTransformer import fusion_layer
from .models import *                    # noqa: F401,F403  (models.py: CNN, Highway, MultiCNNTransformer variants)
from .evaluation import evaluate          # noqa: F401
from .batching import DeviceCorpus, generateTrainBatch          # noqa: F401
