"""CPU restatement of the reference batcher: generateInputChunkHelper / generateTrainBatch, MFT/train.py:59-108.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/batcher.json: the batches produced by the reference's own,
unmodified function bodies (extracted from MFT/train.py with `ast` -- importing that file would start logging to disk -- and
executed by oracle/make_golden_batcher.py in the build container).
"""
from random import shuffle

import numpy as np


def generate_train_batch(input_data, input_target, input_length, batch_size=25, on_eval=False):
    """Yields (data dict mod -> float32 array [B,T,K,D], target [B,T,1], mask [B,T,1], lengths list)."""
    n = len(input_data[list(input_data.keys())[0]])                          # :78
    index = list(range(n))                                                   # :79
    if not on_eval:
        shuffle(index)                                                       # :80-81
    for c0 in range(0, n, batch_size):                                       # chunks(), :52-55,82
        chunk = index[c0:c0 + batch_size]
        length_chunk = [input_length[i] for i in chunk]                      # :89
        max_length = max(length_chunk)                                       # :91
        order = sorted(range(len(chunk)), key=lambda i: length_chunk[i], reverse=True)      # :61-62 (stable, longest first)
        data = {}
        for mod in input_data:                                               # :93-99
            rows = [np.asarray(input_data[mod][chunk[i]], dtype=np.float32) for i in order]
            data[mod] = np.stack(rows)[:, :max_length]
        target = np.stack([np.asarray(input_target[chunk[i]], dtype=np.float32) for i in order])[:, :max_length, None]   # :101-103,108
        lengths = sorted(length_chunk, reverse=True)                         # :106
        mask = np.zeros((len(chunk), max_length, 1), np.float32)             # :105
        for i, l in enumerate(lengths):
            mask[i, :l] = 1                                                  # :107-108
        yield data, target, mask, lengths
