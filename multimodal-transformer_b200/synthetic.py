"""Deterministic, version-stable synthetic SEND-shaped inputs and weight fills (numpy legacy RandomState streams
are frozen by numpy policy), shared by bench.py, the golden-vector script and the tests: torch-free and
reference-free, so the same numbers are regenerated on the GPU box without shipping multi-MB tensors.
Shapes follow SURVEY 8(d): features N(0,1) [B,T,88|256|300], sorted-descending lengths, float mask [B,T,1],
U(0,1) targets zero-padded (MFT/train.py:62-63,96,103-106,507-514)."""
import zlib

import numpy as np


def _rs(seed, key):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF)


def fill_array(key, shape, seed=1):
    """Weights: U(-a, a), a = 1/sqrt(fan_in) for matrices; LayerNorm gains near 1; biases small."""
    shape = tuple(int(s) for s in shape)
    rs = _rs(seed, key)
    if key.endswith('a_2'):
        return (1.0 + 0.1 * rs.uniform(-1, 1, shape)).astype(np.float32)
    if key.endswith('b_2') or 'bias' in key or key.endswith('dec_h0') or key.endswith('dec_c0'):
        return (0.1 * rs.uniform(-1, 1, shape)).astype(np.float32)
    fan_in = shape[-1] if len(shape) >= 2 else shape[0]
    a = 1.0 / np.sqrt(fan_in)
    return rs.uniform(-a, a, shape).astype(np.float32)


def fill_state(shapes, seed=1):
    """shapes: dict key -> shape.  Returns dict key -> float32 ndarray."""
    return {k: fill_array(k, s, seed) for k, s in shapes.items()}


def make_lengths(B, T, seed=1):
    """Sorted-descending lengths, lengths[0] = T, others U[T/4, T]  (MFT/train.py:62-63,96)."""
    rs = _rs(seed, 'lengths')
    lo = max(1, T // 4)
    ls = [T] + [int(v) for v in rs.randint(lo, T + 1, size=B - 1)]
    ls.sort(reverse=True)
    return ls


def make_batch(B, T, dims, seed=1):
    """dims: dict mod -> feature width.  Returns (inputs dict [B,T,D] f32, mask [B,T,1] f32,
    target [B,T,1] f32 zero-padded, lengths list)."""
    lengths = make_lengths(B, T, seed)
    inputs = {m: _rs(seed, 'in_' + m).standard_normal((B, T, d)).astype(np.float32) for m, d in dims.items()}
    mask = np.zeros((B, T, 1), np.float32)
    for b, l in enumerate(lengths):
        mask[b, :l] = 1.0
    target = _rs(seed, 'target').uniform(0, 1, (B, T, 1)).astype(np.float32) * mask
    return inputs, mask, target, lengths


RAW_SHAPES = {'linguistic': (33, 300), 'acoustic': (2, 88), 'image': (2, 1000), 'emotient': (2, 20)}   # (K vectors, D) per window


def make_raw_batch(B, T, shapes, seed=1):
    """Raw-level SEND-shaped batch in front of the window CNNs (MFT/train.py:571): shapes: mod -> (K, D).
    inputs[mod] [B,T,K,D] f32, zero rows for padded windows; windows with K > 2 carry a random number (>= 2) of leading
    non-zero vectors, the rest zero (short windows are zero-padded to the longest one).  Returns (inputs, mask, target, lengths)."""
    lengths = make_lengths(B, T, seed)
    mask = np.zeros((B, T, 1), np.float32)
    for b, l in enumerate(lengths):
        mask[b, :l] = 1.0
    inputs = {}
    for m, (K, D) in shapes.items():
        x = _rs(seed, 'raw_' + m).standard_normal((B, T, K, D)).astype(np.float32)
        if K > 2:
            nvec = _rs(seed, 'nvec_' + m).randint(2, K + 1, size=(B, T))
            x *= (np.arange(K)[None, None, :] < nvec[:, :, None]).astype(np.float32)[..., None]
        inputs[m] = x * mask[..., None]
    target = _rs(seed, 'target').uniform(0, 1, (B, T, 1)).astype(np.float32) * mask
    return inputs, mask, target, lengths
