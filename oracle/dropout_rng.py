"""CPU mirror of the counter-based dropout generator used by the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws dropout masks from PyTorch's Philox stream
(``nn.Dropout`` at MFT/multiTransformer.py:17,45,101,158-174), which a fused
kernel cannot reproduce bit-for-bit.  The CUDA path therefore defines its own
stateless generator -- one 32-bit ``hash(seed, site, element_index >> 1)`` per PAIR of elements, 16 bits each,
``keep = half >= (p * 2**32) >> 16`` --
and train-mode parity is "same result given the same masks": this file
restates that generator (csrc/mt_common.cuh: mt_draw32) with torch int64
arithmetic so the oracle can apply *identical* masks.
"""
import torch

_M32 = 0xFFFFFFFF


def _mix32(x):
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def draw32(seed, site, pair):
    """csrc/mt_common.cuh: mt_drop_resolve (key) + mt_draw32.  pair: int64 tensor of PAIR indices (>= 0).
    Returns int64 tensor in [0, 2**32): low 16 bits decide the even element of the pair, high 16 bits the odd one."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    seed_lo, seed_hi = seed & _M32, (seed >> 32) & _M32
    key0 = seed_lo ^ ((0x9E3779B9 * (int(site) + 1)) & _M32)
    key = (int(_mix32(torch.tensor(key0, dtype=torch.int64))) + seed_hi * 0x85EBCA6B) & _M32
    lo = pair & _M32
    hi = (pair >> 32) & _M32
    return _mix32(lo ^ key ^ ((hi * 0xC2B2AE35) & _M32))


def threshold(p):
    import numpy as np
    t = int(float(np.float32(p)) * 4294967296.0)      # the kernels receive p as a C float
    return max(0, min(t, _M32))


def keep_mask(seed, site, shape, p):
    """Boolean keep-mask of `shape` (row-major element index e: pair e >> 1, low half for even e)."""
    n = 1
    for s in shape:
        n *= int(s)
    idx = torch.arange(n, dtype=torch.int64)
    bits = draw32(seed, site, idx >> 1)
    v = torch.where((idx & 1) == 1, (bits >> 16) & 0xFFFF, bits & 0xFFFF)
    return (v >= (threshold(p) >> 16)).reshape(tuple(shape))


def attn_keep_mask(seed, site, shape, p):
    """Keep-mask of attention probabilities [B,h,T,T] (csrc/mt_common.cuh: mt_attn_drop_factor): one 32-bit draw per PAIR of
    adjacent keys of a query row -- pair index row * ceil(T/2) + (j >> 1), even key = low 16 bits, odd key = high 16 bits,
    threshold = threshold(p) >> 16."""
    *lead, T, Tk = [int(s) for s in shape]
    rows = 1
    for s in lead:
        rows *= s
    rows *= T
    P2 = (Tk + 1) // 2
    idx = torch.arange(rows * P2, dtype=torch.int64)
    bits = draw32(seed, site, idx).reshape(rows, P2)
    lo, hi = bits & 0xFFFF, (bits >> 16) & 0xFFFF
    both = torch.stack([lo, hi], dim=-1).reshape(rows, 2 * P2)[:, :Tk]
    return (both >= (threshold(p) >> 16)).reshape(tuple(shape))


class Dropper:
    """Applies dropout exactly as the CUDA kernels do.

    seed=None  -> identity (eval mode / p = 0)
    """

    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, x, p, site):
        if self.seed is None or p <= 0.0:
            return x
        keep = keep_mask(self.seed, site, x.shape, p).to(x.device)
        return x * keep.to(x.dtype) * (1.0 / (1.0 - p))

    def attn(self, p_attn, p, site):
        """Dropout on attention probabilities [B,h,T,T] (pair generator, see attn_keep_mask)."""
        if self.seed is None or p <= 0.0:
            return p_attn
        keep = attn_keep_mask(self.seed, site, p_attn.shape, p).to(p_attn.device)
        return p_attn * keep.to(p_attn.dtype) * (1.0 / (1.0 - p))


# ---- site ids (shared with csrc/mt_common.cuh) -------------------------------------------
# encoder stack `s`, layer `l`:  base = (s * 64 + l) * 8
SITE_ATTN_P = 0      # attention probabilities  [B, h, T, T]
SITE_SUB0 = 1        # sublayer-0 output        [B*T, d]
SITE_FFN_H = 2       # FFN hidden               [B*T, d_ff]
SITE_SUB1 = 3        # sublayer-1 output        [B*T, d]
# MFN
SITE_MFN_G1 = 0x4000  # gamma1 hidden  [T, B, 64]
SITE_MFN_G2 = 0x4001  # gamma2 hidden  [T, B, 64]
SITE_MFN_OUT = 0x4002  # out hidden    [T, B, 64]
# SFT embed input dropout [B*T, in]
SITE_SFT_EMBED = 0x5000


def enc_site(stack, layer, k):
    return (stack * 64 + layer) * 8 + k
# window front-end output dropout [B*T, E], one site per modality (index in the mods list)
SITE_FRONT = 0x6000
