"""Shared helpers for the test-suite: golden loading, deterministic state dicts, digests."""
import json
import os

import numpy as np
import torch

from oracle import fill

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def meta():
    with open(os.path.join(GOLD, 'meta.json')) as f:
        return json.load(f)


def key_inventory():
    with open(os.path.join(GOLD, 'state_dict_keys.json')) as f:
        return json.load(f)


def filled_sd(shapes, seed, dtype=torch.float32, device='cpu', requires_grad=False):
    sd = {}
    for k, v in fill.fill_state(shapes, seed).items():
        t = torch.from_numpy(v).to(dtype).to(device)
        if requires_grad:
            t.requires_grad_(True)
        sd[k] = t
    return sd


def grad_digest(g):
    g = g.detach().double().cpu().reshape(-1)
    n = g.numel()
    idx = torch.linspace(0, n - 1, steps=min(n, 16)).long()
    return np.concatenate([[g.norm().item(), g.sum().item()], g[idx].numpy()])


def assert_digest_close(got, want, rtol, name='', atol=1e-6):
    """Compare gradient digests: norm (relative), sum and samples (relative to the norm)."""
    # atol: some gradients are analytically zero (e.g. the key-projection bias: softmax is invariant
    # to a per-query constant), so both sides hold only fp32 round-off noise there.
    scale = max(abs(want[0]), 1e-12)
    assert abs(got[0] - want[0]) <= rtol * scale + atol, f'{name}: norm {got[0]} vs {want[0]}'
    err = np.abs(got[1:] - want[1:]).max()
    # the sum over n elements accumulates ~sqrt(n) rounding; samples are bounded by the norm
    assert err <= 50 * rtol * scale + 10 * atol, f'{name}: digest err {err} (norm {scale})'


def strip_prefix(shapes, prefix):
    return {k[len(prefix):]: v for k, v in shapes.items() if k.startswith(prefix)}


def mods_shapes(inv_name, N=None):
    """Shapes of a reference model's state_dict, optionally truncated to N encoder layers."""
    inv = key_inventory()[inv_name]
    out = {}
    for k, s in inv.items():
        if N is not None and '.layers.' in k:
            l = int(k.split('.layers.')[1].split('.')[0])
            if l >= N:
                continue
        out[k] = tuple(s)
    return out
