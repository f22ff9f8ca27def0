"""CPU restatement of the window front-end that feeds the hot path, and of the MultiCNNTransformer variants built on it.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Paths relative to /root/reference/transformer/.
Pinned by tests/golden/front_*.npz (outputs of the imported reference `models.py` classes, oracle/make_golden_frontend.py).
"""
import torch

from . import mt_oracle as O
from .dropout_rng import SITE_FRONT, Dropper

_NO_DROP = Dropper(None)
FIXED_EMBED = {'linguistic': 300, 'emotient': 20, 'acoustic': 256, 'image': 256}      # SFT/B2/B3 models.py:90


def cnn(sd, name, x):
    """CNN.forward MFT/models.py:68-79 on row-major windows.  x [n, K, D] (the reference receives the permuted [n, D, K] view,
    models.py:126); conv1d.weight [E, D, k].  y[n, f, j] = b[f] + sum_{d, i} w[f, d, i] * x[n, j + i, d]; max over every j
    (MaxPool1d(L, stride=3) with L = the full conv length yields exactly one output, :76-77)."""
    w, b = sd[name + '.conv1d.weight'], sd[name + '.conv1d.bias']
    E, D, k = w.shape
    n, K, _ = x.shape
    L = K - k + 1
    y = b.view(1, 1, E).expand(n, L, E).clone()
    for i in range(k):
        y = y + x[:, i:i + L, :] @ w[:, :, i].t()
    return y.max(dim=1).values


def highway(sd, name, c):
    """Highway.forward MFT/models.py:51-54."""
    proj = O.linear(sd, name + '.linear_projection', c)
    gate = torch.sigmoid(O.linear(sd, name + '.linear_gate', c))
    return gate * proj + (1 - gate) * c


def front(sd, inputs, mods, drop=_NO_DROP, p=0.3):
    """MultiCNNTransformer.forward MFT/models.py:117-132 without the per-narrative loop: every window is independent, so the B
    iterations of [T, K, D] are one pass over [B*T, K, D].  Dropout(0.3) :105,129 on the [B*T, E] embeddings."""
    out = {}
    for i, m in enumerate(mods):
        x = inputs[m]
        B, T, K, D = x.shape
        c = cnn(sd, f'cnn_{m}', x.reshape(B * T, K, D))
        e = highway(sd, f'highway_{m}', c)
        out[m] = drop(e, p, SITE_FRONT + i).reshape(B, T, -1)
    return out


def mcnn_mft(sd, inputs, mask, mods, N=6, drop=_NO_DROP, use_encoder=True):
    """MFT/models.py:111-136 (use_encoder=False: B3-MFN/models.py)."""
    emb = front(sd, inputs, mods, drop)
    return O.multi_transformer(sd, 'Transformer.', emb, mask, mods, N=N, drop=drop, use_encoder=use_encoder)


def mcnn_sft(sd, inputs, mask, mods, N=6, drop=_NO_DROP):
    """SFT/models.py:113-142."""
    emb = front(sd, inputs, mods, drop)
    return O.sft_hot_path(sd, [emb[m] for m in mods], mask, N=N, drop=drop)


def mcnn_b2(sd, inputs, mask, mods, N=6, drop=_NO_DROP):
    """B2-Trans/models.py:105-133."""
    emb = front(sd, inputs, mods, drop)
    return O.uni_full_transformer(sd, 'Transformer.', torch.cat([emb[m] for m in mods], 2), mask, N=N, drop=drop)


def mcnn_uni(sd, inputs, mask, mods, N=6, drop=_NO_DROP):
    """One modality: MFT/models.py:102-105,134-135 -- the front-end feeds UniTransformer."""
    emb = front(sd, inputs, mods, drop)
    return O.uni_transformer(sd, 'Transformer.', emb[mods[0]], mask, N=N, drop=drop)
