"""tcgen05 flash attention for long sequences / 64-wide heads (csrc/mt_attention_flash.cu) against fp64 torch on the same bf16 operands and
the same pair-hash dropout masks (oracle/dropout_rng.py: attn_keep_mask), and against the mma.sync tile kernels of the same library:
attention() of MFT/multiTransformer.py:22-34 with the query-ROW mask, live padded keys and dropout on the probabilities."""
import math

import pytest
import torch

from multimodal_transformer_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _reference(qkv, mask, h, p, seed, site):
    from oracle.dropout_rng import attn_keep_mask
    B, T, d3 = qkv.shape
    d = d3 // 3
    dk = d // h
    q, k, v = [t.double().view(B, T, h, dk).transpose(1, 2) for t in qkv.cpu().split(d, dim=-1)]
    s = q @ k.transpose(-1, -2) / math.sqrt(dk)
    if mask is not None:
        s = s.masked_fill(mask.cpu().view(B, 1, T, 1) == 0, -1e9)
    pr = torch.softmax(s, dim=-1)
    lse = torch.logsumexp(s, dim=-1)
    if mask is not None:                       # masked rows: every score is -1e9 -> uniform; the kernels report log(T) for them
        lse = torch.where(mask.cpu().view(B, 1, T) == 0, torch.full_like(lse, math.log(T)), lse)
    if p > 0:
        pr = pr * attn_keep_mask(seed, site, (B, h, T, T), p).double() / (1.0 - float(torch.tensor(p, dtype=torch.float32)))
    o = (pr @ v).transpose(1, 2).reshape(B, T, d)
    return o, lse


@pytest.mark.parametrize('B,T,p', [(2, 300, 0.0), (1, 513, 0.1), (3, 256, 0.1), (1, 1024, 0.0)])
def test_flash_forward_vs_fp64(B, T, p):
    L = _lib.lib()
    d, h = 512, 8
    g = torch.Generator().manual_seed(B * 100 + T)
    qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.7).bfloat16().to(DEV)
    mask = torch.ones(B, T)
    mask[0, T - 7:] = 0
    if B > 1:
        mask[1, 5] = 0
    mask = mask.to(DEV)
    seed, site = 99, 3
    want, want_lse = _reference(qkv, mask, h, p, seed, site)
    out = torch.full((B, T, d), float('nan'), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((B, h, T), float('nan'), device=DEV)
    _lib.check(L.mt_attention_fwd(1, B, T, d, h, _lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(out), _lib.ptr(lse), p, seed, site, None))
    torch.cuda.synchronize()
    err = (out.double().cpu() - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), err
    assert (lse.double().cpu() - want_lse).abs().max().item() <= 2e-3
    # the mma.sync tile kernels on the same inputs
    old = L.mt_attention_force_no_tc(1)
    try:
        out2 = torch.empty_like(out); lse2 = torch.empty_like(lse)
        _lib.check(L.mt_attention_fwd(1, B, T, d, h, _lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(out2), _lib.ptr(lse2), p, seed, site, None))
        torch.cuda.synchronize()
    finally:
        L.mt_attention_force_no_tc(old)
    assert (out.float() - out2.float()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    assert (lse - lse2).abs().max().item() <= 2e-3


@pytest.mark.parametrize('B,T,p', [(2, 300, 0.0), (1, 513, 0.1), (2, 256, 0.1), (1, 1024, 0.1)])
def test_flash_backward_vs_fp64_autograd(B, T, p):
    """dqkv of the tcgen05 flash backward (dK / dV accumulated in TMEM over the query tiles, dQ by fp32 reductions over the key tiles)
    against fp64 autograd through the reference formula with the same dropout masks, and against the mma.sync tile kernels."""
    from oracle.dropout_rng import attn_keep_mask
    L = _lib.lib()
    d, h = 512, 8
    dk = d // h
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.7).bfloat16()
    dout = torch.randn(B, T, d, generator=g).bfloat16()
    mask = torch.ones(B, T)
    mask[0, T - 9:] = 0
    mask[B - 1, 3] = 0
    seed, site = 123, 5
    x = qkv.double().requires_grad_(True)
    q, k, v = [t.view(B, T, h, dk).transpose(1, 2) for t in x.split(d, dim=-1)]
    s = (q @ k.transpose(-1, -2) / math.sqrt(dk)).masked_fill(mask.view(B, 1, T, 1) == 0, -1e9)
    pr = torch.softmax(s, dim=-1)
    if p > 0:
        pr = pr * attn_keep_mask(seed, site, (B, h, T, T), p).double() / (1.0 - float(torch.tensor(p, dtype=torch.float32)))
    o = (pr @ v).transpose(1, 2).reshape(B, T, d)
    o.backward(dout.double())
    want = x.grad
    qd, dd, md = qkv.to(DEV), dout.to(DEV), mask.to(DEV)
    res = []
    for force in (0, 1):
        old = L.mt_attention_force_no_tc(force)
        try:
            out = torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16); lse = torch.empty(B, h, T, device=DEV)
            _lib.check(L.mt_attention_fwd(1, B, T, d, h, _lib.ptr(qd), _lib.ptr(md), _lib.ptr(out), _lib.ptr(lse), p, seed, site, None))
            dqkv = torch.full((B, T, 3 * d), float('nan'), device=DEV, dtype=torch.bfloat16)
            ws = torch.empty(L.mt_attention_bwd_ws_bytes(B, T, h), dtype=torch.uint8, device=DEV)
            _lib.check(L.mt_attention_bwd(1, B, T, d, h, _lib.ptr(qd), _lib.ptr(md), _lib.ptr(out), _lib.ptr(lse), _lib.ptr(dd), _lib.ptr(dqkv), p, seed, site,
                                          _lib.ptr(ws), ws.numel(), None))
            torch.cuda.synchronize()
            res.append(dqkv.double().cpu())
        finally:
            L.mt_attention_force_no_tc(old)
    scale = want.abs().max().item()
    for name, got in (('flash', res[0]), ('mma tiles', res[1])):
        assert torch.isfinite(got).all(), name
        for sl, nm in ((slice(0, d), 'dq'), (slice(d, 2 * d), 'dk'), (slice(2 * d, 3 * d), 'dv')):
            e = (got[..., sl] - want[..., sl]).abs().max().item()
            assert e <= 3e-2 * max(want[..., sl].abs().max().item(), 1e-2 * scale), (name, nm, e)
