"""Cluster / tensor-core forward of the step-wise LSTM decoder (csrc/mt_lstm_head_mma.cu; SFT/multiTransformer.py:465-483) against the FFMA
kernel of the same library on the same bf16 weights: every cluster shape (8 / 16 / 32 narratives per cluster, ragged last tile), forward
values and -- through the shared backward kernel, which consumes the stash the forward wrote -- every parameter gradient.  Parity with
the fp64 oracle is tests/test_gpu_parity.py::test_bf16_mode_other_models_within_2e2 and test_sft_golden."""
import numpy as np
import pytest
import torch

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('B,T', [(3, 5), (25, 33), (150, 12), (300, 9)])
def test_decoder_cluster_forward_matches_ffma_kernel(B, T):
    mtb.set_compute_dtype('bf16')
    try:
        torch.manual_seed(B * 7 + T)
        model = mtb.NLPTransformer(64, N=1).eval()
        with torch.no_grad():
            model.dec_h0.normal_(0, 0.3); model.dec_c0.normal_(0, 0.3)
        x = torch.randn(B, T, 64, device=DEV)
        lengths = [T - (i % 3) for i in range(B)]
        mask = torch.zeros(B, T, 1, device=DEV)
        for b, l in enumerate(lengths):
            mask[b, :l] = 1
        target = torch.rand(B, T, 1, device=DEV) * mask
        res = []
        for force in (0, 1):
            old = _lib.lib().mt_lstm_head_force_ffma(force)
            try:
                model.zero_grad()
                pred = model(x, mask, lengths)
                (((pred - target) ** 2).sum() / sum(lengths)).backward()
                res.append((pred.detach().float().cpu(), {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}))
            finally:
                _lib.lib().mt_lstm_head_force_ffma(old)
        (p_new, g_new), (p_old, g_old) = res
        assert torch.isfinite(p_new).all()
        assert (p_new * (1 - mask.cpu())).abs().max().item() == 0.0
        assert (p_new - p_old).abs().max().item() <= 2e-2 * max(1.0, p_old.abs().max().item())
        gmax = max(v.abs().max().item() for v in g_old.values())
        for k in g_old:
            e = (g_new[k] - g_old[k]).abs().max().item()
            assert e <= 6e-2 * g_old[k].abs().max().item() + 4e-3 * gmax, (k, e)
    finally:
        mtb.set_compute_dtype('fp32')
