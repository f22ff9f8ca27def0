"""CPU: the oracle restatement reproduces the imported reference (tests/golden, made by
oracle/make_golden.py) -- forward to fp32 round-off, gradients via digests -- and its hand-derived
backward equations match autograd in fp64."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fill, mt_oracle as O
from oracle.ccc import eval_ccc
from tests import util

MODS = ['acoustic', 'image', 'linguistic']


def t(x):
    return torch.from_numpy(x)


def test_ccc_known_answers():
    with open(os.path.join(util.GOLD, 'ccc_kat.json')) as f:
        kat = json.load(f)
    assert len(kat) == 4
    for k in kat:
        assert abs(eval_ccc(k['actual'], k['pred']) - k['ccc']) < 2e-7, (k['model'], k['vid'])   # CSV traces are rounded float32 prints


def test_layer_norm():
    g = util.gold('ln')
    sd = util.filled_sd({'a_2': (256,), 'b_2': (256,)}, 3)
    x = t(fill.fill_array('ln_x', (3, 5, 256), 3)) * 20.0 + 1.5
    y = O.layer_norm(x, sd['a_2'], sd['b_2'])
    np.testing.assert_allclose(y.numpy(), g['y'], rtol=1e-5, atol=1e-5)
    # the "standard" LayerNorm (biased variance, eps inside sqrt) is NOT the reference's
    y_std = torch.nn.functional.layer_norm(x, (256,), sd['a_2'], sd['b_2'], 1e-6)
    assert (y_std - t(g['y'])).abs().max() > 1e-3


def test_mha_row_mask():
    g = util.gold('mha'); m = util.meta()['mha']
    shapes = {f'linears.{i}.{p}': s for i in range(4) for p, s in (('weight', (256, 256)), ('bias', (256,)))}
    sd = util.filled_sd(shapes, 4)
    inputs, mask, _, lengths = fill.make_batch(3, 7, {'x': 256}, 4)
    assert lengths == m['lengths']
    sd2 = {'a.' + k: v for k, v in sd.items()}
    y = O.mha(sd2, 'a', t(inputs['x']), t(mask), 8)
    np.testing.assert_allclose(y.numpy(), g['y'], rtol=2e-5, atol=2e-6)
    # padded query rows are uniform 1/T over ALL keys
    for b, l in enumerate(lengths):
        if l < 7:
            np.testing.assert_allclose(g['attn'][b, :, l:, :], 1.0 / 7, rtol=1e-6)


def _enc_shapes(N):
    return util.strip_prefix(util.mods_shapes('MFT.MultiTransformer', N), 'transformer_acoustic.')


def test_encoder_fwd_bwd():
    g = util.gold('encoder')
    sd = util.filled_sd(_enc_shapes(2), 5, requires_grad=True)
    inputs, mask, _, _ = fill.make_batch(3, 9, {'x': 256}, 5)
    x = t(inputs['x']).requires_grad_(True)
    y = O.encoder({'e.' + k: v for k, v in sd.items()}, 'e', x, t(mask), 2, 8)
    np.testing.assert_allclose(y.detach().numpy(), g['y'], rtol=2e-5, atol=2e-5)
    w = t(fill.fill_array('enc_w', (3, 9, 256), 5))
    (y * w).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g['dx'], rtol=1e-4, atol=1e-5)
    n = 0
    for k, v in sd.items():
        if 'grad:' + k in g:
            util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 1e-4, k); n += 1
    assert n == 2 * 16 + 2


def test_mfn_fwd_bwd():
    g = util.gold('mfn')
    shapes = util.strip_prefix(util.mods_shapes('MFT.MultiTransformer'), 'mfn.')
    sd = util.filled_sd(shapes, 6, requires_grad=True)
    inputs, _, _, _ = fill.make_batch(3, 6, {m: 256 for m in MODS}, 6)
    xin = {m: t(inputs[m]).permute(1, 0, 2).contiguous().requires_grad_(True) for m in MODS}
    y = O.mfn({'mfn.' + k: v for k, v in sd.items()}, 'mfn', xin, MODS)
    np.testing.assert_allclose(y.detach().numpy(), g['y'], rtol=2e-5, atol=2e-6)
    w = t(fill.fill_array('mfn_w', (3, 6, 1), 6))
    (y * w).sum().backward()
    for m in MODS:
        np.testing.assert_allclose(xin[m].grad.numpy(), g['dx_' + m], rtol=1e-4, atol=1e-7)
    for k, v in sd.items():
        util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 1e-4, k)


@pytest.mark.parametrize('name,inv,use_enc', [('mft_n2', 'MFT.MultiTransformer', True),
                                              ('mft_n6', 'MFT.MultiTransformer', True),
                                              ('b3', 'B3.MultiTransformer', False)])
def test_multitransformer(name, inv, use_enc):
    g = util.gold(name); m = util.meta()[name]
    N = m.get('N', 6)
    shapes = util.mods_shapes(inv, N)
    # the B3 inventory was taken with its own input widths
    sd = util.filled_sd(shapes, m['seed'], requires_grad=True)
    inputs, mask, target, lengths = fill.make_batch(m['B'], m['T'], m['dims'], m['seed'])
    assert lengths == m['lengths']
    pred = O.multi_transformer(sd, '', {k: t(v) for k, v in inputs.items()}, t(mask), MODS, N=N, use_encoder=use_enc)
    np.testing.assert_allclose(pred.detach().numpy(), g['pred'], rtol=1e-4, atol=2e-6)
    if 'loss' in g:
        loss = O.train_loss(pred, t(target), lengths)
        assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
        loss.backward()
        checked = 0
        for k, v in sd.items():
            if 'grad:' + k in g:
                util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 2e-4, k); checked += 1
            else:
                # orphan templates attn{mod}/ff{mod} never receive a gradient
                assert v.grad is None and (k.startswith('attn') or k.startswith('ff')), k
        assert checked > 20


def test_sft_path():
    g = util.gold('sft'); m = util.meta()['sft']
    shapes = {'Transformer.' + k: v for k, v in util.mods_shapes('SFT.NLPTransformer', 2).items()}
    shapes.update({'fusionLayer.weight': (512, 556), 'fusionLayer.bias': (512,)})
    sd = util.filled_sd(shapes, 10, requires_grad=True)
    inputs, mask, target, lengths = fill.make_batch(3, 8, m['dims'], 10)
    pred = O.sft_hot_path(sd, [t(inputs['image']), t(inputs['linguistic'])], t(mask), N=2)
    np.testing.assert_allclose(pred.detach().numpy(), g['pred'], rtol=1e-4, atol=2e-6)
    loss = O.train_loss(pred, t(target), lengths); loss.backward()
    assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    for k, v in sd.items():
        util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 2e-4, k)


@pytest.mark.parametrize('name,fn,inv,fin', [('unifull', O.uni_full_transformer, 'MFT.UniFullTransformer', 556),
                                             ('uni', O.uni_transformer, 'MFT.UniTransformer', 300)])
def test_uni(name, fn, inv, fin):
    g = util.gold(name); m = util.meta()[name]
    sd = util.filled_sd(util.mods_shapes(inv, 2), m['seed'], requires_grad=True)
    inputs, mask, target, lengths = fill.make_batch(3, 8, {'x': fin}, m['seed'])
    pred = fn(sd, '', t(inputs['x']), t(mask), N=2)
    np.testing.assert_allclose(pred.detach().numpy(), g['pred'], rtol=1e-4, atol=2e-6)
    loss = O.train_loss(pred, t(target), lengths); loss.backward()
    for k, v in sd.items():
        util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 2e-4, k)


# ---- hand-derived backward equations vs autograd (fp64) ------------------------------------------

def test_manual_layer_norm_bwd():
    torch.manual_seed(0)
    x = torch.randn(4, 6, 64, dtype=torch.float64, requires_grad=True)
    a = torch.randn(64, dtype=torch.float64, requires_grad=True); b = torch.randn(64, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(4, 6, 64, dtype=torch.float64)
    O.layer_norm(x, a, b).backward(dy)
    dx, da, db = O.layer_norm_bwd(x.detach(), a.detach(), dy)
    assert (dx - x.grad).abs().max() < 1e-12 and (da - a.grad).abs().max() < 1e-12 and (db - b.grad).abs().max() < 1e-12


def test_manual_attention_bwd():
    torch.manual_seed(0)
    q, k, v = [torch.randn(2, 3, 7, 8, dtype=torch.float64, requires_grad=True) for _ in range(3)]
    mask = torch.ones(2, 1, 7, 1, dtype=torch.float64); mask[1, :, 4:] = 0
    do = torch.randn(2, 3, 7, 8, dtype=torch.float64)
    out, _ = O.attention(q, k, v, mask); out.backward(do)
    dq, dk, dv = O.attention_bwd(q.detach(), k.detach(), v.detach(), mask, do)
    for a, b in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        assert (a - b).abs().max() < 1e-12
    assert dq[1, :, 4:].abs().max() == 0       # masked query rows get no score gradient


def test_manual_mfn_bwd():
    shapes = util.strip_prefix(util.mods_shapes('MFT.MultiTransformer'), 'mfn.')
    sd = util.filled_sd({'mfn.' + k: v for k, v in shapes.items()}, 2, dtype=torch.float64, requires_grad=True)
    inputs, _, _, _ = fill.make_batch(2, 5, {m: 256 for m in MODS}, 2)
    xin = {m: t(inputs[m]).double().permute(1, 0, 2).contiguous().requires_grad_(True) for m in MODS}
    dout = t(fill.fill_array('d', (2, 5, 1), 2)).double()
    y = O.mfn(sd, 'mfn', xin, MODS); y.backward(dout)
    with torch.no_grad():
        y2, dX, G = O.mfn_fwd_bwd_manual({k: v.detach() for k, v in sd.items()}, 'mfn',
                                         {m: v.detach() for m, v in xin.items()}, MODS, dout)
    assert (y2 - y).abs().max() < 1e-12
    for m in MODS:
        assert (dX[m] - xin[m].grad).abs().max() < 1e-12, m
    for k, v in sd.items():
        assert (G[k] - v.grad).abs().max() < 1e-11, k


def test_dropper_statistics_and_determinism():
    from oracle.dropout_rng import keep_mask
    m1 = keep_mask(1234, 7, (512, 512), 0.1); m2 = keep_mask(1234, 7, (512, 512), 0.1)
    assert torch.equal(m1, m2)
    assert abs(m1.float().mean().item() - 0.9) < 3e-3
    assert not torch.equal(m1, keep_mask(1235, 7, (512, 512), 0.1))
    assert not torch.equal(m1, keep_mask(1234, 8, (512, 512), 0.1))
    assert keep_mask(5, 1, (100,), 0.0).all()


# ---- structural properties of the restated path (what the full-size GPU tests rely on) ------------------------------------------------
def _small_mft(N=1, seed=51):
    dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', N), seed)
    return dims, sd


def test_oracle_narratives_are_independent_units():
    """No cross-sample operator on the path: permuting the batch permutes the predictions, and a narrative forwarded
    alone at the same padded length gives the same prediction -- the basis of sharding narratives over GPUs without a collective."""
    dims, sd = _small_mft()
    inputs, mask, _, lengths = fill.make_batch(4, 9, dims, 52)
    x = {k: t(v) for k, v in inputs.items()}
    with torch.no_grad():
        full = O.multi_transformer(sd, '', x, t(mask), MODS, N=1)
        perm = torch.tensor([2, 0, 3, 1])
        p = O.multi_transformer(sd, '', {k: v[perm] for k, v in x.items()}, t(mask)[perm], MODS, N=1)
        one = O.multi_transformer(sd, '', {k: v[1:2] for k, v in x.items()}, t(mask)[1:2], MODS, N=1)
    np.testing.assert_allclose(p.numpy(), full[perm].numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(one.numpy(), full[1:2].numpy(), rtol=1e-5, atol=1e-6)


def test_oracle_padded_windows_are_live_keys_but_mfn_is_causal():
    """Appendix A.12 and A.5: perturbing a PADDED window changes valid predictions (padded rows are attention keys), while the
    B3 variant (no encoder: LSTHM + memory recurrence only) is causal -- a change at window t0 leaves predictions before t0 untouched."""
    dims, sd = _small_mft()
    inputs, mask, _, lengths = fill.make_batch(3, 10, dims, 53)
    b = max(range(3), key=lambda i: -lengths[i] if lengths[i] < 10 else -99)       # a narrative with padding
    assert lengths[b] < 10
    x = {k: t(v).clone() for k, v in inputs.items()}
    x2 = {k: v.clone() for k, v in x.items()}
    x2['image'][b, lengths[b]:] += 1.0                                           # touch padded windows only
    with torch.no_grad():
        y, y2 = (O.multi_transformer(sd, '', xi, t(mask), MODS, N=1) for xi in (x, x2))
    assert (y[b, :lengths[b]] - y2[b, :lengths[b]]).abs().max() > 1e-6
    others = [i for i in range(3) if i != b]
    np.testing.assert_allclose(y[others].numpy(), y2[others].numpy(), rtol=0, atol=0)
    b3dims = {'acoustic': 256, 'image': 256, 'linguistic': 300}
    sd3 = util.filled_sd(util.mods_shapes('B3.MultiTransformer'), 54)
    inputs, mask, _, lengths = fill.make_batch(2, 8, b3dims, 54)
    x = {k: t(v).clone() for k, v in inputs.items()}
    x2 = {k: v.clone() for k, v in x.items()}
    t0 = 5
    x2['linguistic'][:, t0:] += 0.5
    with torch.no_grad():
        y, y2 = (O.multi_transformer(sd3, '', xi, torch.ones(2, 8, 1), MODS, use_encoder=False) for xi in (x, x2))
    np.testing.assert_allclose(y[:, :t0].numpy(), y2[:, :t0].numpy(), rtol=0, atol=0)
    assert (y[:, t0:] - y2[:, t0:]).abs().max() > 1e-6


def test_oracle_loss_gradients_sum_over_shards():
    """The reference loss is a SUM over narratives divided by the global sum of lengths (MFT/train.py:135-139): gradients of two shards
    normalised by the GLOBAL norm add up to the full-batch gradient -- why data parallelism all-reduces with SUM, not mean."""
    dims, sd0 = _small_mft()
    inputs, mask, target, lengths = fill.make_batch(4, 7, dims, 55)

    def grads(rows):
        sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
        pred = O.multi_transformer(sd, '', {k: t(v)[rows] for k, v in inputs.items()}, t(mask)[rows], MODS, N=1)
        (((pred - t(target)[rows]) ** 2).sum() / float(sum(lengths))).backward()
        return {k: v.grad for k, v in sd.items() if v.grad is not None}

    full, a, b = grads([0, 1, 2, 3]), grads([0, 2]), grads([1, 3])
    for k in full:
        np.testing.assert_allclose((a[k] + b[k]).numpy(), full[k].numpy(), rtol=2e-4, atol=1e-7, err_msg=k)
