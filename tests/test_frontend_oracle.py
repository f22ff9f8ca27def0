"""CPU: the front-end oracle (oracle/frontend_oracle.py) reproduces the imported reference `models.py` classes
(tests/golden/front_*.npz, made by oracle/make_golden_frontend.py), and the product's MultiCNNTransformer mirrors keep the
reference's state_dict keys / shapes / constructor signatures (no compute calls: there is no GPU here)."""
import json
import os

import numpy as np
import pytest
import torch

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import models as M
from oracle import fill, frontend_oracle as FO, mt_oracle as O
from tests import util


def t(x):
    return torch.from_numpy(x)


def front_meta():
    with open(os.path.join(util.GOLD, 'front_meta.json')) as f:
        return json.load(f)


def front_inventory():
    with open(os.path.join(util.GOLD, 'front_state_dict_keys.json')) as f:
        return json.load(f)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_cnn_and_highway_match_reference(tag):
    g = util.gold('front_prims'); m = front_meta()['prim_' + tag]
    n, K, D, E, k = m['n'], m['K'], m['D'], m['E'], m['k']
    # fill keys carry no module prefix in the golden script (load_filled on the bare module)
    ref_sd = util.filled_sd({'conv1d.weight': (E, D, k), 'conv1d.bias': (E,)}, 20)
    ref_hw = util.filled_sd({'linear_projection.weight': (E, E), 'linear_projection.bias': (E,), 'linear_gate.weight': (E, E),
                             'linear_gate.bias': (E,)}, 21)
    sd = {('cnn.' + k_): v.clone().requires_grad_(True) for k_, v in ref_sd.items()}
    sd.update({('hw.' + k_): v.clone().requires_grad_(True) for k_, v in ref_hw.items()})
    x = t(fill.fill_array('front_x_' + tag, (n, K, D), 20) * 3.0)
    w = t(fill.fill_array('front_w_' + tag, (n, E), 20))
    c = FO.cnn(sd, 'cnn', x)
    np.testing.assert_allclose(c.detach().numpy(), g[tag + '_c'], rtol=1e-5, atol=1e-5)
    c2 = t(g[tag + '_c']).clone().requires_grad_(True)
    y = FO.highway(sd, 'hw', c2)
    np.testing.assert_allclose(y.detach().numpy(), g[tag + '_y'], rtol=1e-5, atol=1e-5)
    (y * w).sum().backward()
    (c * w).sum().backward()
    np.testing.assert_allclose(c2.grad.numpy(), g[tag + '_dc'], rtol=1e-4, atol=1e-5)
    for k_, v in sd.items():
        want = g[f'{tag}_grad:{k_.split(".", 1)[1]}']
        np.testing.assert_allclose(v.grad.numpy(), want, rtol=1e-4, atol=1e-5, err_msg=k_)


def _run_oracle(name, fn, **kw):
    g = util.gold(name); m = front_meta()[name]
    inv = front_inventory()[{'front_mft': 'MFT.MultiCNNTransformer', 'front_sft': 'SFT.MultiCNNTransformer', 'front_b2': 'B2.MultiCNNTransformer',
                             'front_b3': 'B3.MultiCNNTransformer', 'front_uni': 'MFT.MultiCNNTransformer.single'}[name]]
    sd = util.filled_sd({k: tuple(s) for k, s in inv.items()}, m['seed'], requires_grad=True)
    shapes = {k: tuple(v) for k, v in m['shapes'].items()}
    inputs, mask, target, lengths = fill.make_raw_batch(m['B'], m['T'], shapes, m['seed'])
    assert lengths == m['lengths']
    pred = fn(sd, {k: t(v) for k, v in inputs.items()}, t(mask), m['mods'], **kw)
    np.testing.assert_allclose(pred.detach().numpy(), g['pred'], rtol=1e-4, atol=2e-6)
    loss = O.train_loss(pred, t(target), lengths)
    assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    loss.backward()
    checked = 0
    for k, v in sd.items():
        if 'grad:' + k in g:
            util.assert_digest_close(util.grad_digest(v.grad), g['grad:' + k], 2e-4, k); checked += 1
        else:
            assert v.grad is None and k.startswith(('Transformer.attn', 'Transformer.ff')), k
    assert checked >= (12 if len(m["mods"]) > 1 else 6)


def test_multicnn_mft_matches_reference():
    _run_oracle('front_mft', FO.mcnn_mft)


def test_multicnn_b3_matches_reference():
    _run_oracle('front_b3', FO.mcnn_mft, use_encoder=False)


def test_multicnn_sft_matches_reference():
    _run_oracle('front_sft', FO.mcnn_sft)


def test_multicnn_b2_matches_reference():
    _run_oracle('front_b2', FO.mcnn_b2)


def test_multicnn_single_modality_matches_reference():
    _run_oracle('front_uni', FO.mcnn_uni)


MODS = ['acoustic', 'image', 'linguistic']
DIMS = {'acoustic': 88, 'image': 1000, 'linguistic': 300}
SDIMS = {'image': 1000, 'linguistic': 300}


@pytest.mark.parametrize('inv_name,ctor', [
    ('MFT.MultiCNNTransformer', lambda: M.MultiCNNTransformer(MODS, DIMS, {'acoustic': 88, 'image': 256, 'linguistic': 300})),
    ('SFT.MultiCNNTransformer', lambda: M.SFTMultiCNNTransformer(['image', 'linguistic'], SDIMS)),
    ('B2.MultiCNNTransformer', lambda: M.B2MultiCNNTransformer(['image', 'linguistic'], SDIMS)),
    ('B3.MultiCNNTransformer', lambda: M.B3MultiCNNTransformer(MODS, DIMS)),
    ('MFT.MultiCNNTransformer.single', lambda: M.MultiCNNTransformer(['linguistic'], {'linguistic': 300}, {'linguistic': 300})),
])
def test_multicnn_state_dict_keys_match_reference(inv_name, ctor):
    inv = front_inventory()[inv_name]
    sd = ctor().state_dict()
    assert list(sd.keys()) == list(inv.keys())
    for k, s in inv.items():
        assert list(sd[k].shape) == s, k


def test_front_end_has_no_cpu_fallback():
    m = M.MultiCNNTransformer(MODS, DIMS, {'acoustic': 88, 'image': 256, 'linguistic': 300})
    x = {'acoustic': torch.zeros(1, 3, 2, 88), 'image': torch.zeros(1, 3, 2, 1000), 'linguistic': torch.zeros(1, 3, 4, 300)}
    with pytest.raises(RuntimeError, match='CUDA'):
        m(x, [3], torch.ones(1, 3, 1))
    with pytest.raises(RuntimeError, match='CUDA'):
        mtb.functional.ccc_batched(torch.zeros(2, 4), torch.zeros(2, 4), [4, 2])


def test_raw_batch_generator_shapes_and_padding():
    shapes = {'linguistic': (5, 30), 'image': (2, 10)}
    inputs, mask, target, lengths = fill.make_raw_batch(4, 9, shapes, 3)
    assert inputs['linguistic'].shape == (4, 9, 5, 30) and inputs['image'].shape == (4, 9, 2, 10)
    assert lengths[0] == 9 and lengths == sorted(lengths, reverse=True)
    for b, l in enumerate(lengths):
        assert not inputs['linguistic'][b, l:].any() and not inputs['image'][b, l:].any()
        assert inputs['linguistic'][b, :l, :2].all(axis=-1).all()        # at least two real vectors per valid window
    assert (target * (1 - mask) == 0).all()


# ---- batcher -----------------------------------------------------------------------------------------------------------------------
def _digest(a):
    import hashlib
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return [list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()]


def batcher_gold():
    with open(os.path.join(util.GOLD, 'batcher.json')) as f:
        return json.load(f)


@pytest.mark.parametrize('tag,bs,on_eval', [('train_bs4', 4, False), ('eval_bs1', 1, True), ('eval_bs5', 5, True)])
def test_batcher_oracle_matches_reference_batches(tag, bs, on_eval):
    """The restated batcher yields bit-identical batches to the reference's own generateTrainBatch (same python RNG seed)."""
    import random
    from oracle.batcher_oracle import generate_train_batch
    from oracle.make_golden_batcher import corpus
    data, target, lengths = corpus()
    random.seed(123)
    got = list(generate_train_batch(data, target, lengths, batch_size=bs, on_eval=on_eval))
    want = batcher_gold()[tag]
    assert len(got) == len(want)
    for (d, tg, mask, ln), w in zip(got, want):
        assert ln == w['lengths']
        assert _digest(tg) == w['target'] and _digest(mask) == w['mask']
        for m_, v in d.items():
            assert _digest(v) == w['data'][m_], m_


def test_device_corpus_needs_a_gpu():
    from oracle.make_golden_batcher import corpus
    data, target, lengths = corpus()
    with pytest.raises(RuntimeError, match='CUDA'):
        mtb.DeviceCorpus(data, target, lengths, device='cpu')


# ---- host logic of the inference helpers (no GPU) ------------------------------------------------------------------------------------
def test_ragged_batch_context_nests_and_restores():
    from multimodal_transformer_b200 import functional as K
    assert K._state['key_len'] is None
    with mtb.ragged_batch([3, 2]):
        assert K._state['key_len'] == [3, 2]
        with mtb.ragged_batch([5]):
            assert K._state['key_len'] == [5]
        assert K._state['key_len'] == [3, 2]
    assert K._state['key_len'] is None
    with pytest.raises(ValueError):
        with mtb.ragged_batch([1]):
            raise ValueError('propagates')
    assert K._state['key_len'] is None


def test_model_call_convention_dispatch():
    """models.py classes take forward(inputs, length, mask); multiTransformer.py classes forward(inputs, mask, lengths)."""
    from multimodal_transformer_b200.evaluation import _call
    seen = {}

    class Front(M._FrontEnd):
        def forward(self, inputs, length, mask=None):
            seen['front'] = (inputs, length, mask)

    class Body(torch.nn.Module):
        def forward(self, inputs, mask, lengths):
            seen['body'] = (inputs, mask, lengths)

    _call(Front(), 'x', 'm', [1]); _call(Body(), 'x', 'm', [1])
    assert seen['front'] == ('x', [1], 'm') and seen['body'] == ('x', 'm', [1])


def test_grad_mode_is_visible_to_function_forward():
    """functional._apply hands the caller's grad mode to Function.forward (autograd always runs forward with grad mode off)."""
    from multimodal_transformer_b200 import functional as K

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            Probe.seen = K._need_grad(ctx)
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            return g

    x = torch.ones(2, requires_grad=True)
    K._apply(Probe, x); assert Probe.seen is True
    with torch.no_grad():
        K._apply(Probe, x); assert Probe.seen is False
    K._apply(Probe, torch.ones(2)); assert Probe.seen is False
    assert K._state['grad'] is True
