"""GPU (B200): the window front-end (CNN + max-pool + Highway + dropout: mt_window_cnn_fwd / _bwd), the MultiCNNTransformer mirrors
built on it and the batched CCC kernel -- through the drop-in modules, i.e. through the C ABI -- against the CPU oracle on
identical weights / inputs and against the golden outputs of the imported reference (tests/golden/front_*.npz).

Tolerances: fp32 mode 1e-5 relative (forward) / 3e-4 (gradients); bf16 mode 2e-2 absolute on valence."""
import json
import os

import numpy as np
import pytest
import torch

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib, functional as K, models as M
from oracle import fill, frontend_oracle as FO, mt_oracle as O
from oracle.ccc import eval_ccc
from oracle.dropout_rng import SITE_FRONT, Dropper
from tests import util
from tests.test_gpu_parity import assert_close, grad_floor, t

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(autouse=True)
def _fp32_mode():
    mtb.set_compute_dtype('fp32')
    mtb.fix_seed(None)
    yield
    mtb.set_compute_dtype('fp32')
    mtb.fix_seed(None)


def front_meta():
    with open(os.path.join(util.GOLD, 'front_meta.json')) as f:
        return json.load(f)


def front_inventory():
    with open(os.path.join(util.GOLD, 'front_state_dict_keys.json')) as f:
        return json.load(f)


def _prim_sd(E, D, k):
    sd = {'cnn.' + k_: v for k_, v in util.filled_sd({'conv1d.weight': (E, D, k), 'conv1d.bias': (E,)}, 20).items()}
    sd.update({'hw.' + k_: v for k_, v in util.filled_sd({'linear_projection.weight': (E, E), 'linear_projection.bias': (E,),
                                                         'linear_gate.weight': (E, E), 'linear_gate.bias': (E,)}, 21).items()})
    return sd


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_cnn_and_highway_modules_match_reference_golden(tag):
    """CNN.forward / Highway.forward (stage 1 / stage 2 of the kernel) in fp32 mode against the imported reference's outputs
    and gradients, incl. widths that are not multiples of 4 and a kernel size of 3."""
    g = util.gold('front_prims'); m = front_meta()['prim_' + tag]
    n, Kv, D, E, k = m['n'], m['K'], m['D'], m['E'], m['k']
    sd = _prim_sd(E, D, k)
    cnn = M.CNN(D, E, k).to(DEV); cnn.load_state_dict({k_[4:]: v for k_, v in sd.items() if k_.startswith('cnn.')})
    hw = M.Highway(E).to(DEV); hw.load_state_dict({k_[3:]: v for k_, v in sd.items() if k_.startswith('hw.')})
    x = t(fill.fill_array('front_x_' + tag, (n, Kv, D), 20) * 3.0).to(DEV)
    w = t(fill.fill_array('front_w_' + tag, (n, E), 20)).to(DEV)
    c = cnn(x.permute(0, 2, 1))                       # the reference hands the channel-major view to CNN.forward
    assert_close(c, t(g[tag + '_c']), 1e-5, 'cnn out')
    c2 = t(g[tag + '_c']).to(DEV).requires_grad_(True)
    y = hw(c2)
    assert_close(y, t(g[tag + '_y']), 1e-5, 'highway out')
    (y * w).sum().backward()
    (c * w).sum().backward()
    assert_close(c2.grad, t(g[tag + '_dc']), 1e-4, 'dc')
    for name, mod in (('cnn', cnn), ('hw', hw)):
        for k_, p in mod.named_parameters():
            assert_close(p.grad, t(g[f'{tag}_grad:{k_}']), 1e-4, k_, 1e-6)


@pytest.mark.parametrize('n,Kv,D,E,k', [(300, 33, 300, 300, 2), (513, 2, 1000, 256, 2), (257, 2, 88, 88, 2), (64, 7, 36, 20, 3)])
@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_window_cnn_fused_train_mode_same_masks(n, Kv, D, E, k, mode):
    """All three stages in one call, train mode (dropout 0.3 with the oracle's masks), at the real modality shapes (linguistic 33 x 300,
    image 2 x 1000, acoustic 2 x 88) and an odd one: fp32 mode against the fp64 oracle at 1e-5 / 3e-4; bf16 mode (tcgen05 GEMM over
    overlapping rows) within bf16 operand rounding, gradients by cosine."""
    seed = 99
    sd = _prim_sd(E, D, k)
    sd['cnn.conv1d.weight'] = sd['cnn.conv1d.weight'] * (1.0 / D) ** 0.5      # nn.Conv1d's default scale (fan_in = D * k), O(1) features
    rs = np.random.RandomState(n)
    x = rs.standard_normal((n, Kv, D)).astype(np.float32)
    if Kv > 2:
        nvec = rs.randint(k, Kv + 1, size=n)
        x *= (np.arange(Kv)[None, :] < nvec[:, None]).astype(np.float32)[..., None]
    w = rs.standard_normal((n, E)).astype(np.float32)
    sdr = {k_: v.double().requires_grad_(True) for k_, v in sd.items()}
    cr = FO.cnn(sdr, 'cnn', t(x).double())
    yr = Dropper(seed)(FO.highway(sdr, 'hw', cr), 0.3, SITE_FRONT)
    (yr * t(w).double()).sum().backward()
    mtb.set_compute_dtype(mode)
    mtb.fix_seed(seed)
    ps = {k_: v.to(DEV).requires_grad_(True) for k_, v in sd.items()}
    y = K.window_cnn(t(x).to(DEV), ps['cnn.conv1d.weight'], ps['cnn.conv1d.bias'], ps['hw.linear_projection.weight'],
                     ps['hw.linear_projection.bias'], ps['hw.linear_gate.weight'], ps['hw.linear_gate.bias'], p_drop=0.3, site=SITE_FRONT)
    (y * t(w).to(DEV)).sum().backward()
    # identical dropout masks: the zero patterns agree exactly
    assert torch.equal((y == 0).cpu(), (yr == 0)) or mode == 'bf16'
    fl = grad_floor([v.grad for v in sdr.values()])
    if mode == 'fp32':
        assert_close(y, yr, 1e-5, 'out')
        for k_, p in ps.items():
            assert_close(p.grad, sdr[k_].grad, 3e-4, k_, fl)
    else:
        assert_close(y, yr, 2e-2, 'out')
        for k_, p in ps.items():
            cos = torch.nn.functional.cosine_similarity(p.grad.double().cpu().flatten(), sdr[k_].grad.flatten(), dim=0).item()
            assert cos > 0.97, (k_, cos)           # bf16 near-ties may pick another arg-max position than the fp64 oracle


def _load(model, name):
    inv = front_inventory()[{'front_mft': 'MFT.MultiCNNTransformer', 'front_sft': 'SFT.MultiCNNTransformer', 'front_b2': 'B2.MultiCNNTransformer',
                             'front_b3': 'B3.MultiCNNTransformer', 'front_uni': 'MFT.MultiCNNTransformer.single'}[name]]
    m = front_meta()[name]
    sd = util.filled_sd({k: tuple(s) for k, s in inv.items()}, m['seed'])
    model.load_state_dict(sd)
    shapes = {k: tuple(v) for k, v in m['shapes'].items()}
    inputs, mask, target, lengths = fill.make_raw_batch(m['B'], m['T'], shapes, m['seed'])
    return sd, inputs, mask, target, lengths


@pytest.mark.parametrize('name,ctor', [
    ('front_mft', lambda m: M.MultiCNNTransformer(m['mods'], {k: v[1] for k, v in m['shapes'].items()}, m['embed_dims'])),
    ('front_b3', lambda m: M.B3MultiCNNTransformer(m['mods'], {k: v[1] for k, v in m['shapes'].items()})),
    ('front_sft', lambda m: M.SFTMultiCNNTransformer(m['mods'], {k: v[1] for k, v in m['shapes'].items()})),
    ('front_b2', lambda m: M.B2MultiCNNTransformer(m['mods'], {k: v[1] for k, v in m['shapes'].items()})),
    ('front_uni', lambda m: M.MultiCNNTransformer(m['mods'], {k: v[1] for k, v in m['shapes'].items()}, m['embed_dims'])),
])
def test_multicnn_models_match_reference_golden(name, ctor):
    """The four MultiCNNTransformer variants end to end (raw windows -> prediction), eval mode, fp32: prediction, loss and every
    parameter gradient against the imported reference; then bf16 mode within 2e-2 on valence."""
    g = util.gold(name); m = front_meta()[name]
    model = ctor(m).eval()
    sd, inputs, mask, target, lengths = _load(model, name)
    xin = {k: t(v).to(DEV) for k, v in inputs.items()}
    pred = model(xin, lengths, t(mask).to(DEV))
    assert_close(pred, t(g['pred']), 1e-5, 'pred', 1e-7)
    loss = ((pred - t(target).to(DEV)) ** 2).sum() / sum(lengths)
    assert abs(loss.item() - float(g['loss'])) <= 2e-5 * abs(float(g['loss']))
    loss.backward()
    checked = 0
    for k, p in model.named_parameters():
        if 'grad:' + k in g:
            util.assert_digest_close(util.grad_digest(p.grad), g['grad:' + k], 3e-4, k); checked += 1
        else:
            assert p.grad is None and k.startswith(('Transformer.attn', 'Transformer.ff')), k
    assert checked >= (12 if len(m['mods']) > 1 else 6)
    mtb.set_compute_dtype('bf16')
    with torch.no_grad():
        p16 = model(xin, lengths, t(mask).to(DEV))
    assert (p16.float().cpu() - t(g['pred'])).abs().max().item() < 2e-2


def test_multicnn_mft_train_mode_same_masks_both_dtypes():
    """Train mode through the front-end AND the hot path with the oracle's dropout masks (front-end sites 0x6000 + modality)."""
    name, seed = 'front_mft', 1234
    m = front_meta()[name]
    dims = {k: v[1] for k, v in m['shapes'].items()}
    model = M.MultiCNNTransformer(m['mods'], dims, m['embed_dims']).train()
    sd, inputs, mask, target, lengths = _load(model, name)
    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    predr = FO.mcnn_mft(sdr, {k: t(v).double() for k, v in inputs.items()}, t(mask).double(), m['mods'], drop=Dropper(seed))
    O.train_loss(predr, t(target).double(), lengths).backward()
    fl = grad_floor([v.grad for v in sdr.values()])
    for mode in ('fp32', 'bf16'):
        mtb.set_compute_dtype(mode)
        model.zero_grad(set_to_none=True)
        mtb.fix_seed(seed)
        pred = model({k: t(v).to(DEV) for k, v in inputs.items()}, lengths, t(mask).to(DEV))
        (((pred - t(target).to(DEV)) ** 2).sum() / sum(lengths)).backward()
        if mode == 'fp32':
            assert_close(pred, predr, 2e-5, 'pred')
            for k, p in model.named_parameters():
                if sdr[k].grad is not None:
                    assert_close(p.grad, sdr[k].grad, 3e-4, k, fl)
        else:
            assert (pred.detach().float().cpu() - predr.float()).abs().max().item() < 2e-2


def test_window_cnn_rejects_bad_arguments():
    L = _lib.lib()
    import ctypes
    cfg = _lib.MtWindowCnnCfg(0, 4, 3, 8, 8, 4, 3, 0, 0.0, 0, 0)          # conv kernel longer than the window
    assert L.mt_window_cnn_ws_bytes(ctypes.byref(cfg)) == 0
    cfg = _lib.MtWindowCnnCfg(0, 4, 3, 8, 8, 2, 0, 0, 0.0, 0, 0)          # no stage selected
    assert L.mt_window_cnn_ws_bytes(ctypes.byref(cfg)) == 0
    x = torch.zeros(2, 1, 8, device=DEV)
    cnn = M.CNN(8, 4, 2).to(DEV)
    with pytest.raises(RuntimeError, match='shorter'):
        cnn(x.permute(0, 2, 1))
    with pytest.raises(RuntimeError, match='float32'):
        K.highway(torch.zeros(2, 4, device=DEV, dtype=torch.bfloat16), *[torch.zeros(4, 4, device=DEV), torch.zeros(4, device=DEV)] * 2)


# ---- batched evaluation metrics ----------------------------------------------------------------------------------------------------
def test_ccc_batched_known_answers_and_oracle():
    """The reference's published known-answers (PredSave -> PerfSave, tests/golden/ccc_kat.json) as ONE padded batch, plus random
    ragged narratives against the eval_ccc restatement and scipy's pearsonr."""
    from scipy.stats import pearsonr
    with open(os.path.join(util.GOLD, 'ccc_kat.json')) as f:
        kat = json.load(f)
    T = max(len(k['pred']) for k in kat)
    pred = np.zeros((len(kat), T), np.float32); act = np.zeros((len(kat), T), np.float32)
    for i, k in enumerate(kat):
        pred[i, :len(k['pred'])] = k['pred']; act[i, :len(k['actual'])] = k['actual']
        pred[i, len(k['pred']):] = 7.0                      # padding must not leak into the statistics
    lengths = [len(k['pred']) for k in kat]
    ccc, pr, se = K.ccc_batched(t(pred).to(DEV), t(act).to(DEV).unsqueeze(-1), lengths)
    for i, k in enumerate(kat):
        assert abs(ccc[i].item() - k['ccc']) < 2e-7, (k['model'], k['vid'])
    rs = np.random.RandomState(5)
    B, T = 37, 300
    lengths = [T] + [int(v) for v in rs.randint(2, T + 1, size=B - 1)]
    a = rs.uniform(0, 1, (B, T)).astype(np.float32)
    p = (0.7 * a + 0.3 * rs.uniform(0, 1, (B, T))).astype(np.float32)
    ccc, pr, se = K.ccc_batched(t(p).to(DEV), t(a).to(DEV), torch.tensor(lengths))
    want_se = 0.0
    for b, l in enumerate(lengths):
        assert abs(ccc[b].item() - eval_ccc(a[b, :l], p[b, :l])) < 1e-12
        assert abs(pr[b].item() - pearsonr(p[b, :l].astype(np.float64), a[b, :l].astype(np.float64))[0]) < 1e-12
        want_se += ((p[b, :l].astype(np.float64) - a[b, :l].astype(np.float64)) ** 2).sum()
    assert abs(se.item() - want_se) < 1e-9 * want_se


# ---- ragged inference: a padded batch that reproduces one-narrative-at-a-time evaluation -------------------------------------------
@pytest.mark.parametrize('mode,force,T', [('fp32', None, 37), ('fp32', None, 150), ('bf16', None, 128), ('bf16', None, 100), ('bf16', 'tiled', 128),
                                          ('bf16', None, 200), ('bf16', 'ffma', 90)])
def test_ragged_attention_equals_per_narrative_attention(mode, force, T):
    """Every attention engine (fp32 FFMA, bf16 FFMA, bf16 tiled tensor-core, bf16 whole-head T <= 128 incl. the T == 128 case):
    with key_len, rows < len of narrative b equal the attention of that narrative alone (T = len, no mask)."""
    B, d, h = 5, 256, 8
    L = _lib.lib()
    mtb.set_compute_dtype(mode)
    dt = torch.float32 if mode == 'fp32' else torch.bfloat16
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, T, 3 * d, generator=g).to(DEV).to(dt)
    lengths = [T, T - 1, max(1, T // 2), 9, 1]
    mask = torch.zeros(B, T, 1, device=DEV)
    for b, l in enumerate(lengths):
        mask[b, :l] = 1
    old = L.mt_attention_force_ffma(1) if force == 'ffma' else (L.mt_attention_force_tiled(1) if force == 'tiled' else None)
    try:
        with torch.no_grad():
            with mtb.ragged_batch(lengths):
                got = K.attention_packed(qkv, mask, h)
            for b, l in enumerate(lengths):
                alone = K.attention_packed(qkv[b:b + 1, :l].contiguous(), None, h)
                assert_close(got[b, :l], alone[0], 1e-5 if mode == 'fp32' else 1e-2, f'narrative {b} (len {l})')
                q, k, v = [x.float().view(l, h, d // h).transpose(0, 1) for x in qkv[b, :l].split(d, dim=-1)]
                want = (torch.softmax(q @ k.transpose(1, 2) / (d // h) ** 0.5, -1) @ v).transpose(0, 1).reshape(l, d)
                assert_close(got[b, :l], want, 2e-5 if mode == 'fp32' else 2e-2, f'narrative {b} vs torch')
    finally:
        if force == 'ffma':
            L.mt_attention_force_ffma(old)
        elif force == 'tiled':
            L.mt_attention_force_tiled(old)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_ragged_batch_equals_one_at_a_time_and_differs_from_padded(mode):
    """MFT body, eval: inside ragged_batch the valid part of every prediction equals a forward of that narrative alone (the
    reference's evaluation, batch_size = 1); the plain padded batch (training semantics: padded windows are live keys) does not."""
    N, B, T = 2, 6, 40
    dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
    mods = ['acoustic', 'image', 'linguistic']
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', N), 41)
    inputs, mask, _, lengths = fill.make_batch(B, T, dims, 41)
    model = mtb.MultiTransformer(mods, dims, N=N).eval(); model.load_state_dict(sd)
    mtb.set_compute_dtype(mode)
    xin = {k: t(v).to(DEV) for k, v in inputs.items()}
    tol = 1e-5 if mode == 'fp32' else 2e-2
    with torch.no_grad():
        padded = model(xin, t(mask).to(DEV), lengths)
        with mtb.ragged_batch(lengths):
            ragged = model(xin, t(mask).to(DEV), lengths)
        worst_pad = 0.0
        for b, l in enumerate(lengths):
            alone = model({k: v[b:b + 1, :l].contiguous() for k, v in xin.items()}, torch.ones(1, l, 1, device=DEV), [l])
            if mode == 'fp32':
                want = O.multi_transformer(sd, '', {k: t(v[b:b + 1, :l]) for k, v in inputs.items()}, torch.ones(1, l, 1), mods, N=N)
                assert_close(alone, want, 1e-5, f'alone vs oracle {b}', 1e-7)
            err = (ragged[b, :l] - alone[0]).abs().max().item()
            assert err <= tol * max(1.0, alone.abs().max().item()) * (1 if mode == 'bf16' else 1), (b, l, err)
            assert not ragged[b, l:].any()                            # the output mask still zeroes the padding
            if l < T:
                worst_pad = max(worst_pad, (padded[b, :l] - alone[0]).abs().max().item())
    if mode == 'fp32':
        assert worst_pad > 1e-4                                       # padded keys DO change valid outputs without ragged_batch
    model.train()
    with pytest.raises(RuntimeError, match='inference only'):
        with mtb.ragged_batch(lengths):
            model(xin, t(mask).to(DEV), lengths)


def test_batched_evaluation_equals_one_at_a_time():
    """evaluate(): MultiCNNTransformer on raw windows, 9 narratives in batches of 4 with on-device CCC, against the reference's
    procedure -- one narrative per forward, eval_ccc / pearsonr on the host (MFT/train.py:203-257)."""
    from scipy.stats import pearsonr
    m = front_meta()['front_mft']
    shapes = {k: tuple(v) for k, v in m['shapes'].items()}
    dims = {k: v[1] for k, v in shapes.items()}
    model = M.MultiCNNTransformer(m['mods'], dims, m['embed_dims'])
    inv = front_inventory()['MFT.MultiCNNTransformer']
    model.load_state_dict(util.filled_sd({k: tuple(s) for k, s in inv.items()}, 5))
    N, T = 9, 24
    inputs, mask, target, lengths = fill.make_raw_batch(N, T, shapes, 77)
    order = np.random.RandomState(1).permutation(N)                   # evaluation sets are not sorted by length
    inputs = {k: v[order] for k, v in inputs.items()}; mask = mask[order]; lengths = [lengths[i] for i in order]
    xin = {k: t(v).to(DEV) for k, v in inputs.items()}
    model.eval()
    ones, want_pred = [], []
    with torch.no_grad():
        for b, l in enumerate(lengths):
            o = model({k: v[b:b + 1, :l].contiguous() for k, v in xin.items()}, [l], torch.ones(1, l, 1, device=DEV))
            want_pred.append(o.reshape(-1).cpu().numpy())
    rs = np.random.RandomState(2)
    target = np.zeros((N, T, 1), np.float32)
    for b, l in enumerate(lengths):                                   # targets correlated with the predictions: CCC needs signal
        p = want_pred[b]
        target[b, :l, 0] = 0.5 + 5.0 * (p - p.mean()) + 0.01 * rs.standard_normal(l)
    preds, loss, stats, (bo, bt, bi) = mtb.evaluate(model, xin, t(target).to(DEV), t(mask).to(DEV), lengths, batch_size=4)
    assert not model.training
    cccs, corrs, se = [], [], 0.0
    for b, l in enumerate(lengths):
        np.testing.assert_allclose(preds[b], want_pred[b], rtol=1e-5, atol=1e-6)
        cccs.append(eval_ccc(want_pred[b], target[b, :l, 0])); corrs.append(pearsonr(want_pred[b], target[b, :l, 0])[0])
        se += ((want_pred[b].astype(np.float64) - target[b, :l, 0]) ** 2).sum()
    assert abs(stats['ccc'] - np.mean(cccs)) < 1e-5 and abs(stats['ccc_std'] - np.std(cccs)) < 1e-5
    assert abs(stats['corr'] - np.mean(corrs)) < 1e-5 and abs(stats['corr_std'] - np.std(corrs)) < 1e-5
    assert abs(stats['max_ccc'] - max(cccs)) < 1e-5 and bi == int(np.argmax(cccs)) + 1
    assert abs(loss - se / sum(lengths)) < 1e-5 * (se / sum(lengths))
    np.testing.assert_allclose(bt, target[bi - 1, :lengths[bi - 1], 0])
    # the same through a device-resident corpus: evaluate(model, corpus)
    corpus = mtb.DeviceCorpus(inputs, target[..., 0], lengths)
    preds2, loss2, stats2, _ = mtb.evaluate(model, corpus, batch_size=4)
    assert abs(loss2 - loss) < 1e-12 and stats2 == stats and all(np.array_equal(a, b) for a, b in zip(preds, preds2))


# ---- GPU-side batcher -----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag,bs,on_eval', [('train_bs4', 4, False), ('eval_bs1', 1, True), ('eval_bs5', 5, True)])
def test_device_batcher_yields_the_reference_batches(tag, bs, on_eval):
    """DeviceCorpus / generateTrainBatch (index gather on the device) against the golden digests of the reference's own
    generateTrainBatch: bit-exact data, target, mask and lengths, same batch order under the same python RNG seed."""
    import hashlib
    import random
    from oracle.make_golden_batcher import corpus
    with open(os.path.join(util.GOLD, 'batcher.json')) as f:
        want = json.load(f)[tag]

    def dg(x):
        a = np.ascontiguousarray(x.cpu().numpy().astype(np.float32))
        return [list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()]

    data, target, lengths = corpus()
    random.seed(123)
    got = list(mtb.generateTrainBatch(data, target, lengths, None, batch_size=bs, onEval=on_eval))
    assert len(got) == len(want)
    for (d, tg, mask, ln), w in zip(got, want):
        assert ln == w['lengths'] and dg(tg) == w['target'] and dg(mask) == w['mask']
        assert all(v.is_cuda for v in d.values()) and tg.is_cuda and mask.is_cuda
        for m_, v in d.items():
            assert dg(v) == w['data'][m_], m_


def test_device_batcher_large_rows_and_odd_sizes():
    """Vector and scalar copy paths, many narratives, prefix shorter than the row: against the oracle batcher."""
    import random
    from oracle.batcher_oracle import generate_train_batch
    rs = np.random.RandomState(8)
    n, t_max = 70, 33
    lengths = [int(v) for v in rs.randint(1, t_max + 1, size=n)]
    data = {'a': rs.standard_normal((n, t_max, 3, 7)).astype(np.float32), 'b': rs.standard_normal((n, t_max, 2, 88)).astype(np.float32)}
    target = rs.uniform(0, 1, (n, t_max)).astype(np.float32)
    corpus = mtb.DeviceCorpus(data, target, lengths)
    random.seed(5)
    got = list(corpus.generateTrainBatch(batch_size=32))
    random.seed(5)
    want = list(generate_train_batch(data, target, lengths, batch_size=32))
    assert len(got) == len(want) == 3
    for (d, tg, mask, ln), (dw, tw, mw, lw) in zip(got, want):
        assert ln == lw
        assert np.array_equal(tg.cpu().numpy(), tw) and np.array_equal(mask.cpu().numpy(), mw)
        for m_ in d:
            assert np.array_equal(d[m_].cpu().numpy(), dw[m_]), m_


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_graphed_train_step_with_window_front_end(mode):
    """GraphedTrainStep over a MultiCNNTransformer (raw windows in, front-end + hot path + loss + backward + Adam in one CUDA graph)
    replays to the same losses / parameters as eager steps with FlatAdam (dropout off: both deterministic)."""
    from multimodal_transformer_b200.training import FlatAdam, GraphedTrainStep, train_step_loss
    m = front_meta()['front_mft']
    shapes = {k: tuple(v) for k, v in m['shapes'].items()}
    dims = {k: v[1] for k, v in shapes.items()}
    inv = front_inventory()['MFT.MultiCNNTransformer']
    sd = util.filled_sd({k: tuple(s) for k, s in inv.items()}, 9)
    B, T = 4, 10
    batches = [fill.make_raw_batch(B, T, shapes, 80 + i) for i in range(3)]
    mtb.set_compute_dtype(mode)

    def fresh():
        model = M.MultiCNNTransformer(m['mods'], dims, m['embed_dims']); model.load_state_dict(sd)
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if hasattr(mod, 'p_drop'):
                mod.p_drop = 0.0
        return model, FlatAdam(model, lr=1e-3, weight_decay=1e-4)

    def no_dropout(model):                             # MFN / encoder dropout probabilities live in their cfg dicts
        for mod in model.modules():
            for attr in ('cfgd', '_cfgd'):
                d = getattr(mod, attr, None)
                if isinstance(d, dict):
                    for k in list(d):
                        if k.startswith('p_'):
                            d[k] = 0.0

    mtb.fix_seed(77)                                   # same dropout masks on both paths would need the graph's seed offsets: switch sites off instead
    model_e, opt_e = fresh(); no_dropout(model_e)
    losses_e = []
    for inputs, mask, target, lengths in batches:
        model_e.eval()                                  # eval(): every dropout off, gradients still flow -- deterministic reference for the graph
        pred = model_e({k: t(v).to(DEV) for k, v in inputs.items()}, lengths, t(mask).to(DEV))
        losses_e.append(train_step_loss(pred, t(target).to(DEV), float(sum(lengths))).item())
        opt_e.step(); opt_e.zero_grad()
    del pred

    model_g, opt_g = fresh(); no_dropout(model_g)

    class EvalStep(GraphedTrainStep):                   # the captured step calls model.train(); keep it in eval() for the comparison
        def _step(self):
            tr = self.model.train
            self.model.train = lambda *a, **k: self.model
            try:
                return super()._step()
            finally:
                self.model.train = tr

    model_g.eval()
    gstep = EvalStep(model_g, opt_g, B, T, shapes, torch.device(DEV), warmup=2)
    losses_g = [gstep({k: t(v) for k, v in inputs.items()}, t(mask), t(target), lengths).item() for inputs, mask, target, lengths in batches]
    assert opt_g.step_count == 3
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= (1e-4 if mode == 'fp32' else 3e-2) * abs(a), (losses_e, losses_g)
    if mode == 'fp32':          # bf16: three Adam steps at lr 1e-3 move a weight by at most 3e-3 whatever the gradient noise -- the losses are the check
        pe = dict(model_e.named_parameters())
        for k, p in model_g.named_parameters():
            if k.startswith(('Transformer.attn', 'Transformer.ff')):
                continue
            # fp32 split-K weight gradients are summed with atomics (order varies between the eager and the captured run); Adam's g / sqrt(v)
            # turns a sign flip of a round-off-level gradient into a +-lr step, so the floor is a small fraction of what three steps at
            # lr 1e-3 can move a weight (seen: 4.4e-5 on highway_acoustic.linear_projection.weight in 1 of 4 runs)
            assert_close(p, pe[k], 2e-4, k, 2e-4)


def test_batch_into_static_buffers_equals_batch():
    """DeviceCorpus.batch_into (gather straight into preallocated [B, T, ...] buffers, T >= the batch's longest narrative) gives the
    same tensors as batch() zero-extended to T."""
    rs = np.random.RandomState(3)
    n, t_max, T = 20, 17, 15
    lengths = [int(v) for v in rs.randint(1, T + 1, size=n)]
    data = {'a': rs.standard_normal((n, t_max, 3, 8)).astype(np.float32), 'b': rs.standard_normal((n, t_max, 2, 5)).astype(np.float32)}
    target = rs.uniform(0, 1, (n, t_max)).astype(np.float32)
    for i, l in enumerate(lengths):
        for v in data.values():
            v[i, l:] = 0
        target[i, l:] = 0
    corpus = mtb.DeviceCorpus(data, target, lengths)
    chunk = [int(v) for v in rs.permutation(n)[:8]]
    d, tg, mask, ln = corpus.batch(chunk)
    x = {m_: torch.full((8, T) + v.shape[2:], 7.0, device=DEV) for m_, v in data.items()}
    tgt, msk = torch.full((8, T, 1), 7.0, device=DEV), torch.full((8, T, 1), 7.0, device=DEV)
    ln2 = corpus.batch_into(chunk, x, tgt, msk)
    Tb = ln[0]
    assert ln2 == ln
    for m_ in d:
        assert torch.equal(x[m_][:, :Tb], d[m_]) and not x[m_][:, Tb:].any()
    assert torch.equal(tgt[:, :Tb], tg) and not tgt[:, Tb:].any()
    assert torch.equal(msk[:, :Tb], mask) and not msk[:, Tb:].any()
    with pytest.raises(RuntimeError, match='do not fit'):
        corpus.batch_into(chunk[:5], x, tgt, msk)


def test_graphed_forward_with_window_front_end():
    """GraphedForward over a MultiCNNTransformer (forward(inputs, length, mask) argument order, raw windows in) equals the eager eval()
    forward of the same model."""
    from multimodal_transformer_b200.training import GraphedForward
    m = front_meta()['front_mft']
    shapes = {k: tuple(v) for k, v in m['shapes'].items()}
    dims = {k: v[1] for k, v in shapes.items()}
    inv = front_inventory()['MFT.MultiCNNTransformer']
    sd = util.filled_sd({k: tuple(s) for k, s in inv.items()}, 9)
    B, T = 3, 9
    inputs, mask, target, lengths = fill.make_raw_batch(B, T, shapes, 41)
    model = M.MultiCNNTransformer(m['mods'], dims, m['embed_dims']); model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        ref = model({k: t(v).to(DEV) for k, v in inputs.items()}, [T] * B, t(mask).to(DEV)).clone()
    gf = GraphedForward(model, B, T, shapes, torch.device(DEV), warmup=1)
    out = gf({k: t(v) for k, v in inputs.items()}, t(mask))
    torch.cuda.synchronize()
    assert_close(out, ref, 1e-5, 'pred', 1e-7)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_modality_concat_kernel_equals_torch_cat_and_launches_no_framework_copy(mode):
    """SFT/models.py:136-138, B2-Trans/models.py:130-132: torch.cat(outputs, 2) in front of the fusion / embed Linear is assembled by the
    library's strided cast kernel (mt_concat_fwd / _bwd): same values and gradients as torch.cat, at the reference's odd widths."""
    mtb.set_compute_dtype(mode)
    try:
        torch.manual_seed(4)
        dt = torch.float32 if mode == 'fp32' else torch.bfloat16
        widths = [88, 256, 300]
        feats = [torch.randn(5, 7, w, device='cuda').to(dt).requires_grad_(True) for w in widths]
        L = _lib.lib()
        n0 = L.mt_launch_count()
        out = K.concat_features(feats)
        assert L.mt_launch_count() - n0 == len(widths)
        ref = torch.cat([f.detach() for f in feats], 2)
        assert out.dtype == dt and torch.equal(out, ref)
        g = torch.randn_like(out)
        n0 = L.mt_launch_count()
        out.backward(g)
        assert L.mt_launch_count() - n0 == len(widths)
        o = 0
        for f, w in zip(feats, widths):
            assert f.grad.is_contiguous() and torch.equal(f.grad, g[..., o:o + w])
            o += w
        # mixed input dtypes (fp32 features into a bf16 operand) are converted on the way
        if mode == 'bf16':
            mixed = [feats[0].detach().float(), feats[1].detach()]
            assert torch.equal(K.concat_features(mixed), torch.cat([mixed[0].bfloat16(), mixed[1]], 2))
    finally:
        mtb.set_compute_dtype('fp32')
