"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes (imported from
/root/reference, which exists only in the build container) on deterministic weights and inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden

Weights and inputs are produced by oracle/fill.py from (key, shape, seed) alone, so the fixtures only
store the reference's OUTPUTS (predictions, loss, per-parameter gradient digests), a few KB each.
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

from . import fill

REF = '/root/reference/transformer'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _load(dirname, alias):
    """Import <REF>/<dirname>/multiTransformer.py under a private name (matplotlib is not installed:
    the reference imports matplotlib.pyplot at multiTransformer.py:7 but never uses it on this path)."""
    if 'matplotlib' not in sys.modules:
        m = types.ModuleType('matplotlib'); mp = types.ModuleType('matplotlib.pyplot'); m.pyplot = mp
        sys.modules['matplotlib'] = m; sys.modules['matplotlib.pyplot'] = mp
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, dirname, 'multiTransformer.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_filled(module, seed=1):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = fill.fill_state(shapes, seed)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return shapes


def grad_digest(g):
    g = g.detach().double().reshape(-1)
    n = g.numel()
    idx = torch.linspace(0, n - 1, steps=min(n, 16)).long()
    return np.concatenate([[g.norm().item(), g.sum().item()], g[idx].numpy()])


def digest_params(module):
    out = {}
    for k, p in module.named_parameters():
        if p.grad is not None:
            out['grad:' + k] = grad_digest(p.grad)
    return out


def t(x):
    return torch.from_numpy(x)


def main():
    warnings.filterwarnings('ignore')
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    mft = _load('MFT', 'ref_mft')
    b3 = _load('B3-MFN', 'ref_b3')
    sft = _load('SFT', 'ref_sft')
    meta = {}

    # ---- primitives -------------------------------------------------------------------------
    ln = mft.LayerNorm(256); load_filled(ln, 3)
    x = t(fill.fill_array('ln_x', (3, 5, 256), 3)) * 20.0 + 1.5
    np.savez(os.path.join(OUT, 'ln.npz'), y=ln(x).detach().numpy())

    mha = mft.MultiHeadedAttention(8, 256).eval(); load_filled(mha, 4)
    inputs, mask, _, lengths = fill.make_batch(3, 7, {'x': 256}, 4)
    xx = t(inputs['x']); y = mha(xx, xx, xx, t(mask))
    np.savez(os.path.join(OUT, 'mha.npz'), y=y.detach().numpy(), attn=mha.attn.detach().numpy())
    meta['mha'] = dict(B=3, T=7, lengths=lengths, seed=4)

    c = __import__('copy').deepcopy
    enc = mft.Encoder(mft.EncoderLayer(256, c(mft.MultiHeadedAttention(8, 256)),
                                        c(mft.PositionwiseFeedForward(256, 128, 0.1)), 0.1), 2).eval()
    load_filled(enc, 5)
    inputs, mask, _, lengths = fill.make_batch(3, 9, {'x': 256}, 5)
    xx = t(inputs['x']).requires_grad_(True)
    y = enc(xx, t(mask)); w = t(fill.fill_array('enc_w', (3, 9, 256), 5)); (y * w).sum().backward()
    np.savez(os.path.join(OUT, 'encoder.npz'), y=y.detach().numpy(), dx=xx.grad.numpy(), **digest_params(enc))
    meta['encoder'] = dict(B=3, T=9, N=2, lengths=lengths, seed=5)

    # ---- MFN alone --------------------------------------------------------------------------
    mods = ['acoustic', 'image', 'linguistic']
    dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
    mfn = mft.MFN(mods, {m: 256 for m in mods}, 1).eval(); load_filled(mfn, 6)
    inputs, mask, _, lengths = fill.make_batch(3, 6, {m: 256 for m in mods}, 6)
    xin = {m: t(inputs[m]).permute(1, 0, 2).contiguous().requires_grad_(True) for m in mods}
    y = mfn(xin); w = t(fill.fill_array('mfn_w', (3, 6, 1), 6)); (y * w).sum().backward()
    np.savez(os.path.join(OUT, 'mfn.npz'), y=y.detach().numpy(),
             **{'dx_' + m: xin[m].grad.numpy() for m in mods}, **digest_params(mfn))
    meta['mfn'] = dict(B=3, T=6, mods=mods, seed=6)

    # ---- MFT (N=2 with grads; N=6 forward only) ------------------------------------------------
    def run_model(model, inputs, mask, target, lengths, with_grad=True):
        model.eval()
        pred = model(inputs, t(mask), lengths)
        res = {'pred': pred.detach().numpy()}
        if with_grad:
            loss = ((pred - t(target)) ** 2).sum() / sum(lengths)       # MFT/train.py:135-139
            loss.backward()
            res['loss'] = np.array(loss.item())
            res.update(digest_params(model))
        return res

    for name, N, B, T, seed, wg in [('mft_n2', 2, 3, 10, 7, True), ('mft_n6', 6, 2, 12, 8, False)]:
        m = mft.MultiTransformer(mods, dims, N=N); load_filled(m, seed)
        inputs, mask, target, lengths = fill.make_batch(B, T, dims, seed)
        res = run_model(m, {k: t(v) for k, v in inputs.items()}, mask, target, lengths, wg)
        np.savez(os.path.join(OUT, name + '.npz'), **res)
        meta[name] = dict(B=B, T=T, N=N, mods=mods, dims=dims, lengths=lengths, seed=seed)

    # ---- B3-MFN ------------------------------------------------------------------------------
    b3dims = {'acoustic': 256, 'image': 256, 'linguistic': 300}           # B3-MFN/models.py:90
    m = b3.MultiTransformer(mods, b3dims); load_filled(m, 9)
    inputs, mask, target, lengths = fill.make_batch(3, 10, b3dims, 9)
    res = run_model(m, {k: t(v) for k, v in inputs.items()}, mask, target, lengths)
    np.savez(os.path.join(OUT, 'b3.npz'), **res)
    meta['b3'] = dict(B=3, T=10, mods=mods, dims=b3dims, lengths=lengths, seed=9)

    # ---- SFT hot path: fusionLayer + tanh + NLPTransformer (SFT/models.py:98,136-139) ------------
    class SFTPath(torch.nn.Module):
        def __init__(self, total, N):
            super().__init__()
            self.fusionLayer = torch.nn.Linear(total, 512)
            self.Transformer = sft.NLPTransformer(512, N=N)

        def forward(self, feats, mask, lengths):
            fused = torch.tanh(self.fusionLayer(torch.cat(feats, 2)))
            return self.Transformer(fused, mask, lengths)

    sdims = {'image': 256, 'linguistic': 300}                             # SFT/train.py:533, SFT/models.py:90
    m = SFTPath(556, 2); load_filled(m, 10)
    inputs, mask, target, lengths = fill.make_batch(3, 8, sdims, 10)
    res = run_model(m, [t(inputs['image']), t(inputs['linguistic'])], mask, target, lengths)
    np.savez(os.path.join(OUT, 'sft.npz'), **res)
    meta['sft'] = dict(B=3, T=8, N=2, dims=sdims, lengths=lengths, seed=10)

    # ---- B2-Trans body (UniFullTransformer) and UniTransformer -------------------------------
    m = mft.UniFullTransformer(556, N=2); load_filled(m, 11)
    inputs, mask, target, lengths = fill.make_batch(3, 8, {'x': 556}, 11)
    res = run_model(m, t(inputs['x']), mask, target, lengths)
    np.savez(os.path.join(OUT, 'unifull.npz'), **res)
    meta['unifull'] = dict(B=3, T=8, N=2, lengths=lengths, seed=11)

    m = mft.UniTransformer(300, N=2); load_filled(m, 12)
    inputs, mask, target, lengths = fill.make_batch(3, 8, {'x': 300}, 12)
    res = run_model(m, t(inputs['x']), mask, target, lengths)
    np.savez(os.path.join(OUT, 'uni.npz'), **res)
    meta['uni'] = dict(B=3, T=8, N=2, lengths=lengths, seed=12)

    # ---- state_dict key/shape inventory (checkpoint-compat contract) -------------------
    inv = {}
    m = mft.MultiTransformer(mods, dims); inv['MFT.MultiTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    m = b3.MultiTransformer(mods, b3dims); inv['B3.MultiTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    m = sft.NLPTransformer(512); inv['SFT.NLPTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    m = mft.UniFullTransformer(556); inv['MFT.UniFullTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    m = mft.UniTransformer(300); inv['MFT.UniTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, 'state_dict_keys.json'), 'w') as f:
        json.dump(inv, f)

    # ---- eval_ccc known answers (the reference's only published KAT) ---------------------------
    import csv
    kat = []
    for model, vid, line in [('MFT', '173_4', 535), ('MFT', '165_2', 574), ('SFT', '173_4', 1307), ('SFT', '165_2', 1346)]:
        with open(os.path.join(REF, 'PredSave', f'{model}{vid}.csv')) as f:
            rows = list(csv.DictReader(f))
        with open(os.path.join(REF, 'PerfSave', f'{model}.csv')) as f:
            perf = f.read().splitlines()[line - 1].split(',')
        assert perf[2] == vid, perf
        kat.append(dict(model=model, vid=vid, perf_line=line, ccc=float(perf[4]),
                        pred=[float(r['pred']) for r in rows], actual=[float(r['actual']) for r in rows]))
    with open(os.path.join(OUT, 'ccc_kat.json'), 'w') as f:
        json.dump(kat, f)
    with open(os.path.join(OUT, 'meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
