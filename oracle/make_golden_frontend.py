"""Generate tests/golden/front_*.npz by running the UNMODIFIED reference `models.py` classes (CNN, Highway and the
MultiCNNTransformer of MFT / SFT / B2-Trans / B3-MFN, imported from /root/reference, build container only) on deterministic
weights and raw window inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden_frontend
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

from . import fill
from .make_golden import OUT, REF, digest_params, load_filled, t


def _load_models(dirname, alias):
    """Import <REF>/<dirname>/models.py; it does `from multiTransformer import ...`, so that directory's multiTransformer.py is
    made importable under that bare name for the duration of the import (matplotlib stubbed: not installed, not used)."""
    if 'matplotlib' not in sys.modules:
        m = types.ModuleType('matplotlib'); mp = types.ModuleType('matplotlib.pyplot'); m.pyplot = mp
        sys.modules['matplotlib'] = m; sys.modules['matplotlib.pyplot'] = mp
    d = os.path.join(REF, dirname)
    sys.modules.pop('multiTransformer', None)
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location(alias, os.path.join(d, 'models.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
        sys.modules.pop('multiTransformer', None)
    return mod


def _run(model, inputs, mask, target, lengths):
    model.eval()
    pred = model({k: t(v) for k, v in inputs.items()}, lengths, t(mask))
    loss = ((pred - t(target)) ** 2).sum() / sum(lengths)
    loss.backward()
    res = {'pred': pred.detach().numpy(), 'loss': np.array(loss.item())}
    res.update(digest_params(model))
    return res


def single_modality():
    """MFT MultiCNNTransformer with ONE modality (MFT/models.py:102-105: UniTransformer body) -> front_uni.npz + inventory entry."""
    warnings.filterwarnings('ignore')
    torch.set_num_threads(8)
    with contextlib.redirect_stdout(io.StringIO()):
        mft = _load_models('MFT', 'ref_models_mft1')
        m = mft.MultiCNNTransformer(['linguistic'], {'linguistic': 300}, {'linguistic': 300}, device=torch.device('cpu'))
    load_filled(m, 26)
    shapes = {'linguistic': (4, 300)}
    inputs, mask, target, lengths = fill.make_raw_batch(2, 6, shapes, 26)
    np.savez(os.path.join(OUT, 'front_uni.npz'), **_run(m, inputs, mask, target, lengths))
    with open(os.path.join(OUT, 'front_meta.json')) as f:
        meta = json.load(f)
    meta['front_uni'] = dict(B=2, T=6, mods=['linguistic'], shapes=shapes, embed_dims={'linguistic': 300}, lengths=lengths, seed=26)
    with open(os.path.join(OUT, 'front_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(OUT, 'front_state_dict_keys.json')) as f:
        inv = json.load(f)
    inv['MFT.MultiCNNTransformer.single'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, 'front_state_dict_keys.json'), 'w') as f:
        json.dump(inv, f)
    print('single-modality golden written')


def main():
    warnings.filterwarnings('ignore')
    torch.set_num_threads(8)
    cpu = torch.device('cpu')
    meta = {}
    with contextlib.redirect_stdout(io.StringIO()):
        mft = _load_models('MFT', 'ref_models_mft')
        sft = _load_models('SFT', 'ref_models_sft')
        b2 = _load_models('B2-Trans', 'ref_models_b2')
        b3 = _load_models('B3-MFN', 'ref_models_b3')

    # ---- CNN and Highway alone (kernel sizes 2 and 3, widths that are not multiples of 4 / 8) ---------------------
    prim = {}
    for tag, (n, K, D, E, k) in {'a': (7, 5, 12, 8, 2), 'b': (5, 6, 10, 6, 3), 'c': (4, 2, 88, 88, 2), 'd': (3, 4, 9, 7, 2)}.items():
        cnn = mft.CNN(D, E, k); load_filled(cnn, 20)
        hw = mft.Highway(E); load_filled(hw, 21)
        x = t(fill.fill_array('front_x_' + tag, (n, K, D), 20) * 3.0).requires_grad_(False)
        c = cnn(x.permute(0, 2, 1))
        c2 = c.detach().clone().requires_grad_(True)
        y = hw(c2)
        w = t(fill.fill_array('front_w_' + tag, (n, E), 20))
        (y * w).sum().backward()
        (cnn(x.permute(0, 2, 1)) * w).sum().backward()
        prim[tag + '_c'] = c.detach().numpy(); prim[tag + '_y'] = y.detach().numpy(); prim[tag + '_dc'] = c2.grad.numpy()
        for k_, p in list(cnn.named_parameters()) + list(hw.named_parameters()):
            prim[f'{tag}_grad:{k_}'] = p.grad.numpy()
        meta['prim_' + tag] = dict(n=n, K=K, D=D, E=E, k=k)
    np.savez(os.path.join(OUT, 'front_prims.npz'), **prim)

    # ---- full MultiCNNTransformer variants on raw windows -----------------------------------------------------
    mods = ['acoustic', 'image', 'linguistic']
    shapes = {'acoustic': (2, 88), 'image': (2, 1000), 'linguistic': (5, 300)}
    dims = {m: s[1] for m, s in shapes.items()}
    embed_dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}                  # MFT/train.py:552
    with contextlib.redirect_stdout(io.StringIO()):
        m = mft.MultiCNNTransformer(mods, dims, embed_dims, device=cpu)
    load_filled(m, 22)
    inputs, mask, target, lengths = fill.make_raw_batch(2, 6, shapes, 22)
    np.savez(os.path.join(OUT, 'front_mft.npz'), **_run(m, inputs, mask, target, lengths))
    meta['front_mft'] = dict(B=2, T=6, mods=mods, shapes=shapes, embed_dims=embed_dims, lengths=lengths, seed=22)
    inv = {'MFT.MultiCNNTransformer': {k: list(v.shape) for k, v in m.state_dict().items()}}

    smods = ['image', 'linguistic']                                               # SFT/train.py:533
    sshapes = {'image': (2, 1000), 'linguistic': (4, 300)}
    sdims = {m_: s[1] for m_, s in sshapes.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        m = sft.MultiCNNTransformer(smods, sdims, device=cpu)
    load_filled(m, 23)
    inputs, mask, target, lengths = fill.make_raw_batch(2, 5, sshapes, 23)
    np.savez(os.path.join(OUT, 'front_sft.npz'), **_run(m, inputs, mask, target, lengths))
    meta['front_sft'] = dict(B=2, T=5, mods=smods, shapes=sshapes, lengths=lengths, seed=23)
    inv['SFT.MultiCNNTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}

    with contextlib.redirect_stdout(io.StringIO()):
        m = b2.MultiCNNTransformer(smods, sdims, device=cpu)
    load_filled(m, 24)
    inputs, mask, target, lengths = fill.make_raw_batch(2, 5, sshapes, 24)
    np.savez(os.path.join(OUT, 'front_b2.npz'), **_run(m, inputs, mask, target, lengths))
    meta['front_b2'] = dict(B=2, T=5, mods=smods, shapes=sshapes, lengths=lengths, seed=24)
    inv['B2.MultiCNNTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}

    with contextlib.redirect_stdout(io.StringIO()):
        m = b3.MultiCNNTransformer(mods, dims, device=cpu)
    load_filled(m, 25)
    inputs, mask, target, lengths = fill.make_raw_batch(2, 7, shapes, 25)
    np.savez(os.path.join(OUT, 'front_b3.npz'), **_run(m, inputs, mask, target, lengths))
    meta['front_b3'] = dict(B=2, T=7, mods=mods, shapes=shapes, lengths=lengths, seed=25)
    inv['B3.MultiCNNTransformer'] = {k: list(v.shape) for k, v in m.state_dict().items()}

    with open(os.path.join(OUT, 'front_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    with open(os.path.join(OUT, 'front_state_dict_keys.json'), 'w') as f:
        json.dump(inv, f)
    print('front-end golden vectors written to', OUT)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'single':
        single_modality()            # adds front_uni.npz without rewriting the other fixtures
    else:
        main()
        single_modality()
