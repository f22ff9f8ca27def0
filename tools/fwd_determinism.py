"""Run-to-run determinism of the bf16 train-mode MFT forward (same seed, same masks) under mt_tune presets: prints, per preset, the largest
difference between repeated forwards and the first one.  Every preset except 'no PDL' switches programmatic dependent launch ON (mt_tune
key 3 = 1; the library default is off) and then narrows it by kernel family (key 14) or removes the attention forward's early trigger
(key 15) -- this is the tool that bisected the launch race described in csrc/mt_common.cuh.
Usage: python tools/fwd_determinism.py [reps]   (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib
from oracle import fill
from tests import util

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
MODS = ['acoustic', 'image', 'linguistic']
dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
L = _lib.lib()
t = torch.from_numpy
for N, B, T in ((1, 40, 128), (2, 9, 128)):
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', N), 23)
    inputs, mask, target, lengths = fill.make_batch(B, T, dims, 23)
    presets = [('PDL, every family', {}), ('no PDL', {3: 0})]
    fam = {'ln_fwd': 1, 'ln_bwd': 2, 'gemm_rs': 4, 'gemm_tc': 8, 'attn_fwd': 16, 'attn_bwd': 32, 'misc': 64}
    presets += [(f'PDL only {k}', {14: v}) for k, v in fam.items()]
    presets += [(f'PDL all but {k}', {14: 0xff ^ v}) for k, v in fam.items() if k in ('ln_fwd', 'gemm_rs', 'attn_fwd')]
    presets += [('PDL, attn fwd never triggers early', {15: 1})]
    presets = [(n_, {**{3: 1}, **t_}) if n_ not in ('no PDL',) else (n_, t_) for n_, t_ in presets]      # the default is off: switch it on
    if N != 1:
        presets = presets[:2]
    for name, tune in presets:
        old = {k: L.mt_tune(k, v) for k, v in tune.items()}
        mtb.set_compute_dtype('bf16')
        model = mtb.MultiTransformer(MODS, dims, N=N).to('cuda:0').train(); model.load_state_dict(sd)
        x = {k: t(v).to('cuda:0') for k, v in inputs.items()}; m = t(mask).to('cuda:0')
        first, worst, nbad = None, 0.0, 0
        for r in range(reps):
            mtb.fix_seed(4711)
            pred = model(x, m, lengths)
            (((pred - t(target).to('cuda:0')) ** 2).sum() / sum(lengths)).backward()
            model.zero_grad()
            p = pred.detach().float().cpu()
            if first is None:
                first = p
            else:
                dlt = (p - first).abs().max().item()
                worst = max(worst, dlt); nbad += dlt > 0
        mtb.set_compute_dtype('fp32')
        for k, v in old.items():
            L.mt_tune(k, v)
        print(f'N={N} B={B} T={T} {name:24s} worst |d pred| over {reps} forwards: {worst:.3e}  ({nbad} differ)', flush=True)
