"""GPU (B200): the drop-in claim, demonstrated with the reference's OWN files.  baseline/_ref (tools/install_ref.sh) holds verbatim copies
of MFT/models.py and MFT/train.py; they are imported with the bare module name `multiTransformer` bound to this repository's module --
the one-line switch INTEGRATION.md describes -- and the reference's own train() (MFT/train.py:110-155) then drives the reference's own
MultiCNNTransformer (MFT/models.py:81-138) whose hot path runs on libmt_b200.so.  The same code with the reference's own
multiTransformer.py on the CPU is the comparator."""
import argparse
import random

import numpy as np
import pytest
import torch

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib
from multimodal_transformer_b200 import multiTransformer as ours
from oracle import ref_loader as R
from multimodal_transformer_b200 import synthetic as fill

pytestmark = pytest.mark.gpu
MODS = ['acoustic', 'image', 'linguistic']
RAW = {'acoustic': (2, 88), 'image': (2, 1000), 'linguistic': (33, 300)}       # (K vectors, D) per window, MFT/train.py:571
EMBED = {'acoustic': 88, 'image': 256, 'linguistic': 300}                      # window_embed_size, MFT/train.py:552


@pytest.fixture(autouse=True)
def _fp32_mode():
    # the front-end of this test is the REFERENCE's torch CNN / Highway running on the GPU: keep cuDNN / cuBLAS off TF32 so that the
    # comparison measures this repository's hot path, not torch's default conv precision
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mtb.set_compute_dtype('fp32')
    yield
    mtb.set_compute_dtype('fp32')
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def _corpus(n, T, seed):
    """python-list corpus in the layout generateTrainBatch consumes: input_data[mod][narrative] = [T][K][D] nested lists (zero rows
    past the narrative's length, padInput MFT/train.py:456-505), input_target[narrative] = [T], lengths."""
    inputs, mask, target, lengths = fill.make_raw_batch(n, T, RAW, seed)
    order = list(range(n))
    random.Random(seed).shuffle(order)                # the corpus is not length-sorted; the batcher sorts each chunk
    data = {m: [inputs[m][i].tolist() for i in order] for m in MODS}
    tgt = [target[i, :, 0].tolist() for i in order]
    return data, tgt, [lengths[i] for i in order]


@pytest.mark.skipif(not R.available(), reason='baseline/_ref missing: run tools/install_ref.sh in the build container')
def test_reference_train_py_drives_the_b200_hot_path():
    ref_models = R.load('MFT', 'models')                               # reference models.py + reference multiTransformer.py
    drop_models = R.load('MFT', 'models', hot_path=ours)               # reference models.py + THIS repository's multiTransformer
    ref_train = R.load('MFT', 'train')                                 # reference train.py: train(), generateTrainBatch()
    assert drop_models.MultiTransformer is ours.MultiTransformer
    assert ref_models.MultiTransformer.__module__ == 'multiTransformer' and ref_models.MultiTransformer is not ours.MultiTransformer

    dims = {m: RAW[m][1] for m in MODS}
    torch.manual_seed(1)                                               # MFT/train.py:524
    ref = R.cpu_instance(ref_models.MultiCNNTransformer(MODS, dims, EMBED))
    drop = drop_models.MultiCNNTransformer(MODS, dims, EMBED, device=torch.device('cuda:0'))
    assert type(drop.Transformer) is ours.MultiTransformer
    assert list(drop.state_dict()) == list(ref.state_dict())           # checkpoint compatibility (MFT/train.py:345-351)
    drop.load_state_dict(ref.state_dict())
    _no_dropout(ref); _no_dropout(drop)                                # dropout streams differ by construction: compare the deterministic path

    data, tgt, lengths = _corpus(6, 10, 3)
    crit = torch.nn.MSELoss(reduction='sum')                           # MFT/train.py:536
    launches0 = _lib.lib().mt_launch_count()

    def evaluate(model, dev):
        model.eval()
        random.seed(5)
        outs = []
        with torch.no_grad():
            for d, t_, mask, lens in ref_train.generateTrainBatch(data, tgt, list(lengths), None, batch_size=4, onEval=True):
                outs.append(model({k: v.to(dev) for k, v in d.items()}, lens, mask.to(dev)).cpu())
        return outs

    e_ref, e_drop = evaluate(ref, 'cpu'), evaluate(drop, 'cuda:0')
    for a, b in zip(e_drop, e_ref):
        assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item(), 'eval() forward through the reference models.py differs'

    # the reference's train() for two epochs (one batch each: its batch size is the generateTrainBatch default 25) on both
    losses = {}
    for name, model, dev in (('ref', ref, 'cpu'), ('drop', drop, 'cuda:0')):
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)       # MFT/train.py:557
        args = argparse.Namespace(device=torch.device(dev))
        random.seed(11)
        losses[name] = [float(ref_train.train(data, tgt, list(lengths), model, crit, opt, ep, args)) for ep in range(2)]
    assert _lib.lib().mt_launch_count() > launches0                    # the drop-in really ran libmt_b200 kernels
    for a, b in zip(losses['drop'], losses['ref']):
        assert abs(a - b) <= 1e-4 * abs(b), losses
    assert losses['ref'][1] != losses['ref'][0]
    # after two Adam steps both models still agree (Adam's 1/sqrt(v) amplifies round-off of near-zero gradients: a few lr per weight)
    e_ref2, e_drop2 = evaluate(ref, 'cpu'), evaluate(drop, 'cuda:0')
    for a, b, b0 in zip(e_drop2, e_ref2, e_ref):
        assert (b - b0).abs().max().item() > 0                         # training moved the model
        assert (a - b).abs().max().item() <= 2e-3 * b.abs().max().item()
    sd_r, sd_d = ref.state_dict(), drop.state_dict()
    for k in sd_r:
        assert (sd_d[k].cpu() - sd_r[k]).abs().max().item() <= 4.5e-4, k     # <= 2 steps x lr 1e-4 x (1 + sign flips) + decay
