"""Run the raw-window eager train steps of tests/test_gpu_frontend.py::test_graphed_train_step_with_window_front_end twice and report the
largest parameter difference between the two runs (atomic reduction order is the only source of nondeterminism)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import models as M
from multimodal_transformer_b200.training import FlatAdam, train_step_loss
from oracle import fill
from tests import util
from tests.test_gpu_frontend import front_meta, front_inventory

DEV = 'cuda:0'
t = lambda a: torch.from_numpy(np.asarray(a))
m = front_meta()['front_mft']
shapes = {k: tuple(v) for k, v in m['shapes'].items()}
dims = {k: v[1] for k, v in shapes.items()}
inv = front_inventory()['MFT.MultiCNNTransformer']
sd = util.filled_sd({k: tuple(s) for k, s in inv.items()}, 9)
B, T = 4, 10
batches = [fill.make_raw_batch(B, T, shapes, 80 + i) for i in range(3)]
for grouped in (True, False):
    mtb.set_grouped_stacks(grouped)
    runs = []
    for rep in range(3):
        model = M.MultiCNNTransformer(m['mods'], dims, m['embed_dims']); model.load_state_dict(sd)
        opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
        model.eval()
        for inputs, mask, target, lengths in batches:
            pred = model({k: t(v).to(DEV) for k, v in inputs.items()}, lengths, t(mask).to(DEV))
            train_step_loss(pred, t(target).to(DEV), float(sum(lengths)))
            opt.step(); opt.zero_grad()
        runs.append({k: p.detach().clone() for k, p in model.named_parameters()})
    worst = max(((runs[0][k] - runs[r][k]).abs().max().item() / max(runs[0][k].abs().max().item(), 1e-9), k) for k in runs[0] for r in (1, 2))
    print('grouped' if grouped else 'per-stack', 'worst relative parameter difference between identical eager runs:', worst)
