"""The reference's training loop (MFT/train.py:109-155) on the drop-in modules, with synthetic SEND-shaped batches.

Part 1 is the reference loop VERBATIM in structure -- model(data, mask, lengths) -> MSELoss(sum) / sum(lengths) -> backward ->
torch.optim.Adam -- only the import of MultiTransformer changed (and the CNN front-end is skipped: the hot path starts at the
window-level features).  Part 2 is the same step through GraphedTrainStep (one CUDA graph per step, fused loss + Adam,
input copy pipelined behind compute).

    python examples/train_mft_synthetic.py [--bf16] [--steps 20] [--batch 25]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import synthetic
from multimodal_transformer_b200.training import FlatAdam, GraphedTrainStep

ap = argparse.ArgumentParser()
ap.add_argument('--bf16', action='store_true')
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--batch', type=int, default=25)          # MFT/train.py:74
ap.add_argument('--seq', type=int, default=128)
args = ap.parse_args()

device = torch.device('cuda:0')
torch.manual_seed(1)                                        # MFT/train.py:524
mods = ['acoustic', 'image', 'linguistic']                  # MFT/train.py:544-549
dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}    # window_embed_size, MFT/train.py:552
mtb.set_compute_dtype('bf16' if args.bf16 else 'fp32')
model = mtb.MultiTransformer(mods, dims, device=device)
criterion = nn.MSELoss(reduction='sum')                     # MFT/train.py:556
optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)     # :557

batches = [synthetic.make_batch(args.batch, args.seq, dims, 100 + i) for i in range(4)]
to = lambda a: torch.from_numpy(a).to(device)

model.train()
t0 = time.perf_counter()
for step in range(args.steps):
    inputs, mask, target, lengths = batches[step % len(batches)]
    data = {k: to(v) for k, v in inputs.items()}
    mask_t, target_t = to(mask), to(target)
    out = model(data, mask_t, lengths)                      # :133  (forward(inputs, mask, lengths))
    loss = criterion(out, target_t) / sum(lengths)          # :135-139
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()                                        # :141-143
    if step % 5 == 0 or step == args.steps - 1:
        print(f'[reference-style loop] step {step:3d} loss {loss.item():.5f}')
torch.cuda.synchronize()
print(f'[reference-style loop] {args.steps / (time.perf_counter() - t0) * args.batch:.0f} narratives/s (eager, torch.optim.Adam)')

# Drop every reference to the eager iterations' autograd graph (out / loss keep AccumulateGrad nodes alive that are bound to the
# default stream; CUDA-graph capture runs on its own stream and must not touch the legacy stream).
del out, loss
optimizer.zero_grad(set_to_none=True)

opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
gstep = GraphedTrainStep(model, opt, args.batch, args.seq, dims, device)
pinned = [({k: torch.from_numpy(v).pin_memory() for k, v in i.items()}, torch.from_numpy(m).pin_memory(), torch.from_numpy(t).pin_memory(), l)
          for i, m, t, l in batches]
gstep.prefetch(*pinned[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for step in range(args.steps):
    gstep.prefetch(*pinned[(step + 1) % len(pinned)])
    loss = gstep.step_prefetched()
    if step % 5 == 0 or step == args.steps - 1:
        print(f'[captured step]        step {step:3d} loss {loss.item():.5f}')
torch.cuda.synchronize()
print(f'[captured step]        {args.steps / (time.perf_counter() - t0) * args.batch:.0f} narratives/s (CUDA graph, fused loss + Adam)')
