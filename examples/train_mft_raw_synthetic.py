"""The reference's epoch loop (MFT/train.py:109-155 train, :203-257 evaluate) on the drop-in `models.py` level: raw window
vectors -> MultiCNNTransformer (window CNN + Highway front-end, encoder stacks, MFN), batches from the GPU-side batcher, batched
evaluation with on-device CCC.  Synthetic SEND-shaped corpus (a learnable target: a fixed random projection of the windows).

    python examples/train_mft_raw_synthetic.py [--bf16] [--epochs 3] [--narratives 100] [--batch 25]
"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import multimodal_transformer_b200 as mtb

ap = argparse.ArgumentParser()
ap.add_argument('--bf16', action='store_true')
ap.add_argument('--epochs', type=int, default=3)
ap.add_argument('--narratives', type=int, default=100)
ap.add_argument('--batch', type=int, default=25)           # MFT/train.py:74
ap.add_argument('--seq', type=int, default=64)
args = ap.parse_args()

device = torch.device('cuda:0')
torch.manual_seed(1); random.seed(1)                        # MFT/train.py:524-526
mods = ['acoustic', 'image', 'linguistic']                  # MFT/train.py:544-549
shapes = {'acoustic': (2, 88), 'image': (2, 1000), 'linguistic': (12, 300)}      # (vectors per window, dim); :571
dims = {m: s[1] for m, s in shapes.items()}
embed_dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}                    # :552
mtb.set_compute_dtype('bf16' if args.bf16 else 'fp32')

# ---- a padded corpus in the reference's layout (padInput / padRating output), packed ONCE onto the device ----------------------
N, T = args.narratives, args.seq
lengths = [T] + torch.randint(T // 4, T + 1, (N - 1,)).tolist()
data = {m: torch.randn(N, T, K, D) for m, (K, D) in shapes.items()}
proj = torch.randn(88) / 88 ** 0.5
target = torch.sigmoid(data['acoustic'].mean(2) @ proj)     # ratings in (0, 1), a function of the acoustic windows
for b, l in enumerate(lengths):
    for m in mods:
        data[m][b, l:] = 0
    target[b, l:] = 0
n_train = int(0.8 * N)
train = mtb.DeviceCorpus({m: v[:n_train] for m, v in data.items()}, target[:n_train], lengths[:n_train], device)
valid = mtb.DeviceCorpus({m: v[n_train:] for m, v in data.items()}, target[n_train:], lengths[n_train:], device)

model = mtb.MultiCNNTransformer(mods=mods, dims=dims, embed_dims=embed_dims, device=device)      # MFT/train.py:553
criterion = nn.MSELoss(reduction='sum')                     # :556
optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)                     # :557

for epoch in range(1, args.epochs + 1):
    model.train()
    t0 = time.perf_counter()
    loss_sum, n_points = 0.0, 0
    for batch, tgt, mask, lens in mtb.generateTrainBatch(train, None, None, None, batch_size=args.batch):     # :117
        output = model(batch, lens, mask)                   # :133
        loss = criterion(output, tgt) / sum(lens)           # :135-139
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()                                    # :141-143
        loss_sum += loss.item() * sum(lens); n_points += sum(lens)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # evaluation: many narratives per forward (ragged_batch), CCC / Pearson on the device -- evaluate() MFT/train.py:203-257
    _, vloss, stats, best = mtb.evaluate(model, valid, batch_size=64)
    print(f'epoch {epoch}: train loss {loss_sum / n_points:.5f} ({n_train / dt:.0f} narratives/s)  '
          f'valid loss {vloss:.5f} corr {stats["corr"]:.3f} ccc {stats["ccc"]:.4f} max ccc {stats["max_ccc"]:.4f}')
