"""One eager MFT train step (B=256, T=128, N=6, bf16) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import synthetic as fill
from multimodal_transformer_b200.training import FlatAdam, train_step_loss

MODS = ['acoustic', 'image', 'linguistic']; DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
B, T, N = int(os.environ.get('B', 256)), int(os.environ.get('T', 128)), 6
dev = torch.device('cuda', 0)
mtb.set_compute_dtype('bf16')
torch.manual_seed(1)
model = mtb.MultiTransformer(MODS, DIMS, N=N, device=dev).to(dev)
opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 1)
x = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items()}
m, tg = torch.from_numpy(mask).to(dev), torch.from_numpy(target).to(dev)
norm = float(sum(lengths))

def step():
    model.train()
    pred = model(x, m, lengths)
    loss = train_step_loss(pred, tg, norm)
    opt.step(); opt.zero_grad()

for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
