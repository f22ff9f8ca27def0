"""Is the train step launch-bound?  Host enqueue time per step (no sync inside) vs device time per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib, synthetic as fill
from multimodal_transformer_b200.training import FlatAdam, train_step_loss

MODS = ['acoustic', 'image', 'linguistic']; DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
B, T, N = 256, 128, 6
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

for _ in range(3): step()
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = _lib.lib().mt_launch_count()
t0 = time.perf_counter(); e0.record()
for _ in range(K): step()
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
nl = (_lib.lib().mt_launch_count() - l0) / K
print(f'launches/step {nl:.0f}  host enqueue {1e3*(t1-t0)/K:.2f} ms/step  device {e0.elapsed_time(e1)/K:.2f} ms/step  wall incl. drain {1e3*(t2-t0)/K:.2f} ms/step'
      f'  -> {1e6*(t1-t0)/K/nl:.2f} us of host time per launch')
