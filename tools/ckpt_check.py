"""Checkpoint round trip after eager training with torch.optim.Adam: state_dict of the arena-bound parameters saves / reloads exactly."""
import os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import synthetic
mods = ['acoustic', 'image', 'linguistic']; dims = {'acoustic': 88, 'image': 256, 'linguistic': 300}
dev = torch.device('cuda:0')
model = mtb.MultiTransformer(mods, dims, N=2, device=dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
inputs, mask, target, lengths = synthetic.make_batch(4, 16, dims, 3)
x = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items()}; m = torch.from_numpy(mask).to(dev); tg = torch.from_numpy(target).to(dev)
for _ in range(3):
    out = model(x, m, lengths); loss = ((out - tg) ** 2).sum() / sum(lengths)
    opt.zero_grad(); loss.backward(); opt.step()
buf = io.BytesIO(); torch.save({'model': model.state_dict()}, buf); print('checkpoint bytes', buf.tell())
buf.seek(0); ck = torch.load(buf, map_location='cpu')
model2 = mtb.MultiTransformer(mods, dims, N=2, device=dev); model2.load_state_dict(ck['model'])
model.eval(); model2.eval()
with torch.no_grad():
    a = model(x, m, lengths); b = model2(x, m, lengths)
print('reload max diff', (a - b).abs().max().item(), 'keys', len(ck['model']))
assert torch.equal(a, b)
