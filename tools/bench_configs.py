"""Secondary measurements of the other BASELINE.json configurations (not the bench.py headline line):
C4  B3-MFN (no encoder) over 1024-window sequences, B = 256: persistent-kernel recurrence stress, train step + inference
C5' one encoder stack at d_model 512, 8 heads, d_ff 256, N = 6, T = 4096 (the any-T tensor-core attention engine), forward + backward
One JSON object on stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import synthetic as fill
from multimodal_transformer_b200.training import FlatAdam, GraphedForward, GraphedTrainStep

dev = torch.device('cuda', 0)
mtb.set_compute_dtype('bf16')
MODS = ['acoustic', 'image', 'linguistic']
out = {}


def timed(fn, warm=3, steps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ---- C4 -------------------------------------------------------------------------------------------------
B, T = 256, 1024
dims = {'acoustic': 256, 'image': 256, 'linguistic': 300}
torch.manual_seed(1)
model = mtb.B3MultiTransformer(MODS, dims, device=dev).to(dev)
opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
inputs, mask, target, lengths = fill.make_batch(B, T, dims, 1)
x = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items()}
m, tg = torch.from_numpy(mask).to(dev), torch.from_numpy(target).to(dev)
g = GraphedTrainStep(model, opt, B, T, dims, dev)
g.load(x, m, tg, lengths); g.capture()
ms = timed(g.replay)
gf = GraphedForward(model, B, T, dims, dev); gf.load(x, m); gf.capture()
ms_inf = timed(gf.graph.replay)
out['C4_b3_mfn_T1024_B256'] = dict(train_ms=ms, train_narratives_per_s=B / ms * 1e3, inference_ms=ms_inf, inference_narratives_per_s=B / ms_inf * 1e3,
                                   recurrence_steps_per_s_train=T / ms * 1e3)
del g, gf, model, opt
torch.cuda.empty_cache()

# ---- C5' ------------------------------------------------------------------------------------------------
from multimodal_transformer_b200.multiTransformer import _make_encoder
d, dff, N, T, B = 512, 256, 6, 4096, 8
torch.manual_seed(1)
enc = _make_encoder(d, dff, 8, 0.1, N).to(dev).train()
xin = torch.randn(B, T, d, device=dev, requires_grad=True)
msk = torch.ones(B, T, 1, device=dev); msk[:, 3 * T // 4:] = 0
def step():
    y = enc(xin, msk)
    y.float().sum().backward()
ms = timed(step, 2, 5)
flops = 3 * B * T * N * (8 * d * d + 4 * T * d + 4 * d * dff)
out['C5_encoder_d512_T4096_B8'] = dict(fwd_bwd_ms=ms, tokens_per_s=B * T / ms * 1e3, tflops=flops / ms / 1e9)
print(json.dumps(out))
