"""Long-sequence attention (d = 512, h = 8, T = 4096: BASELINE configuration 5): tcgen05 flash kernels against the mma.sync tile kernels.
CUDA events, forward (and backward when the library has it).  Usage: python tools/attn_flash_probe.py [B] [T]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib
L = _lib.lib()
dev = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d, h = 512, 8
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.7).bfloat16()
mask = torch.ones(B, T, device=dev)
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, h, T, device=dev)
dout = torch.randn(B, T, d, device=dev).bfloat16(); dqkv = torch.empty_like(qkv)
ws = torch.empty(L.mt_attention_bwd_ws_bytes(B, T, h) + 64 * B * T * d + 1024, dtype=torch.uint8, device=dev)
res = {}
for name, force in (('tcgen05 flash', 0), ('mma.sync tiles', 1)):
    old = L.mt_attention_force_no_tc(force)
    def fwd():
        _lib.check(L.mt_attention_fwd(1, B, T, d, h, _lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(out), _lib.ptr(lse), 0.1, 7, 2, _lib.stream()))
    def bwd():
        _lib.check(L.mt_attention_bwd(1, B, T, d, h, _lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(out), _lib.ptr(lse), _lib.ptr(dout), _lib.ptr(dqkv), 0.1, 7, 2,
                                      _lib.ptr(ws), ws.numel(), _lib.stream()))
    for fn, nm, fl in ((fwd, 'fwd', 4.0), (bwd, 'bwd', 10.0)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[f'{name} {nm}'] = {'ms': round(ms, 3), 'tflops': round(fl * B * T * T * d / ms / 1e9, 1)}
        print(name, nm, res[f'{name} {nm}'], flush=True)
    L.mt_attention_force_no_tc(old)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/attn_flash_probe.json', 'w'), indent=1)
