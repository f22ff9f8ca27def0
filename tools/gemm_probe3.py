"""GEMM engine timed inside a CUDA graph (no host launch gaps): fixed cost per launch vs per-tile cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib
L = _lib.lib(); dev = 'cuda:0'
REPS = 20
if len(sys.argv) > 1: print('mode was', L.mt_gemm_tc_mode(int(sys.argv[1])), '-> now', sys.argv[1])
def run(m, n, k, cf=0, akm=1, bkm=1, split=1):
    A = torch.randn((m, k) if akm else (k, m), device=dev).bfloat16(); B = torch.randn((n, k) if bkm else (k, n), device=dev).bfloat16()
    C = torch.zeros(m, n, device=dev, dtype=torch.float32 if cf else torch.bfloat16); bias = torch.randn(n, device=dev)
    def go():
        _lib.check(L.mt_gemm(1, m, n, k, _lib.ptr(A), k if akm else m, akm, _lib.ptr(B), k if bkm else n, bkm, _lib.ptr(C), n, cf,
                             None if split > 1 else _lib.ptr(bias), 0, split, _lib.stream()))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        go(); go()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(REPS): go()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / REPS
for (n, k, cf) in ((768, 256, 0), (256, 256, 1), (128, 256, 0), (256, 128, 1)):
    for m in (128, 32768, 131072):
        t = run(m, n, k, cf)
        tiles = ((m + 127) // 128) * ((n + 127) // 128)
        byts = (m * k + n * k) * 2 + m * n * (4 if cf else 2)
        print(f'M={m:6d} N={n} K={k} cf={cf} tiles={tiles:5d} ({tiles/148:5.2f}/SM): {t:7.1f} us/launch  {byts/t/1e3:7.0f} GB/s')
for (m, n) in ((768, 256), (256, 256), (128, 256), (256, 128)):
    t = run(m, n, 32768, 1, 0, 0, 8)
    print(f'wgrad M={m} N={n} K=32768: {t:7.1f} us/launch  {(m+n)*32768*2/t/1e3:7.0f} GB/s')
