"""GEMM engine: fixed cost vs per-tile cost (N=256, K=256, bf16 out), 10 launches back to back between two events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib
L = _lib.lib(); dev = 'cuda:0'
def run(m, n, k, cf=0, reps=10):
    A = torch.randn(m, k, device=dev).bfloat16(); B = torch.randn(n, k, device=dev).bfloat16()
    C = torch.zeros(m, n, device=dev, dtype=torch.float32 if cf else torch.bfloat16); bias = torch.randn(n, device=dev)
    def go():
        _lib.check(L.mt_gemm(1, m, n, k, _lib.ptr(A), k, 1, _lib.ptr(B), k, 1, _lib.ptr(C), n, cf, _lib.ptr(bias), 0, 1, _lib.stream()))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for m in (128, 128 * 37, 128 * 74, 128 * 148, 128 * 296, 32768, 65536):
    for (n, k) in ((256, 256), (128, 256), (768, 256)):
        t = run(m, n, k)
        tiles = ((m + 127) // 128) * ((n + 127) // 128)
        print(f'M={m:6d} N={n} K={k} tiles={tiles:5d} ({tiles/148:5.2f}/SM): {t:7.1f} us/launch')
