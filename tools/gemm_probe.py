"""Times the GEMM engine on the encoder's shapes through the C ABI (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib

L = _lib.lib()
dev = 'cuda:0'
M = 32768
shapes = [  # (M, N, K, akm, bkm, c_f32, residual, split)
    (M, 768, 256, 1, 1, 0, 0, 1), (M, 256, 256, 1, 1, 1, 1, 1), (M, 128, 256, 1, 1, 0, 0, 1), (M, 256, 128, 1, 1, 1, 1, 1),
    (M, 256, 768, 1, 0, 0, 0, 1), (M, 256, 256, 1, 0, 0, 0, 1), (768, 256, M, 0, 0, 1, 0, 8), (256, 256, M, 0, 0, 1, 0, 8)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (m, n, k, akm, bkm, cf, res, split) in shapes:
    A = torch.randn((m, k) if akm else (k, m), device=dev).bfloat16()
    B = torch.randn((n, k) if bkm else (k, n), device=dev).bfloat16()
    C = torch.zeros(m, n, device=dev, dtype=torch.float32 if cf else torch.bfloat16)
    bias = torch.randn(n, device=dev)
    ts = []
    for r in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.mt_gemm(1, m, n, k, _lib.ptr(A), k if akm else m, akm, _lib.ptr(B), k if bkm else n, bkm, _lib.ptr(C), n, cf,
                             None if split > 1 else _lib.ptr(bias), 0, split, _lib.stream()))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    byts = (m * k + n * k) * 2 + m * n * (4 if cf else 2)
    print(f'm{m} n{n} k{k} {"K" if akm else "M"}{"K" if bkm else "M"} c_f32={cf} split={split}: {t:8.1f} us  {2*m*n*k/t/1e6:7.1f} TFLOP/s  {byts/t/1e3:7.1f} GB/s')
