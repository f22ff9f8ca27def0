"""Per-tile timeline of CTA 0 of one tcgen05 GEMM (mt_gemm_debug_trace): clock64 deltas between pipeline events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib
L = _lib.lib(); dev = 'cuda:0'
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(m, k, device=dev).bfloat16(); B = torch.randn(n, k, device=dev).bfloat16()
C = torch.zeros(m, n, device=dev, dtype=torch.bfloat16); bias = torch.randn(n, device=dev)
buf = torch.zeros(64 * 8 + 512, dtype=torch.int64, device=dev)
def go():
    _lib.check(L.mt_gemm(1, m, n, k, _lib.ptr(A), k, 1, _lib.ptr(B), k, 1, _lib.ptr(C), n, 0, _lib.ptr(bias), 0, 1, _lib.stream()))
go(); go(); torch.cuda.synchronize()
_lib.check(L.mt_gemm_debug_trace(_lib.ptr(buf)))
go(); torch.cuda.synchronize()
_lib.check(L.mt_gemm_debug_trace(None))
kbt = buf.cpu()[512:640].view(16, 4, 2)
iss = buf.cpu()[640:704].view(16, 4)
t = buf.cpu()[:512].view(64, 8)
t0 = int(t[0, 0])
print('tile: tma_issue mma_start ops_landed last_kb_landed epi_sees_acc acc_released last_pass   (clocks since first TMA issue)')
for i in range(64):
    if int(t[i, 0]) == 0: break
    print(i, ' '.join(f'{int(v) - t0:8d}' for v in t[i, :7]))
print('per k-block (tma issue, landed, mma issued+committed) of the first tiles:')
for i in range(6):
    print(i, ' '.join(f'({int(kbt[i, k, 0]) - t0:6d},{int(kbt[i, k, 1]) - t0:6d},{int(iss[i, k]) - t0:6d})' for k in range(4)))
