"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): smoke() in both dtypes, one bf16 MFT train step at T = 128 (tcgen05
attention forward / backward, tcgen05 GEMMs in every mode, tensor-core recurrences, fused loss + Adam) and one many-tiles-per-CTA GEMM.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib, synthetic as fill
from multimodal_transformer_b200.training import FlatAdam, train_step_loss

g.smoke()
L = _lib.lib()
MODS = ['acoustic', 'image', 'linguistic']; DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
dev = torch.device('cuda:0')
mtb.set_compute_dtype('bf16')
torch.manual_seed(1)
for (B, T) in ((8, 128), (3, 50)):
    model = mtb.MultiTransformer(MODS, DIMS, N=2, device=dev).to(dev).train()
    opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
    inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 3)
    x = {k: torch.from_numpy(v).to(dev) for k, v in inputs.items()}
    for _ in range(2):
        pred = model(x, torch.from_numpy(mask).to(dev), lengths)
        loss = train_step_loss(pred, torch.from_numpy(target).to(dev), float(sum(lengths)))
        opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    print(f'train step B={B} T={T}: loss {loss.item():.5f}')
for mode in (2, 1, 0):
    L.mt_gemm_tc_mode(mode)
    M, N, K = 20000, 768, 256
    A = torch.randn(M, K, device=dev).bfloat16(); Bm = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
    _lib.check(L.mt_gemm(1, M, N, K, _lib.ptr(A), K, 1, _lib.ptr(Bm), K, 1, _lib.ptr(C), N, 0, _lib.ptr(bias), 0, 1, _lib.stream()))
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().t() + bias
    print(f'gemm mode {mode}: rel err {((C.float() - ref).abs().max() / ref.abs().max()).item():.2e}')
L.mt_gemm_tc_mode(2)
print('sanitize target done')
