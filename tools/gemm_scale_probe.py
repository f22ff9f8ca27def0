"""How the tcgen05 GEMM engine scales with M on the encoder's shapes (one modality stack = 32768 rows, three grouped = 98304), per
engine mode (mt_gemm_tc_mode: 2 = default, 1 = one weight-resident CTA per SM).  CUDA events, back-to-back launches on buffers
larger than L2.  Usage: python tools/gemm_scale_probe.py [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib

L = _lib.lib()
dev = 'cuda:0'
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
# (name, N, K, a_kmajor, b_kmajor, c_f32, wgrad)
shapes = [('qkv', 768, 256, 1, 1, 0, 0), ('oproj', 256, 256, 1, 1, 1, 0), ('ffn1', 128, 256, 1, 1, 0, 0), ('ffn2', 256, 128, 1, 1, 1, 0),
          ('dgrad_qkv', 256, 768, 1, 0, 0, 0), ('dgrad_o', 256, 256, 1, 0, 0, 0), ('wgrad_qkv', 768, 256, 0, 0, 1, 1), ('wgrad_o', 256, 256, 0, 0, 1, 1),
          ('wgrad_w1', 128, 256, 0, 0, 1, 1)]
out = {}
for mode in (2, 1):
    L.mt_gemm_tc_mode(mode)
    for (name, n, k, akm, bkm, cf, wg) in shapes:
        for M in (32768, 98304):
            if wg:      # dW[n, k] = dy^T x: contraction over the M tokens
                m_, n_, k_ = n, k, M
            else:
                m_, n_, k_ = M, n, k
            nbuf = 4      # rotate buffers so that consecutive launches do not hit L2
            As = [torch.randn((m_, k_) if akm else (k_, m_), device=dev).bfloat16() for _ in range(nbuf)]
            B = torch.randn((n_, k_) if bkm else (k_, n_), device=dev).bfloat16()
            Bs = [B] if not wg else [torch.randn((k_, n_), device=dev).bfloat16() for _ in range(nbuf)]
            Cs = [torch.zeros(m_, n_, device=dev, dtype=torch.float32 if cf else torch.bfloat16) for _ in range(nbuf if not wg else 1)]
            bias = torch.randn(n_, device=dev)
            def run(i):
                A = As[i % nbuf]; Bm = Bs[i % len(Bs)]; C = Cs[i % len(Cs)]
                _lib.check(L.mt_gemm(1, m_, n_, k_, _lib.ptr(A), k_ if akm else m_, akm, _lib.ptr(Bm), k_ if bkm else n_, bkm, _lib.ptr(C), n_, cf,
                                     None if wg else _lib.ptr(bias), 0, 8 if wg else 1, _lib.stream()))
            for i in range(4): run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps): run(i)
            e1.record(); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e3 / reps
            byts = (m_ * k_ + n_ * k_) * 2 + m_ * n_ * (4 if cf else 2)
            out[f'mode{mode}_{name}_M{M}'] = {'us': round(t, 2), 'tflops': round(2.0 * m_ * n_ * k_ / t / 1e6, 1), 'gbs': round(byts / t / 1e3, 1)}
            print(f'mode{mode} {name:10s} M{M:6d}: {t:8.2f} us  {2.0*m_*n_*k_/t/1e6:7.1f} TFLOP/s  {byts/t/1e3:7.1f} GB/s', flush=True)
L.mt_gemm_tc_mode(2)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/gemm_scale_probe.json', 'w'), indent=1)
