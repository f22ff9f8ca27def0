"""Window front-end (CNN + max-pool + Highway + dropout) at the MFT-VAL raw-level batch: B narratives x T windows,
per modality (K vectors x D) -> E.  Device-timed forward (eval) and forward + backward (train) per modality, a per-kernel
breakdown from the library's launch profiler, and the reference-style CPU loop (oracle, one narrative per iteration as
MFT/models.py:117-132) on a bounded sample.  Writes one JSON object to stdout.

    python tools/bench_frontend.py [--B 256] [--T 128] [--dtype bf16] [--iters 10]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_transformer_b200 as mtb                      # noqa: E402
from multimodal_transformer_b200 import _lib, functional as K   # noqa: E402

MODS = {'acoustic': (2, 88, 88), 'image': (2, 1000, 256), 'linguistic': (33, 300, 300)}     # K, D, E  (MFT/train.py:552,571)


def params(D, E, k, dev):
    g = torch.Generator().manual_seed(D + E)
    u = lambda *s, a: ((torch.rand(*s, generator=g) * 2 - 1) * a).to(dev).requires_grad_(True)
    return [u(E, D, k, a=(1.0 / (D * k)) ** 0.5), u(E, a=0.1), u(E, E, a=(1.0 / E) ** 0.5), u(E, a=0.1), u(E, E, a=(1.0 / E) ** 0.5), u(E, a=0.1)]


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=256)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--cpu-narratives', type=int, default=4)
    ap.add_argument('--gemm-mode', type=int, default=-1, help='mt_gemm_tc_mode override (tuning experiments)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    mtb.set_compute_dtype(args.dtype)
    L = _lib.lib()
    if args.gemm_mode >= 0:
        L.mt_gemm_tc_mode(args.gemm_mode)
    out = {'workload': f'window front-end, B={args.B} T={args.T}, {args.dtype}', 'mods': {}}
    tot_f = tot_t = 0.0
    for mod, (Kv, D, E) in MODS.items():
        x = torch.randn(args.B, args.T, Kv, D, device=dev)
        ps = params(D, E, 2, dev)
        w = torch.randn(args.B, args.T, E, device=dev)

        def fwd():
            with torch.no_grad():
                return K.window_cnn(x, *ps, p_drop=0.0)

        def train():
            y = K.window_cnn(x, *ps, p_drop=0.3)
            y.backward(w)
            for p in ps:
                p.grad = None

        ms_f, ms_t = timed(fwd, args.iters), timed(train, args.iters)
        tot_f += ms_f; tot_t += ms_t
        # per-kernel breakdown of one train iteration
        train(); torch.cuda.synchronize()
        _lib.check(L.mt_spin(20.0, _lib.stream()))
        _lib.check(L.mt_prof_start(2000, _lib.stream()))
        train()
        torch.cuda.synchronize()
        n = L.mt_prof_stop()
        name = ctypes.create_string_buffer(128)
        ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
        kern = []
        for i in range(n):
            _lib.check(L.mt_prof_get(i, name, 128, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by)))
            ent = {'site': name.value.decode(), 'ms': round(ms.value, 4)}
            if fl.value > 0:
                ent['tflops'] = round(fl.value / (ms.value * 1e-3) / 1e12, 1)
            if by.value > 0:
                ent['gbs'] = round(by.value / (ms.value * 1e-3) / 1e9, 1)
            kern.append(ent)
        n_win = args.B * args.T
        flops_f = 2.0 * n_win * ((Kv - 1) * 2 * D * E + 2 * E * E)          # conv positions x (k*D) x E + the two Highway linears
        out['mods'][mod] = {'K': Kv, 'D': D, 'E': E, 'fwd_ms': round(ms_f, 4), 'train_ms': round(ms_t, 4),
                            'input_bytes': x.numel() * 4, 'fwd_read_gbs': round(x.numel() * 4 / (ms_f * 1e-3) / 1e9, 1),
                            'fwd_tflops': round(flops_f / (ms_f * 1e-3) / 1e12, 1), 'kernels': kern}
        del x, w, ps
        torch.cuda.empty_cache()
    out['fwd_ms_all_mods'] = round(tot_f, 4)
    out['train_ms_all_mods'] = round(tot_t, 4)
    out['narratives_per_s_train_front_end_only'] = round(args.B / (tot_t * 1e-3), 1)

    # reference-style CPU loop on a bounded sample: one narrative per iteration, Conv1d -> max over positions -> Highway with the ATen
    # operators the reference calls (MFT/models.py:68-79, 51-54), fp32 torch CPU
    import torch.nn.functional as F
    nb = args.cpu_narratives
    t_cpu = 0.0
    for mod, (Kv, D, E) in MODS.items():
        x = torch.randn(nb, args.T, Kv, D)
        cw, cb, pw, pb, gw, gb = [p.detach().cpu() for p in params(D, E, 2, 'cpu')]
        t0 = time.perf_counter()
        with torch.no_grad():
            for b in range(nb):
                c = F.conv1d(x[b].permute(0, 2, 1), cw, cb).max(dim=2).values
                g = torch.sigmoid(F.linear(c, gw, gb))
                _ = g * F.linear(c, pw, pb) + (1 - g) * c
        t_cpu += time.perf_counter() - t0
    out['cpu_baseline'] = {'fwd_narratives_per_s': round(nb / t_cpu, 2), 'cores': torch.get_num_threads(), 'kind': 'port',
                           'sample': f'{nb} narratives x T={args.T}, forward only, per-narrative loop'}
    out['gpu_fwd_narratives_per_s'] = round(args.B / (tot_f * 1e-3), 1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
