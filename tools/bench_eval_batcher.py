"""Measurement for the SURVEY 8(f) rows 3 and 4 on one B200 (writes one JSON object to stdout):

  batcher     DeviceCorpus.generateTrainBatch (index gather on the device) over one epoch of a SEND-sized raw-window corpus, against
              the reference's way of building a batch -- torch.tensor(nested python lists), MFT/train.py:59-68 -- on a bounded sample
  evaluation  mtb.evaluate (ragged batches + on-device CCC) against the reference's procedure with the same kernels: one narrative
              per forward, CCC / Pearson on the host (MFT/train.py:203-257)

    python tools/bench_eval_batcher.py [--narratives 192] [--T 128] [--dtype bf16]
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_transformer_b200 as mtb                      # noqa: E402

SHAPES = {'acoustic': (2, 88), 'image': (2, 1000), 'linguistic': (33, 300)}
FEAT = {'acoustic': 88, 'image': 256, 'linguistic': 300}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--narratives', type=int, default=192)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--batch', type=int, default=64)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    mtb.set_compute_dtype(args.dtype)
    N, T = args.narratives, args.T
    rs = np.random.RandomState(0)
    lengths = [T] + [int(v) for v in rs.randint(T // 4, T + 1, size=N - 1)]
    out = {'narratives': N, 'T_max': T, 'dtype': args.dtype}

    # ---------------- batcher ----------------
    data = {m: torch.randn(N, T, K, D) for m, (K, D) in SHAPES.items()}
    target = torch.rand(N, T)
    corpus = mtb.DeviceCorpus(data, target, lengths)
    def epoch():
        nbytes, nb = 0, 0
        for d, tg, mask, ln in corpus.generateTrainBatch(batch_size=args.batch):
            nbytes += sum(v.numel() for v in d.values()) * 4 + tg.numel() * 4
            nb += 1
        return nbytes, nb

    for _ in range(3):
        epoch()
    torch.cuda.synchronize()
    times = []
    for _ in range(7):                                   # steady state: median of 7 epochs (the allocator's cache is warm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        nbytes, nb = epoch()
        e1.record()
        torch.cuda.synchronize()
        times.append((e0.elapsed_time(e1), time.perf_counter() - t0))
    times.sort()
    ms, wall = times[len(times) // 2]
    # the gather kernels alone (library launch profiler: algorithmic bytes / event time)
    import ctypes
    from multimodal_transformer_b200 import _lib
    L = _lib.lib()
    _lib.check(L.mt_spin(5.0, _lib.stream()))
    _lib.check(L.mt_prof_start(100, _lib.stream()))
    corpus.batch(list(range(args.batch)))
    torch.cuda.synchronize()
    nrec = L.mt_prof_stop()
    name = ctypes.create_string_buffer(128)
    pms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
    kern = []
    for i in range(nrec):
        _lib.check(L.mt_prof_get(i, name, 128, ctypes.byref(pms), ctypes.byref(fl), ctypes.byref(by)))
        kern.append({'site': name.value.decode(), 'ms': round(pms.value, 4), 'gbs': round(by.value / (pms.value * 1e-3) / 1e9, 1) if by.value else None})
    out['batcher_kernels_one_batch'] = kern
    out['batcher'] = {'batch_size': args.batch, 'batches': nb, 'epoch_ms_device': round(ms, 3), 'epoch_ms_wall': round(wall * 1e3, 3),
                      'gathered_bytes': nbytes, 'copy_gbs': round(2 * nbytes / (ms * 1e-3) / 1e9, 1),
                      'narratives_per_s': round(N / wall, 1)}
    # the reference's batch assembly on a bounded sample: nested python lists -> torch.tensor (MFT/train.py:68)
    ns = 2
    nested = {m: v[:ns].tolist() for m, v in data.items()}
    t0 = time.perf_counter()
    for m in nested:
        torch.tensor(nested[m], dtype=torch.float)
    t_ref = time.perf_counter() - t0
    out['batcher']['cpu_baseline'] = {'narratives_per_s': round(ns / t_ref, 2), 'kind': 'port', 'cores': 1,
                                      'sample': f'torch.tensor(nested lists) of {ns} narratives x T={T}, three modalities'}
    del corpus, data
    torch.cuda.empty_cache()

    # ---------------- evaluation (hot-path model on window features) ----------------
    mods = list(FEAT)
    model = mtb.MultiTransformer(mods, FEAT).eval()
    feats = {m: torch.randn(N, T, d, device=dev) for m, d in FEAT.items()}
    mask = torch.zeros(N, T, 1, device=dev)
    for b, l in enumerate(lengths):
        mask[b, :l] = 1
    tgt = torch.rand(N, T, 1, device=dev) * mask

    def batched():
        return mtb.evaluate(model, feats, tgt, mask, lengths, batch_size=args.batch)

    def one_at_a_time():
        def eval_ccc(y_true, y_pred):                # the reference's host-side statistic (MFT/train.py:42-50), numpy as there
            tm, pm = y_true.mean(), y_pred.mean()
            cov = ((y_true - tm) * (y_pred - pm)).mean()
            return 2 * cov / (y_true.var() + y_pred.var() + (pm - tm) ** 2)
        cc = []
        with torch.no_grad():
            for b, l in enumerate(lengths):
                o = model({m: v[b:b + 1, :l].contiguous() for m, v in feats.items()}, torch.ones(1, l, 1, device=dev), [l])
                cc.append(eval_ccc(tgt[b, :l, 0].double().cpu().numpy(), o.reshape(-1).double().cpu().numpy()))
        return cc

    batched(); torch.cuda.synchronize()
    t0 = time.perf_counter(); _, loss, stats, _ = batched(); torch.cuda.synchronize(); t_b = time.perf_counter() - t0
    one_at_a_time(); torch.cuda.synchronize()
    t0 = time.perf_counter(); cc = one_at_a_time(); torch.cuda.synchronize(); t_1 = time.perf_counter() - t0
    out['evaluation'] = {'batch_size': args.batch, 'batched_s': round(t_b, 4), 'one_at_a_time_s': round(t_1, 4),
                         'narratives_per_s_batched': round(N / t_b, 1), 'narratives_per_s_one_at_a_time': round(N / t_1, 1),
                         'mean_ccc_batched': stats['ccc'], 'mean_ccc_one_at_a_time': float(np.mean(cc)),
                         'abs_diff_mean_ccc': abs(stats['ccc'] - float(np.mean(cc)))}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
