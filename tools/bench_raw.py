"""Whole-model train step from RAW windows (SURVEY 8(d) 'raw level', secondary): MultiCNNTransformer = window front-end (f-1) + MFT hot
path, batches drawn from a device-resident corpus by the GPU batcher (f-3), the step captured in one CUDA graph.  One JSON object.

    python tools/bench_raw.py [--B 256] [--T 128] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodal_transformer_b200 as mtb                      # noqa: E402
from multimodal_transformer_b200.training import FlatAdam, GraphedTrainStep   # noqa: E402

SHAPES = {'acoustic': (2, 88), 'image': (2, 1000), 'linguistic': (33, 300)}
EMBED = {'acoustic': 88, 'image': 256, 'linguistic': 300}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=256)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--dtype', default='bf16')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    mtb.set_compute_dtype(args.dtype)
    B, T = args.B, args.T
    mods = list(SHAPES)
    model = mtb.MultiCNNTransformer(mods, {m: s[1] for m, s in SHAPES.items()}, EMBED)
    opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
    g = torch.Generator(device=dev).manual_seed(1)
    lengths = sorted([T] + torch.randint(T // 4, T + 1, (B - 1,)).tolist(), reverse=True)
    # device-resident corpus of exactly one batch (every narrative padded to T); the batcher gathers a fresh permutation per step
    corpus = mtb.DeviceCorpus({m: torch.randn(B, T, K, D, generator=g, device=dev) for m, (K, D) in SHAPES.items()},
                              torch.rand(B, T, generator=g, device=dev), lengths)
    step = GraphedTrainStep(model, opt, B, T, SHAPES, dev)
    batches = lambda: corpus.generateTrainBatch(batch_size=B)
    for data, target, mask, ln in batches():
        step(data, mask, target, ln)                  # capture + first replay
    for _ in range(3):
        for data, target, mask, ln in batches():
            step(data, mask, target, ln)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        for data, target, mask, ln in batches():
            loss = step(data, mask, target, ln)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # the batcher gathering straight into the graph's static inputs (no intermediate batch)
    import random
    index = list(range(B))
    for _ in range(2):
        random.shuffle(index); step.step_from_corpus(corpus, index)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        random.shuffle(index)
        loss = step.step_from_corpus(corpus, index)
    e1.record()
    torch.cuda.synchronize()
    ms_d = e0.elapsed_time(e1) / args.steps
    # graph replay alone (inputs already in the static buffers)
    e0.record()
    for _ in range(args.steps):
        step.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_r = e0.elapsed_time(e1) / args.steps
    print(json.dumps({'workload': f'MFT MultiCNNTransformer raw-level train step, B={B} T={T} {args.dtype}, windows (K x D): {SHAPES}',
                      'ms_per_step_batcher_plus_step': round(ms, 3), 'ms_per_step_gather_into_static_inputs': round(ms_d, 3),
                      'ms_per_step_replay_only': round(ms_r, 3), 'narratives_per_s': round(B / (ms_d * 1e-3), 1), 'loss': float(loss.item()),
                      'raw_input_bytes_per_step': sum(B * T * K * D * 4 for K, D in SHAPES.values())}))


if __name__ == '__main__':
    main()
