"""CPU, world_size 2 over gloo: the data-parallel host logic of the training path -- length-sorted
round-robin sharding, the GLOBAL sum-of-lengths loss normaliser, and the SUM all-reduce of the flat gradient buffers --
reproduces the single-process full-batch gradient.  The per-rank compute here is the CPU oracle (tests may use it);
on GPUs the same helpers feed libmt_b200.so and NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_transformer_b200.training import all_reduce_flat_, shard_batch
from oracle import fill, mt_oracle as O
from tests import util

MODS = ['acoustic', 'image', 'linguistic']
DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
N, B, T = 1, 6, 8


def _grads(sd_np, inputs, mask, target, lengths, norm):
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd_np.items()}
    pred = O.multi_transformer(sd, '', {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in inputs.items()},
                               torch.from_numpy(np.ascontiguousarray(mask)), MODS, N=N)
    loss = ((pred - torch.from_numpy(np.ascontiguousarray(target))) ** 2).sum() / norm
    loss.backward()
    live = [k for k, v in sd.items() if v.grad is not None]
    return pred.detach(), torch.cat([sd[k].grad.reshape(-1) for k in live]), live


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    sd_np = fill.fill_state(util.mods_shapes('MFT.MultiTransformer', N), 9)
    inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 9)
    x, m, tg, ls, norm = shard_batch(inputs, mask, target, lengths, rank, world)
    pred, flat, live = _grads(sd_np, x, m, tg, ls, norm)
    all_reduce_flat_([flat])
    if rank == 0:
        torch.save({'flat': flat, 'live': live, 'pred': pred, 'ls': ls, 'norm': norm}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_batch_deals_sorted_round_robin():
    inputs, mask, target, lengths = fill.make_batch(7, 8, DIMS, 3)
    lengths = [3, 8, 5, 8, 2, 6, 4]                      # unsorted on purpose
    seen = []
    for r in range(2):
        x, m, tg, ls, norm = shard_batch(inputs, mask, target, lengths, r, 2)
        assert norm == float(sum(lengths)) and ls == sorted(ls, reverse=True)
        assert x['image'].shape[1] == 8 and m.shape[0] == len(ls) == tg.shape[0]
        seen += ls
    assert sorted(seen) == sorted(lengths)
    l0 = shard_batch(inputs, mask, target, lengths, 0, 2)[3]; l1 = shard_batch(inputs, mask, target, lengths, 1, 2)[3]
    assert abs(sum(l0) - sum(l1)) <= max(lengths)


def test_shard_batch_takes_raw_window_batches():
    """The same partition applies one level up (MultiCNNTransformer inputs [B,T,K,D]): shards keep the global T and the window axes."""
    shapes = {'linguistic': (5, 30), 'image': (2, 10)}
    inputs, mask, target, lengths = fill.make_raw_batch(6, 9, shapes, 4)
    parts = [shard_batch(inputs, mask, target, lengths, r, 3) for r in range(3)]
    assert sorted(l for p in parts for l in p[3]) == sorted(lengths)
    for x, m, tg, ls, norm in parts:
        assert x['linguistic'].shape == (2, 9, 5, 30) and x['image'].shape == (2, 9, 2, 10) and m.shape == (2, 9, 1)
        assert norm == float(sum(lengths))
        for b, l in enumerate(ls):
            assert m[b, :l].all() and not m[b, l:].any() and not x['linguistic'][b, l:].any()


def test_subtract_ranges_leaves_what_the_overlapped_all_reduce_did_not_cover():
    from multimodal_transformer_b200.training import subtract_ranges
    # one arena of three stacks (stride 100 floats) whose tails [60, 100) were reduced early; a second arena untouched
    base, other = 4096, 1 << 20
    done = [(base + 4 * (100 * g + 60), 40) for g in range(3)]
    rest = subtract_ranges([(base, 300), (other, 17)], done)
    assert rest == [(base, 60), (base + 400, 60), (base + 800, 60), (other, 17)]
    assert subtract_ranges([(base, 300)], []) == [(base, 300)]
    assert subtract_ranges([(base, 40)], [(base, 40)]) == []
    assert sum(c for _, c in rest) + sum(c for _, c in done) == 317


def test_all_reduce_is_noop_without_process_group():
    g = torch.ones(5)
    assert all_reduce_flat_([g])[0] is g and torch.equal(g, torch.ones(5))


def test_two_rank_gloo_gradients_equal_full_batch(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    sd_np = fill.fill_state(util.mods_shapes('MFT.MultiTransformer', N), 9)
    inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 9)
    pred, flat, live = _grads(sd_np, inputs, mask, target, lengths, float(sum(lengths)))
    assert got['live'] == live and got['norm'] == float(sum(lengths))
    # orphan templates never get a gradient and are never packed
    assert not any(k.startswith(('attn', 'ff')) for k in live)
    scale = flat.abs().max().item()
    assert (got['flat'] - flat).abs().max().item() <= 2e-5 * scale
    order = sorted(range(B), key=lambda i: -lengths[i])[0::2]
    assert torch.allclose(got['pred'], pred[order], rtol=1e-5, atol=1e-6)
