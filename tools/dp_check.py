"""Data-parallel correctness on real GPUs (torchrun, N ranks): sharded backward + FlatAdam.all_reduce_grads (the C-ABI grouped NCCL
all-reduce) reproduces the full-batch gradient computed on every rank.  Prints max relative error per arena; exit code 1 on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import synthetic as fill
from multimodal_transformer_b200.training import FlatAdam, shard_batch, train_step_loss

rank, world, lr_ = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr_)
dev = torch.device('cuda', lr_)
dist.init_process_group('nccl', device_id=dev)
MODS = ['acoustic', 'image', 'linguistic']; DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
B, T, N = 8 * world, 32, 2
mtb.set_compute_dtype('fp32')
torch.manual_seed(1)
model = mtb.MultiTransformer(MODS, DIMS, N=N, device=dev).to(dev).eval()
inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 5)


def grads(x, m, tg, ls, norm, reduce):
    opt = FlatAdam(model)
    for p in model.parameters():
        p.grad = None
    pred = model({k: torch.from_numpy(v).to(dev) for k, v in x.items()}, torch.from_numpy(m).to(dev), ls)
    train_step_loss(pred, torch.from_numpy(tg).to(dev), norm)
    if reduce:
        opt.all_reduce_grads()
    torch.cuda.synchronize()
    out = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    used_c = opt._comm_handle is not None
    opt.close()
    return out, used_c


full, _ = grads(inputs, mask, target, lengths, float(sum(lengths)), False)
x, m, tg, ls, norm = shard_batch(inputs, mask, target, lengths, rank, world)
part, used_c = grads(x, m, tg, ls, norm, True)
worst = 0.0
gmax = max(v.abs().max().item() for v in full.values())
for k, v in full.items():
    err = (part[k] - v).abs().max().item() / max(v.abs().max().item(), 1e-3 * gmax)
    worst = max(worst, err)
print(f'rank {rank}/{world}: C-ABI all-reduce used: {used_c}; worst relative gradient error vs full batch: {worst:.2e}', flush=True)

# overlapped form: one optimizer over two steps -- the first creates the communicator, zero_grad() arms the second, whose grouped encoder
# backward all-reduces its upper layers on the communication stream (mt_comm_overlap_arm / _join); the sum must not change
opt = FlatAdam(model)
xs = {k: torch.from_numpy(v).to(dev) for k, v in x.items()}
worst2, ranges = 0.0, 0
for step in range(2):
    pred = model(xs, torch.from_numpy(m).to(dev), ls)
    train_step_loss(pred, torch.from_numpy(tg).to(dev), norm)
    opt.all_reduce_grads()
    torch.cuda.synchronize()
    if step == 1:
        ranges = opt.last_overlap_ranges
        for k, p in model.named_parameters():
            if p.grad is not None:
                worst2 = max(worst2, (p.grad - full[k]).abs().max().item() / max(full[k].abs().max().item(), 1e-3 * gmax))
    opt.zero_grad()
print(f'rank {rank}/{world}: overlapped all-reduce ranges: {ranges}; worst relative gradient error vs full batch: {worst2:.2e}', flush=True)
opt.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if worst < 1e-4 and used_c and worst2 < 1e-4 and ranges == len(MODS) else 1)
