"""Times the MFN recurrence (forward, forward+backward) alone through the module API: B=256, T=128 by default."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dtype = sys.argv[4] if len(sys.argv) > 4 else 'bf16'
mtb.set_compute_dtype(dtype)
mods = ['acoustic', 'image', 'linguistic']
torch.manual_seed(0)
mfn = mtb.MFN(mods, {m: 256 for m in mods}, 1).cuda()
adt = torch.bfloat16 if dtype == 'bf16' else torch.float32
xs = [torch.randn(B, T, 256, device='cuda').to(adt).requires_grad_(True) for _ in mods]
mask = torch.ones(B, T, 1, device='cuda')


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fwd_eval():
    mfn.eval()
    with torch.no_grad():
        return mfn._run(xs, mask, t_major=False)


def fwd_bwd():
    mfn.train()
    out = mfn._run(xs, mask, t_major=False)
    out.backward(torch.ones_like(out))


print(f'MFN B={B} T={T} {dtype}: eval fwd {timed(fwd_eval):.3f} ms   train fwd+bwd {timed(fwd_bwd):.3f} ms')
