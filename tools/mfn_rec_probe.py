"""Per-kernel times of the four MFN recurrence kernels (bf16 mode) through the library's per-launch profiler, optionally under the
timing-experiment switches of mt_tune key 8 (bit 0: no gate stash, 1: no per-step state stores, 2: no input feed, 3: no mma).
    python tools/mfn_rec_probe.py [B] [T] [dbg,dbg,...]"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dbgs = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]
mtb.set_compute_dtype('bf16')
L = _lib.lib()
mods = ['acoustic', 'image', 'linguistic']
torch.manual_seed(0)
mfn = mtb.MFN(mods, {m: 256 for m in mods}, 1).cuda().train()
xs = [torch.randn(B, T, 256, device='cuda').bfloat16().requires_grad_(True) for _ in mods]
mask = torch.ones(B, T, 1, device='cuda')


def step():
    out = mfn._run(xs, mask, t_major=False)
    out.backward(torch.ones_like(out))


for dbg in dbgs:
    L.mt_tune(8, dbg)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.check(L.mt_spin(30.0, _lib.stream()))
    _lib.check(L.mt_prof_start(4000, _lib.stream()))
    reps = 3
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    n = L.mt_prof_stop()
    name = ctypes.create_string_buffer(128)
    ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
    agg = {}
    for i in range(n):
        _lib.check(L.mt_prof_get(i, name, 128, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by)))
        k = name.value.decode()
        agg[k] = agg.get(k, 0.0) + ms.value / reps
    rec = {k.split(':')[0].replace('mt_mfn_mma_', ''): v for k, v in agg.items() if 'mt_mfn_mma' in k}
    tot = sum(agg.values())
    print(f'dbg {dbg:2d}  B={B} T={T}: ' + '  '.join(f'{k} {v * 1e3:7.1f} us ({v * 1e3 / T:5.2f}/step)' for k, v in sorted(rec.items()))
          + f'   all MFN launches {tot:.3f} ms', flush=True)
L.mt_tune(8, 0)
if os.environ.get('MT_REC_TRACE'):
    import numpy as np
    L.mt_tune(8, 32 | int(os.environ['MT_REC_TRACE']))
    step(); torch.cuda.synchronize()
    L.mt_tune(8, 0)
    buf = np.zeros((4, 128, 8), dtype=np.uint64)
    fn = ctypes.CDLL(_lib.LIB_PATH).mt_mfn_rec_trace
    fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
    assert fn(buf.ctypes.data) == 0
    for k, nm, ns in ((0, 'lstm_fwd', 8), (1, 'lstm_bwd', 6)):
        tr = buf[k].astype(np.int64)
        print(nm, 'stamps relative to the loop top, clocks (steps 40..47):')
        for s in range(40, 48):
            print('  step', s, ' '.join(f'{int(tr[s, j] - tr[s, 0]):6d}' for j in range(ns)), '  next top', int(tr[s + 1, 0] - tr[s, 0]))
        d = tr[41:120, 0] - tr[40:119, 0]
        print('  mean step', d.mean(), 'clocks; mean per-slot offsets', [float((tr[40:120, j] - tr[40:120, 0]).mean()) for j in range(ns)])
    buf2 = np.zeros((8, 128, 8), dtype=np.uint64)
    fn2 = ctypes.CDLL(_lib.LIB_PATH).mt_mfn_rec_trace2
    fn2.argtypes = [ctypes.c_void_p]; fn2.restype = ctypes.c_int
    assert fn2(buf2.ctypes.data) == 0
    tr = buf2.astype(np.int64)
    base = tr[0, 40:120, 0]
    print('lstm_fwd v2, all compute warps: mean stamp offsets relative to warp 0 loop top (steps 40..119)')
    for w in range(8):
        print('  warp', w, ' '.join(f'{float((tr[w, 40:120, j] - base).mean()):7.0f}' for j in range(8)))
    print('  step period', float((tr[0, 41:120, 0] - tr[0, 40:119, 0]).mean()))
