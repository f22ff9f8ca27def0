"""Row-stream GEMM engine (csrc/mt_gemm_rs.cu) on the encoder's shapes, three modality stacks grouped (G = 3, 32768 rows each), against
the streaming tcgen05 engine on the same total rows.  CUDA events, back-to-back launches over rotating buffers larger than L2.
Usage: python tools/gemm_rs_probe.py [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_transformer_b200 import _lib

L = _lib.lib()
dev = 'cuda:0'
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
for a in sys.argv[3:]:
    if a.startswith('tune'):
        k, v = a[4:].split('=')
        L.mt_tune(int(k), int(v))
G, Mg = 3, 32768
rows = G * Mg
HBM = 6549.4
# name, N, K, b_kmajor, c_f32, bias, act, p, gate, res, colsum, ln
cases = [('qkv', 768, 256, 1, 0, 1, 0, 0.0, 0, 0, 0, 0),
         ('oproj', 256, 256, 1, 1, 1, 0, 0.1, 0, 1, 0, 0),
         ('ffn1', 128, 256, 1, 0, 1, 1, 0.1, 0, 0, 0, 0),
         ('ffn2', 256, 128, 1, 1, 1, 0, 0.1, 0, 1, 0, 0),
         ('ffn2+ln', 256, 128, 1, 1, 1, 0, 0.1, 0, 1, 0, 1),
         ('dgrad_w2', 128, 256, 0, 0, 0, 0, 0.0, 1, 0, 1, 0),
         ('dgrad_w1', 256, 128, 0, 0, 0, 0, 0.0, 0, 0, 0, 0),
         ('dgrad_o', 256, 256, 0, 0, 0, 0, 0.0, 0, 0, 0, 0)]
out = {}
nbuf = 3
for (name, N, K, bkm, cf, bias, act, p, gate, res, colsum, ln) in cases:
    if only and name not in only:
        continue
    As = [torch.randn(rows, K, device=dev).bfloat16() for _ in range(nbuf)]
    W = (torch.randn(G, N, K, device=dev) / K ** 0.5).bfloat16()
    Cs = [torch.empty(rows, N, device=dev, dtype=torch.float32 if cf else torch.bfloat16) for _ in range(nbuf)]
    b = torch.randn(G, N, device=dev) if bias else None
    gt = [torch.randn(rows, N, device=dev).bfloat16() for _ in range(nbuf)] if gate else None
    r = [torch.randn(rows, N, device=dev) for _ in range(nbuf)] if res else None
    cs = torch.zeros(G, N, device=dev) if colsum else None
    lo = [torch.empty(rows, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)] if ln else None
    la = torch.ones(G, N, device=dev) if ln else None
    lb = torch.zeros(G, N, device=dev) if ln else None
    P = _lib.ptr

    def run(i):
        j = i % nbuf
        _lib.check(L.mt_gemm_rs(G, Mg, N, K, P(As[j]), P(W), bkm, P(Cs[j]), cf, P(b), act, p, 5, 8, P(gt[j]) if gate else None, 1.1,
                                P(r[j]) if res else None, P(cs), P(lo[j]) if ln else None, P(la), P(lb), _lib.stream()))
    for i in range(4): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): run(i)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / reps
    byts = rows * K * 2 + G * N * K * 2 + rows * N * (4 if cf else 2) + (rows * N * 4 if res else 0) + (rows * N * 2 if gate else 0) + (rows * N * 2 if ln else 0)
    rec = {'us': round(t, 2), 'tflops': round(2.0 * rows * N * K / t / 1e6, 1), 'gbs': round(byts / t / 1e3, 1), 'hbm_frac': round(byts / t / 1e3 / HBM, 3),
           'hbm_floor_us': round(byts / HBM / 1e3, 2)}
    # the streaming engine on the same rows (one weight matrix: a lower bound for three separate launches), plain epilogue
    if not (gate or res or ln):
        C2 = Cs[0]
        def run_old(i):
            j = i % nbuf
            _lib.check(L.mt_gemm(1, rows, N, K, P(As[j]), K, 1, P(W), K if bkm else N, bkm, P(Cs[j]), N, cf, P(b[0]) if bias else None, act, 1, _lib.stream()))
        for i in range(4): run_old(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps): run_old(i)
        e1.record(); torch.cuda.synchronize()
        rec['streaming_engine_us'] = round(e0.elapsed_time(e1) * 1e3 / reps, 2)
    out[name] = rec
    print(name, rec, flush=True)
os.makedirs('gpurun_out', exist_ok=True)
if not only:
    json.dump(out, open('gpurun_out/gemm_rs_probe.json', 'w'), indent=1)

# per-tile timeline of CTA 0 (clock64): where a tile's time goes
if only and 'trace' in sys.argv[3:]:
    for (name, N, K, bkm, cf, bias, act, p, gate, res, colsum, ln) in cases:
        if name not in only:
            continue
        tb = torch.zeros(512, dtype=torch.int64, device=dev)
        A = torch.randn(rows, K, device=dev).bfloat16(); W = (torch.randn(G, N, K, device=dev) / K ** 0.5).bfloat16()
        C = torch.empty(rows, N, device=dev, dtype=torch.float32 if cf else torch.bfloat16)
        b = torch.randn(G, N, device=dev) if bias else None
        r = torch.randn(rows, N, device=dev) if res else None
        gt = torch.randn(rows, N, device=dev).bfloat16() if gate else None
        cs = torch.zeros(G, N, device=dev) if colsum else None
        L.mt_gemm_rs_trace(_lib.ptr(tb))
        _lib.check(L.mt_gemm_rs(G, Mg, N, K, _lib.ptr(A), _lib.ptr(W), bkm, _lib.ptr(C), cf, _lib.ptr(b), act, p, 5, 8, _lib.ptr(gt), 1.1, _lib.ptr(r),
                                _lib.ptr(cs), None, None, None, _lib.stream()))
        torch.cuda.synchronize()
        L.mt_gemm_rs_trace(None)
        t = tb.cpu().view(32, 16)
        t0 = int(t[0, 0])
        print('trace', name, '(cycles since first TMA issue): issue kb0-3 | landed kb0-3 | acc free | epi0 start end | epi1 start end')
        for i in range(20):
            if int(t[i, 0]) == 0:
                break
            print(i, [int(x) - t0 for x in t[i, :13]])
