#!/usr/bin/env python
"""Benchmark of the MFT hot path (BASELINE.json configs[1]: MFT default hyper-parameters, bf16 train step, batch 256
narratives per GPU, T = 128 windows, synthetic SEND-shaped inputs, deterministic random-init weights).

    python bench.py --gpus N --steps K --warmup W              # our sm_100a path (torchrun for N > 1)
    python bench.py --impl reference ...                        # the CPU restatement of the reference, same metric

One JSON line on stdout (rank 0).  `value` = narratives/s of the TRAIN step (forward + fused loss + backward +
gradient all-reduce when N > 1 + fused Adam) with inputs resident in HBM; `e2e` = the same through the module API
with pinned-host inputs copied in and the loss read back every step; `inference` = eval() forward only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', ''):
    os.environ['NCCL_DEBUG'] = 'WARN'
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODS = ['acoustic', 'image', 'linguistic']
DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}
METRIC = 'MFT narratives/sec (train step + inference)'
FLOP_PER_TOKEN_FWD = 15836288          # SURVEY 8(d), T = 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='narratives per GPU')
    ap.add_argument('--seq', type=int, default=128)
    ap.add_argument('--layers', type=int, default=6)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--cpu-sample', type=int, default=32, help='narratives per CPU-baseline step (the reference trains with batch 25, MFT/train.py:74)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--gemm-mode', type=int, default=0, help='tcgen05 GEMM CTAs per SM (tuning; 0 = library default)')
    ap.add_argument('--tune', default='', help='library tuning knobs, e.g. 0=2,1=2 (mt_tune key=value)')
    ap.add_argument('--serial-stacks', action='store_true', help='run the modality stacks on one stream')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p['bf16_tflops_sustained'], src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
def cpu_train_step_factory(n_layers, B, T, seed=1):
    """The oracle (CPU restatement of the reference path) doing one MFT train step: fwd + loss + backward + Adam."""
    import torch
    from oracle import fill, mt_oracle as O
    from tests import util
    torch.set_num_threads(os.cpu_count() or 1)
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', n_layers), seed, requires_grad=True)
    live = [v for k, v in sd.items() if not k.startswith(('attn', 'ff'))]
    opt = torch.optim.Adam(live, lr=1e-4, weight_decay=1e-4)
    inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, seed)
    xin = {k: torch.from_numpy(v) for k, v in inputs.items()}
    mk, tg = torch.from_numpy(mask), torch.from_numpy(target)

    def step():
        pred = O.multi_transformer(sd, '', xin, mk, MODS, N=n_layers)
        loss = O.train_loss(pred, tg, lengths)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss)

    return step


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    step = cpu_train_step_factory(args.layers, args.cpu_sample, args.seq)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    k = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    val = args.cpu_sample / dt
    cores = torch.get_num_threads()
    sample = f'{args.cpu_sample} narratives x T={args.seq} per step, {k} timed steps (requested {args.steps}), fp32, torch CPU'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'narratives/s', 'n_gpus': args.gpus, 'steps': k, 'warmup': min(args.warmup, 2),
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'MFT-VAL N={args.layers} train step (fwd+loss+bwd+Adam), T={args.seq}; CPU arm runs a bounded sample',
                   'batch_per_step': args.cpu_sample, 'seq_len': args.seq},
        'cpu_baseline': {'value': val, 'unit': 'narratives/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'narratives/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    import multimodal_transformer_b200 as mtb
    from multimodal_transformer_b200 import _lib
    from multimodal_transformer_b200.training import FlatAdam, GraphedForward, GraphedTrainStep, train_step_loss
    from multimodal_transformer_b200 import synthetic as fill      # deterministic synthetic inputs (torch-free)

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L = _lib.lib()
    mtb.set_compute_dtype(args.dtype)
    if args.gemm_mode:
        L.mt_gemm_tc_mode(args.gemm_mode)
    for kv in filter(None, args.tune.split(',')):
        k_, v_ = kv.split('=')
        L.mt_tune(int(k_), int(v_))
    if args.serial_stacks:
        mtb.set_parallel_stacks(False)
    B, T, N = args.batch, args.seq, args.layers

    torch.manual_seed(1)                                            # MFT/train.py:524; default (random) init
    model = mtb.MultiTransformer(MODS, DIMS, N=N, device=dev)
    model.to(dev)
    opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)

    inputs, mask, target, lengths = fill.make_batch(B, T, DIMS, 1 + rank)
    norm = float(sum(sum(fill.make_lengths(B, T, 1 + r)) for r in range(world)))      # GLOBAL sum of lengths
    host = {k: torch.from_numpy(v).pin_memory() for k, v in inputs.items()}
    host_mask, host_target = torch.from_numpy(mask).pin_memory(), torch.from_numpy(target).pin_memory()
    res = {k: v.to(dev) for k, v in host.items()}
    res_mask, res_target = host_mask.to(dev), host_target.to(dev)
    h2d = sum(v.numel() * 4 for v in host.values()) + host_mask.numel() * 4 + host_target.numel() * 4
    loss_host = torch.zeros(1).pin_memory()
    pred_host = torch.zeros(B, T, 1).pin_memory()

    # eager path (one launch at a time; used for the per-kernel breakdown and reported as `eager`)
    def train_step(e2e):
        x, m, tg = res, res_mask, res_target
        model.train()
        pred = model(x, m, lengths)
        loss = train_step_loss(pred, tg, norm)
        opt.step()
        opt.zero_grad()
        return loss

    # the product path: the whole step captured once into a CUDA graph (multimodal_transformer_b200.training)
    gstep = GraphedTrainStep(model, opt, B, T, DIMS, dev, norm_fn=lambda _l: norm)
    for _ in range(2):
        train_step(False)
    gstep.load(res, res_mask, res_target, lengths)
    l0 = L.mt_launch_count()
    gstep.capture()
    launches_per_step = (L.mt_launch_count() - l0) // (gstep.warmup + 1)
    gfwd = GraphedForward(model, B, T, DIMS, dev)
    gfwd.load(res, res_mask)
    gfwd.capture()

    # e2e: every step copies its inputs from pinned host memory and reads its result back.  The public API pipelines the
    # input copy: batch k + 1 crosses PCIe on a copy stream while batch k computes (GraphedTrainStep.prefetch).
    pending = {'train': 0, 'infer': 0}

    def g_train(e2e):
        if not e2e:
            return gstep.replay()
        if pending['train'] == 0:
            gstep.prefetch(host, host_mask, host_target, lengths); pending['train'] += 1
        gstep.prefetch(host, host_mask, host_target, lengths)          # next step's inputs: pinned host -> staging, async
        loss = gstep.step_prefetched()
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return loss

    def g_infer(e2e):
        if not e2e:
            gfwd.graph.replay()
            return gfwd.pred
        if pending['infer'] == 0:
            gfwd.prefetch(host, host_mask); pending['infer'] += 1
        gfwd.prefetch(host, host_mask)
        pred = gfwd.forward_prefetched()
        pred_host.copy_(pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pred

    def timed(fn, e2e, warm, steps):
        for _ in range(warm):
            fn(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def note(msg):
        if os.environ.get('MT_BENCH_VERBOSE'):
            print(f'[bench rank {rank}] {msg}', file=sys.stderr, flush=True)

    W, K = max(3, args.warmup), args.steps
    note('captured')
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_train = timed(g_train, False, W, K)
    clocks = sampler.stop() if sampler else None
    launches = launches_per_step * K
    note('train timed')
    ms_inf = timed(g_infer, False, W, K)
    note('inference timed')
    ms_train_e2e = timed(g_train, True, W, K)
    note('train e2e timed')
    ms_inf_e2e = timed(g_infer, True, W, K)
    note('inference e2e timed')
    ms_eager = timed(train_step, False, 2, max(2, min(K, 5)))
    note('eager timed')

    # ---- per-kernel breakdown of the train step (after the timed regions; events after every launch) ---------
    roofline, kernels = None, None
    nprof = 1
    if not args.no_profile:
        # EVERY rank runs these eager steps (they all-reduce when N > 1); only rank 0 records.
        # The host needs ~30 us per launch; a busy-wait kernel in front lets it run ahead so that consecutive event
        # records bracket a kernel's true duration instead of the host's enqueue gap
        torch.cuda.synchronize()
        mtb.set_parallel_stacks(False)              # the per-launch profiler times consecutive launches of ONE stream
        train_step(False)                           # un-profiled: lets the caching allocator settle for the one-stream schedule
        torch.cuda.synchronize()
        if rank == 0:
            _lib.check(L.mt_spin(60.0, _lib.stream()))
            _lib.check(L.mt_prof_start(20000, _lib.stream()))
        for _ in range(nprof):
            train_step(False)
        torch.cuda.synchronize()
        n = L.mt_prof_stop() if rank == 0 else 0
    if rank == 0 and not args.no_profile:
        P = peaks()
        import ctypes
        agg = {}
        name = ctypes.create_string_buffer(128)
        ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
        for i in range(n):
            _lib.check(L.mt_prof_get(i, name, 128, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by)))
            a = agg.setdefault(name.value.decode(), [0.0, 0.0, 0.0, 0])
            a[0] += ms.value; a[1] += fl.value; a[2] += by.value; a[3] += 1
        tot = sum(a[0] for a in agg.values())
        kernels = []
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            ent = dict(site=k, launches_per_step=a[3] / nprof, ms_per_step=a[0] / nprof, share=a[0] / nprof / ms_train)
            if a[1] > 0:
                ent['tflops'] = a[1] / (a[0] * 1e-3) / 1e12
            if a[2] > 0:
                ent['gbs'] = a[2] / (a[0] * 1e-3) / 1e9
            kernels.append(ent)
        top = kernels[0]
        a = agg[top['site']]
        # a kernel is judged on the roof that binds it: compare its two fractions and report the larger
        f_t = (top.get('tflops', 0.0)) / P['tf_sustained']
        f_h = (top.get('gbs', 0.0)) / P['hbm']
        if f_t >= f_h:
            roofline = dict(bound='tensor', achieved=top.get('tflops', 0.0), peak=P['tf_sustained'], unit='TFLOP/s', frac=f_t)
        else:
            roofline = dict(bound='hbm', achieved=top.get('gbs', 0.0), peak=P['hbm'], unit='GB/s', frac=f_h)
        traffic = None
        try:                                        # DRAM bytes per launch of that kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                traffic = json.load(f).get(top['site'].split(':')[0])
        except (OSError, ValueError):
            pass
        roofline.update(kernel=top['site'], avg_launch_ms=a[0] / a[3], share_of_step=top['share'], peak_source=P['src'], traffic=traffic,
                        algorithmic_bytes_per_launch=a[2] / a[3], algorithmic_flops_per_launch=a[1] / a[3],
                        how='algorithmic work annotated at the launch site / CUDA-event duration between consecutive launches on the launching stream')
        kernels = kernels[:40]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step = cpu_train_step_factory(N, args.cpu_sample, T)
        step()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            step()
        dt = (time.perf_counter() - t0) / reps
        cpu = dict(value=args.cpu_sample / dt, unit='narratives/s', cores=torch.get_num_threads(), kind='port',
                   sample=f'oracle MFT train step on {args.cpu_sample} narratives x T={T}, 1 warm-up + {reps} timed, fp32 torch CPU, '
                          f'os.cpu_count()={os.cpu_count()}')

    if rank == 0:
        gb = world * B
        stash_gb = 3 * N * B * T * (256 * (4 + 2 + 6 + 2 + 4 + 2) + 128 * 2) / 1e9
        out = {
            'metric': METRIC, 'value': gb / (ms_train * 1e-3), 'unit': 'narratives/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_train, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
            'data': 'synthetic',
            'config': {'workload': f'MFT-VAL (acoustic 88 / image 256 / linguistic 300), N={N} d=256 h=8 dff=128, train step = fwd + '
                                   f'MSE/sum(lengths) + bwd + {"NCCL all-reduce + " if world > 1 else ""}Adam(lr 1e-4, wd 1e-4), dropout on',
                       'batch_per_gpu': B, 'global_batch': gb, 'seq_len': T, 'parallelism': f'dp{world}',
                       'l2': f'no explicit flush: each step streams ~{stash_gb:.1f} GB of activations per GPU, far above the 126 MB L2',
                       'value_is': 'train step, inputs resident in HBM'},
            'inference': {'value': gb / (ms_inf * 1e-3), 'unit': 'narratives/s', 'ms_per_step': ms_inf,
                          'e2e_value': gb / (ms_inf_e2e * 1e-3), 'd2h_bytes_per_step': B * T * 4},
            'e2e': {'value': gb / (ms_train_e2e * 1e-3), 'unit': 'narratives/s', 'ms_per_step': ms_train_e2e, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches),
            'launch_mode': f'one CUDA graph per step ({int(launches_per_step)} kernel nodes from libmt_b200.so, captured once); '
                           f'eager launch of the same step: {ms_eager:.2f} ms',
            'eager': {'ms_per_step': ms_eager, 'value': gb / (ms_eager * 1e-3)},
            'clocks': clocks,
            'model_flops_utilisation': {'train_tflops_per_gpu': 3 * FLOP_PER_TOKEN_FWD * B * T / (ms_train * 1e-3) / 1e12 if T == 128 and N == 6 else None},
            'roofline': roofline, 'kernels': kernels, 'cpu_baseline': cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # tear down in order: captured graphs hold NCCL kernels, so drop them before the communicator; and never let a stuck
        # communicator teardown hang the job after the result line is out
        torch.cuda.synchronize()
        dist.barrier()
        gstep.graph = None
        gfwd.graph = None
        torch.cuda.synchronize()
        opt.close()                                  # the library's own NCCL communicator (mt_comm_*)
        watchdog = threading.Timer(20.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


def main():
    args = parse()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.gpus > 1 and world == 1 and 'RANK' not in os.environ:
        # convenience: relaunch under torchrun, one process per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', '29533', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == 'reference':
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
