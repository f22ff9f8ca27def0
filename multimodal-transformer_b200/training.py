"""Training-step plumbing around the drop-in modules: fused MSE loss (+ its gradient), one-launch-per-arena Adam with
L2 weight decay (torch.optim.Adam semantics, MFT/train.py:557), and the data-parallel gradient all-reduce over the
flat gradient buffers (one process per GPU, torch.distributed; the reference has no distributed code at all).

    opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
    pred = model(inputs, mask, lengths)
    loss = train_step_loss(pred, target, global_sum_lengths)      # backward() included
    opt.step()                                                    # all-reduces first when torch.distributed is up
"""
import os

import torch
import torch.distributed as dist

from . import functional as K


def _arenas_of(model):
    seen, out = set(), []
    for m in model.modules():
        for getter in ('arena', 'group_arena'):
            if hasattr(m, getter) and callable(getattr(m, getter)):
                if getter == 'group_arena' and not getattr(m, 'use_encoder', True):
                    continue
                a = getattr(m, getter)()
                if a is not None and id(a) not in seen:
                    seen.add(id(a)); out.append(a)
        a = getattr(m, '_dec_arena', None)
        if a is not None and id(a) not in seen:
            seen.add(id(a)); out.append(a)
    # the same parameters can sit in a per-stack arena AND in the group arena of MultiTransformer; only the arena whose flat buffer the
    # parameters currently live in (the one the last forward bound) owns them
    bound = [a for a in out if a.bound()]
    owned = {id(p) for a in bound for p in a.params}
    return [a for a in out if a.bound() or not all(id(p) in owned for p in a.params)]


def shard_batch(inputs, mask, target, lengths, rank, world):
    """Data-parallel partition of one global batch: narratives sorted by length (descending, stable --
    what generateTrainBatch does, MFT/train.py:62-63) are dealt round-robin to the ranks, so every rank sees a balanced
    sum of lengths; each shard keeps the GLOBAL padded length T, which keeps kernels shape-identical across ranks.
    Works on numpy arrays or torch tensors.  Returns (inputs, mask, target, lengths, global_norm) for `rank`;
    global_norm = sum of ALL lengths is the loss normaliser every rank must use (gradients are then SUMMED)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    mine = order[rank::world]
    take = (lambda a: a[mine])
    return ({m: take(v) for m, v in inputs.items()}, take(mask), take(target), [lengths[i] for i in mine], float(sum(lengths)))


def subtract_ranges(bufs, done):
    """bufs, done: lists of (address, n_floats).  Returns the parts of `bufs` not covered by `done` (fp32 elements)."""
    out = []
    for p, n in bufs:
        lo, hi = p, p + 4 * n
        cuts = sorted((max(lo, q), min(hi, q + 4 * c)) for q, c in done if q < hi and q + 4 * c > lo)
        cur = lo
        for a, b in cuts:
            if a > cur:
                out.append((cur, (a - cur) // 4))
            cur = max(cur, b)
        if hi > cur:
            out.append((cur, (hi - cur) // 4))
    return out


def all_reduce_flat_(bufs, group=None):
    """SUM-all-reduce every flat gradient buffer in place (one collective per arena; NCCL on GPUs, gloo in the CPU tests).
    No-op without an initialised process group or with a single rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bufs
    for g in bufs:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return bufs


def train_step_loss(pred, target, norm):
    """loss = sum((pred - target)^2) / norm (MFT/train.py:135-139) computed and differentiated by one kernel, then
    pred.backward(dloss/dpred).  `norm` is the GLOBAL sum of lengths under data parallelism.  Returns loss [1]."""
    loss, dpred = K.mse_loss_sum_normalised(pred.detach(), target, norm)
    pred.backward(dpred.view_as(pred))
    return loss


class FlatAdam:
    """Adam over the models' flat parameter arenas: one mt_adam_step launch per arena, gradients read in place from
    the flat buffers backward produced.  Parameters outside any arena (embed / fusion Linears) are gathered into one
    more arena the first time they are seen with a gradient.  Parameters that never receive a gradient (the orphan
    attn{mod}/ff{mod} templates) are left untouched, like torch.optim.Adam does."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        self.step_count = 0
        self.state = {}          # id(arena) -> (m, v)
        self.misc = None
        self._comm_tried, self._comm_handle = False, None
        self.last_overlap_ranges = 0
        # opt-in (MT_AR_OVERLAP=1): all-reduce the upper encoder layers under the rest of the backward (prepare_backward).  Measured on 2 and 4
        # B200s it does not pay: 6.70 vs 6.66 ms and 6.57 vs 6.53 ms per step -- the step's kernels are persistent and sized to every SM, so
        # the NCCL kernel that runs under them takes as much from the backward as it saves after it (31 MB cost 0.17 ms un-overlapped)
        self.overlap = os.environ.get('MT_AR_OVERLAP', '0') == '1'
        self.param_groups = [dict(lr=lr)]      # ReduceLROnPlateau-style schedulers poke this

    def _all_arenas(self):
        arenas = _arenas_of(self.model)
        in_arena = {id(p) for a in arenas for p in a.params}
        if self.misc is None:
            rest = [p for p in self.model.parameters() if id(p) not in in_arena and p.grad is not None]
            if rest:
                self.misc = K.Arena(rest)
                self.misc.bind()
        return arenas + ([self.misc] if self.misc is not None else [])

    def _flat_grad(self, arena):
        g = arena.flat_grad()
        if g is None:
            if any(p.grad is None for p in arena.params):
                return None
            g = torch.cat([p.grad.reshape(-1) for p in arena.params])
        return g

    def _comm(self, device):
        """The library's own NCCL communicator (C ABI: mt_comm_init / mt_allreduce_grads), created on first use from a unique id
        that rank 0 broadcasts through torch.distributed.  None when unavailable (no NCCL in the process, sub-groups, CPU)."""
        if self._comm_tried:
            return self._comm_handle
        self._comm_tried = True
        import ctypes
        from . import _lib
        L = _lib.lib()
        if self.group is not None or device.type != 'cuda' or dist.get_backend() != 'nccl' or not L.mt_comm_available():
            return None
        idbuf = ctypes.create_string_buffer(128)
        if dist.get_rank() == 0:
            _lib.check(L.mt_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0)
        idbuf = ctypes.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        handle = ctypes.c_void_p()
        torch.cuda.synchronize(device)
        _lib.check(L.mt_comm_init(idbuf, dist.get_rank(), dist.get_world_size(), ctypes.byref(handle)))
        self._comm_handle = handle
        return handle

    def close(self):
        """Destroy the library communicator (call before torch.distributed.destroy_process_group)."""
        if getattr(self, '_comm_handle', None) is not None:
            from . import _lib
            _lib.lib().mt_comm_destroy(self._comm_handle)
            self._comm_handle = None

    def all_reduce_grads(self):
        """SUM the gradients over ranks (the loss is already normalised by the global sum of lengths): ONE grouped NCCL all-reduce
        over every flat gradient arena through the C ABI (mt_allreduce_grads), torch.distributed otherwise."""
        K.join_deferred()                             # weight gradients still in flight on a side stream (set_deferred_weight_grads)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        flats, gathered = [], []
        for a in self._all_arenas():
            g = self._flat_grad(a)
            if g is None:
                continue
            (flats if a.flat_grad() is not None else gathered).append((a, g))
        comm = self._comm(flats[0][1].device) if flats else None
        if comm is not None:
            import ctypes
            from . import _lib
            # ranges the armed backward already reduced on the communication stream (prepare_backward): join it, reduce only the rest
            done_p, done_n, nd = (ctypes.c_void_p * 4)(), (ctypes.c_size_t * 4)(), ctypes.c_int(0)
            _lib.check(_lib.lib().mt_comm_overlap_join(_lib.stream(), done_p, done_n, ctypes.byref(nd)))
            self.last_overlap_ranges = nd.value        # how many ranges the backward had already reduced (0: not armed / not grouped)
            rest = subtract_ranges([(g.data_ptr(), g.numel()) for _, g in flats], [(done_p[i], done_n[i]) for i in range(nd.value)])
            n = len(rest)
            bufs = (ctypes.c_void_p * n)(*[p for p, _ in rest])
            counts = (ctypes.c_size_t * n)(*[c for _, c in rest])
            _lib.check(_lib.lib().mt_allreduce_grads(comm, bufs, counts, n, _lib.stream()))
        else:
            all_reduce_flat_([g for _, g in flats], self.group)
        for a, g in gathered:                         # gradients that do not live in one flat buffer: reduce a packed copy, scatter back
            all_reduce_flat_([g], self.group)
            for p, v in zip(a.params, a.grad_views(g)):
                p.grad.copy_(v)

    @torch.no_grad()
    def step(self):
        self.all_reduce_grads()
        self.step_count += 1
        lr = self.param_groups[0]['lr']
        for a in self._all_arenas():
            g = self._flat_grad(a)
            if g is None:
                continue
            flat = a.bind()
            st = self.state.get(id(a))
            if st is None:
                st = (torch.zeros_like(flat), torch.zeros_like(flat))
                self.state[id(a)] = st
            K.adam_step_flat(flat, g, st[0], st[1], self.step_count, lr, self.betas, self.eps, self.weight_decay)
            a._lp_key = None             # the kernel wrote the arena behind autograd's back: refresh the bf16 shadow

    @torch.no_grad()
    def step_dev(self, step_t, lr_t=None):
        """step() with the step count (and optionally lr) in device memory -- the form a captured CUDA graph replays.
        The bf16 shadow of each arena is refreshed by the same launch."""
        self.all_reduce_grads()
        lr = self.param_groups[0]['lr']
        for a in self._all_arenas():
            g = self._flat_grad(a)
            if g is None:
                continue
            flat = a.bind()
            st = self.state.get(id(a))
            if st is None:
                st = (torch.zeros_like(flat), torch.zeros_like(flat))
                self.state[id(a)] = st
            lp = a.lp if (a.lp is not None and a._lp_key is not None) else None
            K.adam_step_flat_dev(flat, g, st[0], st[1], step_t, lr_t, lr, self.betas, self.eps, self.weight_decay, lp)
            if lp is None:
                a._lp_key = None

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None
        self.prepare_backward()

    def prepare_backward(self):
        """Data-parallel runs: arm the overlapped all-reduce for the next backward (C ABI mt_comm_overlap_arm): the grouped encoder
        backward hands the gradients of its upper layers to the communication stream as soon as they are enqueued, all_reduce_grads
        joins and reduces the rest.  Called by zero_grad(), i.e. once per step of the usual loop; a step that was not armed simply
        reduces everything at the end.  No-op on one rank or without the library communicator."""
        if self.overlap and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1 and self._comm_handle is not None:
            from . import _lib
            _lib.check(_lib.lib().mt_comm_overlap_arm(self._comm_handle, -1))


class GraphedTrainStep:
    """The whole train step -- forward, fused loss, backward, gradient all-reduce (when torch.distributed is up) and
    Adam -- captured ONCE into a CUDA graph for a fixed (B, T) and replayed per batch: the ~530 kernel launches of a
    step cost ~16 ms of host time when issued one by one, more than the kernels themselves take.  Per-step scalars
    (1/sum(lengths), Adam step count, lr, dropout seed offset) live in device memory so a replay sees fresh values.

        step = GraphedTrainStep(model, opt, B, T, {mod: in_dim}, device)
        loss = step(inputs, mask, target, lengths)          # tensors on host (pinned) or device; returns loss [1] on device

    Re-create it after load_state_dict() or a change of model.train()/dtype (the graph bakes addresses and modes).
    If the model was trained eagerly on the default stream before, drop every reference to those iterations' outputs / losses first
    (they keep autograd AccumulateGrad nodes alive that are bound to the legacy stream, which a capturing stream may not touch)."""

    def __init__(self, model, opt, B, T, in_dims, device, norm_fn=None, warmup=3, input_dtype=torch.float32):
        from . import _lib
        self.model, self.opt, self.B, self.T, self.device = model, opt, B, T, device
        self.mods = list(in_dims)
        # in_dims: mod -> feature width (hot-path models, inputs [B,T,d]) or mod -> (K, D) (MultiCNNTransformer variants, raw windows
        # [B,T,K,D]: the window front-end is then part of the captured step).  input_dtype = torch.bfloat16 (bf16 compute mode, hot-path
        # models): the window features are held, copied from the host and fed to the embed GEMMs as bf16 -- the same values the embed
        # would round them to anyway, half the PCIe bytes per step and no cast pass
        self.x = {m: torch.zeros((B, T) + (tuple(d) if isinstance(d, (tuple, list)) else (d,)), device=device, dtype=input_dtype)
                  for m, d in in_dims.items()}
        self.mask = torch.zeros(B, T, 1, device=device)
        self.target = torch.zeros(B, T, 1, device=device)
        self.inv_norm = torch.ones(1, device=device)
        self.lr = torch.full((1,), float(opt.param_groups[0]['lr']), device=device)
        self.step_t = torch.full((1,), int(opt.step_count), dtype=torch.int64, device=device)
        self.seed_off = torch.zeros(1, dtype=torch.int64, device=device)
        self.lengths = [T] * B
        self.norm_fn = norm_fn or (lambda lengths: float(sum(lengths)))
        self.warmup = warmup
        self.graph = None
        self.loss = None
        self._lib = _lib

    def load(self, inputs, mask, target, lengths):
        for m in self.mods:
            self.x[m].copy_(inputs[m], non_blocking=True)
        self.mask.copy_(mask.reshape(self.B, self.T, 1), non_blocking=True)
        self.target.copy_(target.reshape(self.B, self.T, 1), non_blocking=True)
        self.inv_norm.fill_(1.0 / self.norm_fn(lengths))
        self._refresh_lr()

    def _refresh_lr(self):
        """The captured Adam reads lr from device memory: mirror opt.param_groups[0]['lr'] (ReduceLROnPlateau and friends change it
        between steps) on EVERY path that feeds the graph -- load(), step_from_corpus() and step_prefetched()."""
        lr = float(self.opt.param_groups[0]['lr'])
        if lr != getattr(self, '_lr_seen', None):
            self.lr.fill_(lr); self._lr_seen = lr

    def step_from_corpus(self, corpus, chunk):
        """One train step on the narratives `chunk` (corpus indices) of a batching.DeviceCorpus: the GPU batcher gathers them straight
        into the graph's static inputs (no intermediate batch, no second copy), then the captured step replays."""
        lengths = corpus.batch_into(chunk, self.x, self.target, self.mask)
        self.inv_norm.fill_(1.0 / self.norm_fn(lengths))
        self._refresh_lr()
        if self.graph is None:
            self.capture()
        return self.replay()

    def _step(self):
        self.seed_off.add_(1)
        self.step_t.add_(1)
        self.model.train()
        from .evaluation import _call                  # models.py classes take (inputs, length, mask), multiTransformer.py's (inputs, mask, lengths)
        pred = _call(self.model, self.x, self.mask, self.lengths)
        loss, dpred = K.mse_loss_sum_normalised_dev(pred.detach(), self.target, self.inv_norm)
        old = K.set_deferred_weight_grads(True)        # MFN weight gradients overlap the encoder stacks' backward
        try:
            pred.backward(dpred.view_as(pred))
        finally:
            K.set_deferred_weight_grads(old)
        self.opt.step_dev(self.step_t, self.lr)        # joins the deferred batches first (FlatAdam.all_reduce_grads)
        self.opt.zero_grad()
        return loss

    def capture(self):
        """Warm up eagerly on a side stream (lazy initialisation: arena binding, Adam state, kernel attributes), undo
        the warm-up's parameter / optimizer updates, then capture one step.  Capturing executes nothing."""
        L = self._lib.lib()
        opt = self.opt
        warm = max(1, int(self.warmup))
        snap = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        had_state = {k: (m.clone(), v.clone()) for k, (m, v) in opt.state.items()}
        step0, count0 = self.step_t.clone(), opt.step_count
        self._lib.check(L.mt_set_seed_offset_ptr(self._lib.ptr(self.seed_off)))
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(warm):
                    self._step()
                with torch.no_grad():
                    for k, v in self.model.state_dict().items():
                        v.copy_(snap[k])
                    for k, (m, v) in opt.state.items():
                        if k in had_state:
                            m.copy_(had_state[k][0]); v.copy_(had_state[k][1])
                        else:
                            m.zero_(); v.zero_()
                    self.step_t.copy_(step0)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            opt.step_count = count0
            self.graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self.graph):
                    self.loss = self._step()
            except RuntimeError as e:
                self.graph = None
                if 'legacy stream' in str(e) or 'StreamCapture' in str(e):
                    raise RuntimeError('CUDA-graph capture of the train step failed because an autograd graph of an earlier EAGER iteration '
                                       'is still alive (its AccumulateGrad nodes are bound to the default stream): delete the references to '
                                       'earlier outputs / losses (del out, loss) and call zero_grad(set_to_none=True) before capturing') from e
                raise
        finally:
            self._lib.check(L.mt_set_seed_offset_ptr(None))      # eager calls keep their value seeds
        return self

    def replay(self):
        self.graph.replay()
        self.opt.step_count += 1
        return self.loss

    def __call__(self, inputs, mask, target, lengths):
        self.load(inputs, mask, target, lengths)
        if self.graph is None:
            self.capture()
        return self.replay()

    # ---- pipelined input path: the host->device copy of batch k + 1 runs on a copy stream while batch k computes ------
    def prefetch(self, inputs, mask, target, lengths):
        """Enqueue the host->device copy of the NEXT batch (pinned host tensors) on the copy stream into one of two staging
        sets and return at once.  `step_prefetched()` then moves it into the graph's static inputs with device-to-device
        copies (3 TB/s) and replays.  PCIe time (84 MB = 1.6 ms at B = 256) disappears behind the previous step."""
        if not hasattr(self, '_stage'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            mk = lambda: dict(x={m: torch.empty_like(v) for m, v in self.x.items()}, mask=torch.empty_like(self.mask),
                              target=torch.empty_like(self.target), ready=torch.cuda.Event(), free=torch.cuda.Event(), lengths=None)
            self._stage = [mk(), mk()]
            self._next_in, self._next_out = 0, 0
            for s_ in self._stage:
                s_['free'].record(torch.cuda.current_stream(self.device))
        st = self._stage[self._next_in]
        self._next_in ^= 1
        cs = self._copy_stream
        cs.wait_event(st['free'])                    # the device-to-device copies that last read this set are done
        with torch.cuda.stream(cs):
            for m in self.mods:
                st['x'][m].copy_(inputs[m], non_blocking=True)
            st['mask'].copy_(mask.reshape(self.B, self.T, 1), non_blocking=True)
            st['target'].copy_(target.reshape(self.B, self.T, 1), non_blocking=True)
            st['ready'].record(cs)
        st['lengths'] = lengths

    def step_prefetched(self):
        """Run the train step on the oldest prefetched batch; returns the loss [1] on the device."""
        st = self._stage[self._next_out]
        self._next_out ^= 1
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(st['ready'])
        for m in self.mods:
            self.x[m].copy_(st['x'][m], non_blocking=True)
        self.mask.copy_(st['mask'], non_blocking=True)
        self.target.copy_(st['target'], non_blocking=True)
        st['free'].record(cur)
        self.inv_norm.fill_(1.0 / self.norm_fn(st['lengths']))
        self._refresh_lr()
        if self.graph is None:
            self.capture()
        return self.replay()


class ResultPipe:
    """Read a step's result (loss, predictions) back to the host WITHOUT draining the stream every step: push(t) enqueues the
    device -> pinned-host copy of t behind the step that produced it and returns the host copy of the PREVIOUS push (complete by its
    own event, usually long since), so the host enqueues step k + 1 while step k runs.  drain() returns the last one.  Every step's
    result still reaches the host, one step late -- what a training loop that logs its losses does (MFT/train.py:143 reads loss.item()
    for the epoch average only)."""

    def __init__(self, like, device):
        self.host = [torch.empty(like.shape, dtype=like.dtype).pin_memory() for _ in range(2)]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.device = device
        self.n = 0

    def push(self, t):
        k = self.n & 1
        self.host[k].copy_(t, non_blocking=True)
        self.done[k].record(torch.cuda.current_stream(self.device))
        self.n += 1
        if self.n == 1:
            return None
        self.done[k ^ 1].synchronize()
        return self.host[k ^ 1]

    def drain(self):
        if self.n == 0:
            return None
        k = (self.n - 1) & 1
        self.done[k].synchronize()
        return self.host[k]


class GraphedForward:
    """eval() forward captured into a CUDA graph for a fixed (B, T); returns the static prediction buffer [B,T,1]."""

    def __init__(self, model, B, T, in_dims, device, warmup=2, input_dtype=torch.float32):
        self.model, self.B, self.T, self.device = model, B, T, device
        self.mods = list(in_dims)
        self._input_dtype = input_dtype
        # in_dims: mod -> feature width (hot-path models, inputs [B,T,d]) or mod -> (K, D) (MultiCNNTransformer variants, raw windows
        # [B,T,K,D]: the window front-end is then part of the captured step)
        self.x = {m: torch.zeros((B, T) + (tuple(d) if isinstance(d, (tuple, list)) else (d,)), device=device, dtype=input_dtype)
                  for m, d in in_dims.items()}
        self.mask = torch.zeros(B, T, 1, device=device)
        self.lengths = [T] * B
        self.warmup = warmup
        self.graph = None
        self.pred = None

    def load(self, inputs, mask):
        for m in self.mods:
            self.x[m].copy_(inputs[m], non_blocking=True)
        self.mask.copy_(mask.reshape(self.B, self.T, 1), non_blocking=True)

    def _fwd(self):
        self.model.eval()
        from .evaluation import _call                  # models.py classes take (inputs, length, mask), multiTransformer.py's (inputs, mask, lengths)
        with torch.no_grad():
            return _call(self.model, self.x, self.mask, self.lengths)

    def capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._fwd()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.pred = self._fwd()
        return self

    def __call__(self, inputs, mask):
        self.load(inputs, mask)
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.pred

    def prefetch(self, inputs, mask):
        """Host->device copy of the NEXT batch on a copy stream (see GraphedTrainStep.prefetch)."""
        if not hasattr(self, '_stage'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            mk = lambda: dict(x={m: torch.empty_like(v) for m, v in self.x.items()}, mask=torch.empty_like(self.mask),
                              ready=torch.cuda.Event(), free=torch.cuda.Event())
            self._stage = [mk(), mk()]
            self._next_in, self._next_out = 0, 0
            for s_ in self._stage:
                s_['free'].record(torch.cuda.current_stream(self.device))
        st = self._stage[self._next_in]
        self._next_in ^= 1
        cs = self._copy_stream
        cs.wait_event(st['free'])
        with torch.cuda.stream(cs):
            for m in self.mods:
                st['x'][m].copy_(inputs[m], non_blocking=True)
            st['mask'].copy_(mask.reshape(self.B, self.T, 1), non_blocking=True)
            st['ready'].record(cs)

    def forward_prefetched(self):
        st = self._stage[self._next_out]
        self._next_out ^= 1
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(st['ready'])
        for m in self.mods:
            self.x[m].copy_(st['x'][m], non_blocking=True)
        self.mask.copy_(st['mask'], non_blocking=True)
        st['free'].record(cur)
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.pred
