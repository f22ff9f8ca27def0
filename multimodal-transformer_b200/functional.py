"""torch.autograd glue over the C ABI: every forward/backward here is ONE call into libmt_b200.so.

Compute dtype is a process-wide switch (`set_compute_dtype`): 'fp32' (FFMA GEMMs, the 1e-5 parity mode, default)
or 'bf16' (bf16 GEMM operands, fp32 accumulation / residual stream / statistics / recurrent state).
Dropout is counter-based: keep = hash(seed, site, element) >= p * 2^32 (csrc/mt_common.cuh), one fresh seed per
forward call drawn from `next_seed()`; `manual_seed` makes runs reproducible and lets tests inject known masks.
"""
import ctypes

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, MT_BF16, MT_F32, check, lib, ptr, require, stream

_SITE_RESIDUAL = 0x7000      # stand-alone SublayerConnection dropout: its own site range (0x6000.. belongs to the window front-end)

import os as _os

_state = {'dtype': MT_F32, 'seed': 0x5EED0000, 'counter': 0, 'fixed_seed': None, 'parallel_stacks': True, 'defer_wgrad': False,
          'pending': [], 'wgrad_stream': {}, 'key_len': None, 'grouped_stacks': _os.environ.get('MT_GROUPED_STACKS', '1') != '0'}


def set_compute_dtype(name):
    name = {'float32': 'fp32', 'bfloat16': 'bf16'}.get(str(name).replace('torch.', ''), name)
    if name not in ('fp32', 'bf16'):
        raise ValueError("compute dtype must be 'fp32' or 'bf16'")
    _state['dtype'] = MT_F32 if name == 'fp32' else MT_BF16


def get_compute_dtype():
    return 'fp32' if _state['dtype'] == MT_F32 else 'bf16'


def set_parallel_stacks(on):
    """Run the per-modality encoder stacks of MultiTransformer on side CUDA streams (default on).  Returns the previous
    setting.  The per-launch profiler (mt_prof_*) assumes one stream: switch this off while profiling."""
    old = _state['parallel_stacks']
    _state['parallel_stacks'] = bool(on)
    return old


def parallel_stacks():
    return _state['parallel_stacks']


def set_grouped_stacks(on):
    """Run the per-modality encoder stacks of MultiTransformer as ONE grouped call (mt_encoder_group_fwd / _bwd: one launch per
    projection / LayerNorm / weight gradient serves all stacks; default on) instead of one call per stack.  Returns the previous
    setting."""
    old = _state['grouped_stacks']
    _state['grouped_stacks'] = bool(on)
    return old


def grouped_stacks():
    return _state['grouped_stacks']


def set_deferred_weight_grads(on):
    """While on, MFN.backward enqueues its batched weight / bias gradients on a side stream and returns as soon as the input
    gradients are enqueued, so the encoder stacks' backward overlaps them.  The caller MUST call `join_deferred()` on the
    stream that consumes the parameter gradients (GraphedTrainStep / FlatAdam do) -- plain torch.optim users leave it off."""
    old = _state['defer_wgrad']
    _state['defer_wgrad'] = bool(on)
    return old


def join_deferred():
    """Make the current stream wait for every deferred weight-gradient batch enqueued so far."""
    cur = torch.cuda.current_stream()
    for ev in _state['pending']:
        cur.wait_event(ev)
    _state['pending'].clear()


class ragged_batch:
    """Context manager for INFERENCE on a padded batch whose narratives have different lengths:

        with mtb.ragged_batch(lengths):
            pred = model(inputs, mask, lengths)

    Inside it, narrative b only HAS its first lengths[b] windows: attention keys beyond them are excluded, so the valid part of
    every prediction equals a forward of that narrative alone -- what the reference's evaluation computes with batch_size = 1
    (MFT/train.py:169,218).  Without it a padded batch follows the reference's TRAINING semantics, where padded windows are
    live attention keys.  No gradients, no dropout: a model in train() mode raises."""

    def __init__(self, lengths):
        self.lengths = lengths
        self.old = None

    def __enter__(self):
        self.old = _state['key_len']
        _state['key_len'] = self.lengths
        return self

    def __exit__(self, *exc):
        _state['key_len'] = self.old
        return False


def _key_len(B, T, device, need_grad, p_drop):
    """The active ragged_batch lengths as a device int32 [B] (None outside the context manager)."""
    kl = _state['key_len']
    if kl is None:
        return None
    if need_grad or p_drop > 0:
        raise RuntimeError('ragged_batch is inference only: call model.eval() and run under torch.no_grad()')
    if not torch.is_tensor(kl) or kl.device != device or kl.dtype != torch.int32:
        kl = (kl.to(device=device, dtype=torch.int32) if torch.is_tensor(kl) else
              torch.tensor([int(v) for v in kl], dtype=torch.int32).to(device))
        _state['key_len'] = kl          # converted once per context, reused by every stack / layer call
    if kl.numel() != B:
        raise RuntimeError(f'ragged_batch: {kl.numel()} lengths for a batch of {B} narratives')
    return kl.contiguous()


def manual_seed(seed):
    _state['seed'] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state['counter'] = 0


def fix_seed(seed):
    """Every subsequent forward uses exactly this dropout seed (tests); None restores the counter."""
    _state['fixed_seed'] = None if seed is None else int(seed)


def next_seed():
    if _state['fixed_seed'] is not None:
        return _state['fixed_seed']
    # consecutive calls are a large odd stride apart: a captured graph adds +1 per replay to every seed (seed_off), so a unit stride
    # would hand call k at replay r the seed of call k + 1 at replay r - 1 (identical masks on shared sites)
    _state['counter'] += 1
    return (_state['seed'] * 0x9E3779B97F4A7C15 + _state['counter'] * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def _apply(fn, *args):
    """fn.apply(*args) with the caller's grad mode made visible to fn.forward: autograd runs every Function.forward with grad mode off
    and reports `needs_input_grad` from `requires_grad` alone, so without this an inference call under torch.no_grad() would keep the
    whole backward stash."""
    old = _state.get('grad', True)
    _state['grad'] = torch.is_grad_enabled()
    try:
        return fn.apply(*args)
    finally:
        _state['grad'] = old


def _need_grad(ctx):
    return _state.get('grad', True) and any(ctx.needs_input_grad)


def _adt():
    return torch.float32 if _state['dtype'] == MT_F32 else torch.bfloat16


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def mask2d(mask, B, T, device):
    """[B,T,1] / [B,T] mask of any dtype -> contiguous fp32 [B,T] (None stays None)."""
    if mask is None:
        return None
    m = mask.reshape(B, T)
    if m.dtype != torch.float32:
        m = m.float()
    if m.device != device:
        raise RuntimeError('mask must live on the same device as the inputs')
    return m.contiguous()


# ---------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = act(dropout_in(x) W^T + b) [* rowmask]   (nn.Linear call sites; mt_linear_fwd / mt_linear_bwd)."""

    @staticmethod
    def forward(ctx, x, W, b, act, rowmask, in_drop_p, out_f32):
        dt = _state['dtype']
        shp = x.shape
        K = shp[-1]
        N = W.shape[0]
        x2 = x.reshape(-1, K)
        x2 = require(x2 if x2.is_contiguous() else x2.contiguous(), name='x')
        if x2.dtype not in (torch.float32, torch.bfloat16) or (dt == MT_F32 and x2.dtype != torch.float32):
            raise RuntimeError(f'Linear input dtype {x2.dtype} does not match compute dtype {get_compute_dtype()}')
        require(W, torch.float32, 'weight')
        if b is not None:
            require(b, torch.float32, 'bias')
        _lib.check_device(x2.device.index)
        M = x2.shape[0]
        x_f32 = int(x2.dtype == torch.float32)
        y_f32 = int(out_f32 or dt == MT_F32)
        y = torch.empty((M, N), dtype=torch.float32 if y_f32 else torch.bfloat16, device=x2.device)
        seed = next_seed() if in_drop_p > 0 else 0
        L = lib()
        ws = _ws(L.mt_linear_ws_bytes(dt, M, N, K, x_f32, in_drop_p), x2.device)
        check(L.mt_linear_fwd(dt, M, N, K, ptr(x2), x_f32, ptr(W), ptr(b), ptr(y), y_f32, act, ptr(rowmask), in_drop_p, seed, 0x5000,
                              ptr(ws), ws.numel(), stream()))
        ctx.save_for_backward(x2, W, y if act != ACT_NONE else None, rowmask)
        ctx.meta = (dt, M, N, K, x_f32, y_f32, act, in_drop_p, seed, shp, b is not None)
        return y.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, W, y, rowmask = ctx.saved_tensors
        dt, M, N, K, x_f32, y_f32, act, in_drop_p, seed, shp, has_b = ctx.meta
        dy2 = dy.reshape(M, N)
        dy2 = require(dy2 if dy2.is_contiguous() else dy2.contiguous(), name='dy')
        dy_f32 = int(dy2.dtype == torch.float32)
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty((M, K), dtype=torch.float32 if dt == MT_F32 else torch.bfloat16, device=dy2.device) if need_dx else None
        dW = torch.empty_like(W)
        db = torch.empty(N, dtype=torch.float32, device=dy2.device) if has_b else None
        L = lib()
        ws = _ws(L.mt_linear_bwd_ws_bytes(dt, M, N, K, x_f32, in_drop_p), dy2.device)
        check(L.mt_linear_bwd(dt, M, N, K, ptr(x2), x_f32, ptr(W), ptr(y), y_f32, ptr(dy2), dy_f32, act, ptr(rowmask), in_drop_p, seed,
                              0x5000, ptr(dx), ptr(dW), ptr(db), ptr(ws), ws.numel(), stream()))
        if dx is not None:
            dx = dx.view(shp)
            if x_f32 and dx.dtype != torch.float32:
                dx = dx.float()
        return dx, dW, db, None, None, None, None


class ConcatFn(torch.autograd.Function):
    """torch.cat(feats, dim=-1) in front of the fusion / embed Linear (SFT/models.py:136-138, B2-Trans/models.py:130-132), assembled by
    the library's strided cast kernel in the compute dtype (mt_concat_fwd / mt_concat_bwd)."""

    @staticmethod
    def forward(ctx, *feats):
        import ctypes
        dt = _state['dtype']
        shp = feats[0].shape[:-1]
        widths = [int(f.shape[-1]) for f in feats]
        xs = []
        for f in feats:
            if f.shape[:-1] != shp:
                raise RuntimeError('concat: leading dimensions differ')
            f2 = f.reshape(-1, f.shape[-1])
            f2 = require(f2 if f2.is_contiguous() else f2.contiguous(), name='concat input')
            if f2.dtype not in (torch.float32, torch.bfloat16):
                raise RuntimeError(f'concat input dtype {f2.dtype}')
            xs.append(f2)
        _lib.check_device(xs[0].device.index)
        M, n, Kt = xs[0].shape[0], len(xs), sum(widths)
        out = torch.empty((M, Kt), dtype=torch.float32 if dt == MT_F32 else torch.bfloat16, device=xs[0].device)
        srcs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
        wd = (ctypes.c_int * n)(*widths)
        f32 = (ctypes.c_int * n)(*[int(x.dtype == torch.float32) for x in xs])
        check(lib().mt_concat_fwd(M, n, srcs, wd, f32, ptr(out), Kt, int(out.dtype == torch.float32), stream()))
        ctx.meta = (M, widths, [x.dtype for x in xs], [f.shape for f in feats])
        return out.view(*shp, Kt)

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        M, widths, dtypes, shapes = ctx.meta
        Kt = sum(widths)
        dy2 = dy.reshape(M, Kt)
        dy2 = require(dy2 if dy2.is_contiguous() else dy2.contiguous(), name='dy')
        n = len(widths)
        outs = [torch.empty((M, w), dtype=d, device=dy2.device) if ctx.needs_input_grad[i] else None for i, (w, d) in enumerate(zip(widths, dtypes))]
        dsts = (ctypes.c_void_p * n)(*[o.data_ptr() if o is not None else None for o in outs])
        wd = (ctypes.c_int * n)(*widths)
        f32 = (ctypes.c_int * n)(*[int(d == torch.float32) for d in dtypes])
        check(lib().mt_concat_bwd(M, n, dsts, wd, f32, ptr(dy2), Kt, int(dy2.dtype == torch.float32), stream()))
        return tuple(o.view(s) if o is not None else None for o, s in zip(outs, shapes))


def concat_features(feats):
    return ConcatFn.apply(*feats)


def linear(x, W, b=None, act=ACT_NONE, rowmask=None, in_drop_p=0.0, out_f32=False):
    return LinearFn.apply(x, W, b, act, rowmask, float(in_drop_p), out_f32)


# ---------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """a_2 * (x - mean) / (std_unbiased + eps) + b_2   (LayerNorm.forward MFT/multiTransformer.py:88-91)."""

    @staticmethod
    def forward(ctx, x, a, b, eps):
        d = x.shape[-1]
        x2 = x.reshape(-1, d)
        x2 = require(x2 if x2.is_contiguous() else x2.contiguous(), torch.float32, 'x')
        require(a, torch.float32, 'a_2'); require(b, torch.float32, 'b_2')
        _lib.check_device(x2.device.index)
        y = torch.empty_like(x2)
        check(lib().mt_layernorm_fwd(MT_F32, x2.shape[0], d, ptr(x2), ptr(a), ptr(b), eps, ptr(y), 1, stream()))
        ctx.save_for_backward(x2, a)
        ctx.eps = eps
        ctx.shp = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, a = ctx.saved_tensors
        d = x2.shape[1]
        dy2 = dy.reshape(-1, d)
        dy2 = require(dy2 if dy2.is_contiguous() else dy2.contiguous(), torch.float32, 'dy')
        dx = torch.empty_like(x2)
        da = torch.zeros_like(a); db = torch.zeros_like(a)
        check(lib().mt_layernorm_bwd(MT_F32, x2.shape[0], d, ptr(x2), ptr(a), ctx.eps, ptr(dy2), 1, None, ptr(dx), ptr(da), ptr(db),
                                     stream()))
        return dx.view(ctx.shp), da, db, None


def layer_norm(x, a, b, eps=1e-6):
    return LayerNormFn.apply(x, a, b, float(eps))


# ---------------------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """Fused scores / query-row mask / softmax / dropout / PV / head merge on packed qkv [B,T,3d]."""

    @staticmethod
    def forward(ctx, qkv, mask, h, p_drop):
        dt = _state['dtype']
        B, T, d3 = qkv.shape
        d = d3 // 3
        require(qkv, _adt(), 'qkv')
        _lib.check_device(qkv.device.index)
        m = mask2d(mask, B, T, qkv.device)
        out = torch.empty((B, T, d), dtype=qkv.dtype, device=qkv.device)
        kl = _key_len(B, T, qkv.device, _need_grad(ctx), p_drop)
        if kl is not None:
            check(lib().mt_attention_ragged_fwd(dt, B, T, d, h, ptr(qkv), ptr(m), ptr(kl), ptr(out), stream()))
            return out
        lse = torch.empty((B, h, T), dtype=torch.float32, device=qkv.device)
        seed = next_seed() if p_drop > 0 else 0
        check(lib().mt_attention_fwd(dt, B, T, d, h, ptr(qkv), ptr(m), ptr(out), ptr(lse), p_drop, seed, 0, stream()))
        ctx.save_for_backward(qkv, m, out, lse)
        ctx.meta = (dt, B, T, d, h, p_drop, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, m, out, lse = ctx.saved_tensors
        dt, B, T, d, h, p_drop, seed = ctx.meta
        dout = require(dout if dout.is_contiguous() else dout.contiguous(), qkv.dtype, 'dout')
        dqkv = torch.empty_like(qkv)
        L = lib()
        ws = _ws(L.mt_attention_bwd_ws_bytes(B, T, h), qkv.device)
        check(L.mt_attention_bwd(dt, B, T, d, h, ptr(qkv), ptr(m), ptr(out), ptr(lse), ptr(dout), ptr(dqkv), p_drop, seed, 0, ptr(ws),
                                 ws.numel(), stream()))
        return dqkv, None, None, None


def attention_packed(qkv, mask, h, p_drop=0.0):
    return _apply(AttentionFn, qkv, mask, h, float(p_drop))


def attention_probs(qkv, mask, h):
    B, T, d3 = qkv.shape
    require(qkv, _adt(), 'qkv')
    m = mask2d(mask, B, T, qkv.device)
    probs = torch.empty((B, h, T, T), dtype=torch.float32, device=qkv.device)
    check(lib().mt_attention_probs(_state['dtype'], B, T, d3 // 3, h, ptr(qkv), ptr(m), ptr(probs), stream()))
    return probs


# ---------------------------------------------------------------------------------------------------------
class ResidualDropoutFn(torch.autograd.Function):
    """x + dropout(y)   (SublayerConnection, stand-alone path); fp32."""

    @staticmethod
    def forward(ctx, x, y, p):
        require(x, torch.float32, 'x')
        y = require(y if y.is_contiguous() else y.contiguous(), torch.float32, 'y')
        out = torch.empty_like(x)
        seed = next_seed() if p > 0 else 0
        check(lib().mt_residual_dropout_fwd(ptr(x), ptr(y), ptr(out), x.numel(), p, seed, _SITE_RESIDUAL, stream()))
        ctx.meta = (p, seed)
        return out

    @staticmethod
    def backward(ctx, g):
        p, seed = ctx.meta
        g = require(g if g.is_contiguous() else g.contiguous(), torch.float32, 'g')
        if p <= 0:
            return g, g, None
        gy = torch.empty_like(g)
        check(lib().mt_dropout_bwd(ptr(g), ptr(gy), g.numel(), p, seed, _SITE_RESIDUAL, stream()))
        return g, gy, None


def residual_dropout(x, y, p):
    if not x.is_contiguous():
        x = x.contiguous()
    return ResidualDropoutFn.apply(x, y, float(p))


# ---------------------------------------------------------------------------------------------------------
class Arena:
    """A flat fp32 buffer that OWNS the storage of a list of nn.Parameters (each parameter's .data is a view),
    in the canonical order of a C-ABI parameter block, plus its bf16 shadow."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.offsets = []
        o = 0
        for n in self.sizes:
            self.offsets.append(o)
            o += n
        self.total = o
        self.flat = None
        self.lp = None
        self._lp_key = None

    def bound(self):
        f = self.flat
        if f is None:
            return False
        base = f.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                return False
        return True

    def bind(self):
        """(Re)point every parameter at the flat buffer; values are preserved.  Cheap no-op when already bound."""
        if self.bound():
            return self.flat
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('parameters must be on a CUDA device (no CPU fallback); call .to("cuda") first')
        flat = torch.empty(self.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(self.params, self.offsets, self.sizes):
                v = flat[o:o + n].view(p.shape)
                v.copy_(p.data)
                p.data = v
        self.flat = flat
        self.lp = None
        self._lp_key = None
        return flat

    def shadow(self):
        """bf16 copy of the arena, refreshed when any parameter changed (version counters)."""
        key = sum(p._version for p in self.params)
        if self.lp is None or self.lp.device != self.flat.device:
            self.lp = torch.empty(self.total, dtype=torch.bfloat16, device=self.flat.device)
            self._lp_key = None
        if key != self._lp_key:
            check(lib().mt_cast_f32_to_bf16(ptr(self.flat), ptr(self.lp), self.total, stream()))
            self._lp_key = key
        return self.lp

    def grad_views(self, gflat):
        return [gflat[o:o + n].view(p.shape) for p, o, n in zip(self.params, self.offsets, self.sizes)]

    def flat_grad(self):
        """The flat gradient buffer if every .grad is still the view handed out by backward, else None."""
        g0 = self.params[0].grad
        if g0 is None:
            return None
        base = g0.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + 4 * o:
                return None
        if (g0.storage_offset() + self.total) * 4 > g0.untyped_storage().nbytes():
            return None          # adjacent by accident of the allocator, not one buffer
        return g0.as_strided((self.total,), (1,), g0.storage_offset())


# ---------------------------------------------------------------------------------------------------------
class EncoderFn(torch.autograd.Function):
    """Whole encoder stack (N pre-norm layers + final LayerNorm) in one C call each way."""

    @staticmethod
    def forward(ctx, x, mask, arena, cfgd, *params):
        dt = _state['dtype']
        B, T, d = x.shape
        x = require(x if x.is_contiguous() else x.contiguous(), torch.float32, 'encoder input (fp32 residual stream)')
        _lib.check_device(x.device.index)
        flat = arena.bind()
        lp = arena.shadow() if dt == MT_BF16 else None
        m = mask2d(mask, B, T, x.device)
        need_grad = _need_grad(ctx)
        p_drop = float(cfgd['p_drop'])
        cfg = _lib.MtEncoderCfg(B, T, d, cfgd['h'], cfgd['dff'], cfgd['n_layers'], dt, int(need_grad), p_drop,
                                next_seed() if p_drop > 0 else 0, cfgd['stack_id'], int(cfgd.get('y_f32', dt == MT_F32)),
                                int(cfgd.get('grid_share', 0)), None)
        kl = _key_len(B, T, x.device, need_grad, p_drop)
        if kl is not None:
            cfg.key_len = kl.data_ptr()
        L = lib()
        nws = L.mt_encoder_ws_bytes(ctypes.byref(cfg))
        if nws == 0:
            raise RuntimeError(f'unsupported encoder configuration {cfgd} for input {tuple(x.shape)}')
        ws = _ws(nws, x.device)
        y = torch.empty((B, T, d), dtype=torch.float32 if cfg.y_f32 else torch.bfloat16, device=x.device)
        check(L.mt_encoder_fwd(ctypes.byref(cfg), ptr(flat), ptr(lp), ptr(x), ptr(m), ptr(y), ptr(ws), ws.numel(), stream()))
        if need_grad:
            ctx.save_for_backward(x, m, ws, flat, lp)
            ctx.cfg = cfg
            ctx.arena = arena
        return y

    @staticmethod
    def backward(ctx, dy):
        x, m, ws, flat, lp = ctx.saved_tensors
        cfg = ctx.cfg
        dy = dy if dy.is_contiguous() else dy.contiguous()
        want = torch.float32 if cfg.y_f32 else torch.bfloat16
        if dy.dtype != want:
            dy = dy.to(want)
        dx = torch.empty_like(x)
        g = torch.empty(ctx.arena.total, dtype=torch.float32, device=x.device)
        check(lib().mt_encoder_bwd(ctypes.byref(cfg), ptr(flat), ptr(lp), ptr(x), ptr(m), ptr(dy), ptr(dx), ptr(g), ptr(ws), ws.numel(),
                                   stream()))
        return (dx, None, None, None, *ctx.arena.grad_views(g))


def encoder_stack(x, mask, arena, cfgd):
    return _apply(EncoderFn, x, mask, arena, cfgd, *arena.params)


class EncoderGroupFn(torch.autograd.Function):
    """The per-modality embed Linears (MFT/multiTransformer.py:270,296) and encoder stacks (:278-299) of MultiTransformer as ONE grouped
    call each way: every stack has the same shape, so their rows are laid out back to back and one launch per projection / LayerNorm /
    weight gradient serves all of them (mt_encoder_group_fwd / _bwd).  Inputs: G raw inputs, then (weight, bias) of the G embeds, then
    the parameters of the group arena.  Returns G views [B,T,d] of one output block."""

    @staticmethod
    def forward(ctx, mask, arena, cfgd, G, *tensors):
        dt = _state['dtype']
        xs, emb = list(tensors[:G]), tensors[G:3 * G]
        B, T = xs[0].shape[0], xs[0].shape[1]
        d = emb[0].shape[0]
        M = B * T
        dev = xs[0].device
        _lib.check_device(dev.index)
        L = lib()
        flat = arena.bind()
        lp = arena.shadow() if dt == MT_BF16 else None
        pstride = arena.total // G
        m = mask2d(mask, B, T, dev)
        need_grad = _need_grad(ctx)
        p_drop = float(cfgd['p_drop'])
        seeds = [next_seed() if p_drop > 0 else 0 for _ in range(G)]
        y_f32 = int(cfgd.get('y_f32', dt == MT_F32))
        cfg = _lib.MtEncoderCfg(B, T, d, cfgd['h'], cfgd['dff'], cfgd['n_layers'], dt, int(need_grad), p_drop, seeds[0],
                                cfgd['stack_ids'][0], y_f32, 0, None)
        kl = _key_len(B, T, dev, need_grad, p_drop)
        if kl is not None:
            cfg.key_len = kl.data_ptr()
        # embeds straight into the stacked residual-stream block
        x0 = torch.empty((G, B, T, d), dtype=torch.float32, device=dev)
        xin = []
        for g in range(G):
            x2 = xs[g].reshape(M, -1)
            x2 = require(x2 if x2.is_contiguous() else x2.contiguous(), name='x')
            if x2.dtype not in (torch.float32, torch.bfloat16) or (dt == MT_F32 and x2.dtype != torch.float32):
                raise RuntimeError(f'embed input dtype {x2.dtype} does not match compute dtype {get_compute_dtype()}')
            W, b = emb[2 * g], emb[2 * g + 1]
            require(W, torch.float32, 'embed weight'); require(b, torch.float32, 'embed bias')
            Kg = x2.shape[1]
            x_f32 = int(x2.dtype == torch.float32)
            ws = _ws(L.mt_linear_ws_bytes(dt, M, d, Kg, x_f32, 0.0), dev)
            check(L.mt_linear_fwd(dt, M, d, Kg, ptr(x2), x_f32, ptr(W), ptr(b), ctypes.c_void_p(x0[g].data_ptr()), 1, ACT_NONE, None, 0.0, 0,
                                  0x5000, ptr(ws), ws.numel(), stream()))
            xin.append(x2)
        nws = L.mt_encoder_group_ws_bytes(ctypes.byref(cfg), G)
        if nws == 0:
            raise RuntimeError(f'unsupported grouped encoder configuration {cfgd} for {G} x {(B, T, d)}')
        ws = _ws(nws, dev)
        y = torch.empty((G, B, T, d), dtype=torch.float32 if y_f32 else torch.bfloat16, device=dev)
        c_seeds = (ctypes.c_uint64 * G)(*seeds)
        c_ids = (ctypes.c_int * G)(*[int(i) for i in cfgd['stack_ids']])
        check(L.mt_encoder_group_fwd(ctypes.byref(cfg), G, c_seeds, c_ids, ptr(flat), ptr(lp), pstride, ptr(x0), ptr(m), ptr(y), ptr(ws),
                                     ws.numel(), stream()))
        if need_grad:
            ctx.save_for_backward(x0, m, ws, flat, lp, *xin, *emb)
            ctx.cfg, ctx.arena, ctx.G, ctx.seeds, ctx.ids, ctx.pstride = cfg, arena, G, c_seeds, c_ids, pstride
            ctx.x_need = ctx.needs_input_grad[4:4 + G]
            ctx.x_shapes = [tuple(x.shape) for x in xs]
        return tuple(y.unbind(0))

    @staticmethod
    def backward(ctx, *dys):
        x0, m, ws, flat, lp, *rest = ctx.saved_tensors
        G, cfg = ctx.G, ctx.cfg
        xin, emb = rest[:G], rest[G:]
        dt = cfg.dtype
        M, d = cfg.B * cfg.T, cfg.d
        dev = x0.device
        want = torch.float32 if cfg.y_f32 else torch.bfloat16
        # the G output gradients, back to back: in place when they already are one block (MfnFn.backward hands them out that way)
        step = M * d * (4 if cfg.y_f32 else 2)
        ok = all(t is not None and t.dtype == want and t.is_contiguous() for t in dys)
        if ok and all(dys[g].data_ptr() == dys[0].data_ptr() + g * step for g in range(G)):
            dy = dys[0]
        else:
            dy = torch.stack([torch.zeros((cfg.B, cfg.T, d), dtype=want, device=dev) if t is None else t.to(want).reshape(cfg.B, cfg.T, d)
                              for t in dys])
        dx = torch.empty_like(x0)
        g_all = torch.empty(ctx.arena.total, dtype=torch.float32, device=dev)
        L = lib()
        check(L.mt_encoder_group_bwd(ctypes.byref(cfg), G, ctx.seeds, ctx.ids, ptr(flat), ptr(lp), ctx.pstride, ptr(x0), ptr(m), ptr(dy),
                                     ptr(dx), ptr(g_all), ptr(ws), ws.numel(), stream()))
        emb_grads, in_grads = [], []
        # one block for the embed gradients, in parameter order: an optimizer arena over the embeds reads it in place
        eg = torch.empty(sum(t_.numel() for t_ in emb), dtype=torch.float32, device=dev)
        eo = 0
        for g in range(G):
            W = emb[2 * g]
            Kg = W.shape[1]
            x_f32 = int(xin[g].dtype == torch.float32)
            dW = eg[eo:eo + W.numel()].view(W.shape)
            db = eg[eo + W.numel():eo + W.numel() + d]
            eo += W.numel() + d
            dxin = torch.empty((M, Kg), dtype=torch.float32 if dt == MT_F32 else torch.bfloat16, device=dev) if ctx.x_need[g] else None
            wsl = _ws(L.mt_linear_bwd_ws_bytes(dt, M, d, Kg, x_f32, 0.0), dev)
            check(L.mt_linear_bwd(dt, M, d, Kg, ptr(xin[g]), x_f32, ptr(W), None, 1, ctypes.c_void_p(dx[g].data_ptr()), 1, ACT_NONE, None, 0.0,
                                  0, 0x5000, ptr(dxin), ptr(dW), ptr(db), ptr(wsl), wsl.numel(), stream()))
            emb_grads += [dW, db]
            if dxin is not None:
                dxin = dxin.view(ctx.x_shapes[g])
                if x_f32 and dxin.dtype != torch.float32:
                    dxin = dxin.float()
            in_grads.append(dxin)
        return (None, None, None, None, *in_grads, *emb_grads, *ctx.arena.grad_views(g_all))


def encoder_stack_group(xs, mask, embeds, arena, cfgd):
    """xs: G raw inputs [B,T,K_g]; embeds: G nn.Linear modules; arena: the group arena (stack-major).  Returns G tensors [B,T,d]."""
    G = len(xs)
    emb = []
    for e in embeds:
        emb += [e.weight, e.bias]
    return _apply(EncoderGroupFn, mask, arena, cfgd, G, *xs, *emb, *arena.params)


# ---------------------------------------------------------------------------------------------------------
class MfnFn(torch.autograd.Function):
    """MFN.forward (hoisted input projections + persistent recurrence kernel + head) and its BPTT."""

    @staticmethod
    def forward(ctx, mask, arena, cfgd, t_major, n_mods, *rest):
        dt = _state['dtype']
        xs = list(rest[:n_mods])
        adt = _adt()
        xs = [require(x if x.is_contiguous() else x.contiguous(), adt, 'MFN input') for x in xs]
        if t_major:
            T, B = xs[0].shape[0], xs[0].shape[1]
        else:
            B, T = xs[0].shape[0], xs[0].shape[1]
        dev = xs[0].device
        _lib.check_device(dev.index)
        flat = arena.bind()
        lp = arena.shadow() if dt == MT_BF16 else None
        m = mask2d(mask, B, T, dev)
        need_grad = _need_grad(ctx)
        cfg = _lib.MtMfnCfg()
        cfg.B, cfg.T, cfg.n_mods = B, T, n_mods
        for i in range(n_mods):
            cfg.in_dim[i] = xs[i].shape[2]
            cfg.hid[i] = cfgd['hid'][i]
        cfg.mem_dim, cfg.h_att1, cfg.h_att2, cfg.h_gamma, cfg.h_out = cfgd['mem'], cfgd['a1'], cfgd['a2'], cfgd['g'], cfgd['o']
        cfg.dtype, cfg.training = dt, int(need_grad)
        cfg.p_gamma, cfg.p_out = float(cfgd['p_gamma']), float(cfgd['p_out'])
        cfg.seed = next_seed() if (cfg.p_gamma > 0 or cfg.p_out > 0) else 0
        L = lib()
        nws = L.mt_mfn_ws_bytes(ctypes.byref(cfg))
        if nws == 0 or L.mt_mfn_param_count(ctypes.byref(cfg)) != arena.total:
            raise RuntimeError('unsupported MFN configuration')
        ws = _ws(nws, dev)
        xp = (ctypes.c_void_p * n_mods)(*[x.data_ptr() for x in xs])
        if t_major:
            sb = (ctypes.c_int64 * n_mods)(*[x.shape[2] for x in xs])
            st = (ctypes.c_int64 * n_mods)(*[x.shape[2] * B for x in xs])
        else:
            sb = (ctypes.c_int64 * n_mods)(*[x.shape[2] * T for x in xs])
            st = (ctypes.c_int64 * n_mods)(*[x.shape[2] for x in xs])
        out = torch.empty((B, T), dtype=torch.float32, device=dev)
        Hs = sum(cfgd['hid'])
        h_last = torch.empty((B, Hs), dtype=torch.float32, device=dev)
        c_last = torch.empty((B, Hs), dtype=torch.float32, device=dev)
        mem_last = torch.empty((B, cfgd['mem']), dtype=torch.float32, device=dev)
        check(L.mt_mfn_fwd(ctypes.byref(cfg), ptr(flat), ptr(lp), xp, sb, st, ptr(m), ptr(out), ptr(h_last), ptr(c_last), ptr(mem_last),
                           ptr(ws), ws.numel(), stream()))
        if need_grad:
            ctx.save_for_backward(m, ws, flat, lp, *xs)
            ctx.cfg, ctx.arena, ctx.t_major, ctx.n_mods = cfg, arena, t_major, n_mods
            ctx.x_need = ctx.needs_input_grad[5:5 + n_mods]
        ctx.mark_non_differentiable(h_last, c_last, mem_last)
        return out.unsqueeze(-1), h_last, c_last, mem_last

    @staticmethod
    def backward(ctx, dout, *_unused):
        m, ws, flat, lp, *xs = ctx.saved_tensors
        cfg, n_mods = ctx.cfg, ctx.n_mods
        B, T = cfg.B, cfg.T
        dev = xs[0].device
        dout = dout.reshape(B, T)
        dout = dout if dout.is_contiguous() else dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        if all(ctx.x_need) and all(x.shape == xs[0].shape and x.dtype == xs[0].dtype for x in xs):
            # one block for all modalities: the grouped encoder backward reads it in place (no stacking copy)
            dxs = list(torch.empty((n_mods,) + tuple(xs[0].shape), dtype=xs[0].dtype, device=dev).unbind(0))
        else:
            dxs = [torch.empty_like(x) if need else None for x, need in zip(xs, ctx.x_need)]
        dxp = (ctypes.c_void_p * n_mods)(*[None if d is None else d.data_ptr() for d in dxs])
        xp = (ctypes.c_void_p * n_mods)(*[x.data_ptr() for x in xs])
        if ctx.t_major:
            sb = (ctypes.c_int64 * n_mods)(*[x.shape[2] for x in xs])
            st = (ctypes.c_int64 * n_mods)(*[x.shape[2] * B for x in xs])
        else:
            sb = (ctypes.c_int64 * n_mods)(*[x.shape[2] * T for x in xs])
            st = (ctypes.c_int64 * n_mods)(*[x.shape[2] for x in xs])
        g = torch.empty(ctx.arena.total, dtype=torch.float32, device=dev)
        if _state['defer_wgrad']:
            # phase 1 (recurrences + input gradients) here; phase 2 (batched weight / bias gradients, which nothing downstream
            # reads) on a side stream, joined by join_deferred() before the optimizer touches the gradients
            cur = torch.cuda.current_stream(dev)
            side = _state['wgrad_stream'].get(dev)
            if side is None:
                side = _state['wgrad_stream'][dev] = torch.cuda.Stream(device=dev)
            cfg.bwd_phase = 1
            check(lib().mt_mfn_bwd(ctypes.byref(cfg), ptr(flat), ptr(lp), xp, sb, st, ptr(m), ptr(dout), dxp, ptr(g), ptr(ws), ws.numel(),
                                   stream()))
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cfg.bwd_phase = 2
                check(lib().mt_mfn_bwd(ctypes.byref(cfg), ptr(flat), ptr(lp), xp, sb, st, ptr(m), ptr(dout), dxp, ptr(g), ptr(ws), ws.numel(),
                                       stream()))
                ev = torch.cuda.Event()
                ev.record(side)
            cfg.bwd_phase = 0
            for t_ in (g, ws, flat, m, dout, *xs) + ((lp,) if lp is not None else ()):
                t_.record_stream(side)
            _state['pending'].append(ev)
        else:
            cfg.bwd_phase = 0
            check(lib().mt_mfn_bwd(ctypes.byref(cfg), ptr(flat), ptr(lp), xp, sb, st, ptr(m), ptr(dout), dxp, ptr(g), ptr(ws), ws.numel(),
                                   stream()))
        return (None, None, None, None, None, *dxs, *ctx.arena.grad_views(g))


def mfn_forward(xs, mask, arena, cfgd, t_major):
    return _apply(MfnFn, mask, arena, cfgd, bool(t_major), len(xs), *xs, *arena.params)


# ---------------------------------------------------------------------------------------------------------
def mse_loss_sum_normalised(pred, target, norm):
    """sum((pred-target)^2) / norm and d/dpred, in one kernel (MFT/train.py:135-139).  Returns (loss[1], dpred)."""
    p = require(pred if pred.is_contiguous() else pred.contiguous(), torch.float32, 'pred')
    t = require(target if target.is_contiguous() else target.contiguous(), torch.float32, 'target')
    loss = torch.zeros(1, dtype=torch.float32, device=p.device)
    dp = torch.empty_like(p)
    check(lib().mt_mse_loss_fwd_bwd(ptr(p), ptr(t), p.numel(), 1.0 / float(norm), ptr(loss), ptr(dp), stream()))
    return loss, dp


def mse_loss_sum_normalised_dev(pred, target, inv_norm_t):
    """Same as mse_loss_sum_normalised with 1/norm read from a device scalar (captured-graph train step)."""
    p = require(pred if pred.is_contiguous() else pred.contiguous(), torch.float32, 'pred')
    t = require(target if target.is_contiguous() else target.contiguous(), torch.float32, 'target')
    require(inv_norm_t, torch.float32, 'inv_norm')
    loss = torch.zeros(1, dtype=torch.float32, device=p.device)
    dp = torch.empty_like(p)
    check(lib().mt_mse_loss_fwd_bwd_dev(ptr(p), ptr(t), p.numel(), ptr(inv_norm_t), ptr(loss), ptr(dp), stream()))
    return loss, dp


def adam_step_flat_dev(p, g, m, v, step_t, lr_t, lr=0.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, p_lp=None):
    """Adam with the step count (int64 device scalar) and optionally lr (fp32 device scalar) read on the device; writes the
    bf16 shadow `p_lp` in the same launch when given."""
    for t_, nm in ((p, 'p'), (g, 'g'), (m, 'm'), (v, 'v')):
        require(t_, torch.float32, nm)
    require(step_t, torch.int64, 'step')
    check(lib().mt_adam_step_dev(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(lr_t), lr, betas[0], betas[1], eps, weight_decay,
                                 ptr(step_t), ptr(p_lp), stream()))


def adam_step_flat(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    for t_, nm in ((p, 'p'), (g, 'g'), (m, 'm'), (v, 'v')):
        require(t_, torch.float32, nm)
    check(lib().mt_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, betas[0], betas[1], eps, weight_decay, int(step), stream()))


# ---------------------------------------------------------------------------------------------------------
class LstmHeadFn(torch.autograd.Function):
    """Step-wise LSTM decoder with output feedback + MLP head (SFT/multiTransformer.py:465-483) in one persistent
    kernel each way."""

    @staticmethod
    def forward(ctx, enc, mask, arena, E, Hd, *params):
        dt = _state['dtype']
        B, T, _ = enc.shape
        enc = require(enc if enc.is_contiguous() else enc.contiguous(), _adt(), 'decoder input')
        _lib.check_device(enc.device.index)
        flat = arena.bind()
        lp = arena.shadow() if dt == MT_BF16 else None
        m = mask2d(mask, B, T, enc.device)
        need_grad = _need_grad(ctx)
        cfg = _lib.MtLstmHeadCfg(B, T, E, Hd, dt, int(need_grad))
        L = lib()
        nws = L.mt_lstm_head_ws_bytes(ctypes.byref(cfg))
        if nws == 0 or L.mt_lstm_head_param_count(ctypes.byref(cfg)) != arena.total:
            raise RuntimeError('unsupported LSTM decoder configuration')
        ws = _ws(nws, enc.device)
        out = torch.empty((B, T), dtype=torch.float32, device=enc.device)
        check(L.mt_lstm_head_fwd(ctypes.byref(cfg), ptr(flat), ptr(lp), ptr(enc), ptr(m), ptr(out), ptr(ws), ws.numel(), stream()))
        if need_grad:
            ctx.save_for_backward(enc, m, ws, flat, lp)
            ctx.cfg, ctx.arena = cfg, arena
        return out.unsqueeze(-1)

    @staticmethod
    def backward(ctx, dout):
        enc, m, ws, flat, lp = ctx.saved_tensors
        cfg = ctx.cfg
        dout = dout.reshape(cfg.B, cfg.T)
        dout = dout if dout.is_contiguous() else dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        denc = torch.empty_like(enc) if ctx.needs_input_grad[0] else None
        g = torch.empty(ctx.arena.total, dtype=torch.float32, device=enc.device)
        check(lib().mt_lstm_head_bwd(ctypes.byref(cfg), ptr(flat), ptr(lp), ptr(enc), ptr(m), ptr(dout), ptr(denc), ptr(g), ptr(ws),
                                     ws.numel(), stream()))
        return (denc, None, None, None, None, *ctx.arena.grad_views(g))


def lstm_head(enc, mask, arena, E, Hd):
    return _apply(LstmHeadFn, enc, mask, arena, E, Hd, *arena.params)


# ---------------------------------------------------------------------------------------------------------
class WindowCnnFn(torch.autograd.Function):
    """CNN (Conv1d over the K vectors of a window + global max-pool, MFT/models.py:57-79) and / or Highway + dropout
    (MFT/models.py:27-55,129) for all windows of a batch: one mt_window_cnn_fwd / _bwd call each way.
    stages: 1 = CNN only (x [..., K, D] -> [..., E]), 2 = Highway + dropout only (x [..., E]), 3 = both."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, wproj, bproj, wgate, bgate, stages, p_drop, site):
        dt = _state['dtype']
        if x.dtype != torch.float32:
            raise RuntimeError(f'window front-end input must be float32 (got {x.dtype})')
        x = require(x if x.is_contiguous() else x.contiguous(), torch.float32, 'window input')
        _lib.check_device(x.device.index)
        for nm, w in (('conv1d.weight', conv_w), ('conv1d.bias', conv_b), ('linear_projection.weight', wproj),
                      ('linear_projection.bias', bproj), ('linear_gate.weight', wgate), ('linear_gate.bias', bgate)):
            if w is not None:
                require(w, torch.float32, nm)
        if stages & 1:
            E, D, k = conv_w.shape
            K = x.shape[-2]
            if x.shape[-1] != D:
                raise RuntimeError(f'window vectors are {x.shape[-1]} wide, conv1d expects {D}')
            if K < k:
                raise RuntimeError(f'a window of {K} vectors is shorter than the conv kernel ({k})')
            lead = x.shape[:-2]
        else:
            E = wproj.shape[0]
            K = D = k = 0
            if x.shape[-1] != E:
                raise RuntimeError(f'Highway input is {x.shape[-1]} wide, expected {E}')
            lead = x.shape[:-1]
        n_win = 1
        for s in lead:
            n_win *= int(s)
        need_grad = _need_grad(ctx)
        seed = next_seed() if p_drop > 0 else 0
        cfg = _lib.MtWindowCnnCfg(dt, n_win, K, D, E, k, stages, int(need_grad), float(p_drop), seed, site)
        L = lib()
        nws = L.mt_window_cnn_ws_bytes(ctypes.byref(cfg))
        if nws == 0:
            raise RuntimeError('unsupported window front-end configuration')
        ws = _ws(nws, x.device)
        out = torch.empty((*lead, E), dtype=torch.float32, device=x.device)
        check(L.mt_window_cnn_fwd(ctypes.byref(cfg), ptr(x), ptr(conv_w), ptr(conv_b), ptr(wproj), ptr(bproj), ptr(wgate), ptr(bgate),
                                  ptr(out), ptr(ws), ws.numel(), stream()))
        if need_grad:
            ctx.save_for_backward(x, ws)
            ctx.cfg = cfg
            ctx.shapes = tuple(None if w is None else w.shape for w in (conv_w, conv_b, wproj, bproj, wgate, bgate))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, ws = ctx.saved_tensors
        cfg = ctx.cfg
        dout = dout if dout.is_contiguous() else dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        dev = x.device
        grads = [None if s is None else torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        dx = torch.empty_like(x) if (cfg.stages == 2 and ctx.needs_input_grad[0]) else None
        check(lib().mt_window_cnn_bwd(ctypes.byref(cfg), ptr(x), ptr(dout), ptr(dx), *[ptr(g) for g in grads], ptr(ws), ws.numel(),
                                      stream()))
        return (dx, *grads, None, None, None)


def window_cnn(x, conv_w, conv_b, wproj, bproj, wgate, bgate, p_drop=0.0, site=0x6000):
    """CNN + Highway + dropout of one modality: [..., K, D] -> [..., E]."""
    return _apply(WindowCnnFn, x, conv_w, conv_b, wproj, bproj, wgate, bgate, 3, float(p_drop), int(site))


def conv_maxpool(x, conv_w, conv_b):
    return _apply(WindowCnnFn, x, conv_w, conv_b, None, None, None, None, 1, 0.0, 0)


def highway(x, wproj, bproj, wgate, bgate, p_drop=0.0, site=0x6000):
    return _apply(WindowCnnFn, x, None, None, wproj, bproj, wgate, bgate, 2, float(p_drop), int(site))


def ccc_batched(pred, target, lengths):
    """Per-narrative CCC, Pearson r and the summed squared error of a padded batch in one launch (eval_ccc MFT/train.py:42-50;
    the reference evaluates one narrative per forward).  pred / target [B,T] or [B,T,1] fp32 CUDA; lengths list or int tensor.
    Returns (ccc [B] float64, pearson [B] float64, sq_err [] float64) on the device."""
    B, T = pred.shape[0], pred.shape[1]
    p = require(pred.reshape(B, T).contiguous(), torch.float32, 'pred')
    t = require(target.reshape(B, T).contiguous(), torch.float32, 'target')
    if torch.is_tensor(lengths):
        ln = lengths.to(device=p.device, dtype=torch.int32).contiguous()
    else:
        ln = torch.tensor([int(v) for v in lengths], dtype=torch.int32).to(p.device)
    if ln.numel() != B:
        raise RuntimeError('lengths must have one entry per narrative')
    _lib.check_device(p.device.index)
    ccc = torch.empty(B, dtype=torch.float64, device=p.device)
    pr = torch.empty(B, dtype=torch.float64, device=p.device)
    se = torch.empty((), dtype=torch.float64, device=p.device)
    check(lib().mt_ccc_batched(ptr(p), ptr(t), ptr(ln), B, T, ptr(ccc), ptr(pr), ptr(se), stream()))
    return ccc, pr, se
