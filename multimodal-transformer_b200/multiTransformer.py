"""Drop-in replacements for the model classes of frankaging/Multimodal-Transformer's `multiTransformer.py`
(MFT/, SFT/, B2-Trans/, B3-MFN/ copies), same class names, constructor and forward() signatures and state_dict
keys -- so `train.py` / `Performance-Eval` can `from multiTransformer import *` unchanged -- but every forward
and backward runs hand-written sm_100a kernels through libmt_b200.so.  There is no PyTorch / CPU fallback:
CPU tensors raise.  Reference lines are cited per class (paths relative to /root/reference/transformer/).
"""
import copy
import math

import torch
import torch.nn as nn

from . import functional as K
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH

__all__ = ['PositionwiseFeedForward', 'attention', 'MultiHeadedAttention', 'Encoder', 'clones', 'LayerNorm',
           'SublayerConnection', 'EncoderLayer', 'MFN', 'MultiTransformer', 'B3MultiTransformer', 'UniTransformer',
           'UniFullTransformer', 'NLPTransformer']


def _pick_device(device):
    # same convention as the reference constructors (MFT/multiTransformer.py:176-178): fall back to CPU for
    # construction / state_dict handling when no GPU is visible; forward() on CPU tensors raises.
    return device if torch.cuda.is_available() else torch.device('cpu')


def clones(module, N):
    """MFT/multiTransformer.py:78-79."""
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class LayerNorm(nn.Module):
    """MFT/multiTransformer.py:81-91 -- unbiased std, eps added to std; parameters a_2 / b_2."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(features))
        self.b_2 = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        return K.layer_norm(x.float() if x.dtype != torch.float32 else x, self.a_2, self.b_2, self.eps)


class PositionwiseFeedForward(nn.Module):
    """MFT/multiTransformer.py:9-20: w_2(dropout(relu(w_1(x))))."""

    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        hid = K.linear(x, self.w_1.weight, self.w_1.bias, act=ACT_RELU)
        p = self.dropout.p if self.training else 0.0
        return K.linear(hid, self.w_2.weight, self.w_2.bias, in_drop_p=p, out_f32=True)


def attention(query, key, value, mask=None, dropout=None):
    """MFT/multiTransformer.py:22-34.  query/key/value [B,h,T,dk]; mask broadcastable [B,1,T,1] (query-row mask).
    Returns (out [B,h,T,dk], p_attn [B,h,T,T]) like the reference; p_attn is materialised on demand only here."""
    B, h, T, dk = query.shape
    if key.shape != query.shape or value.shape != query.shape:
        raise RuntimeError('attention(): only self-attention shapes (q, k, v of equal shape) are supported on this path')
    packed = torch.cat([t.transpose(1, 2).reshape(B, T, h * dk) for t in (query, key, value)], dim=-1)
    adt = torch.float32 if K.get_compute_dtype() == 'fp32' else torch.bfloat16
    packed = packed.to(adt).contiguous()
    m = None if mask is None else mask.reshape(B, T)
    p = 0.0
    if dropout is not None and getattr(dropout, 'training', False):
        p = float(dropout.p)
    out = K.attention_packed(packed, m, h, p)
    p_attn = K.attention_probs(packed.detach(), m, h)
    return out.view(B, T, h, dk).transpose(1, 2).to(query.dtype), p_attn


class MultiHeadedAttention(nn.Module):
    """MFT/multiTransformer.py:36-65.  linears[0..2] = Q, K, V projections, linears[3] = output projection."""

    def __init__(self, h, d_model, dropout=0.1):
        super().__init__()
        assert d_model % h == 0
        self.d_k = d_model // h
        self.h = h
        self.linears = clones(nn.Linear(d_model, d_model), 4)
        self.dropout = nn.Dropout(p=dropout)
        self._last = None

    @property
    def attn(self):
        """p_attn of the last stand-alone forward ([B,h,T,T]); the reference stores it eagerly (:59), here it is
        recomputed from the saved projections only when somebody reads it."""
        if self._last is None:
            return None
        qkv, mask = self._last
        return K.attention_probs(qkv, mask, self.h)

    def forward(self, query, key, value, mask=None):
        if key.shape != query.shape or value.shape != query.shape:
            raise RuntimeError('MultiHeadedAttention: query/key/value must have equal shapes on this path')
        B, T, _ = query.shape
        q, k, v = [K.linear(x, l.weight, l.bias) for l, x in zip(self.linears, (query, key, value))]
        qkv = torch.cat([q, k, v], dim=-1)
        m = None if mask is None else mask.reshape(B, T)
        self._last = (qkv.detach(), m)
        p = self.dropout.p if self.training else 0.0
        a = K.attention_packed(qkv, m, self.h, p)
        return K.linear(a, self.linears[3].weight, self.linears[3].bias, out_f32=True)


class SublayerConnection(nn.Module):
    """MFT/multiTransformer.py:93-104: x + dropout(sublayer(norm(x)))."""

    def __init__(self, size, dropout):
        super().__init__()
        self.norm = LayerNorm(size)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, sublayer):
        y = sublayer(self.norm(x))
        if y.dtype != torch.float32:
            y = y.float()
        return K.residual_dropout(x, y, self.dropout.p if self.training else 0.0)


class EncoderLayer(nn.Module):
    """MFT/multiTransformer.py:106-116 (stand-alone path; inside an Encoder the whole stack is one fused call)."""

    def __init__(self, size, self_attn, feed_forward, dropout):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.sublayer = clones(SublayerConnection(size, dropout), 2)
        self.size = size

    def forward(self, x, mask):
        x = self.sublayer[0](x, lambda t: self.self_attn(t, t, t, mask))
        return self.sublayer[1](x, self.feed_forward)


def _check_encoder_shape(d_model, h, what):
    """The sm_100a encoder kernels cover d_model % 128 == 0 (<= 1024) with head widths 16 / 32 / 64 -- every configuration the reference
    instantiates except the 16-wide 'emotient' modality (EMBED 16, h = 8 -> d_k = 2, MFT/multiTransformer.py:260).  There is no CPU /
    PyTorch fallback on this path, so an unsupported shape fails HERE, at construction, not at the first forward."""
    dk = d_model // h if h and d_model % h == 0 else 0
    if d_model % 128 != 0 or d_model > 1024 or dk not in (16, 32, 64):
        raise NotImplementedError(
            f'{what}: d_model={d_model}, h={h} (d_k={dk}) is outside the B200 encoder kernels (d_model % 128 == 0, d_model <= 1024, '
            f'd_k in 16/32/64); the reference\'s "emotient" modality (d_model 16) is the one reference configuration not covered -- '
            f'see INTEGRATION.md, "Unsupported reference configurations"')


class Encoder(nn.Module):
    """MFT/multiTransformer.py:67-76: N cloned layers + final LayerNorm, executed by mt_encoder_fwd / mt_encoder_bwd
    on a flat parameter arena (the nn.Parameters below are views into it, so state_dict / optimizers see the
    reference's tensors)."""

    def __init__(self, layer, N):
        super().__init__()
        if isinstance(layer, EncoderLayer) and isinstance(layer.self_attn, MultiHeadedAttention):
            _check_encoder_shape(layer.size, layer.self_attn.h, 'Encoder')
        self.layers = clones(layer, N)
        self.norm = LayerNorm(layer.size)
        self.stack_id = 0          # selects the dropout streams of this stack
        self.out_fp32 = None       # None: follow the compute dtype
        self._arena = None

    def _canonical_params(self):
        ps = []
        for l in self.layers:
            a, f = l.self_attn, l.feed_forward
            ps += [a.linears[0].weight, a.linears[1].weight, a.linears[2].weight,
                   a.linears[0].bias, a.linears[1].bias, a.linears[2].bias,
                   a.linears[3].weight, a.linears[3].bias, f.w_1.weight, f.w_1.bias, f.w_2.weight, f.w_2.bias,
                   l.sublayer[0].norm.a_2, l.sublayer[0].norm.b_2, l.sublayer[1].norm.a_2, l.sublayer[1].norm.b_2]
        return ps + [self.norm.a_2, self.norm.b_2]

    def _fusable(self):
        l0 = self.layers[0]
        ok = all(isinstance(l, EncoderLayer) and isinstance(l.self_attn, MultiHeadedAttention)
                 and isinstance(l.feed_forward, PositionwiseFeedForward) and l.size == l0.size
                 and l.self_attn.h == l0.self_attn.h and l.feed_forward.w_1.out_features == l0.feed_forward.w_1.out_features
                 for l in self.layers)
        ps = {l.sublayer[0].dropout.p for l in self.layers} | {l.sublayer[1].dropout.p for l in self.layers} | \
             {l.self_attn.dropout.p for l in self.layers} | {l.feed_forward.dropout.p for l in self.layers}
        return ok and len(ps) == 1

    def arena(self):
        if self._arena is None:
            self._arena = K.Arena(self._canonical_params())
        return self._arena

    def forward(self, x, mask):
        if not self._fusable():
            for layer in self.layers:
                x = layer(x, mask)
            return self.norm(x)
        l0 = self.layers[0]
        cfgd = dict(h=l0.self_attn.h, dff=l0.feed_forward.w_1.out_features, n_layers=len(self.layers),
                    p_drop=l0.sublayer[0].dropout.p if self.training else 0.0, stack_id=self.stack_id)
        if self.out_fp32 is not None:
            cfgd['y_f32'] = bool(self.out_fp32)
        if getattr(self, 'grid_share', 0) > 1:      # set by MultiTransformer while its stacks run on concurrent streams
            cfgd['grid_share'] = int(self.grid_share)
        if x.dtype != torch.float32:
            x = x.float()
        return K.encoder_stack(x, mask, self.arena(), cfgd)


class MFN(nn.Module):
    """Memory Fusion Network, MFT/multiTransformer.py:118-248.  forward(inputs: dict mod -> [T,B,D]) -> [B,T,out]."""

    HIDDEN = {'linguistic': 88, 'emotient': 16, 'acoustic': 48, 'image': 88}          # :128

    def __init__(self, mods, dims, output_dim, device=torch.device('cuda:0')):
        super().__init__()
        if output_dim != 1:
            raise ValueError('the B200 MFN head is built for output_dim == 1 (the only value the reference uses)')
        self.mods = mods
        self.dims = dims
        self.hidden_dim = dict(MFN.HIDDEN)
        total_h = sum(self.hidden_dim[m] for m in mods)
        self.mem_dim = 128
        att_in = 2 * total_h
        gamma_in = att_in + self.mem_dim
        self.lstm = dict()
        for mod in mods:
            self.lstm[mod] = nn.LSTMCell(dims[mod], self.hidden_dim[mod])
            self.add_module('lstm_{}'.format(mod), self.lstm[mod])
        sizes = dict(att1=(128, 0.0), att2=(256, 0.0), gamma1=(64, 0.2), gamma2=(64, 0.2), out=(64, 0.5))      # :138-147
        self.att1_fc1 = nn.Linear(att_in, sizes['att1'][0]); self.att1_fc2 = nn.Linear(sizes['att1'][0], att_in)
        self.att1_dropout = nn.Dropout(sizes['att1'][1])
        self.att2_fc1 = nn.Linear(att_in, sizes['att2'][0]); self.att2_fc2 = nn.Linear(sizes['att2'][0], self.mem_dim)
        self.att2_dropout = nn.Dropout(sizes['att2'][1])
        self.gamma1_fc1 = nn.Linear(gamma_in, sizes['gamma1'][0]); self.gamma1_fc2 = nn.Linear(sizes['gamma1'][0], self.mem_dim)
        self.gamma1_dropout = nn.Dropout(sizes['gamma1'][1])
        self.gamma2_fc1 = nn.Linear(gamma_in, sizes['gamma2'][0]); self.gamma2_fc2 = nn.Linear(sizes['gamma2'][0], self.mem_dim)
        self.gamma2_dropout = nn.Dropout(sizes['gamma2'][1])
        self.out_fc1 = nn.Linear(total_h + self.mem_dim, sizes['out'][0]); self.out_fc2 = nn.Linear(sizes['out'][0], output_dim)
        self.out_dropout = nn.Dropout(sizes['out'][1])
        self._arena = None
        self.device = _pick_device(device)
        self.to(self.device)

    def arena(self):
        if self._arena is None:
            ps = []
            for mod in self.mods:
                c = self.lstm[mod]
                ps += [c.weight_ih, c.weight_hh, c.bias_ih, c.bias_hh]
            for l in (self.att1_fc1, self.att1_fc2, self.att2_fc1, self.att2_fc2, self.gamma1_fc1, self.gamma1_fc2,
                      self.gamma2_fc1, self.gamma2_fc2, self.out_fc1, self.out_fc2):
                ps += [l.weight, l.bias]
            self._arena = K.Arena(ps)
        return self._arena

    def _run(self, xs, mask, t_major):
        if self.att1_dropout.p != 0 or self.att2_dropout.p != 0 or self.gamma1_dropout.p != self.gamma2_dropout.p:
            raise RuntimeError('MFN kernel supports the reference dropout layout only (att1/att2 = 0, gamma1 == gamma2)')
        tr = self.training
        cfgd = dict(hid=[self.hidden_dim[m] for m in self.mods], mem=self.mem_dim, a1=self.att1_fc1.out_features,
                    a2=self.att2_fc1.out_features, g=self.gamma1_fc1.out_features, o=self.out_fc1.out_features,
                    p_gamma=self.gamma1_dropout.p if tr else 0.0, p_out=self.out_dropout.p if tr else 0.0)
        adt = torch.float32 if K.get_compute_dtype() == 'fp32' else torch.bfloat16
        xs = [x if x.dtype == adt else x.to(adt) for x in xs]
        out, h_last, c_last, mem_last = K.mfn_forward(xs, mask, self.arena(), cfgd, t_major)
        # side-effect attributes of the reference module (:187-198, :224-229)
        hs = [self.hidden_dim[m] for m in self.mods]
        self.h = dict(zip(self.mods, torch.split(h_last, hs, dim=1)))
        self.c = dict(zip(self.mods, torch.split(c_last, hs, dim=1)))
        self.mem = mem_last
        return out

    def forward(self, inputs):
        return self._run([inputs[m] for m in self.mods], None, t_major=True)


class MultiTransformer(nn.Module):
    """MFT model body, MFT/multiTransformer.py:250-313: per-modality Linear embed -> 6-layer encoder -> MFN -> mask.
    `use_encoder=False` gives the B3-MFN variant (B3-MFN/multiTransformer.py:250-307: no encoder modules at all)."""

    EMBED = {'linguistic': 256, 'emotient': 16, 'acoustic': 256, 'image': 256}       # :260

    def __init__(self, mods, window_embed_size, N=6, d_ff=128, h=8, dropout=0.1, n_layers=1,
                 device=torch.device('cuda:0'), use_encoder=True):
        super().__init__()
        self.mods = mods
        self.window_embed_size = window_embed_size
        self.embed_dim = dict(type(self).EMBED)      # subclasses may scale d_model (bench.py --config c5)
        self.use_encoder = use_encoder
        self.embed, self.transformer, self.lstm, self.attn, self.ff = dict(), dict(), dict(), dict(), dict()
        for i, mod in enumerate(mods):
            self.embed[mod] = nn.Linear(window_embed_size[mod], self.embed_dim[mod])
            self.add_module('embed_{}'.format(mod), self.embed[mod])
            if use_encoder:
                # the reference registers the template attention / FFN modules too (:273-276, names without underscore);
                # they are never used in forward but live in every checkpoint, so they are kept for state_dict parity
                self.attn[mod] = MultiHeadedAttention(h, self.embed_dim[mod])
                self.ff[mod] = PositionwiseFeedForward(self.embed_dim[mod], d_ff, dropout)
                self.add_module('attn{}'.format(mod), self.attn[mod])
                self.add_module('ff{}'.format(mod), self.ff[mod])
                enc = Encoder(EncoderLayer(self.embed_dim[mod], copy.deepcopy(self.attn[mod]), copy.deepcopy(self.ff[mod]),
                                           dropout), N)
                enc.stack_id = i
                self.transformer[mod] = enc
                self.add_module('transformer_{}'.format(mod), enc)
        self.mfn = MFN(mods, self.embed_dim, 1)
        self.device = _pick_device(device)
        self.to(self.device)

    def group_arena(self):
        """One arena over the parameters of ALL modality stacks (stack-major), used by the grouped call; None when the stacks differ."""
        if getattr(self, '_group_arena', None) is None:
            encs = [self.transformer[m] for m in self.mods]
            per = [e._canonical_params() for e in encs]
            n = [sum(p.numel() for p in ps) for ps in per]
            if len(set(n)) != 1 or n[0] % 8 != 0:
                return None
            self._group_arena = K.Arena([p for ps in per for p in ps])
        return self._group_arena

    def _groupable(self, inputs):
        if not (K.grouped_stacks() and self.use_encoder and 2 <= len(self.mods) <= 4 and inputs[self.mods[0]].is_cuda):
            return False
        encs = [self.transformer[m] for m in self.mods]
        if not all(e._fusable() for e in encs):
            return False
        sig = {(e.layers[0].size, e.layers[0].self_attn.h, e.layers[0].feed_forward.w_1.out_features, len(e.layers),
                e.layers[0].sublayer[0].dropout.p, e.out_fp32, e.training) for e in encs}
        shapes = {tuple(inputs[m].shape[:2]) for m in self.mods}
        return len(sig) == 1 and len(shapes) == 1 and self.group_arena() is not None

    def _stack(self, mod, inputs, mask):
        e = self.embed[mod]
        x = K.linear(inputs[mod], e.weight, e.bias, out_f32=self.use_encoder)
        if self.use_encoder:
            x = self.transformer[mod](x, mask)
        return x

    def forward(self, inputs, mask, lengths, tgt_init=0.5, target=None):
        if self._groupable(inputs):
            # the stacks have the same shape: ONE grouped call (embeds + encoders), one launch per projection / LayerNorm / weight
            # gradient for all modalities, instead of one call per modality
            e0 = self.transformer[self.mods[0]]
            l0 = e0.layers[0]
            cfgd = dict(h=l0.self_attn.h, dff=l0.feed_forward.w_1.out_features, n_layers=len(e0.layers),
                        p_drop=l0.sublayer[0].dropout.p if e0.training else 0.0,
                        stack_ids=[self.transformer[m].stack_id for m in self.mods])
            if e0.out_fp32 is not None:
                cfgd['y_f32'] = bool(e0.out_fp32)
            xs = K.encoder_stack_group([inputs[m] for m in self.mods], mask, [self.embed[m] for m in self.mods], self.group_arena(), cfgd)
            return self.mfn._run(list(xs), mask, t_major=False)
        if K.parallel_stacks() and self.use_encoder and len(self.mods) > 1 and inputs[self.mods[0]].is_cuda:
            # the modality stacks are independent until the MFN: run them on side streams so the prologue / tail of one
            # stack's kernels overlaps the others' (autograd replays each stack's backward on the same stream); under CUDA
            # graph capture this becomes a fork / join in the graph
            cur = torch.cuda.current_stream()
            dev = inputs[self.mods[0]].device
            if getattr(self, '_side', None) is None or self._side[0].device != dev:
                self._side = [torch.cuda.Stream(device=dev) for _ in self.mods[1:]]
            xs = [None] * len(self.mods)
            for mod in self.mods:                     # concurrent stacks share the SMs: half-size grids co-reside (measured -2 %)
                self.transformer[mod].grid_share = 2
            for i, mod in enumerate(self.mods[1:], 1):
                s = self._side[i - 1]
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    xs[i] = self._stack(mod, inputs, mask)
            xs[0] = self._stack(self.mods[0], inputs, mask)
            for i, s in enumerate(self._side, 1):
                cur.wait_stream(s)
                xs[i].record_stream(cur)
        else:
            if self.use_encoder:
                for mod in self.mods:
                    self.transformer[mod].grid_share = 0
            xs = [self._stack(mod, inputs, mask) for mod in self.mods]
        # [B,T,D] row order goes straight into the recurrence (the reference permutes to [T,B,D] first, :300);
        # the output mask (:310) is applied by the kernel.
        return self.mfn._run(xs, mask, t_major=False)


class B3MultiTransformer(MultiTransformer):
    """B3-MFN/multiTransformer.py:250-307."""

    def __init__(self, mods, window_embed_size, N=6, d_ff=128, h=8, dropout=0.1, n_layers=1, device=torch.device('cuda:0')):
        super().__init__(mods, window_embed_size, N, d_ff, h, dropout, n_layers, device, use_encoder=False)


def _make_encoder(embed_dim, d_ff, h, dropout, N):
    attn = MultiHeadedAttention(h, embed_dim)
    ff = PositionwiseFeedForward(embed_dim, d_ff, dropout)
    return Encoder(EncoderLayer(embed_dim, copy.deepcopy(attn), copy.deepcopy(ff), dropout), N)


class UniFullTransformer(nn.Module):
    """B2-Trans body, MFT/multiTransformer.py:378-420: Linear embed -> encoder -> Linear/ReLU/Linear head -> mask."""

    def __init__(self, window_embed_size, embed_dim=256, h_dim=128, N=6, d_ff=128, h=8, dropout=0.1, n_layers=1,
                 device=torch.device('cuda:0')):
        super().__init__()
        self.embed_dim = embed_dim
        self.h_dim = h_dim
        self.embed = nn.Linear(window_embed_size, embed_dim)
        self.encoder = _make_encoder(embed_dim, d_ff, h, dropout, N)
        self.out = nn.Sequential(nn.Linear(embed_dim, h_dim), nn.ReLU(), nn.Linear(h_dim, 1))
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, mask, lengths, tgt_init=0.5, target=None):
        B, T, _ = inputs.shape
        x = K.linear(inputs, self.embed.weight, self.embed.bias, out_f32=True)
        enc = self.encoder(x, mask)
        hid = K.linear(enc, self.out[0].weight, self.out[0].bias, act=ACT_RELU)
        m = K.mask2d(mask, B, T, inputs.device)
        return K.linear(hid, self.out[2].weight, self.out[2].bias, rowmask=None if m is None else m.reshape(-1), out_f32=True)


class _LstmDecoderMixin:
    """Shared by UniTransformer / NLPTransformer: nn.LSTM(2E -> E) stepped with output feedback + MLP head
    (SFT/multiTransformer.py:465-483, MFT/multiTransformer.py:357-375)."""

    def _build_decoder(self, embed_dim, h_dim, n_layers):
        if n_layers != 1:
            raise ValueError('the B200 decoder kernel implements the single-layer LSTM the reference uses')
        self.decoder = nn.LSTM(2 * embed_dim, embed_dim, n_layers, batch_first=True)
        self.dec_h0 = nn.Parameter(torch.zeros(n_layers, 1, embed_dim))
        self.dec_c0 = nn.Parameter(torch.zeros(n_layers, 1, embed_dim))
        self.out = nn.Sequential(nn.Linear(embed_dim, h_dim), nn.ReLU(), nn.Linear(h_dim, 1))
        self._dec_arena = None

    def _decode(self, enc, mask):
        if self._dec_arena is None:
            d = self.decoder
            self._dec_arena = K.Arena([d.weight_ih_l0, d.weight_hh_l0, d.bias_ih_l0, d.bias_hh_l0, self.dec_h0, self.dec_c0,
                                       self.out[0].weight, self.out[0].bias, self.out[2].weight, self.out[2].bias])
        return K.lstm_head(enc, mask, self._dec_arena, self.embed_dim, self.h_dim)


class UniTransformer(nn.Module, _LstmDecoderMixin):
    """MFT/multiTransformer.py:315-376: Linear embed -> encoder -> step-wise LSTM decoder."""

    def __init__(self, window_embed_size, embed_dim=256, h_dim=128, N=6, d_ff=128, h=8, dropout=0.1, n_layers=1,
                 device=torch.device('cuda:0')):
        super().__init__()
        self.embed_dim = embed_dim
        self.h_dim = h_dim
        self.embed = nn.Linear(window_embed_size, embed_dim)
        self.encoder = _make_encoder(embed_dim, d_ff, h, dropout, N)
        self._build_decoder(embed_dim, h_dim, n_layers)
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, mask, lengths, tgt_init=0.5, target=None):
        x = K.linear(inputs, self.embed.weight, self.embed.bias, out_f32=True)
        return self._decode(self.encoder(x, mask), mask)


class NLPTransformer(nn.Module, _LstmDecoderMixin):
    """SFT model body, SFT/multiTransformer.py:422-484: Dropout(.1) -> Linear -> ReLU embed, encoder, LSTM decoder."""

    def __init__(self, window_embed_size, embed_dim=256, h_dim=128, N=6, d_ff=128, h=8, dropout=0.1, n_layers=1,
                 device=torch.device('cuda:0')):
        super().__init__()
        self.embed_dim = embed_dim
        self.h_dim = h_dim
        self.embed = nn.Sequential(nn.Dropout(0.1), nn.Linear(window_embed_size, embed_dim), nn.ReLU())
        self.encoder = _make_encoder(embed_dim, d_ff, h, dropout, N)
        self._build_decoder(embed_dim, h_dim, n_layers)
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, mask, lengths, tgt_init=0.5, target=None):
        p = self.embed[0].p if self.training else 0.0
        x = K.linear(inputs, self.embed[1].weight, self.embed[1].bias, act=ACT_RELU, in_drop_p=p, out_f32=True)
        return self._decode(self.encoder(x, mask), mask)


def fusion_layer(feats, weight, bias):
    """Early-fusion head of SFT (SFT/models.py:136-138): tanh(Linear(cat(mods, dim=2)))."""
    x = K.concat_features(list(feats))       # the library's strided cast kernel assembles the GEMM operand: no framework copy kernel
    return K.linear(x, weight, bias, act=ACT_TANH, out_f32=True)
