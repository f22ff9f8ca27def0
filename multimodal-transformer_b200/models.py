"""Mirror of the reference's `models.py` classes that sit directly in front of the hot path: CNN, Highway and the four
MultiCNNTransformer variants (MFT/models.py:27-136, SFT/models.py:81-142, B2-Trans/models.py:81-133, B3-MFN/models.py:81-136).

Same constructor arguments, `forward(inputs, length, mask)`, submodule names and state_dict keys (`cnn_{mod}.conv1d.*`,
`highway_{mod}.linear_projection.* / linear_gate.*`, `fusionLayer.*`, `Transformer.*`), so a reference checkpoint loads with
`model.load_state_dict(torch.load(path)['model'])` (Performance-Eval/train.py:556-557).

The reference runs the window CNN once per narrative and modality in a python loop (MFT/models.py:117-132: B x mods
iterations); here all B*T windows of a modality go through ONE mt_window_cnn_fwd call (a tcgen05 GEMM over overlapping
rows of the raw [B,T,K,D] tensor, a segmented max, the stacked Highway GEMM and a fused gate/dropout kernel).
"""
import torch
import torch.nn as nn

from . import functional as K
from .multiTransformer import (B3MultiTransformer, MultiTransformer, NLPTransformer, UniFullTransformer, UniTransformer, _pick_device,
                               fusion_layer)

__all__ = ['CNN', 'Highway', 'MultiCNNTransformer', 'SFTMultiCNNTransformer', 'B2MultiCNNTransformer', 'B3MultiCNNTransformer']

_SITE_FRONT = 0x6000          # dropout sites 0x6000 + modality index (oracle/dropout_rng.py)


class Highway(nn.Module):
    """MFT/models.py:27-55: g = sigmoid(gate(x)); g * projection(x) + (1 - g) * x."""

    def __init__(self, word_embed_size):
        super().__init__()
        self.word_embed_size = word_embed_size
        self.linear_projection = nn.Linear(word_embed_size, word_embed_size, bias=True)
        self.linear_gate = nn.Linear(word_embed_size, word_embed_size, bias=True)

    def forward(self, x_conv_out):
        return K.highway(x_conv_out, self.linear_projection.weight, self.linear_projection.bias, self.linear_gate.weight,
                         self.linear_gate.bias)


class CNN(nn.Module):
    """MFT/models.py:57-79: Conv1d(word_embed_size -> window_embed_size, k) then a max over every conv position."""

    def __init__(self, word_embed_size=300, window_embed_size=128, k=2):
        super().__init__()
        self.k = k
        self.f = window_embed_size
        self.word_embed_size = word_embed_size
        self.window_embed_size = window_embed_size
        self.conv1d = nn.Conv1d(word_embed_size, window_embed_size, k, bias=True)

    def forward(self, x_reshape):
        """x_reshape [batch, word_embed_size, max_window_length] (the reference's channel-major view) -> [batch, window_embed_size]."""
        x = x_reshape.permute(0, 2, 1).contiguous()             # the kernel reads vectors as rows: [batch, K, D]
        return K.conv_maxpool(x, self.conv1d.weight, self.conv1d.bias)


class _FrontEnd(nn.Module):
    """Shared front half of every MultiCNNTransformer variant: per modality CNN -> Highway -> Dropout(0.3)."""

    def _build_front(self, mods, dims, window_embed_size, k):
        self.mods = mods
        self.dims = dims
        self.CNN = dict()
        self.Highway = dict()
        self.window_embed_size = window_embed_size
        total_embed_size = 0
        for mod in mods:
            self.CNN[mod] = CNN(dims[mod], self.window_embed_size[mod], k)
            self.Highway[mod] = Highway(self.window_embed_size[mod])
            self.add_module('cnn_{}'.format(mod), self.CNN[mod])
            self.add_module('highway_{}'.format(mod), self.Highway[mod])
            total_embed_size += self.window_embed_size[mod]
        self.dropout = nn.Dropout(p=0.3)
        return total_embed_size

    def _front(self, inputs):
        """dict mod -> [B, T, K, D] raw window vectors  ->  dict mod -> [B, T, E] window embeddings."""
        p = self.dropout.p if self.training else 0.0
        out = {}
        for i, mod in enumerate(self.mods):
            x = inputs[mod]
            if x.dim() != 4:
                raise RuntimeError(f'{mod}: expected [batch, windows, vectors, dim], got {tuple(x.shape)}')
            cnn, hw = self.CNN[mod], self.Highway[mod]
            out[mod] = K.window_cnn(x, cnn.conv1d.weight, cnn.conv1d.bias, hw.linear_projection.weight, hw.linear_projection.bias,
                                    hw.linear_gate.weight, hw.linear_gate.bias, p_drop=p, site=_SITE_FRONT + i)
        return out


class MultiCNNTransformer(_FrontEnd):
    """MFT/models.py:81-136: window CNNs -> MultiTransformer (MFN fusion); one modality -> UniTransformer."""

    def __init__(self, mods, dims, embed_dims, fuse_embed_size=256, k=2, device=torch.device('cuda:0')):
        super().__init__()
        total = self._build_front(mods, dims, embed_dims, k)
        if len(mods) > 1:
            self.Transformer = MultiTransformer(mods=mods, window_embed_size=self.window_embed_size)
        else:
            assert len(mods) == 1
            self.Transformer = UniTransformer(total)
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, length, mask=None):
        outputs = self._front(inputs)
        if len(outputs) > 1:
            return self.Transformer(outputs, mask, length)
        return self.Transformer(outputs[self.mods[0]], mask, length)


_FIXED_EMBED = {'linguistic': 300, 'emotient': 20, 'acoustic': 256, 'image': 256}       # SFT/B2/B3 models.py:90


class B3MultiCNNTransformer(_FrontEnd):
    """B3-MFN/models.py:81-136: as MFT but fixed window embedding sizes and the encoder-less MultiTransformer."""

    def __init__(self, mods, dims, fuse_embed_size=256, k=2, device=torch.device('cuda:0')):
        super().__init__()
        total = self._build_front(mods, dims, dict(_FIXED_EMBED), k)
        if len(mods) > 1:
            self.Transformer = B3MultiTransformer(mods=mods, window_embed_size=self.window_embed_size)
        else:
            assert len(mods) == 1
            self.Transformer = UniTransformer(total)
        self.device = _pick_device(device)
        self.to(self.device)

    forward = MultiCNNTransformer.forward


class SFTMultiCNNTransformer(_FrontEnd):
    """SFT/models.py:81-142: window CNNs -> cat -> tanh(fusionLayer) -> NLPTransformer."""

    def __init__(self, mods, dims, fuse_embed_size=512, k=2, device=torch.device('cuda:0')):
        super().__init__()
        total = self._build_front(mods, dims, dict(_FIXED_EMBED), k)
        self.fusionLayer = nn.Linear(total, fuse_embed_size)
        if len(mods) > 1:
            self.Transformer = NLPTransformer(fuse_embed_size)
        else:
            assert len(mods) == 1
            self.Transformer = UniTransformer(total)
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, length, mask=None):
        outputs = self._front(inputs)
        if len(outputs) > 1:
            fused = fusion_layer([outputs[m] for m in self.mods], self.fusionLayer.weight, self.fusionLayer.bias)
            return self.Transformer(fused, mask, length)
        return self.Transformer(outputs[self.mods[0]], mask, length)


class B2MultiCNNTransformer(_FrontEnd):
    """B2-Trans/models.py:81-133: window CNNs -> cat -> UniFullTransformer (no fusion layer)."""

    def __init__(self, mods, dims, fuse_embed_size=256, k=2, device=torch.device('cuda:0')):
        super().__init__()
        total = self._build_front(mods, dims, dict(_FIXED_EMBED), k)
        self.Transformer = UniFullTransformer(total)
        self.device = _pick_device(device)
        self.to(self.device)

    def forward(self, inputs, length, mask=None):
        outputs = self._front(inputs)
        if len(outputs) > 1:
            return self.Transformer(K.concat_features([outputs[m] for m in self.mods]), mask, length)
        return self.Transformer(outputs[self.mods[0]], mask, length)
