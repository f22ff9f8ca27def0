"""GPU-side batcher.  The reference builds every batch on the host with `torch.tensor(nested python lists)`
(generateInputChunkHelper, MFT/train.py:59-68 -- seconds per batch at B = 256) and moves it to the GPU afterwards (:120-125).
Here the padded corpus (the output of padInput / padRating, MFT/train.py:456-514) is packed ONCE into device tensors and a batch is an
index gather on the device (mt_batch_gather) plus a length mask (mt_length_mask).

Wire format and ordering are the reference's: chunks of `batch_size` consecutive entries of the (shuffled unless onEval) index list;
inside a chunk narratives are sorted by length, longest first, ties in chunk order (list.sort is stable); everything trimmed to the
chunk's longest narrative; yields `(data: dict mod -> [B,T,K,D], target [B,T,1], mask [B,T,1] float, lengths: list)` (MFT/train.py:74-108).
`random.shuffle` is the reference's shuffler, so a seeded run visits the same batches.
"""
from random import shuffle

import numpy as np
import torch

from ._lib import check, lib, ptr, stream

__all__ = ['DeviceCorpus', 'generateTrainBatch']


class DeviceCorpus:
    """The padded training / evaluation set, resident in HBM.

    input_data: dict mod -> [N][T_max][K][D] (nested lists or an array, as padInput returns); input_target: [N][T_max];
    input_length: N ints.  The SEND corpus is a few hundred narratives, so even the raw-window tensors fit easily in 180 GB."""

    def __init__(self, input_data, input_target, input_length, device='cuda:0'):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('DeviceCorpus needs a CUDA device (there is no CPU fallback; the reference batcher is the CPU path)')
        self.lengths = [int(v) for v in input_length]
        self.N = len(self.lengths)
        self.data, self.row_shape = {}, {}
        for mod, v in input_data.items():
            a = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float32)))
            if a.shape[0] != self.N:
                raise RuntimeError(f'{mod}: {a.shape[0]} narratives, {self.N} lengths')
            self.data[mod] = a.to(self.device, torch.float32).contiguous()
            self.row_shape[mod] = tuple(a.shape[1:])
            if a.shape[1] < max(self.lengths):
                raise RuntimeError(f'{mod}: padded to {a.shape[1]} windows but the longest narrative has {max(self.lengths)}')
        tg = input_target if torch.is_tensor(input_target) else torch.from_numpy(np.asarray(input_target, dtype=np.float32))
        self.target = tg.to(self.device, torch.float32).reshape(self.N, -1).contiguous()
        self.T_max = self.target.shape[1]

    def __len__(self):
        return self.N

    def length_mask(self):
        """[N, T_max, 1] float mask of the whole corpus (1 for the valid windows of every narrative)."""
        ln = torch.tensor(self.lengths, dtype=torch.int32).to(self.device)
        mask = torch.empty((self.N, self.T_max, 1), dtype=torch.float32, device=self.device)
        check(lib().mt_length_mask(ptr(ln), self.N, self.T_max, ptr(mask), stream()))
        return mask

    def _alloc(self, B, per_win, T, t_cap):
        return torch.empty(B * t_cap * per_win, dtype=torch.float32, device=self.device)[:B * T * per_win]

    def batch(self, chunk):
        """One batch from a list of corpus indices (the reference's `chunk`)."""
        order = sorted(range(len(chunk)), key=lambda i: -self.lengths[chunk[i]])          # stable: ties keep chunk order
        rows = [chunk[i] for i in order]
        lengths = [self.lengths[r] for r in rows]
        T, B = lengths[0], len(rows)
        meta = torch.tensor(rows + lengths, dtype=torch.int32).to(self.device, non_blocking=True)      # one small H2D per batch
        idx, ln = meta[:B], meta[B:]
        L = lib()
        data = {}
        for mod, src in self.data.items():
            per_win = int(np.prod(self.row_shape[mod][1:])) if len(self.row_shape[mod]) > 1 else 1
            # always allocate for T_max and hand out the contiguous prefix: constant block sizes hit the caching allocator every time
            # (a fresh cudaMalloc of a few hundred MB per batch costs milliseconds, more than the gather itself)
            t_mod = self.row_shape[mod][0]                      # this modality's padded window count (>= every length)
            out = self._alloc(B, per_win, T, t_mod).view((B, T) + self.row_shape[mod][1:])
            check(L.mt_batch_gather(ptr(src), t_mod * per_win, ptr(idx), B, T * per_win, ptr(out), stream()))
            data[mod] = out
        target = self._alloc(B, 1, T, self.T_max).view(B, T, 1)
        check(L.mt_batch_gather(ptr(self.target), self.T_max, ptr(idx), B, T, ptr(target), stream()))
        mask = self._alloc(B, 1, T, self.T_max).view(B, T, 1)
        check(L.mt_length_mask(ptr(ln), B, T, ptr(mask), stream()))
        return data, target, mask, lengths

    def batch_into(self, chunk, x, target, mask):
        """Gather one batch straight into preallocated [B, T, ...] tensors (e.g. the static inputs of a captured train step,
        GraphedTrainStep.x / .target / .mask) -- no intermediate batch tensors, no second copy.  Every narrative is taken up to the
        buffers' T windows (the corpus is zero-padded beyond a narrative's length, so the tail arrives as zeros).  Returns lengths."""
        order = sorted(range(len(chunk)), key=lambda i: -self.lengths[chunk[i]])
        rows = [chunk[i] for i in order]
        lengths = [self.lengths[r] for r in rows]
        B, T = target.shape[0], target.shape[1]
        if len(rows) != B or lengths[0] > T or T > self.T_max:
            raise RuntimeError(f'batch_into: {len(rows)} narratives up to {lengths[0]} windows do not fit buffers of [{B}, {T}]')
        meta = torch.tensor(rows + lengths, dtype=torch.int32).to(self.device, non_blocking=True)
        idx, ln = meta[:B], meta[B:]
        L = lib()
        for mod, src in self.data.items():
            per_win = int(np.prod(self.row_shape[mod][1:])) if len(self.row_shape[mod]) > 1 else 1
            dst = x[mod]
            if tuple(dst.shape) != (B, T) + self.row_shape[mod][1:] or not dst.is_contiguous() or dst.dtype != torch.float32:
                raise RuntimeError(f'batch_into: buffer for {mod} must be contiguous float32 {(B, T) + self.row_shape[mod][1:]}')
            check(L.mt_batch_gather(ptr(src), self.row_shape[mod][0] * per_win, ptr(idx), B, T * per_win, ptr(dst), stream()))
        check(L.mt_batch_gather(ptr(self.target), self.T_max, ptr(idx), B, T, ptr(target), stream()))
        check(L.mt_length_mask(ptr(ln), B, T, ptr(mask), stream()))
        return lengths

    def generateTrainBatch(self, batch_size=25, onEval=False):
        index = list(range(self.N))
        if not onEval:
            shuffle(index)
        for i in range(0, self.N, batch_size):
            yield self.batch(index[i:i + batch_size])


def generateTrainBatch(input_data, input_target, input_length, args=None, batch_size=25, onEval=False):
    """Drop-in for generateTrainBatch (MFT/train.py:74-108): same arguments, same batches, already on the device.  Pass a DeviceCorpus
    as `input_data` (then input_target / input_length are ignored) to pack the corpus once instead of once per epoch."""
    corpus = input_data if isinstance(input_data, DeviceCorpus) else DeviceCorpus(
        input_data, input_target, input_length, getattr(args, 'device', 'cuda:0') if args is not None else 'cuda:0')
    return corpus.generateTrainBatch(batch_size=batch_size, onEval=onEval)
