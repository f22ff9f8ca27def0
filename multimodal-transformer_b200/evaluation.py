"""Batched evaluation: what `evaluate` / `evaluateOnEval` of MFT/train.py:157-257 compute -- per-narrative
CCC and Pearson r, their means / stds, the best narrative, and the summed-MSE loss per valid time-point -- with many
narratives per forward instead of the reference's batch_size = 1 (MFT/train.py:169,218), and the statistics reduced on the
device by one kernel per batch (mt_ccc_batched) instead of a host round trip per narrative.

Exactness: the reference forwards every narrative ALONE (no padded window exists).  A padded batch through the training-mode
semantics would differ, because padded windows are live attention keys there; `ragged_batch` removes them, so the numbers here
equal the one-at-a-time evaluation to round-off (tests/test_gpu_frontend.py::test_batched_evaluation_equals_one_at_a_time).
"""
import torch

from . import functional as K

__all__ = ['evaluate']


def _call(model, inputs, mask, lengths):
    from .models import _FrontEnd
    if isinstance(model, _FrontEnd):                      # models.py signature: forward(inputs, length, mask)
        return model(inputs, lengths, mask)
    return model(inputs, mask, lengths)                   # multiTransformer.py signature: forward(inputs, mask, lengths)


def evaluate(model, inputs, target=None, mask=None, lengths=None, batch_size=64):
    """`evaluate(model, corpus)` with a batching.DeviceCorpus, or explicit tensors --
    inputs: dict mod -> [N, T, ...] (or one tensor [N, T, F] for the single-input models), target / mask [N, T, 1], lengths: N ints
    (any order; every narrative is padded to the common T).  All tensors on the model's GPU.

    Returns (predictions, loss, stats, (best_output, best_target, best_index)) like evaluate() MFT/train.py:203-257:
      predictions  list of N float32 numpy arrays, the valid part of every prediction (as evaluateOnEval collects them, :181)
      loss         sum of squared errors over valid time-points / number of valid time-points               (:229-231,249)
      stats        {'corr', 'corr_std', 'ccc', 'ccc_std', 'max_ccc'}                                       (:251-252)
      best_*       prediction / target of the narrative with the highest CCC and its 1-based position       (:240-245)"""
    from .batching import DeviceCorpus
    if isinstance(inputs, DeviceCorpus):
        corpus = inputs
        inputs, lengths = corpus.data, corpus.lengths
        target = corpus.target.unsqueeze(-1)
        mask = corpus.length_mask()
    lengths = [int(v) for v in lengths]
    N = len(lengths)
    was_training = model.training
    model.eval()
    ccc_all, pr_all, preds = [], [], []
    se_total = torch.zeros((), dtype=torch.float64, device=target.device)
    try:
        with torch.no_grad():
            for i in range(0, N, batch_size):
                ln = lengths[i:i + batch_size]
                Tm = max(ln)
                sl = (lambda x: x[i:i + batch_size, :Tm].contiguous())
                xin = {m: sl(v) for m, v in inputs.items()} if isinstance(inputs, dict) else sl(inputs)
                with K.ragged_batch(ln):
                    out = _call(model, xin, sl(mask), ln)
                ccc, pr, se = K.ccc_batched(out, sl(target), ln)
                ccc_all.append(ccc); pr_all.append(pr); preds.append((out.reshape(len(ln), Tm), ln))
                se_total += se
    finally:
        model.train(was_training)
    ccc = torch.cat(ccc_all); pr = torch.cat(pr_all)
    best = int(torch.argmax(ccc).item())
    stats = {'corr': pr.mean().item(), 'corr_std': pr.std(unbiased=False).item(), 'ccc': ccc.mean().item(),
             'ccc_std': ccc.std(unbiased=False).item(), 'max_ccc': ccc[best].item()}
    predictions = []
    for out, ln in preds:
        o = out.float().cpu().numpy()
        predictions.extend(o[b, :l].copy() for b, l in enumerate(ln))
    loss = (se_total / float(sum(lengths))).item()
    best_target = target[best, :lengths[best]].reshape(-1).float().cpu().numpy()
    return predictions, loss, stats, (predictions[best], best_target, best + 1)
