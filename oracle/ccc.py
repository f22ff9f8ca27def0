"""Concordance correlation coefficient, restating eval_ccc (MFT/train.py:42-50).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by the reference's only published
known-answers: PredSave/{MFT,SFT}{173_4,165_2}.csv -> PerfSave/{MFT,SFT}.csv (tests/golden/ccc_kat.json).
"""
import numpy as np


def eval_ccc(y_true, y_pred):
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    tm, pm = y_true.mean(), y_pred.mean()
    tv, pv = y_true.var(), y_pred.var()
    cov = ((y_true - tm) * (y_pred - pm)).mean()
    return 2 * cov / (tv + pv + (pm - tm) ** 2)
