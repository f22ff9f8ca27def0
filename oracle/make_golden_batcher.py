"""Generate tests/golden/batcher.json by executing the reference's UNMODIFIED generateTrainBatch / generateInputChunkHelper / chunks
function bodies (MFT/train.py:52-108).  They are extracted with `ast` and compiled in a namespace holding exactly the names they use
(torch, shuffle, itemgetter): importing MFT/train.py itself would open a log file and pull in the dataset readers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.make_golden_batcher
"""
import ast
import hashlib
import json
import os
import random
from operator import itemgetter
from random import shuffle

import numpy as np
import torch

from .make_golden import OUT, REF


def corpus(seed=31, n=11, t_max=9, shapes=(('linguistic', 3, 5), ('image', 2, 4))):
    """A tiny padded corpus in the reference's format: nested python lists (padInput / padRating output)."""
    rs = np.random.RandomState(seed)
    lengths = [int(v) for v in rs.randint(2, t_max + 1, size=n)]
    lengths[3] = lengths[5] = t_max                                   # ties and a full-length narrative
    data = {}
    for mod, K, D in shapes:
        x = rs.standard_normal((n, t_max, K, D)).astype(np.float32)
        for i, l in enumerate(lengths):
            x[i, l:] = 0
        data[mod] = x.tolist()
    target = rs.uniform(0, 1, (n, t_max)).astype(np.float32)
    for i, l in enumerate(lengths):
        target[i, l:] = 0
    return data, target.tolist(), lengths


def digest(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return [list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()]


def main():
    src = open(os.path.join(REF, 'MFT', 'train.py')).read()
    tree = ast.parse(src)
    want = {'chunks', 'generateInputChunkHelper', 'generateTrainBatch'}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {f.name for f in fns} == want
    ns = {'torch': torch, 'shuffle': shuffle, 'itemgetter': itemgetter}
    exec(compile(ast.Module(body=fns, type_ignores=[]), 'MFT/train.py', 'exec'), ns)
    data, target, lengths = corpus()
    out = {}
    for tag, bs, on_eval in [('train_bs4', 4, False), ('eval_bs1', 1, True), ('eval_bs5', 5, True)]:
        random.seed(123)
        batches = []
        for d, tg, mask, ln in ns['generateTrainBatch'](data, target, list(lengths), None, batch_size=bs, onEval=on_eval):
            batches.append({'lengths': ln, 'target': digest(tg.numpy()), 'mask': digest(mask.numpy()),
                            'data': {m: digest(v.numpy()) for m, v in d.items()}})
        out[tag] = batches
    with open(os.path.join(OUT, 'batcher.json'), 'w') as f:
        json.dump(out, f)
    print('batcher golden written:', {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
