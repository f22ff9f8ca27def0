"""CPU: the reference arm of bench.py (the reference's own classes from baseline/_ref -- or, without them / with --port, the oracle
port -- timed on host cores) prints ONE JSON line with the contract's keys, and the product arm refuses to run without a GPU instead
of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize('extra,kind', [([], None), (['--port'], 'port'), (['--config', 'c4'], None), (['--config', 'c1'], None)])
def test_reference_arm_prints_one_contract_line(extra, kind):
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if kind is None:
        if not ref_loader.available() and extra:
            pytest.skip('baseline/_ref missing: only c2 has a port arm')
        kind = 'reference' if ref_loader.available() else 'port'
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-sample', '2', '--seq', '16', '--layers', '1'] + extra, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'narratives/s' and d['value'] > 0 and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and 'workload' in d['config']


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode != 0 and not any(l.startswith('{') for l in out.stdout.splitlines())
