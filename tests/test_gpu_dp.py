"""GPU (needs >= 2 GPUs; skipped otherwise): data-parallel hardware equivalence -- the sharded backward + the C-ABI grouped NCCL
all-reduce (mt_allreduce_grads) reproduce the full-batch gradient on every rank (tools/dp_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_gradients_equal_full_batch_on_real_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f'needs >= 2 GPUs, have {n}')
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env['MT_AR_OVERLAP'] = '1'      # the overlapped all-reduce is opt-in: cover it here
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
                          '--master-port', '29571', os.path.join(ROOT, 'tools', 'dp_check.py')], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    assert out.stdout.count('C-ABI all-reduce used: True') == world
    assert out.stdout.count('overlapped all-reduce ranges: 3') == world
