"""GPU tier (needs >= 2 GPUs, skipped otherwise): NCCL data parallelism == single-GPU step on the union batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import ROOT


@pytest.mark.parametrize('mode', ['eager', 'graph'])
def test_two_rank_step_equals_union_batch(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29631', os.path.join(ROOT, 'tests', 'ddp_worker.py'), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'ddp ok' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
