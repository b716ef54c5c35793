# -*- coding: utf-8 -*-
"""
Multi-GPU (``-m gpu``, skipped with fewer than 2 devices): a run sharded over ranks with NCCL
(torchrun, one process per GPU) must equal the single-GPU run exactly -- counters, images and
found histories -- for an analytic, a plasma and a mesh scene.  The script does the comparison.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_run_equals_single_gpu_run():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29541',
           os.path.join(ROOT, 'tests', 'scripts', 'dist_history_check.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0 and 'DIST_HISTORY_OK' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
