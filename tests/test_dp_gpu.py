"""Data-parallel path on >= 2 GPUs: NCCL all-reduce(SUM) of the gradient arena, bucketed and
overlapped with the backward pass; replicas must stay bit-identical."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('overlap', ['1', '0'])
def test_two_rank_nccl_gradient_sum(overlap):
    env = dict(os.environ, VPD_DP_OVERLAP=overlap)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29621',
           os.path.join(ROOT, 'tests', 'dp_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    m = re.search(r'DP_RESULT rel_vs_separate_run=([\d.]+) grads_identical_across_ranks=(\w+) '
                  r'params_identical_across_ranks=(\w+) loss=([\d.]+) overlapped=(\w+)', res.stdout)
    assert m, res.stdout[-2000:]
    assert m.group(2) == 'True' and m.group(3) == 'True'
    assert float(m.group(1)) < 0.6          # two bf16 runs of the same step (noise floor)
    assert m.group(5) == ('True' if overlap == '1' else 'False')
