"""Data-parallel path on >= 2 GPUs: NCCL all-reduce(SUM) of the gradient arena, bucketed and
overlapped with the backward pass. Per bucket the result must be the sum of the ranks' own
gradients to fp32 rounding; skipping one bucket must be detected; replicas must start and stay
bit-identical. (bench.py --gpus N runs the same `dp_self_check` after its timed region and
prints it as `dp_check`, so the driver's scaling runs carry the evidence too.)"""
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
    m = re.search(r'DP_RESULT (.*)', res.stdout)
    assert m, res.stdout[-2000:]
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, 'dp_check.txt'), 'a') as fp:
        fp.write('overlap={} {}\n'.format(overlap, m.group(1)))
    kv = dict(item.split('=') for item in m.group(1).split())
    assert kv['ok'] == 'True' and float(kv['max_rel']) <= 1e-4, kv
    assert kv['overlapped'] == ('True' if overlap == '1' else 'False')
    assert int(kv['buckets']) == (4 if overlap == '1' else 1)
    assert kv['dropped_ok'] == 'False' and float(kv['dropped_max_rel']) > 1e-2, kv
    assert kv['start_identical'] == 'True' and kv['params_identical'] == 'True'
    assert kv['eval_identical'] == 'True'
