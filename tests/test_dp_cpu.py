"""CPU (gloo, world size 2): the data-parallel exchange steps of the training path
(vpd_b200/dp.py - the functions ModelTrainer calls under NCCL): gradient SUM per bucket and as
one call, replicas staying identical after the same update, the epoch loss over the global
batch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vpd_b200 import dp

assert dp.active() is None                      # not initialised: every call is a no-op
g0 = torch.arange(10, dtype=torch.float32)
dp.sum_gradients(g0)
assert torch.equal(g0, torch.arange(10, dtype=torch.float32))
assert dp.epoch_loss(torch.tensor([6.0], dtype=torch.float64), 3) == 2.0

dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
assert world == 2 and dp.active() is not None
n = 1000
gen = torch.Generator().manual_seed(100 + rank)
local = torch.randn(n, generator=gen)
other = torch.randn(n, generator=torch.Generator().manual_seed(100 + (1 - rank)))
expect = local + other if rank == 0 else other + local

a = local.clone()
dp.sum_gradients(a)                             # one call over the arena
assert torch.allclose(a, expect, rtol=0, atol=1e-6)

b = local.clone()                               # bucket by bucket, last layers first
buckets = [(700, 300), (256, 444), (0, 256)]
dp.sum_gradients(b, buckets)
assert torch.equal(a, b)
try:
    dp.sum_gradients(local.clone(), [(0, 10)])
    raise SystemExit('a partial bucket list must be rejected')
except AssertionError:
    pass

# the per-bucket checker the GPU test and bench.py use: passes on the real sum, and a bucket
# whose exchange was skipped (it still holds the local gradient) is caught - on its own norm,
# even when it is a tiny part of the arena
chk = dp.check_bucket_sums(b, expect, buckets)
assert chk['ok'] and chk['max_rel'] <= 1e-6 and len(chk['per_bucket']) == 3
skipped = b.clone()
skipped[0:256] = local[0:256]                   # bucket (0, 256) never all-reduced
chk = dp.check_bucket_sums(skipped, expect, buckets)
assert not chk['ok'] and chk['max_rel'] > 0.1
assert [r > 0.1 for _, _, r in chk['per_bucket']] == [True, False, False]
tiny = [(0, 4), (4, 996)]
skipped = b.clone()
skipped[0:4] = local[0:4]
assert not dp.check_bucket_sums(skipped, expect, tiny)['ok']
try:
    dp.check_bucket_sums(b, expect, [(0, 500), (600, 400)])
    raise SystemExit('buckets with a gap must be rejected')
except AssertionError:
    pass

# replicas start from rank 0's state whatever their own initialisation was
state = [torch.full((5,), float(rank + 1)), None, torch.arange(3) + 10 * rank]
dp.broadcast_state(state)
assert torch.equal(state[0], torch.ones(5)) and torch.equal(state[2], torch.arange(3))
assert dp.rank() == rank
dp.barrier()

# identical parameters + summed gradients -> identical parameters after the step
p = torch.ones(n)
p -= 0.1 * a
gathered = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
assert torch.equal(gathered[0], gathered[1])

# epoch loss = sum of losses over ranks / sum of frames over ranks, same value on every rank
loss = dp.epoch_loss(torch.tensor([10.0 * (rank + 1)], dtype=torch.float64), 4 + rank)
assert abs(loss - 30.0 / 9.0) < 1e-12
print('RANK%dOK' % rank, flush=True)
dist.destroy_process_group()
'''


def test_two_rank_gloo_gradient_sum_and_epoch_loss(tmp_path):
    script = os.path.join(str(tmp_path), 'dp_gloo.py')
    with open(script, 'w') as fp:
        fp.write(WORKER)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29623', script, ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-2500:])
    assert res.stdout.count('OK') == 2, res.stdout      # the two ranks' lines may interleave


def test_bind_host_to_device_is_best_effort():
    """No GPU / no readable topology: the NUMA binding returns None and leaves the affinity alone."""
    import os
    from vpd_b200 import dp
    before = os.sched_getaffinity(0)
    assert dp.bind_host_to_device(0) is None
    assert os.sched_getaffinity(0) == before
