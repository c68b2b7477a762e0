"""Worker for tests/test_dp_gpu.py (launched under torch.distributed.run, one rank per GPU).

Checks, per all-reduce bucket, that the gradient arena the trainer leaves behind equals the SUM
over ranks of the gradients each rank computes on its own for the same batch
(`ModelTrainer.dp_self_check`), that the check FAILS when one bucket's all-reduce is skipped,
and that replicas stay bit-identical through optimizer steps even though the ranks construct
their models under different seeds (rank 0's state is broadcast)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200 import dp, synth, RGBF_EmbeddingModel, ModelTrainer        # noqa: E402
from vpd_b200.assemble import assemble_batch                             # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    local = int(os.environ['LOCAL_RANK'])
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    world = dist.get_world_size()
    B = 16
    torch.manual_seed(100 + rank)               # DIFFERENT init per rank: the trainer must fix it
    enc = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda')
    tr = ModelTrainer(enc, True)
    opt, _ = tr.get_optimizer(5e-4)
    gp = [torch.empty_like(enc._params) for _ in range(world)]
    dist.all_gather(gp, enc._params)
    start_same = all(torch.equal(gp[0], g) for g in gp)
    rgb, flow = synth.crops(B, seed=10 + rank)
    teach = synth.teacher(B, seed=20 + rank)
    fl = synth.flips(B, seed=30 + rank)
    batch = assemble_batch(rgb.to(dev), flow.to(dev), synth.FS_MEAN_STD, flip=fl.to(dev),
                           teacher=teach.to(dev))
    # 1) sum over ranks, bucket by bucket
    res = tr.dp_self_check(batch['img'], batch['emb'], B)
    # 2) negative control: skip the all-reduce of ONE bucket -> the check must fail
    real = dp.sum_bucket
    calls = {'n': 0}

    def dropping(grads, offset, count, async_op=False):
        calls['n'] += 1
        if calls['n'] == 2:                     # the second exchange of the trainer's step
            return None
        return real(grads, offset, count, async_op=async_op)

    # (dp_self_check's own reference all-reduce goes through dist directly, not sum_bucket)
    if not res['overlapped']:
        calls['n'] = 1                          # single-call path: the one exchange is dropped
    dp.sum_bucket = dropping
    try:
        bad = tr.dp_self_check(batch['img'], batch['emb'], B)
    finally:
        dp.sum_bucket = real
    # 3) a few optimizer steps keep the replicas bit-identical
    for _ in range(3):
        loss = tr.epoch([batch], optimizer=opt)
    dist.all_gather(gp, enc._params)
    params_same = all(torch.equal(gp[0], g) for g in gp)
    ev = tr.epoch([batch])                      # eval loss is reduced over the ranks too
    evs = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(evs, torch.tensor([ev], device=dev, dtype=torch.float64))
    eval_same = all(torch.equal(evs[0], e) for e in evs)
    if rank == 0:
        print('DP_RESULT ok={} max_rel={:.3e} buckets={} overlapped={} dropped_ok={} '
              'dropped_max_rel={:.3e} start_identical={} params_identical={} eval_identical={} '
              'loss={:.4f}'.format(res['ok'], res['max_rel'], res['buckets'], res['overlapped'],
                                   bad['ok'], bad['max_rel'], start_same, params_same, eval_same,
                                   loss), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
