"""Data-parallel exchange steps of the student training path, separated from
the CUDA trainer so the same code runs under `gloo` on CPU tensors in the tests.

Training is pure data parallel: each rank draws its own frames; the only exchange per step
is the SUM of the flat fp32 gradient arena (SUM, not mean: the loss is
`F.mse_loss(reduction='sum')`, train_vpd_model.py:87, so the gradient of the global batch is
the sum of the per-rank gradients), issued per bucket as the backward pass hands ranges out
(last layers first). Per epoch one more pair of scalars is summed: loss and frame count."""
import torch


def active():
    """torch.distributed when it is initialised with more than one rank, else None"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def sum_bucket(grads, offset, count, async_op=False):
    """all-reduce(SUM) grads[offset:offset+count] in place; returns the work handle (or None)"""
    dist = active()
    if dist is None or count <= 0:
        return None
    return dist.all_reduce(grads[offset:offset + count], op=dist.ReduceOp.SUM, async_op=async_op)


def sum_gradients(grads, buckets=None):
    """Whole-arena exchange: one call, or bucket by bucket (`buckets` = [(offset, count)]
    partitioning the arena, in the order the backward pass produces them)."""
    if active() is None:
        return
    if buckets is None:
        sum_bucket(grads, 0, grads.numel())
        return
    covered = sum(c for _, c in buckets)
    assert covered == grads.numel(), 'buckets must partition the gradient arena'
    works = [sum_bucket(grads, o, c, async_op=True) for o, c in buckets]
    for w in works:
        if w is not None:
            w.wait()


def epoch_loss(loss_sum, frames):
    """`sum of losses / number of frames` over all ranks (train_vpd_model.py:93-98 on the global
    batch). loss_sum: 1-element float64 tensor on the training device; frames: int."""
    dist = active()
    if dist is None:
        return loss_sum.item() / frames
    total = loss_sum.clone()
    cnt = torch.tensor([float(frames)], device=loss_sum.device, dtype=torch.float64)
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return total.item() / cnt.item()


def rank():
    dist = active()
    return 0 if dist is None else dist.get_rank()


def barrier():
    dist = active()
    if dist is not None:
        dist.barrier()


def broadcast_state(tensors, src=0):
    """Replicas must START identical: copy rank `src`'s parameters / BN buffers / counters (and
    optimizer moments when resuming) to every rank. The reference has no DP; DDP does the same
    broadcast at construction. No-op without an initialised process group."""
    dist = active()
    if dist is None:
        return
    for t in tensors:
        if t is not None:
            dist.broadcast(t, src=src)


def check_bucket_sums(got, expect, buckets, rtol=1e-4):
    """Per bucket: is `got` (the arena after the trainer's bucketed all-reduce) the SUM over ranks
    `expect` (all-reduce of every rank's local gradients of the same step)? A bucket that was
    never exchanged still holds one rank's local gradient, i.e. a relative error of order 1, so
    every bucket is judged on its own norm, not the arena's.
    -> {'ok', 'max_rel', 'per_bucket': [(offset, count, rel)]}; buckets must tile the arena."""
    covered = sorted((int(o), int(c)) for o, c in buckets)
    pos = 0
    for o, c in covered:
        assert o == pos, 'buckets must partition the gradient arena (gap/overlap at {})'.format(o)
        pos = o + c
    assert pos == got.numel() == expect.numel(), 'buckets cover {} of {}'.format(pos, got.numel())
    per = []
    for o, c in covered:
        e = expect[o:o + c].double()
        rel = ((got[o:o + c].double() - e).norm() / (e.norm() + 1e-300)).item()
        per.append((o, c, rel))
    max_rel = max(r for _, _, r in per)
    return {'ok': bool(max_rel <= rtol), 'max_rel': max_rel, 'per_bucket': per}
