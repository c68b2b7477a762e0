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
