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


def bind_host_to_device(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned
    host batches it allocates afterwards (first touch) and the threads that fill them sit next
    to that GPU's PCIe root. With one process per GPU on a two-socket host, half of the
    host-to-device traffic otherwise crosses the socket interconnect, which is what bounds the
    end-to-end rate of 8 ranks fed fp32 batches (84 MB per step and rank). Best effort: returns
    the CPU list it applied, or None when the topology cannot be read (single node, container
    without sysfs, non-Linux)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = '{:04x}:{:02x}:{:02x}.0'.format(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    try:
        base = '/sys/bus/pci/devices/' + str(bdf).lower()
        with open(base + '/numa_node') as fp:
            node = int(fp.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node{}/cpulist'.format(node)) as fp:
            spec = fp.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None

