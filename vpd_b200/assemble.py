"""Python face of K1 (frame-batch assembly) - see include/vpd_b200.h.

`assemble_batch` / `assemble_apply` reproduce what the reference's Dataset
`__getitem__` + DataLoader collation produce for the deterministic part of
vpd_dataset/single_frame.py:168-206 and :373-400, from uint8 crops already on
the device."""
import torch

from ._lib import lib, stream_ptr


def _check(t, name, dtype, device, ndim=None):
    """The C ABI takes raw pointers: a wrong dtype / device / stride is silently misread there,
    so it is rejected here. None passes (optional arguments)."""
    if t is None or not isinstance(t, torch.Tensor):
        return t
    if t.dtype != dtype:
        raise TypeError('{} must be {} (got {})'.format(name, dtype, t.dtype))
    if t.device != device:
        raise ValueError('{} must live on {} (got {})'.format(name, device, t.device))
    if not t.is_contiguous():
        raise ValueError('{} must be contiguous'.format(name))
    if ndim is not None and t.dim() not in ndim:
        raise ValueError('{} must have {} dimensions (got shape {})'.format(name, ndim, tuple(t.shape)))
    return t


def _check_batch_args(rgb, flow, flip, teacher, index, mask=None, noise_on=None, noise=None):
    dev = rgb.device
    if dev.type != 'cuda':
        raise ValueError('K1 assembles on the GPU: rgb must be a CUDA tensor')
    _check(rgb, 'rgb', torch.uint8, dev, (4,))
    _check(flow, 'flow', torch.uint8, dev, (4,))
    _check(flip, 'flip', torch.uint8, dev, (1,))
    _check(teacher, 'teacher', torch.float32, dev, (2, 3))
    _check(index, 'index', torch.int32, dev, (1,))
    _check(mask, 'mask', torch.uint8, dev, (3,))
    _check(noise_on, 'noise_on', torch.uint8, dev, (1,))
    _check(noise, 'noise', torch.float32, dev, (4,))
    if flow is not None and flow.shape[:3] != rgb.shape[:3]:
        raise ValueError('flow {} does not match rgb {}'.format(tuple(flow.shape), tuple(rgb.shape)))
    B = rgb.shape[0] if index is None else index.numel()
    if flip is not None and flip.numel() != B:
        raise ValueError('flip has {} entries for {} frames'.format(flip.numel(), B))
    if teacher is not None and teacher.shape[0] != rgb.shape[0]:
        raise ValueError('teacher has {} rows for a pool of {} frames'.format(
            teacher.shape[0], rgb.shape[0]))


def check_index(index, pool):
    """Host-side range check of pool indices where they are drawn (the kernels trust them):
    raises for any index outside [0, pool). Accepts a CPU tensor / array; device tensors are
    checked only on request because it costs a synchronisation."""
    idx = torch.as_tensor(index)
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= pool):
        raise IndexError('pool index outside [0, {})'.format(pool))
    return index


def _mean_std(rgb_mean_std):
    mean = torch.tensor([float(v) for v in rgb_mean_std[0]], dtype=torch.float32)
    std = torch.tensor([float(v) for v in rgb_mean_std[1]], dtype=torch.float32)
    return mean, std


RANDOM_NOISE_SD = 0.05 ** 0.5        # vpd_dataset/single_frame.py:21
RANDOM_MASK_PROB = 0.5               # :20


def _noise_args(mask, noise_on, noise, noise_sd, seed):
    if noise is None and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())    # torch's global RNG, like the loader
    return (mask.contiguous(), noise_on, None if noise is None else noise.contiguous(),
            float(noise_sd), int(seed or 0))


def assemble_batch(rgb, flow, rgb_mean_std, flip=None, teacher=None, index=None, mask=None,
                   noise_on=None, noise=None, noise_sd=RANDOM_NOISE_SD, seed=None):
    """-> {'img': fp32 [B,C,H,W], 'emb': fp32 [B,E] (if teacher given)} on the device.

    rgb uint8 [P,H,W,3]; flow uint8 [P,H,W,>=2] or None; flip uint8 [B] or None;
    teacher fp32 [P,2,E] (rows unflipped/flipped) or [P,E]; index int32 [B] or None.
    mask uint8 [P,H,W] (first channel of `<n>.mask.png`) switches on the reference's noise
    augmentation (single_frame.py:179-191) for the frames with noise_on[b] != 0: Gaussian
    noise of sd `noise_sd` on the normalised RGB planes where the mask byte is not 0;
    `noise` fp32 [B,3,H,W] supplies the noise explicitly (else device Philox from `seed`)."""
    _check_batch_args(rgb, flow, flip, teacher, index, mask, noise_on, noise)
    mean, std = _mean_std(rgb_mean_std)
    P, H, W, _ = rgb.shape
    B = P if index is None else index.numel()
    C = 5 if flow is not None else 3
    img = torch.empty((B, 1, C, H, W), device=rgb.device, dtype=torch.float32)
    emb = None
    rows = tdim = 0
    if teacher is not None:
        teacher = teacher.contiguous()
        rows = teacher.shape[1] if teacher.dim() == 3 else 1
        tdim = teacher.shape[-1]
        emb = torch.empty((B, tdim), device=rgb.device, dtype=torch.float32)
    if mask is not None:
        lib().call('vpd_assemble_nchw_noise', rgb, flow, 0 if flow is None else flow.shape[-1],
                   index, flip, teacher, rows, tdim, mean, std, img, emb, B, H, W,
                   *_noise_args(mask, noise_on, noise, noise_sd, seed), stream_ptr(rgb.device))
    else:
        lib().call('vpd_assemble_nchw', rgb, flow, 0 if flow is None else flow.shape[-1], index,
                   flip, teacher, rows, tdim, mean, std, img, emb, B, H, W, 1,
                   stream_ptr(rgb.device))
    out = {'img': img[:, 0]}
    if emb is not None:
        out['emb'] = emb
    return out


def assemble_batch_aug(rgb, flow, rgb_mean_std, params, teacher=None, mask=None,
                       noise_sd=RANDOM_NOISE_SD, seed=None):
    """The reference's `augment=True` training batch (single_frame.py:168-206) on the device:
    ColorJitter -> Normalize -> masked noise -> flow -> flip -> RandomResizedCrop, with the
    draws in `params` (`vpd_b200.augment.draw_batch`, already on rgb's device).
    -> {'img': fp32 [B,C,H,W], 'emb': fp32 [B,E] (if teacher given)}.
    mask uint8 [P,H,W] enables the noise for the frames with params.noise_on; params.noise
    (fp32 [B,3,H,W], host-drawn) makes it exact, else the device generator runs from `seed`."""
    from .augment import check_params
    mean, std = _mean_std(rgb_mean_std)
    P, H, W, _ = rgb.shape
    if not params.crop.is_cuda:      # device-resident draws were validated when drawn (no sync)
        check_params(params, H, W)
    B = params.index.numel()
    C = 5 if flow is not None else 3
    img = torch.empty((B, C, H, W), device=rgb.device, dtype=torch.float32)
    emb = None
    rows = tdim = 0
    if teacher is not None:
        teacher = teacher.contiguous()
        rows = teacher.shape[1] if teacher.dim() == 3 else 1
        tdim = teacher.shape[-1]
        emb = torch.empty((B, tdim), device=rgb.device, dtype=torch.float32)
    jo, jf = params.jitter_order, params.jitter_factor
    has_jitter = getattr(params, 'has_jitter', None)       # set where the draws are made (host)
    if has_jitter is None:
        has_jitter = jo is not None and not bool((jo > 3).all())
    if not has_jitter:
        jo = jf = None
    nz = (None, None, None, 0.0, 0)
    if mask is not None:
        nz = _noise_args(mask, params.noise_on, params.noise, noise_sd, seed)
    lib().call('vpd_assemble_nchw_aug', rgb, flow, 0 if flow is None else flow.shape[-1],
               params.index, params.flip, teacher, rows,
               tdim, mean, std, img, emb, B, H, W, jo, jf, params.crop, *nz,
               stream_ptr(rgb.device))
    out = {'img': img}
    if emb is not None:
        out['emb'] = emb
    return out


def assemble_apply(rgb, flow, rgb_mean_std, flip=True):
    """FrameDataset batch: fp32 [B,k,C,H,W], k = 2 ([orig, flipped]) or 1."""
    _check_batch_args(rgb, flow, None, None, None)
    mean, std = _mean_std(rgb_mean_std)
    B, H, W, _ = rgb.shape
    k = 2 if flip else 1
    C = 5 if flow is not None else 3
    img = torch.empty((B, k, C, H, W), device=rgb.device, dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb, flow, 0 if flow is None else flow.shape[-1], None, None,
               None, 0, 0, mean, std, img, None, B, H, W, k, stream_ptr(rgb.device))
    return img


def assemble_stem(out, rgb, flow, rgb_mean_std, flip=None, teacher=None, index=None, k=1,
                  tgt=None, mask=None, noise_on=None, noise=None, noise_sd=RANDOM_NOISE_SD,
                  seed=None):
    """Fused device pipeline: write the network's own bf16 input layout
    [B*k, (H+7)//2, (W+9)//4, 64] (the padded image, 8 channel slots per pixel, space-to-depth
    2 x 4 - include/vpd_b200.h) straight into `out` (tensor or raw pointer)."""
    _check_batch_args(rgb, flow, flip, teacher, index, mask, noise_on, noise)
    _check(tgt, 'tgt', torch.float32, rgb.device, (2,))
    mean, std = _mean_std(rgb_mean_std)
    P, H, W, _ = rgb.shape
    B = P if index is None else index.numel()
    rows = tdim = 0
    if teacher is not None:
        rows = teacher.shape[1] if teacher.dim() == 3 else 1
        tdim = teacher.shape[-1]
        if tgt is not None and tuple(tgt.shape) != (B, tdim):
            raise ValueError('tgt must be [{}, {}] (got {})'.format(B, tdim, tuple(tgt.shape)))
    if mask is not None:
        assert k == 1, 'the noise augmentation is a training-batch option'
        lib().call('vpd_assemble_stem_noise', rgb, flow, 0 if flow is None else flow.shape[-1],
                   index, flip, teacher, rows, tdim, mean, std, out, tgt, B, H, W,
                   *_noise_args(mask, noise_on, noise, noise_sd, seed), stream_ptr(rgb.device))
        return B
    lib().call('vpd_assemble_stem', rgb, flow, 0 if flow is None else flow.shape[-1], index, flip,
               teacher, rows, tdim, mean, std, out, tgt, B, H, W, k, stream_ptr(rgb.device))
    return B * k
