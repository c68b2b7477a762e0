"""Host side of the augmented training batch: the random DRAWS of the reference's
`GenericDataset.__getitem__` (vpd_dataset/single_frame.py:168-206, augment=True), made in the
reference's order from the generators the reference uses, packed for the device kernel
(`vpd_assemble_nchw_aug`, csrc/augment.cu) which does the pixel work.

Per item the reference consumes, in this order:
  1. `random.choice(self.data)`                       Python `random`   (common.py:107-108)
  2. `_randbool()` = `random.getrandbits(1) > 0`      Python `random`   (single_frame.py:24-25,
     only when the teacher entry has the two rows [unflipped, flipped])            :171-174)
  3. `ColorJitter.get_params`: `torch.randperm(4)` + four `uniform_` draws   torch global RNG
     (inside `_load_image`'s transform, common.py:58,88-92)
  4. `random.random() <= RANDOM_MASK_PROB`            Python `random`   (single_frame.py:179)
     and, if the coin is on AND `<n>.mask.png` exists, `torch.randn(img.shape)`  (:184)
  5. `RandomResizedCrop.get_params`: up to 10 x (2 `uniform_`) + 2 `randint`  torch global RNG
     (common.py:79-80, single_frame.py:204-205)
`draw_batch` does exactly that (it calls torchvision's own `get_params`, so the parameter
arithmetic is the reference's by construction): after `random.seed(s); torch.manual_seed(s)`
it yields the parameters a single-process reference loader yields after the same seeding, which
is how tests/golden/augment.npz pins the whole path against the unmodified reference.
"""
import random

import numpy as np
import torch

JITTER_KWARGS = {'brightness': 0.2, 'contrast': 0.2, 'saturation': 0.05, 'hue': 0.05}  # common.py:11-12
CROP_SCALE = (0.5, 1.)           # common.py:50
CROP_RATIO = (0.9, 1.1)
RANDOM_MASK_PROB = 0.5           # single_frame.py:20
RANDOM_NOISE_SD = 0.05 ** 0.5    # single_frame.py:21


def _jitter_ranges(kw):
    """transforms.ColorJitter.__init__'s `_check_input`: [1-v, 1+v] clipped at 0; hue [-v, v]"""
    out = {}
    for name in ('brightness', 'contrast', 'saturation'):
        v = kw.get(name)
        out[name] = None if not v else [max(0., 1. - v), 1. + v]
    v = kw.get('hue')
    out['hue'] = None if not v else [-v, v]
    return out


class AugmentParams:
    """Draws for one batch, host tensors ready for `assemble.assemble_batch_aug`."""

    def __init__(self, B):
        self.index = torch.zeros(B, dtype=torch.int32)
        self.flip = torch.zeros(B, dtype=torch.uint8)
        self.jitter_order = torch.full((B, 4), 255, dtype=torch.uint8)
        self.jitter_factor = torch.zeros((B, 8), dtype=torch.float32)
        self.jitter = [None] * B             # (fn_idx list, b, c, s, h) as drawn (doubles)
        self.crop = torch.zeros((B, 4), dtype=torch.int32)
        self.noise_on = torch.zeros(B, dtype=torch.uint8)
        self.noise = None                    # fp32 [B,3,H,W] when the host draws the noise
        self.has_jitter = None               # any ColorJitter op active (None: look at the orders)

    def set_jitter(self, b, fn_idx, bf, cf, sf, hf):
        """pack one ColorJitter draw the way the kernel wants it (factors rounded from the
        doubles exactly where torchvision's fp32 tensor ops round them)"""
        fn_idx = [int(v) for v in fn_idx]
        self.jitter[b] = (fn_idx, bf, cf, sf, hf)
        present = {0: bf, 1: cf, 2: sf, 3: hf}
        self.jitter_order[b] = torch.tensor(
            [f if present[f] is not None else 255 for f in fn_idx], dtype=torch.uint8)
        d = lambda v: 1.0 if v is None else float(v)
        self.jitter_factor[b] = torch.tensor(
            [d(bf), d(cf), 1.0 - d(cf), d(sf), 1.0 - d(sf), 0.0 if hf is None else float(hf),
             0.0, 0.0], dtype=torch.float64).to(torch.float32)

    def to(self, device):
        for k in ('index', 'flip', 'jitter_order', 'jitter_factor', 'crop', 'noise_on', 'noise'):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device, non_blocking=True))
        return self


def draw_batch(B, pool_size, H, W, two_rows=True, has_mask=None, host_noise=False, channels=5,
               jitter=JITTER_KWARGS, crop=True):
    """The reference's draws for B consecutive `__getitem__` calls (module docstring).
    has_mask: bool [pool] (does `<n>.mask.png` exist), or None = no masks at all.
    host_noise: draw `torch.randn` on the host like the reference (exact stream; 3*H*W floats
    per noisy frame) instead of leaving the noise to the device generator."""
    from torchvision import transforms
    p = AugmentParams(B)
    ranges = _jitter_ranges(jitter or {})
    use_jitter = any(v is not None for v in ranges.values())
    shape_probe = torch.empty((channels, H, W), device='meta') if crop else None
    if host_noise:
        p.noise = torch.zeros((B, 3, H, W), dtype=torch.float32)
    for b in range(B):
        idx = random.choice(range(pool_size))                                  # 1
        p.index[b] = idx
        if two_rows:
            p.flip[b] = 1 if random.getrandbits(1) > 0 else 0                  # 2
        if use_jitter:                                                         # 3
            fn_idx, bf, cf, sf, hf = transforms.ColorJitter.get_params(
                ranges['brightness'], ranges['contrast'], ranges['saturation'], ranges['hue'])
            p.set_jitter(b, fn_idx.tolist(), bf, cf, sf, hf)
        if random.random() <= RANDOM_MASK_PROB:                                # 4
            if has_mask is not None and bool(has_mask[idx]):
                p.noise_on[b] = 1
                if host_noise:
                    p.noise[b] = torch.randn((3, H, W)) * RANDOM_NOISE_SD
        if crop:                                                               # 5
            i, j, h, w = transforms.RandomResizedCrop.get_params(
                shape_probe, list(CROP_SCALE), list(CROP_RATIO))
            p.crop[b] = torch.tensor([i, j, h, w], dtype=torch.int32)
        else:
            p.crop[b] = torch.tensor([0, 0, H, W], dtype=torch.int32)
    p.has_jitter = bool(use_jitter)
    check_params(p, H, W)
    return p


def check_params(p, H, W):
    """host-side validation of what the kernel cannot check cheaply"""
    c = p.crop.cpu().numpy() if isinstance(p.crop, torch.Tensor) else np.asarray(p.crop)
    ok = ((c[:, 2] >= 1) & (c[:, 3] >= 1) & (c[:, 0] >= 0) & (c[:, 1] >= 0)
          & (c[:, 0] + c[:, 2] <= H) & (c[:, 1] + c[:, 3] <= W))
    assert bool(ok.all()), 'crop box outside the frame'


def draw_batch_fast(B, pool_size, H, W, two_rows=True, has_mask=None, generator=None,
                    jitter=JITTER_KWARGS, crop=True):
    """Vectorised draws with the SAME distributions as `draw_batch` (ColorJitter.get_params,
    RandomResizedCrop.get_params incl. its 10-try rejection loop and central fallback, flip and
    noise coins, sampling with replacement) but from one torch.Generator in batched calls: about
    0.3 ms per 256 frames instead of 22 ms, so the augmenting loader keeps up with the training
    step. Not stream-identical with the reference (use `draw_batch` for that); the noise always
    comes from the device generator."""
    import math
    g = generator
    p = AugmentParams(B)
    p.jitter = None
    p.index = torch.randint(0, pool_size, (B,), generator=g).to(torch.int32)
    if two_rows:
        p.flip = torch.randint(0, 2, (B,), generator=g).to(torch.uint8)
    ranges = _jitter_ranges(jitter or {})
    if any(v is not None for v in ranges.values()):
        order = torch.rand((B, 4), generator=g).argsort(dim=1)             # uniform permutations
        fac = torch.zeros((B, 4), dtype=torch.float64)
        present = torch.zeros(4, dtype=torch.bool)
        for k, name in enumerate(('brightness', 'contrast', 'saturation', 'hue')):
            r = ranges[name]
            if r is not None:
                present[k] = True
                fac[:, k] = (torch.rand(B, generator=g, dtype=torch.float32).double()
                             * (r[1] - r[0]) + r[0])
            elif k < 3:
                fac[:, k] = 1.0
        p.jitter_order = torch.where(present[order], order, torch.full_like(order, 255)).to(torch.uint8)
        jf = torch.zeros((B, 8), dtype=torch.float64)
        jf[:, 0], jf[:, 1], jf[:, 2] = fac[:, 0], fac[:, 1], 1.0 - fac[:, 1]
        jf[:, 3], jf[:, 4], jf[:, 5] = fac[:, 2], 1.0 - fac[:, 2], fac[:, 3]
        p.jitter_factor = jf.to(torch.float32)
    coin = torch.rand(B, generator=g) <= RANDOM_MASK_PROB
    if has_mask is not None:
        p.noise_on = (coin & torch.as_tensor(has_mask, dtype=torch.bool)[p.index.long()]).to(torch.uint8)
    if crop:
        tries = 10
        area = float(H * W)
        ta = area * (torch.rand((B, tries), generator=g, dtype=torch.float32).double()
                     * (CROP_SCALE[1] - CROP_SCALE[0]) + CROP_SCALE[0])
        lr0, lr1 = math.log(CROP_RATIO[0]), math.log(CROP_RATIO[1])
        ar = torch.exp(torch.rand((B, tries), generator=g, dtype=torch.float32).double() * (lr1 - lr0) + lr0)
        w = torch.round(torch.sqrt(ta * ar)).long()
        h = torch.round(torch.sqrt(ta / ar)).long()
        ok = (w > 0) & (w <= W) & (h > 0) & (h <= H)
        first = ok.int().argmax(dim=1)
        any_ok = ok.any(dim=1)
        w = w.gather(1, first[:, None])[:, 0]
        h = h.gather(1, first[:, None])[:, 0]
        # central fallback (transforms.py:951-963)
        in_ratio = float(W) / float(H)
        if in_ratio < min(CROP_RATIO):
            fw, fh = W, int(round(W / min(CROP_RATIO)))
        elif in_ratio > max(CROP_RATIO):
            fh = H
            fw = int(round(fh * max(CROP_RATIO)))
        else:
            fw, fh = W, H
        w = torch.where(any_ok, w, torch.full_like(w, fw))
        h = torch.where(any_ok, h, torch.full_like(h, fh))
        i = (torch.rand(B, generator=g, dtype=torch.float64) * (H - h + 1).double()).floor().long()
        j = (torch.rand(B, generator=g, dtype=torch.float64) * (W - w + 1).double()).floor().long()
        i = torch.where(any_ok, torch.minimum(i, H - h), (H - h) // 2)
        j = torch.where(any_ok, torch.minimum(j, W - w), (W - w) // 2)
        p.crop = torch.stack([i, j, h, w], dim=1).to(torch.int32)
    else:
        p.crop = torch.tensor([[0, 0, H, W]], dtype=torch.int32).repeat(B, 1)
    p.has_jitter = any(v is not None for v in ranges.values())
    check_params(p, H, W)
    return p
