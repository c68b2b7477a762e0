"""Oracle: frame-batch assembly, numpy + torch CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, on already-decoded uint8 arrays (PNG decode is out of scope):
  * `_BaseDataset._load_image`  vpd_dataset/common.py:52-58 + Normalize (:87)
  * `_BaseDataset._load_flow`   vpd_dataset/common.py:63-69
  * `GenericDataset.__getitem__` deterministic part, vpd_dataset/single_frame.py:168-206
    (teacher-row select by flip bit :171-174, cat RGB+flow :197, horizontal
    flip + flow-x negate :199-203); jitter / mask-noise / RandomResizedCrop are
    stochastic and unseeded in the reference and are not restated.
  * `FrameDataset.__getitem__`  vpd_dataset/single_frame.py:373-400 (apply: stack
    [orig, flipped], flipped flow has x negated :394-395), default path
    (jitter 0).
  * `GenericDataset.load_default` motion-target concat, single_frame.py:247-258.

Input convention: `rgb_u8` is the RGB-ordered image (i.e. after
cv2.cvtColor(BGR2RGB), common.py:54) uint8 [H,W,3]; `flow_u8` is the flow PNG
as cv2.imread returns it, uint8 [H,W,3] (or [H,W,2]); channels 0,1 are used.
"""
import numpy as np
import torch


def load_image(rgb_u8, mean, std):
    """common.py:58 `transform(torch.FloatTensor(rgb).permute(2,0,1) / 255.)`
    with transform = Normalize(mean, std, inplace=True): fp32 x/255 then
    sub_(mean) then div_(std), mean/std cast to fp32 first."""
    x = torch.FloatTensor(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1) / 255.
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1)
    x = x.clone()
    x.sub_(m).div_(s)
    return x


def load_flow(flow_u8):
    """common.py:69 `torch.FloatTensor((flow[:, :, :2] / 255) - 0.5).permute(2,0,1)`:
    float64 numpy arithmetic, then one rounding to fp32."""
    f = (np.asarray(flow_u8)[:, :, :2] / 255) - 0.5
    return torch.FloatTensor(f).permute(2, 0, 1)


def lut(mean, std):
    """The five 256-entry fp32 tables equivalent to load_image / load_flow
    (SURVEY §4 [probed]: a LUT reproduces the reference bit for bit)."""
    u = np.arange(256, dtype=np.uint8).reshape(1, 256, 1)
    rgb = load_image(np.repeat(u, 3, axis=2), mean, std).numpy()[:, 0, :]     # [3,256]
    fl = load_flow(np.repeat(u, 3, axis=2)).numpy()[:, 0, :]                   # [2,256]
    return np.concatenate([rgb, fl], axis=0).astype(np.float32)               # [5,256]


def train_item(rgb_u8, flow_u8, teacher, flip, mean, std, mask_u8=None, noise=None):
    """single_frame.py:168-206 without the unseeded draws (ColorJitter, RandomResizedCrop).
    teacher: fp32 [2, E] (rows = unflipped / flipped) or [E]. Returns
    (img fp32 [5,H,W] (or [3,H,W] if flow_u8 is None), emb fp32 [E]).
    mask_u8 [H,W] (first channel of <n>.mask.png) with noise fp32 [3,H,W]: the masked noise
    of single_frame.py:179-191 with the noise tensor given instead of drawn -
    `mask = mask_png[:,:,0] == 0; noise[:, mask] = 0; img += noise`, before the flow is
    stacked and before the flip."""
    emb = np.asarray(teacher, dtype=np.float32)
    if emb.ndim == 2:
        emb = emb[int(flip), :]
    else:
        flip = False
    img = load_image(rgb_u8, mean, std)
    if mask_u8 is not None and noise is not None:
        mask = torch.from_numpy(np.asarray(mask_u8) == 0)
        nz = torch.as_tensor(noise, dtype=torch.float32).clone()
        nz[:, mask] = 0
        img += nz
    if flow_u8 is not None:
        img = torch.cat((img, load_flow(flow_u8)))
    if flip:
        img = torch.flip(img, (2,))
        if flow_u8 is not None:
            img[3, :, :] *= -1
    return img, torch.FloatTensor(emb)


def train_batch(rgb_u8, flow_u8, teacher, flips, mean, std, mask_u8=None, noise=None,
                noise_on=None):
    """Collated batch like the DataLoader does: img [B,5,H,W], emb [B,E]. With mask_u8
    [B,H,W] + noise [B,3,H,W] the frames with noise_on[i] (default all) get the masked noise."""
    imgs, embs = [], []
    for i in range(len(rgb_u8)):
        on = mask_u8 is not None and noise is not None and (noise_on is None or noise_on[i])
        a, b = train_item(rgb_u8[i], None if flow_u8 is None else flow_u8[i],
                          teacher[i], bool(flips[i]), mean, std,
                          mask_u8[i] if on else None, noise[i] if on else None)
        imgs.append(a)
        embs.append(b)
    return torch.stack(imgs), torch.stack(embs)


def apply_item(rgb_u8, flow_u8, mean, std, flip=True):
    """single_frame.py:373-400, jitter_count = 0. Returns fp32 [k,5,H,W],
    k = 2 if flip else 1."""
    img = load_image(rgb_u8, mean, std)
    imgs = [img]
    flip_imgs = [torch.flip(img, (2,))] if flip else None
    if flow_u8 is not None:
        flow = load_flow(flow_u8)
        imgs = [torch.cat((x, flow)) for x in imgs]
        if flip_imgs:
            flip_flow = torch.flip(flow, (2,))
            flip_flow[0, :, :] *= -1
            flip_imgs = [torch.cat((x, flip_flow)) for x in flip_imgs]
    if flip_imgs:
        imgs += flip_imgs
    return torch.stack(imgs)


def apply_batch(rgb_u8, flow_u8, mean, std, flip=True):
    return torch.stack([apply_item(rgb_u8[i], None if flow_u8 is None else flow_u8[i],
                                   mean, std, flip) for i in range(len(rgb_u8))])


def motion_targets(video_embs, min_pose_score=0.5):
    """single_frame.py:208-261 for one video with embed_time=True,
    normalize_target=False: keeps frames whose score passes and whose
    predecessor entry is frame_num-1, target = concat[e, e - e_prev] on the
    last axis. video_embs: list of (frame_num, ndarray [2,D] or [D], meta)."""
    out = []
    for i in range(len(video_embs)):
        frame_num, emb, meta = video_embs[i]
        score = meta.get('dp_score', meta.get('kp_score'))
        if score is None:
            raise NotImplementedError()
        if score < min_pose_score:
            continue
        if i == 0 or video_embs[i - 1][0] != frame_num - 1:
            continue
        prev = video_embs[i - 1][1]
        out.append((frame_num, np.concatenate(
            [emb, emb - prev], axis=0 if emb.ndim == 1 else 1), meta))
    return out
