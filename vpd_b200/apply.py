"""Corpus feature extraction - the B200 side of apply_vpd_model.py:92-179.

Per video: frames -> K1 assembly straight into the network's input layout (original
and horizontally flipped variant, flow-x negated on the flip) -> eval-mode encoder ->
`[(frame_num, np.float32 [2, D] (or [D] with no_flip), {}), ...]` sorted by frame and
pickled to `<out_dir>/<video>.emb.pkl` (util/io.py:35-37, README.md:185-194).

Multi-GPU: videos are partitioned across ranks (greedy longest-first), every rank
writes its own files, there is no collective.
"""
import json
import os
import pickle
import re

import numpy as np
import torch

BATCH_SIZE = 500        # apply_vpd_model.py:15 (frames per batch; x2 images with flip)


def shard_videos(frame_counts, world_size, rank):
    """Indices of the videos this rank processes: longest-processing-time greedy."""
    order = sorted(range(len(frame_counts)), key=lambda i: (-frame_counts[i], i))
    loads = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        loads[r] += frame_counts[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def format_video_embs(frame_nums, embs, flip):
    """(frame_num, ndarray, {}) tuples exactly as apply_vpd_model.py:163-168 builds them."""
    out = []
    for i, f in enumerate(frame_nums):
        out.append((int(f), embs[i, :, :] if flip else embs[i, 0, :], {}))
    out.sort(key=lambda t: t[0])          # frame numbers are unique per video
    return out


def store_pickle(path, obj):
    with open(path, 'wb') as fp:
        pickle.dump(obj, fp)


def embed_frames(model, rgb_u8, flow_u8, rgb_mean_std, flip=True, batch_size=BATCH_SIZE):
    """uint8 crops [n,H,W,3] (+ flow [n,H,W,>=2]) on the model's device ->
    np.float32 [n, k, D], k = 2 if flip else 1."""
    from .assemble import assemble_stem
    n, H, W, _ = rgb_u8.shape
    k = 2 if flip else 1
    out = torch.empty((n, k, model.emb_dim), device=rgb_u8.device, dtype=torch.float32)
    model.eval()
    for s in range(0, n, batch_size):
        e = min(n, s + batch_size)
        net = model._native(H, W, (e - s) * k)
        from ._lib import lib
        stem = lib().call('vpd_net_stem_input', net.handle)
        assemble_stem(stem, rgb_u8[s:e], None if flow_u8 is None else flow_u8[s:e],
                      rgb_mean_std, k=k)
        out[s:e] = model.embed_stem(stem, (e - s) * k, H, W).view(e - s, k, -1)
    return out.cpu().numpy()


def extract_corpus(model, videos, out_dir, rgb_mean_std, flip=True, rank=0, world_size=1,
                   batch_size=BATCH_SIZE):
    """videos: list of (name, frame_nums, rgb_u8 [n,H,W,3], flow_u8 or None) with host or
    device uint8 tensors. Writes this rank's `<name>.emb.pkl`; returns the names written."""
    counts = [len(v[1]) for v in videos]
    written = []
    os.makedirs(out_dir, exist_ok=True)
    for i in shard_videos(counts, world_size, rank):
        name, frame_nums, rgb, flow = videos[i]
        if len(frame_nums) == 0:
            continue
        rgb = rgb.to(model._dev, non_blocking=True)
        flow = None if flow is None else flow.to(model._dev, non_blocking=True)
        embs = embed_frames(model, rgb, flow, rgb_mean_std, flip, batch_size)
        store_pickle(os.path.join(out_dir, '{}.emb.pkl'.format(name)),
                     format_video_embs(frame_nums, embs, flip))
        written.append(name)
    return written


def load_model_dir(model_dir, model_epoch=None, device='cuda'):
    """config.json + `<name>.encoder.pt` as written by train_vpd_model.py:222-228,107-112.
    The reference writes the key 'motion' but its apply script reads 'embed_time'
    (apply_vpd_model.py:102); either is accepted."""
    from .rgb import RGBF_EmbeddingModel
    with open(os.path.join(model_dir, 'config.json')) as fp:
        cfg = json.load(fp)
    name = 'best_epoch' if model_epoch is None else 'epoch{:04d}'.format(model_epoch)
    model = RGBF_EmbeddingModel(cfg['encoder_arch'], cfg['emb_dim'], cfg['use_flow'], device)
    model.load_state_dict(torch.load(os.path.join(model_dir, name + '.encoder.pt'),
                                     map_location='cpu'))
    return model, cfg


def read_crop_dir(crop_dir, flow_img=None, img_dim=128, nested=False):
    """`<crop_dir>/<video>/<n>.png` (+ `<n>.<flow_img>.png`) -> videos list for
    extract_corpus (apply_vpd_model.py:69-89). Host-side PNG decode with cv2.
    nested: the tennis layout `<crop_dir>/<video>/<player>/<n>.png` (single_frame.py:51-57);
    the entries are then named `<video>/<player>`."""
    img_re = re.compile(r'^\d+\.png$')
    names = []
    for video_name in sorted(os.listdir(crop_dir)):
        vdir = os.path.join(crop_dir, video_name)
        if not os.path.isdir(vdir):
            continue
        if nested:
            names.extend('{}/{}'.format(video_name, p) for p in sorted(os.listdir(vdir))
                         if os.path.isdir(os.path.join(vdir, p)))
        else:
            names.append(video_name)
    videos = []
    for video_name in names:
        vdir = os.path.join(crop_dir, video_name)
        frames = sorted(int(os.path.splitext(f)[0]) for f in os.listdir(vdir) if img_re.match(f))
        rgb, flow = _read_frames(vdir, frames, flow_img, img_dim)
        videos.append((video_name, frames, rgb, flow))
    return videos


def _read_frames(vdir, frames, flow_img, img_dim):
    import cv2
    rgb = np.empty((len(frames), img_dim, img_dim, 3), np.uint8)
    flow = np.empty((len(frames), img_dim, img_dim, 3), np.uint8) if flow_img else None
    for i, f in enumerate(frames):
        im = cv2.cvtColor(cv2.imread(os.path.join(vdir, '{}.png'.format(f))), cv2.COLOR_BGR2RGB)
        if im.shape[:2] != (img_dim, img_dim):
            im = cv2.resize(im, (img_dim, img_dim))
        rgb[i] = im
        if flow_img:
            fl = cv2.imread(os.path.join(vdir, '{}.{}.png'.format(f, flow_img)))
            if fl.shape[:2] != (img_dim, img_dim):
                fl = cv2.resize(fl, (img_dim, img_dim))
            flow[i] = fl
    return torch.from_numpy(rgb), None if flow is None else torch.from_numpy(flow)


def read_tennis_crops(crop_dir, clip_names, flow_img=None, img_dim=128, players=('front', 'back')):
    """The tennis layout of apply_vpd_model.get_tennis_dataset (:36-66): clips are named
    `<src_video>_<start>_<end>`, crops live in `<crop_dir>/<src_video>/<player>/<frame>.png` with
    ABSOLUTE frame numbers; every clip yields one output "video" per player,
    `<player>__<clip>`, whose frame numbers are relative to the clip start. Frames without a
    crop are skipped; a player without any crop in the clip still gets an (empty) entry, like
    the reference's video list."""
    videos = []
    for clip in clip_names:
        src_video_name, start_frame, end_frame = clip.rsplit('_', 2)
        start_frame, end_frame = int(start_frame), int(end_frame)
        for player in players:
            vdir = os.path.join(crop_dir, src_video_name, player)
            present = [f for f in range(start_frame, end_frame + 1)
                       if os.path.isfile(os.path.join(vdir, '{}.png'.format(f)))]
            rgb, flow = _read_frames(vdir, present, flow_img, img_dim)
            videos.append(('{}__{}'.format(player, clip), [f - start_frame for f in present],
                           rgb, flow))
    return videos


def apply_crop_dir(model_dir, crop_dir, out_dir, flow_img=None, no_flip=False, model_epoch=None,
                   rank=0, world_size=1):
    """apply_vpd_model.main for a crop directory."""
    model, cfg = load_model_dir(model_dir, model_epoch)
    if cfg['use_flow']:
        assert flow_img is not None, 'No flow image name specified'
    videos = read_crop_dir(crop_dir, flow_img if cfg['use_flow'] else None, cfg['img_dim'])
    return extract_corpus(model, videos, out_dir, cfg['rgb_mean_std'], flip=not no_flip,
                          rank=rank, world_size=world_size)
