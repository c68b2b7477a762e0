"""On-disk ingest: the reference's crop directory
`<crop_dir>/<video>/<frame>.png` (+ `<frame>.<flow_img>.png`; README.md:152-183, writer
extract_square_crops.py:122-135, flow encoding raft/flow.py:80-93) packed once into flat
uint8 shards that are memory-mapped and copied to the GPU as the pools K1 assembles from -
PNG decoding (the reference's DataLoader workers) leaves the training loop entirely.

    pack_crop_dir(crop_dir, out_prefix, flow_img='flow')   # -> <prefix>.rgb.npy, .flow.npy, .json
    shard = load_shard(out_prefix)                         # memory-mapped arrays + index
    pools = shard.to_device('cuda')                        # (rgb_u8, flow_u8) device tensors
    idx = shard.rows_of(data)                              # (video, frame) keys -> pool rows

The decoded bytes are exactly what `_BaseDataset._load_image/_load_flow`
(vpd_dataset/common.py:52-69) see before their float conversion: BGR->RGB for the crop, the
flow PNG's stored channels as is (K1 uses the first two).
"""
import json
import os

import numpy as np

from .apply import read_crop_dir


class Shard:
    def __init__(self, rgb, flow, index, mask=None):
        self.rgb, self.flow, self.index, self.mask = rgb, flow, index, mask
        self._row = {}
        for v in index['videos']:
            for j, f in enumerate(v['frames']):
                self._row[(v['name'], int(f))] = v['first_row'] + j

    def __len__(self):
        return int(self.rgb.shape[0])

    def rows_of(self, data):
        """Pool rows of `(video, frame, ...)` tuples (e.g. vpd_b200.targets data lists);
        raises KeyError for a frame that is not in the shard."""
        return np.array([self._row[(d[0], int(d[1]))] for d in data], dtype=np.int64)

    def to_device(self, device='cuda'):
        import torch
        rgb = torch.from_numpy(np.array(self.rgb)).to(device)
        flow = None if self.flow is None else \
            torch.from_numpy(np.array(self.flow)).to(device)
        return rgb, flow

    def videos(self, materialize=False):
        """-> the `videos` list vpd_b200.apply.extract_corpus takes. The tensors are views of
        the memory-mapped shard (nothing is read until the extraction pipeline copies a chunk
        into its pinned staging buffer), so a corpus larger than RAM streams through;
        materialize=True reads every video into memory now."""
        import warnings
        import torch
        out = []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')         # read-only memmap -> tensor: never written
            for v in self.index['videos']:
                lo, hi = v['first_row'], v['first_row'] + len(v['frames'])
                take = (lambda a: np.array(a[lo:hi])) if materialize else (lambda a: a[lo:hi])
                out.append((v['name'], list(v['frames']), torch.from_numpy(take(self.rgb)),
                            None if self.flow is None else torch.from_numpy(take(self.flow))))
        return out


def _read_masks(crop_dir, videos, img_dim):
    """First channel of `<n>.mask.png` per frame ([n, H, W] uint8); frames without a mask file
    get zeros, which the noise augmentation treats as "nothing to perturb" - the reference
    skips the augmentation for them (single_frame.py:182)."""
    import cv2
    n = sum(len(v[1]) for v in videos)
    out = np.zeros((n, img_dim, img_dim), np.uint8)
    row = 0
    for name, frames, _, _ in videos:
        for f in frames:
            path = os.path.join(crop_dir, name, '{}.mask.png'.format(f))
            if os.path.exists(path):
                m = cv2.imread(path)
                if m.shape[:2] != (img_dim, img_dim):
                    m = cv2.resize(m, (img_dim, img_dim))
                out[row] = m[:, :, 0]
            row += 1
    return out


def pack_crop_dir(crop_dir, out_prefix, flow_img=None, img_dim=128, with_mask=False, nested=False):
    """Decode every `<video>/<n>.png` (and flow / mask PNG) once and write the shard files.
    nested: the tennis layout `<video>/<player>/<n>.png`; shard videos are `<video>/<player>`."""
    videos = read_crop_dir(crop_dir, flow_img, img_dim, nested=nested)
    n = sum(len(v[1]) for v in videos)
    rgb = np.lib.format.open_memmap(out_prefix + '.rgb.npy', mode='w+', dtype=np.uint8,
                                    shape=(n, img_dim, img_dim, 3))
    flow = None
    if flow_img:
        flow = np.lib.format.open_memmap(out_prefix + '.flow.npy', mode='w+', dtype=np.uint8,
                                         shape=(n, img_dim, img_dim, 3))
    index = {'img_dim': img_dim, 'flow_img': flow_img, 'with_mask': bool(with_mask), 'videos': []}
    if with_mask:
        np.save(out_prefix + '.mask.npy', _read_masks(crop_dir, videos, img_dim))
    row = 0
    for name, frames, vrgb, vflow in videos:
        k = len(frames)
        rgb[row:row + k] = vrgb.numpy()
        if flow is not None:
            flow[row:row + k] = vflow.numpy()
        index['videos'].append({'name': name, 'frames': [int(f) for f in frames], 'first_row': row})
        row += k
    rgb.flush()
    if flow is not None:
        flow.flush()
    with open(out_prefix + '.json', 'w') as fp:
        json.dump(index, fp)
    return index


def load_shard(prefix):
    with open(prefix + '.json') as fp:
        index = json.load(fp)
    rgb = np.load(prefix + '.rgb.npy', mmap_mode='r')
    flow = np.load(prefix + '.flow.npy', mmap_mode='r') if index['flow_img'] else None
    mask = np.load(prefix + '.mask.npy', mmap_mode='r') if index.get('with_mask') else None
    return Shard(rgb, flow, index, mask)
