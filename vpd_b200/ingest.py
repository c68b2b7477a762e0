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


def _decode_chunk(task):
    """Worker: decode frames[lo:hi] of one video straight into rows [row0, row0 + k) of the
    shard files (opened memory-mapped here: nothing but the row count travels back)."""
    import cv2
    (crop_dir, prefix, name, frames, row0, flow_img, img_dim, with_mask) = task
    cv2.setNumThreads(1)
    vdir = os.path.join(crop_dir, name)
    rgb = np.load(prefix + '.rgb.npy', mmap_mode='r+')
    flow = np.load(prefix + '.flow.npy', mmap_mode='r+') if flow_img else None
    mask = np.load(prefix + '.mask.npy', mmap_mode='r+') if with_mask else None

    def fit(im):
        return im if im.shape[:2] == (img_dim, img_dim) else cv2.resize(im, (img_dim, img_dim))
    for j, f in enumerate(frames):
        im = cv2.imread(os.path.join(vdir, '{}.png'.format(f)))
        rgb[row0 + j] = fit(cv2.cvtColor(im, cv2.COLOR_BGR2RGB))
        if flow is not None:
            flow[row0 + j] = fit(cv2.imread(os.path.join(vdir, '{}.{}.png'.format(f, flow_img))))
        if mask is not None:
            # first channel of `<n>.mask.png`; frames without one keep zeros, which the noise
            # augmentation treats as "nothing to perturb" - the reference skips it for them
            # (single_frame.py:182)
            path = os.path.join(vdir, '{}.mask.png'.format(f))
            if os.path.exists(path):
                mask[row0 + j] = fit(cv2.imread(path))[:, :, 0]
    for a in (rgb, flow, mask):
        if a is not None:
            a.flush()
    return len(frames)


def pack_crop_dir(crop_dir, out_prefix, flow_img=None, img_dim=128, with_mask=False, nested=False,
                  workers=None, chunk_frames=256):
    """Decode every `<video>/<n>.png` (and flow / mask PNG) once and write the shard files.
    nested: the tennis layout `<video>/<player>/<n>.png`; shard videos are `<video>/<player>`.

    Streaming and parallel: the directory is scanned for names only, the shard files are
    created at their final size as memory maps, and `workers` processes (default: all cores)
    each decode chunks of `chunk_frames` frames straight into their rows - the corpus is never
    held in RAM (the page cache writes rows back as they fill) and PNG decoding, the whole
    cost (~0.3 ms per 128x128 PNG and core), scales with the core count."""
    from .apply import scan_crop_dir
    listing = scan_crop_dir(crop_dir, nested)
    n = sum(len(frames) for _, frames in listing)
    shape = (n, img_dim, img_dim)
    np.lib.format.open_memmap(out_prefix + '.rgb.npy', mode='w+', dtype=np.uint8, shape=shape + (3,)).flush()
    if flow_img:
        np.lib.format.open_memmap(out_prefix + '.flow.npy', mode='w+', dtype=np.uint8,
                                  shape=shape + (3,)).flush()
    if with_mask:
        np.lib.format.open_memmap(out_prefix + '.mask.npy', mode='w+', dtype=np.uint8, shape=shape).flush()
    index = {'img_dim': img_dim, 'flow_img': flow_img, 'with_mask': bool(with_mask), 'videos': []}
    tasks, row = [], 0
    for name, frames in listing:
        index['videos'].append({'name': name, 'frames': [int(f) for f in frames], 'first_row': row})
        for lo in range(0, len(frames), chunk_frames):
            part = frames[lo:lo + chunk_frames]
            tasks.append((crop_dir, out_prefix, name, part, row + lo, flow_img, img_dim, with_mask))
        row += len(frames)
    workers = (os.cpu_count() or 1) if workers is None else workers
    if workers <= 1 or len(tasks) <= 1:
        done = sum(_decode_chunk(t) for t in tasks)
    else:
        import multiprocessing
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(min(workers, len(tasks)),
                                 mp_context=multiprocessing.get_context('fork')) as pool:
            done = sum(pool.map(_decode_chunk, tasks))
    assert done == n, (done, n)
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
