"""Corpus feature extraction - the B200 side of apply_vpd_model.py:92-179.

Per video: frames -> K1 assembly straight into the network's input layout (original
and horizontally flipped variant, flow-x negated on the flip) -> eval-mode encoder ->
`[(frame_num, np.float32 [2, D] (or [D] with no_flip), {}), ...]` sorted by frame and
pickled to `<out_dir>/<video>.emb.pkl` (util/io.py:35-37, README.md:185-194).

Multi-GPU: videos are partitioned across ranks (greedy longest-first), every rank
writes its own files, there is no collective. `extract_corpus` overlaps the host->device
copy of the next chunk, the forward pass and the per-video pickling (reader thread, CUDA
events, writer processes).
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


def plan_chunks(frame_counts, batch_size):
    """Cut the concatenation of the videos (in the given order) into chunks of exactly
    `batch_size` frames (the last one may be shorter). A chunk may span video boundaries, so
    every launch but the last runs the same full batch.
    -> [[(video_pos, lo, hi, offset_in_chunk), ...], ...] with lo/hi frame ranges inside the video."""
    chunks, cur, fill = [], [], 0
    for v, n in enumerate(frame_counts):
        lo = 0
        while lo < n:
            take = min(n - lo, batch_size - fill)
            cur.append((v, lo, lo + take, fill))
            fill += take
            lo += take
            if fill == batch_size:
                chunks.append(cur)
                cur, fill = [], 0
    if cur:
        chunks.append(cur)
    return chunks


def _write_video_pickle(path, frame_nums, embs, flip):
    """Body of a writer process: build the reference's tuples and pickle them."""
    store_pickle(path, format_video_embs(frame_nums, embs, flip))
    return path


class _PickleWriters:
    """Per-video pickling off the critical path. Formatting + pickling a 2,700-frame video is
    ~15 ms of pure Python (5.6 us per (frame, ndarray, {}) tuple), i.e. one core writes about as
    many frames per second as one B200 embeds; it runs in worker PROCESSES so that it neither
    caps the rate nor competes with the launching thread for the interpreter lock.
    workers = 0: write in the calling thread (tests, tiny corpora)."""

    def __init__(self, workers):
        self.pool, self.futures = None, []
        if workers > 0:
            import multiprocessing
            from concurrent.futures import ProcessPoolExecutor
            # fork: the children never touch CUDA (numpy + pickle only) and start in ~1 ms
            self.pool = ProcessPoolExecutor(workers, mp_context=multiprocessing.get_context('fork'))
            self.pool.submit(int).result()  # start the workers now, before any thread of ours runs

    def submit(self, path, frame_nums, embs, flip):
        if self.pool is None:
            _write_video_pickle(path, frame_nums, embs, flip)
        else:
            self.futures.append(self.pool.submit(_write_video_pickle, path, list(frame_nums),
                                                 np.ascontiguousarray(embs), flip))

    def close(self):
        try:
            for f in self.futures:
                f.result()                  # re-raises a worker's exception
        finally:
            if self.pool is not None:
                self.pool.shutdown()


def extract_corpus(model, videos, out_dir, rgb_mean_std, flip=True, rank=0, world_size=1,
                   batch_size=BATCH_SIZE, writers=2, timing=None):
    """videos: list of (name, frame_nums, rgb_u8 [n,H,W,3], flow_u8 or None) with host (ideally
    pinned or memory-mapped) or device uint8 tensors. Writes this rank's `<name>.emb.pkl`;
    returns the names written (apply_vpd_model.py:152-178).

    A three-stage pipeline over fixed-size chunks of the rank's frames (`plan_chunks`):
      reader thread   host frames -> pinned staging slot -> H2D copy on a copy stream into one
                      of the device input slots (98 KB per frame: 15 GB/s at 158 k frames/s);
      this thread     K1 assembly [orig, flipped] -> eval-mode encoder (one CUDA-graph replay
                      per chunk) -> D2H of the chunk's [n, k, D] embeddings into a pinned array;
      writer procs    per finished video: tuples + pickle.
    The stages meet only through CUDA events and two queues, so the copy of chunk i+1 and the
    pickling of earlier videos overlap the forward pass of chunk i. `timing` (dict) receives
    frames / seconds / chunks of this rank."""
    import queue
    import threading
    import time
    from ._lib import lib
    from .assemble import assemble_stem
    t_start = time.perf_counter()
    dev = model._dev
    os.makedirs(out_dir, exist_ok=True)
    mine = [i for i in shard_videos([len(v[1]) for v in videos], world_size, rank)
            if len(videos[i][1]) > 0]
    if not mine:
        return []
    counts = [len(videos[i][1]) for i in mine]
    chunks = plan_chunks(counts, batch_size)
    total = sum(counts)
    k = 2 if flip else 1
    first = videos[mine[0]]
    H, W = first[2].shape[1:3]
    has_flow = first[3] is not None
    fch = first[3].shape[-1] if has_flow else 0
    D = model.emb_dim
    model.eval()
    n_in, n_pin = 2, 3
    # pinned staging slots only when some source is not pinned already (page-locking memory
    # costs ~10 ms per 100 MB: a tenth of a second for a corpus that needs no staging at all)
    need_staging = any(not videos[i][2].is_pinned() for i in mine)
    with torch.cuda.device(dev):
        pin_rgb = [torch.empty((batch_size, H, W, 3), dtype=torch.uint8).pin_memory()
                   for _ in range(n_pin)] if need_staging else None
        pin_flow = [torch.empty((batch_size, H, W, fch), dtype=torch.uint8).pin_memory()
                    for _ in range(n_pin)] if (has_flow and need_staging) else None
        dev_rgb = [torch.empty((batch_size, H, W, 3), device=dev, dtype=torch.uint8)
                   for _ in range(n_in)]
        dev_flow = [torch.empty((batch_size, H, W, fch), device=dev, dtype=torch.uint8)
                    for _ in range(n_in)] if has_flow else None
        dev_out = [torch.empty((batch_size * k, D), device=dev, dtype=torch.float32)
                   for _ in range(2)]
        host_out = torch.empty((total, k, D), dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
    ready = queue.Queue()                    # (chunk index, H2D-done event) or an exception
    consumed = [None] * len(chunks)          # event: the compute that read device slot is done
    consumed_cv = threading.Condition()

    def reader():
        try:
            torch.cuda.set_device(dev)
            pin_free = [None] * n_pin        # event of the H2D that last read the staging slot
            for ci, parts in enumerate(chunks):
                ps, ds = ci % n_pin, ci % n_in
                # frames that already sit in pinned memory go to the device straight from
                # there; pageable / memory-mapped ones pass through a pinned staging slot first
                staged = [not videos[mine[v]][2].is_pinned() for v, _, _, _ in parts]
                if any(staged):
                    if pin_free[ps] is not None:
                        pin_free[ps].synchronize()
                    for (v, lo, hi, off), st in zip(parts, staged):
                        if not st:
                            continue
                        _, _, rgb, flow = videos[mine[v]]
                        pin_rgb[ps][off:off + hi - lo].copy_(rgb[lo:hi])
                        if has_flow:
                            pin_flow[ps][off:off + hi - lo].copy_(flow[lo:hi])
                if ci >= n_in:               # the device slot's previous chunk must be consumed
                    with consumed_cv:
                        consumed_cv.wait_for(lambda: consumed[ci - n_in] is not None)
                    copy_stream.wait_event(consumed[ci - n_in])
                with torch.cuda.stream(copy_stream):
                    for (v, lo, hi, off), st in zip(parts, staged):
                        _, _, rgb, flow = videos[mine[v]]
                        k_ = hi - lo
                        src_r = pin_rgb[ps][off:off + k_] if st else rgb[lo:hi]
                        dev_rgb[ds][off:off + k_].copy_(src_r, non_blocking=True)
                        if has_flow:
                            src_f = pin_flow[ps][off:off + k_] if st else flow[lo:hi]
                            dev_flow[ds][off:off + k_].copy_(src_f, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                if any(staged):
                    pin_free[ps] = ev
                ready.put((ci, ev))
        except BaseException as exc:         # surface in the consuming thread
            ready.put(exc)

    th = threading.Thread(target=reader, name='vpd-corpus-reader', daemon=True)
    th.start()
    writer = _PickleWriters(writers)
    done_ev = [None] * len(chunks)
    last_chunk_of = {}
    for ci, parts in enumerate(chunks):
        for v, _, _, _ in parts:
            last_chunk_of[v] = ci
    first_row = np.concatenate([[0], np.cumsum(counts)])
    next_video = 0
    written = []

    def flush(upto_chunk, block):
        """hand every video whose last chunk is <= upto_chunk and finished to the writers"""
        nonlocal next_video
        while next_video < len(mine) and last_chunk_of[next_video] <= upto_chunk:
            ev = done_ev[last_chunk_of[next_video]]
            if block:
                ev.synchronize()
            elif not ev.query():
                return
            name, frame_nums = videos[mine[next_video]][0], videos[mine[next_video]][1]
            lo, hi = first_row[next_video], first_row[next_video + 1]
            writer.submit(os.path.join(out_dir, '{}.emb.pkl'.format(name)), frame_nums,
                          host_out[lo:hi].numpy(), flip)
            written.append(name)
            next_video += 1

    try:
        with torch.cuda.device(dev):
            row = 0
            for _ in range(len(chunks)):
                item = ready.get()
                if isinstance(item, BaseException):
                    raise item
                ci, ev = item
                parts = chunks[ci]
                n = sum(hi - lo for _, lo, hi, _ in parts)
                ds, os_ = ci % n_in, ci % 2
                main_stream.wait_event(ev)
                net = model._native(H, W, batch_size * k)
                stem = lib().call('vpd_net_stem_input', net.handle)
                assemble_stem(stem, dev_rgb[ds][:n], dev_flow[ds][:n] if has_flow else None,
                              rgb_mean_std, k=k)
                lib().call('vpd_net_forward', net.handle, None, stem, n * k, dev_out[os_],
                           main_stream.cuda_stream)
                cev = torch.cuda.Event()
                cev.record(main_stream)
                with consumed_cv:
                    consumed[ci] = cev
                    consumed_cv.notify_all()
                host_out[row:row + n].copy_(dev_out[os_][:n * k].view(n, k, D), non_blocking=True)
                dev_ev = torch.cuda.Event()
                dev_ev.record(main_stream)
                done_ev[ci] = dev_ev
                row += n
                flush(ci - 1, block=False)
            flush(len(chunks) - 1, block=True)
    finally:
        th.join(timeout=60)
        writer.close()
    if timing is not None:
        timing.update(frames=total, chunks=len(chunks), videos=len(written),
                      seconds=time.perf_counter() - t_start)
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


def scan_crop_dir(crop_dir, nested=False):
    """[(video name, sorted frame numbers)] of `<crop_dir>/<video>/<n>.png` without decoding
    anything. nested: the tennis layout `<crop_dir>/<video>/<player>/<n>.png`
    (single_frame.py:51-57); the entries are then named `<video>/<player>`."""
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
    out = []
    for video_name in names:
        vdir = os.path.join(crop_dir, video_name)
        out.append((video_name, sorted(int(os.path.splitext(f)[0]) for f in os.listdir(vdir)
                                       if img_re.match(f))))
    return out


def read_crop_dir(crop_dir, flow_img=None, img_dim=128, nested=False):
    """`<crop_dir>/<video>/<n>.png` (+ `<n>.<flow_img>.png`) -> videos list for
    extract_corpus (apply_vpd_model.py:69-89), every frame decoded into host memory (cv2).
    For corpora that do not fit in RAM pack them once with vpd_b200.ingest.pack_crop_dir
    (parallel, streaming) and hand `load_shard(...).videos()` to extract_corpus instead."""
    videos = []
    for video_name, frames in scan_crop_dir(crop_dir, nested):
        rgb, flow = _read_frames(os.path.join(crop_dir, video_name), frames, flow_img, img_dim)
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
