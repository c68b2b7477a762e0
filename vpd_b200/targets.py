"""Teacher-target construction for student training.

What the reference does per teacher pickle inside `GenericDataset.load_default` /
`TennisDataset.load_default` (vpd_dataset/single_frame.py:208-273, :90-163), done here per
VIDEO with array operations instead of per frame:

  pickle `<video>.emb.pkl` = [(frame_num, emb [2, D] (rows: as is / flipped) or [D], meta)]
  (README.md:185-194)  ->  frames [n], embeddings [n, (2,) D], scores [n]
  keep      = score >= threshold                      (low-confidence poses dropped)
  embedding = embedding / |embedding|  per row        (normalize_target, off on the CLI)
  --motion:   keep &= frame[i] == frame[i-1] + 1      (needs the frame right before, whether
              target = [e_i, e_i - e_(i-1)]            or not that one passed the score test)

The arithmetic per element is the reference's (same numpy ufuncs on the same dtype), so the
targets are bit-identical to its output (tests/test_targets_cpu.py, golden from the
unmodified reference). The 80/20 split uses sklearn's `train_test_split` on the numpy global
RNG like the reference, so the same seed gives the same split. Host logic, numpy only.
"""
import os
import pickle

import numpy as np

EMB_FILE_SUFFIX = '.emb.pkl'          # vpd_dataset/common.py:9
DEFAULT_MIN_POSE_SCORE = 0.5          # vpd_dataset/single_frame.py:17
_SCORE_KEYS = ('dp_score', 'kp_score')   # DensePose score first, keypoint score second (:35-45)


def _pose_scores(metas):
    """fp64 [n]: the pose-confidence of every frame of a video; a frame without one is an error."""
    out = np.empty(len(metas), np.float64)
    for i, meta in enumerate(metas):
        score = next((meta[k] for k in _SCORE_KEYS if meta.get(k) is not None), None)
        if score is None:
            raise ValueError('Missing pose score in {}'.format(meta))
        out[i] = score
    return out


def _unit_rows(embs):
    """Row-normalise a stack of embeddings with the rounding of `x / np.linalg.norm(x, ...)`:
    [n, 2, D] stacks reduce over the last axis exactly like the reference's axis=1 call on one
    [2, D] entry; flat [n, D] stacks (1-D entries) go through the dot-product norm numpy uses
    for vectors, one row at a time."""
    if embs.ndim == 3:
        return embs / np.linalg.norm(embs, axis=2, keepdims=True)
    norms = np.array([np.linalg.norm(row) for row in embs], dtype=embs.dtype)
    return embs / norms[:, None]


def _video_targets(entries, embed_time, min_pose_score, normalize_target):
    """entries = one pickle's list -> (indices of the surviving frames, their targets)."""
    frames = np.array([e[0] for e in entries], dtype=np.int64)
    embs = np.stack([np.asarray(e[1]) for e in entries])
    keep = _pose_scores([e[2] for e in entries]) >= (
        DEFAULT_MIN_POSE_SCORE if min_pose_score is None else min_pose_score)
    if normalize_target:
        embs = _unit_rows(embs)
    if embed_time:
        consecutive = np.zeros(len(frames), dtype=bool)
        consecutive[1:] = frames[1:] == frames[:-1] + 1
        keep &= consecutive
        delta = np.zeros_like(embs)
        delta[1:] = embs[1:] - embs[:-1]
        embs = np.concatenate([embs, delta], axis=-1)
    idx = np.flatnonzero(keep)
    return idx, embs[idx]


def _teacher_pickles(emb_dir, exclude_prefixes):
    """(file stem, entries) of every teacher pickle, in os.listdir order like the reference."""
    for name in os.listdir(emb_dir):
        if not name.endswith(EMB_FILE_SUFFIX):
            continue
        stem = name.split(EMB_FILE_SUFFIX)[0]
        if exclude_prefixes is not None and stem.startswith(exclude_prefixes):
            continue
        with open(os.path.join(emb_dir, name), 'rb') as fp:
            yield stem, pickle.load(fp)


def _collect(emb_dir, embed_time, min_pose_score, normalize_target, exclude_prefixes, make_row):
    rows, emb_dim = [], None
    for stem, entries in _teacher_pickles(emb_dir, exclude_prefixes):
        dims = {np.asarray(e[1]).shape[-1] for e in entries}
        if emb_dim is None and dims:
            emb_dim = np.asarray(entries[0][1]).shape[-1]
        assert dims <= {emb_dim}, 'Inconsistent emb dims {} != {}'.format(sorted(dims), emb_dim)
        if not entries:
            continue
        idx, tgt = _video_targets(entries, embed_time, min_pose_score, normalize_target)
        rows.extend(make_row(stem, entries[i], tgt[j]) for j, i in enumerate(idx))
    return rows, emb_dim


def load_teacher_targets(emb_dir, embed_time, min_pose_score=None, normalize_target=False,
                         exclude_prefixes=None):
    """-> (all_data, emb_dim): all_data = [(video, frame_num, target ndarray, meta), ...] in the
    reference's order (os.listdir order of the pickles, then frame order); emb_dim is the
    teacher's D (the target's last axis is 2D with embed_time)."""
    return _collect(emb_dir, embed_time, min_pose_score, normalize_target, exclude_prefixes,
                    lambda video, entry, tgt: (video, entry[0], tgt, entry[2]))


def _tennis_key(file_stem):
    """`<player>__<video>_<start>_<end>` (one pickle per player and clip) -> (video, player, start)"""
    player, rest = file_stem.split('__', 1)
    video_name, start_frame, _ = rest.rsplit('_', 2)
    return video_name, player, int(start_frame)


def load_teacher_targets_tennis(emb_dir, embed_time, min_pose_score=None, normalize_target=False,
                                exclude_prefixes=None):
    """The Tennis layout of the same data (single_frame.py:90-163): one pickle per player and
    clip, entries become (video, player, start_frame + frame_num, target, meta) - the crop of
    that frame lives in `<img_dir>/<video>/<player>/<frame>.png` (:51-57)."""
    def row(stem, entry, tgt):
        video, player, start = _tennis_key(stem)
        return (video, player, start + entry[0], tgt, entry[2])
    return _collect(emb_dir, embed_time, min_pose_score, normalize_target, exclude_prefixes, row)


def split_train_val(all_data, test_size=0.2, key_len=2):
    """80/20 split like the reference (`train_test_split` on the numpy global RNG), each part
    ordered by its (video, frame) - Tennis: (video, player, frame) - key."""
    from sklearn.model_selection import train_test_split
    parts = train_test_split(all_data, test_size=test_size)
    return tuple(sorted(part, key=lambda row: row[:key_len]) for part in parts)


def targets_array(data):
    """Stack the targets of a data list into one fp32 array [n, 2, Dt] (or [n, Dt]) - the
    `teacher` operand of vpd_b200.assemble (row 0: unflipped crop, row 1: flipped)."""
    if not data:
        return np.zeros((0,), np.float32)
    col = 3 if isinstance(data[0][1], str) else 2      # Tennis tuples carry the player
    return np.stack([np.asarray(row[col], dtype=np.float32) for row in data])
