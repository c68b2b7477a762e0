"""Teacher-target construction for student training.

Mirror of `GenericDataset.load_default` (reference vpd_dataset/single_frame.py:208-273), minus
the dataset objects it wraps the result in: reads the per-video teacher pickles
`<video>.emb.pkl` = [(frame_num, emb [2, D] (rows: as is / flipped) or [D], meta), ...]
(README.md:185-194), drops low-confidence poses, optionally row-normalises, with
`embed_time` (`--motion`) appends the temporal difference to the previous frame so the target
is [2, 2D], and splits 80/20 with sklearn's `train_test_split` exactly like the reference
(so the same numpy global seed gives the same split). Pure host logic: numpy only.
"""
import os
import pickle

import numpy as np

EMB_FILE_SUFFIX = '.emb.pkl'          # vpd_dataset/common.py:9
DEFAULT_MIN_POSE_SCORE = 0.5          # vpd_dataset/single_frame.py:17


def _normalize_rows(x):               # single_frame.py:28-32
    if len(x.shape) == 1:
        return x / np.linalg.norm(x)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _get_pose_score(meta, default=None):   # single_frame.py:35-45
    score = meta.get('dp_score')
    if score is not None:
        return score
    score = meta.get('kp_score')
    if score is not None:
        return score
    if default is not None:
        return default
    raise ValueError('Missing pose score in {}'.format(meta))


def load_teacher_targets(emb_dir, embed_time, min_pose_score=None, normalize_target=False,
                         exclude_prefixes=None):
    """-> (all_data, emb_dim): all_data = [(video, frame_num, target ndarray, meta), ...] in the
    reference's order (os.listdir order of the pickles, then frame order); emb_dim is the
    teacher's D (the target's last axis is 2D with embed_time)."""
    all_data = []
    emb_dim = None
    for emb_file in os.listdir(emb_dir):
        if not emb_file.endswith(EMB_FILE_SUFFIX):
            continue
        video_name = emb_file.split(EMB_FILE_SUFFIX)[0]
        if exclude_prefixes is not None and video_name.startswith(exclude_prefixes):
            continue
        with open(os.path.join(emb_dir, emb_file), 'rb') as fp:
            video_embs = pickle.load(fp)
        for i in range(len(video_embs)):
            frame_num, emb_target, emb_meta = video_embs[i]
            if emb_dim is not None:
                assert emb_target.shape[-1] == emb_dim, \
                    'Inconsistent emb dims {} != {}'.format(emb_target.shape[-1], emb_dim)
            else:
                emb_dim = emb_target.shape[-1]
            thresh = DEFAULT_MIN_POSE_SCORE if min_pose_score is None else min_pose_score
            if _get_pose_score(emb_meta) < thresh:
                continue
            if normalize_target:
                emb_target = _normalize_rows(emb_target)
            if embed_time:
                # needs the embedding of the frame right before
                if i == 0 or video_embs[i - 1][0] != frame_num - 1:
                    continue
                emb_prev = video_embs[i - 1][1]
                if normalize_target:
                    emb_prev = _normalize_rows(emb_prev)
                emb_target = np.concatenate(
                    [emb_target, emb_target - emb_prev],
                    axis=0 if len(emb_target.shape) == 1 else 1)
            all_data.append((video_name, frame_num, emb_target, emb_meta))
    return all_data, emb_dim


def _tennis_key(file_stem):
    """`<player>__<video>_<start>_<end>` (one pickle per player and clip) -> (video, player, start)"""
    player, rest = file_stem.split('__', 1)
    video_name, start_frame, _ = rest.rsplit('_', 2)
    return video_name, player, int(start_frame)


def load_teacher_targets_tennis(emb_dir, embed_time, min_pose_score=None, normalize_target=False,
                                exclude_prefixes=None):
    """`TennisDataset.load_default` (single_frame.py:90-163): same filtering / motion targets as
    `load_teacher_targets`, but the pickles are per player and clip and the entries become
    (video, player, start_frame + frame_num, target, meta) - the crop of that frame lives in
    `<img_dir>/<video>/<player>/<frame>.png` (:51-57). The reference loads every pickle first
    and takes emb_dim from each file's first entry; the order of the result is the same."""
    all_data = []
    emb_dim = None
    for emb_file in os.listdir(emb_dir):
        if not emb_file.endswith(EMB_FILE_SUFFIX):
            continue
        stem = emb_file.split(EMB_FILE_SUFFIX)[0]
        if exclude_prefixes is not None and stem.startswith(exclude_prefixes):
            continue
        with open(os.path.join(emb_dir, emb_file), 'rb') as fp:
            video_embs = pickle.load(fp)
        if emb_dim is None:
            emb_dim = video_embs[0][1].shape[-1]
        else:
            assert emb_dim == video_embs[0][1].shape[-1]
        video_name, player, start_frame = _tennis_key(stem)
        thresh = DEFAULT_MIN_POSE_SCORE if min_pose_score is None else min_pose_score
        for i, (frame_num, emb_target, emb_meta) in enumerate(video_embs):
            if _get_pose_score(emb_meta) < thresh:
                continue
            if normalize_target:
                emb_target = _normalize_rows(emb_target)
            if embed_time:
                if i == 0 or video_embs[i - 1][0] != frame_num - 1:
                    continue
                emb_prev = video_embs[i - 1][1]
                if normalize_target:
                    emb_prev = _normalize_rows(emb_prev)
                emb_target = np.concatenate(
                    [emb_target, emb_target - emb_prev],
                    axis=0 if len(emb_target.shape) == 1 else 1)
            all_data.append((video_name, player, start_frame + frame_num, emb_target, emb_meta))
    return all_data, emb_dim


def split_train_val(all_data, test_size=0.2, key_len=2):
    """80/20 split like the reference (`train_test_split` on the numpy global RNG, then
    sort). Sorting tuples that contain ndarrays only works while (video, frame) pairs are
    unique - the reference has the same precondition."""
    from sklearn.model_selection import train_test_split
    train_data, val_data = train_test_split(all_data, test_size=test_size)
    train_data.sort(key=lambda x: x[:key_len])         # Tennis entries sort on x[:3] (:150)
    val_data.sort(key=lambda x: x[:key_len])
    return train_data, val_data


def targets_array(data):
    """Stack the targets of a data list into one fp32 array [n, 2, Dt] (or [n, Dt]) - the
    `teacher` operand of vpd_b200.assemble (row 0: unflipped crop, row 1: flipped)."""
    col = 3 if data and isinstance(data[0][1], str) else 2      # Tennis tuples carry the player
    return np.stack([np.asarray(d[col], dtype=np.float32) for d in data]) if data else \
        np.zeros((0,), np.float32)
