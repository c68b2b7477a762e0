"""`apply_vipe_model.py` as a whole: pose files in, teacher embedding pickles out.

Mirrors the host side of the reference's teacher apply script:
  * `normalize_2d_skeleton`  vipe_dataset/dataset_base.py:105-141 (hip-centred, torso-scaled
    COCO skeleton, optional mirror image, confidences shifted by -0.5, optional bone features),
    here for a whole batch of detections at once; same arithmetic element for element
    (bit-exact against the reference function, tests/test_keypoint_cpu.py);
  * `VideoDataset.__getitem__`  apply_vipe_model.py:72-130 (`coco_keypoints.json.gz` layout:
    [[frame_num, [[score, box, 17 x (x, y, conf)], ...]], ...], flat or nested directories);
  * `main`  apply_vipe_model.py:165-204: embed, group per frame, write `<video>.emb.pkl`.
The embedding itself runs on the GPU (`vpd_b200.keypoint.Keypoint_EmbeddingModel.embed`); all
detections of a video go through it in one call instead of chunks of 250.
"""
import gzip
import json
import os

import numpy as np

from . import keypoint

COCO_POINTS_IDXS = [0] + list(range(5, 17))            # dataset_base.py:86-88: no eyes / ears
COCO_FLIP_IDXS = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
COCO_TORSO_POINTS = [5, 6, 11, 12]
_BONES_ORIG = [(a - 1, b - 1) for a, b in [
    (16, 14), (14, 12), (17, 15), (15, 13), (12, 13), (6, 12), (7, 13), (6, 7), (6, 8), (7, 9),
    (8, 10), (9, 11), (2, 3), (1, 2), (1, 3), (2, 4), (3, 5), (4, 6), (5, 7)]]
COCO_BONES = [x for x in _BONES_ORIG if x[0] in COCO_POINTS_IDXS and x[1] in COCO_POINTS_IDXS]


def normalize_2d_skeletons(kp, flip, include_bone_features=False, zero_confs=False):
    """kp float32 [n, 17, 3] (x, y, confidence) -> float32 [n, 13 (+12 bones), 3].
    `flip`: bool or bool [n] (mirror image: left/right joints swapped, x negated)."""
    kp = np.array(kp, dtype=np.float32, copy=True)
    n = kp.shape[0]
    flip = np.broadcast_to(np.asarray(flip, dtype=bool), (n,))
    kp[:, :, :2] -= ((kp[:, 11, :2] + kp[:, 12, :2]) / 2)[:, None, :]
    torso = kp[:, COCO_TORSO_POINTS, :2].astype(np.float64)        # pdist works in double
    dmax = np.zeros(n, dtype=np.float64)
    for a in range(4):
        for b in range(a + 1, 4):
            d = torso[:, a] - torso[:, b]
            dmax = np.maximum(dmax, np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]))
    dmax[dmax == 0] = 1                                             # prevent 0div
    kp[:, :, :2] *= (0.5 / dmax)[:, None, None]                     # double product, rounded to fp32
    if flip.any():
        f = kp[flip][:, COCO_FLIP_IDXS, :]
        f[:, :, 0] *= -1
        kp[flip] = f
    if zero_confs:
        kp[:, :, 2] = 0
    else:
        kp[:, :, 2] -= 0.5
    out = kp[:, COCO_POINTS_IDXS, :]
    if include_bone_features:
        bones = np.zeros((n, len(COCO_BONES), 3), dtype=np.float32)
        for i, (a, b) in enumerate(COCO_BONES):
            bones[:, i, :2] = kp[:, a, :2] - kp[:, b, :2]
            bones[:, i, 2] = (kp[:, a, 2] + kp[:, b, 2]) / 2
        out = np.concatenate((out, bones), axis=1)
    return np.ascontiguousarray(out, dtype=np.float32)


def list_videos(pose_dir):
    """apply_vipe_model.py:77-89: `<video>.json.gz` files or `<video>/coco_keypoints.json.gz`"""
    videos = []
    for name in sorted(os.listdir(pose_dir)):
        if name.endswith('.json.gz'):
            path, name = os.path.join(pose_dir, name), name.split('.json.gz')[0]
        else:
            path = os.path.join(pose_dir, name, 'coco_keypoints.json.gz')
        if os.path.exists(path):
            videos.append((name, path))
    return videos


def load_video_poses(path, embed_bones=False, min_score=0, augment_flip=True, invert=False):
    """apply_vipe_model.py:99-130 for one video -> dict(frame, score, is_flip, pose)"""
    with gzip.open(path, 'rt', encoding='ascii') as fp:
        data = json.load(fp)
    frames, kps = [], []
    for frame_num, pose_data in data:
        for score, _, kp in pose_data:
            if score >= min_score:
                frames.append(frame_num)
                kps.append(np.array(kp, dtype=np.float32))
    if not kps:
        return {'frame': np.array([]), 'score': np.array([]), 'is_flip': np.array([]),
                'pose': np.zeros(0)}
    kps = np.stack(kps)
    if invert:
        kps[:, :, 1] *= -1
    scores = np.mean(kps[:, :, 2], axis=1)
    k = 2 if augment_flip else 1
    is_flip = np.tile(np.arange(k) == 1, len(frames))               # [False, True] per detection
    pose = normalize_2d_skeletons(np.repeat(kps, k, axis=0), is_flip,
                                  include_bone_features=embed_bones)
    return {'frame': np.repeat(np.array(frames), k), 'score': np.repeat(scores, k),
            'is_flip': is_flip, 'pose': pose}


def apply_pose_dir(pose_dir, model_dir, out_dir, model_epoch=None, allow_many_per_frame=False,
                   min_score=0, no_flip=False, invert=False, model=None, log=print):
    """apply_vipe_model.main (:165-204). `model` (anything with `.embed`) may be passed in with
    `embed_bones` taken from the run's config.json; otherwise it is loaded from `model_dir`."""
    if model is None:
        model, embed_bones = keypoint.load_embedding_model(model_dir, model_epoch)
    else:
        with open(os.path.join(model_dir, 'config.json')) as fp:
            embed_bones = json.load(fp)['embed_bones']
    done = []
    for name, path in list_videos(pose_dir):
        v = load_video_poses(path, embed_bones, min_score, not no_flip, invert)
        embs = keypoint.embed_video(model, v['frame'], v['score'], v['is_flip'], v['pose'],
                                    flip=not no_flip, allow_many_per_frame=allow_many_per_frame)
        keypoint.write_embs(out_dir, name, embs)
        done.append((name, len(embs)))
        log('{}: {} entries'.format(name, len(embs)))
    return done
