"""CPU tests of the keypoint (VIPE*) teacher apply path: the oracle restatement against the
golden outputs of the unmodified reference classes, the initialisation, and the host logic
(`mean_embs_by_frame`, the pickle layout the student's target loader reads)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import keypoint_ref as K
from vpd_b200 import init, keypoint, targets

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'keypoint.npz')
CASES = [('d39', 39, 13, 1024, 2), ('d75', 75, 25, 256, 1)]


def _sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().numpy()).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('tag,in_dim,joints,hidden,blocks', CASES)
def test_init_and_oracle_match_the_reference(tag, in_dim, joints, hidden, blocks):
    gold = np.load(GOLD)
    seed = int(gold[tag + '_seed'])
    torch.manual_seed(seed)
    sd = init.fcresnet_state(in_dim, 32, blocks, hidden)
    assert _sd_hash(sd) == str(gold[tag + '_init_sha256'])       # same draws as nn.Module init
    sd = K.perturb_bn(sd, seed + 100)
    poses = K.synth_poses(96, seed + 200, joints)
    emb = K.embed(sd, poses, blocks)
    assert emb.dtype == np.float32 and emb.shape == (96, 32)
    np.testing.assert_allclose(emb, gold[tag + '_emb'], rtol=0, atol=2e-5)
    one = K.embed(sd, poses[3].numpy(), blocks)
    np.testing.assert_allclose(one, gold[tag + '_emb_one'], rtol=0, atol=2e-5)
    assert np.abs(gold[tag + '_emb']).max() > 0.05                # not a degenerate fixture


def test_mean_embs_by_frame_matches_the_reference():
    gold = np.load(GOLD)
    g = torch.Generator().manual_seed(9)
    embs = []
    for frame in (7, 3, 3, 11):
        for fl in (False, True):
            embs.append((frame, torch.randn(32, generator=g).numpy(),
                         {'kp_score': float(torch.rand(1, generator=g)), 'is_mean': False,
                          'is_flip': fl}))
    res = keypoint.mean_embs_by_frame(embs, True)
    assert [r[0] for r in res] == gold['mean_frames'].tolist()
    assert np.array_equal(np.stack([r[1] for r in res]), gold['mean_embs'])
    assert np.array_equal(np.array([r[2]['kp_score'] for r in res]), gold['mean_scores'])
    assert [r[2]['is_mean'] for r in res] == gold['mean_is_mean'].tolist()
    # no flip: one row per frame
    res1 = keypoint.mean_embs_by_frame([e for e in embs if not e[2]['is_flip']], False)
    assert [r[0] for r in res1] == [3, 7, 11] and res1[0][1].shape == (32,)


def test_embed_video_feeds_the_students_target_loader(tmp_path):
    """teacher apply -> `<video>.emb.pkl` -> targets.load_teacher_targets (A13)"""
    class FakeModel:                      # host logic only; the CUDA encoder is tested on the GPU
        def embed(self, pose):
            p = torch.as_tensor(np.asarray(pose)).reshape(len(pose), -1)
            return p[:, :8].numpy().astype(np.float32)
    frames = np.repeat(np.arange(10), 2)
    is_flip = np.tile([False, True], 10)
    scores = np.full(20, 0.9)
    poses = K.synth_poses(20, 1)
    embs = keypoint.embed_video(FakeModel(), frames, scores, is_flip, poses, flip=True)
    assert len(embs) == 10 and embs[0][1].shape == (2, 8) and embs[0][2]['kp_score'] == 0.9
    keypoint.write_embs(str(tmp_path), 'vid0', embs)
    data, emb_dim = targets.load_teacher_targets(str(tmp_path), embed_time=True)
    assert emb_dim == 8 and len(data) == 9          # the first frame has no predecessor
    assert data[0][2].shape == (2, 16)
    assert keypoint.embed_video(FakeModel(), [], [], [], np.zeros(0)) == []


def test_no_cpu_fallback():
    enc = keypoint.FCResNet(39, 32, 2, 128)
    with pytest.raises(Exception):
        enc.to('cpu')
    with pytest.raises(NotImplementedError):
        keypoint.FCResNet(39, None, 2, 128)
    assert list(enc.state_dict())[:2] == ['layers.0.weight', 'layers.0.bias']


def test_normalize_2d_skeletons_bit_exact_vs_reference():
    from vpd_b200 import keypoint_apply as KA
    gold = np.load(GOLD)
    kp = gold['skel_in']
    for fl in (False, True):
        for bones in (False, True):
            got = KA.normalize_2d_skeletons(kp, fl, include_bone_features=bones)
            want = gold['skel_f{}_b{}'.format(int(fl), int(bones))]
            assert got.dtype == np.float32 and got.shape == want.shape
            assert np.array_equal(got, want), (fl, bones, float(np.abs(got - want).max()))
    mixed = KA.normalize_2d_skeletons(kp, np.array([0, 1, 0, 1, 1, 0], bool))
    assert np.array_equal(mixed[1], gold['skel_f1_b0'][1]) and np.array_equal(mixed[2], gold['skel_f0_b0'][2])
    assert len(KA.COCO_BONES) == keypoint.NUM_COCO_BONES == 12


def test_apply_pose_dir_host_pipeline(tmp_path):
    """pose files (flat and nested layout) -> per-video pickles, with a stand-in model"""
    import gzip
    import json
    import pickle
    from vpd_b200 import keypoint_apply as KA
    gold = np.load(GOLD)
    kp = gold['skel_in']
    pose_dir, model_dir, out_dir = (os.path.join(str(tmp_path), d) for d in ('poses', 'model', 'out'))
    os.makedirs(os.path.join(pose_dir, 'nested'))
    os.makedirs(model_dir)
    with open(os.path.join(model_dir, 'config.json'), 'w') as fp:
        json.dump({'embed_bones': False, 'embedding_dim': 8, 'encoder_arch': [2, 128]}, fp)
    dets = [[3, [[0.9, [0, 0, 1, 1], kp[0].tolist()], [0.2, [0, 0, 1, 1], kp[1].tolist()]]],
            [1, [[0.8, [0, 0, 1, 1], kp[2].tolist()]]], [2, []]]
    for path in (os.path.join(pose_dir, 'flat.json.gz'),
                 os.path.join(pose_dir, 'nested', 'coco_keypoints.json.gz')):
        with gzip.open(path, 'wt', encoding='ascii') as fp:
            json.dump(dets, fp)

    class FakeModel:
        def embed(self, pose):
            return np.asarray(pose, dtype=np.float32).reshape(len(pose), -1)[:, :8].copy()
    v = KA.load_video_poses(os.path.join(pose_dir, 'flat.json.gz'), min_score=0.5)
    assert v['frame'].tolist() == [3, 3, 1, 1] and v['is_flip'].tolist() == [False, True] * 2
    assert np.array_equal(v['pose'][0], gold['skel_f0_b0'][0]) and np.array_equal(v['pose'][3], gold['skel_f1_b0'][2])
    assert abs(v['score'][0] - kp[0][:, 2].mean()) < 1e-7
    done = KA.apply_pose_dir(pose_dir, model_dir, out_dir, min_score=0.5, model=FakeModel(),
                             log=lambda *_: None)
    assert done == [('flat', 2), ('nested', 2)]
    with open(os.path.join(out_dir, 'nested.emb.pkl'), 'rb') as fp:
        embs = pickle.load(fp)
    assert [e[0] for e in embs] == [1, 3] and embs[0][1].shape == (2, 8) and embs[0][2]['is_mean'] is False
    assert np.array_equal(embs[1][1][1], gold['skel_f1_b0'][0].reshape(-1)[:8])
    # every detection kept, no flip
    done = KA.apply_pose_dir(pose_dir, model_dir, os.path.join(out_dir, 'all'), no_flip=True,
                             allow_many_per_frame=True, model=FakeModel(), log=lambda *_: None)
    assert done[0] == ('flat', 3)
    data, emb_dim = targets.load_teacher_targets(out_dir, embed_time=False, min_pose_score=0)
    assert emb_dim == 8 and len(data) == 4
