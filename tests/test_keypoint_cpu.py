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
