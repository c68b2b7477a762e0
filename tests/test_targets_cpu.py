"""A13 (teacher-target construction): vpd_b200.targets against golden vectors produced by the
unmodified reference's GenericDataset.load_default (oracle/gen_golden.py::gen_targets)."""
import json
import os

import numpy as np
import pytest

from oracle.gen_golden import write_teacher_pickles
from vpd_b200 import targets


@pytest.fixture()
def sorted_listdir(monkeypatch):
    real = os.listdir
    monkeypatch.setattr(os, 'listdir', lambda d: sorted(real(d)))


CASES = [('motion', dict(embed_time=True)),
         ('plain_norm', dict(embed_time=False, normalize_target=True)),
         ('motion_norm_excl', dict(embed_time=True, normalize_target=True, min_pose_score=0.2,
                                   exclude_prefixes=('skip',)))]


@pytest.mark.parametrize('name,kw', CASES)
def test_targets_match_reference_golden(name, kw, tmp_path, golden_dir, sorted_listdir):
    gold = np.load(os.path.join(golden_dir, 'targets.npz'))
    with open(os.path.join(golden_dir, 'targets.json')) as fp:
        meta = json.load(fp)
    write_teacher_pickles(str(tmp_path))
    data, D = targets.load_teacher_targets(str(tmp_path), **kw)
    assert D == meta[name + '_emb_dim'] == 8
    np.random.seed(5)                                   # same global RNG state as the generator
    train, val = targets.split_train_val(data)
    for part, got in (('train', train), ('val', val)):
        assert [[d[0], int(d[1])] for d in got] == meta['{}_{}_keys'.format(name, part)]
        arr = targets.targets_array(got)
        ref = gold['{}_{}'.format(name, part)]
        assert arr.dtype == np.float32 and arr.shape == ref.shape
        assert np.array_equal(arr, ref)                 # bit-exact: same numpy operations
    # --motion doubles the last axis: [e, e - e_prev]
    if kw['embed_time']:
        assert targets.targets_array(train).shape[-1] == 2 * D


def test_targets_edge_cases(tmp_path):
    import pickle
    # empty directory / only foreign files -> no data, no emb_dim
    (tmp_path / 'readme.txt').write_text('x')
    data, D = targets.load_teacher_targets(str(tmp_path), True)
    assert data == [] and D is None
    # a video whose first frame has no predecessor and a gap: only frame 4 survives --motion
    e = [(2, np.ones((2, 4), np.float32), {'dp_score': 1.0}),
         (3, np.full((2, 4), 2.0, np.float32), {'dp_score': 0.1}),     # low score: dropped
         (4, np.full((2, 4), 5.0, np.float32), {'dp_score': 1.0}),
         (6, np.zeros((2, 4), np.float32), {'dp_score': 1.0})]         # gap: dropped
    with open(tmp_path / 'v.emb.pkl', 'wb') as fp:
        pickle.dump(e, fp)
    data, D = targets.load_teacher_targets(str(tmp_path), True)
    assert D == 4 and [(d[0], d[1]) for d in data] == [('v', 4)]
    # the difference is taken to the previous ENTRY even when that entry itself was filtered
    assert np.array_equal(data[0][2], np.concatenate([e[2][1], e[2][1] - e[1][1]], axis=1))
    with pytest.raises(AssertionError):                                 # inconsistent dims
        with open(tmp_path / 'w.emb.pkl', 'wb') as fp:
            pickle.dump([(0, np.ones((2, 5), np.float32), {'dp_score': 1.0})], fp)
        targets.load_teacher_targets(str(tmp_path), False)


@pytest.mark.parametrize('name,kw', [('tennis_motion', dict(embed_time=True)),
                                     ('tennis_plain', dict(embed_time=False, min_pose_score=0.2))])
def test_tennis_targets_match_reference_golden(name, kw, tmp_path, golden_dir, sorted_listdir):
    """TennisDataset.load_default: per player-and-clip pickles, (video, player, frame) keys"""
    from oracle.gen_golden import write_tennis_pickles
    gold = np.load(os.path.join(golden_dir, 'targets.npz'))
    with open(os.path.join(golden_dir, 'targets.json')) as fp:
        meta = json.load(fp)
    write_tennis_pickles(str(tmp_path))
    data, D = targets.load_teacher_targets_tennis(str(tmp_path), **kw)
    assert D == meta[name + '_emb_dim'] == 8
    np.random.seed(6)
    train, val = targets.split_train_val(data, key_len=3)
    for part, got in (('train', train), ('val', val)):
        assert [[d[0], d[1], int(d[2])] for d in got] == meta['{}_{}_keys'.format(name, part)]
        arr = targets.targets_array(got)
        assert np.array_equal(arr, gold['{}_{}'.format(name, part)])
    assert targets._tennis_key('front__match_b_set_2_7_30') == ('match_b_set_2', 'front', 7)
