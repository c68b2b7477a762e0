"""Host logic of vpd_b200.train.fit (mirror of train_vpd_model.main's epoch loop):
model selection on the moving average, checkpoint names, config.json / loss.json."""
import json
import os

import pytest

from vpd_b200 import train


class FakeTrainer:
    def __init__(self, train_losses, val_losses):
        self.t, self.v = list(train_losses), list(val_losses)
        self.saved = []
        self.calls = 0

    def epoch(self, loader, optimizer=None, scaler=None, progress_cb=None):
        self.calls += 1
        return self.t.pop(0) if optimizer is not None else self.v.pop(0)

    def save_model(self, save_dir, name):
        self.saved.append(name)
        open(os.path.join(save_dir, name + '.encoder.pt'), 'w').close()


CFG = {'num_epochs': 4, 'batch_size': 8, 'learning_rate': 5e-4, 'img_dim': 128, 'use_flow': True,
       'motion': True, 'emb_dim': 32, 'encoder_arch': 'resnet34',
       'rgb_mean_std': [[0.5, 0.5, 0.5], [0.2, 0.2, 0.2]], 'extra': 'not written'}


def test_fit_selection_checkpoints_and_files(tmp_path):
    out = str(tmp_path / 'run')
    # window 2: moving averages of val = 4, 3.5, 3.5, 2.5 -> best at epochs 1, 2, 4
    tr = FakeTrainer([9, 8, 7, 6], [4, 3, 4, 1])
    hist = train.fit(tr, [1], [1], out, CFG, 4, optimizer=object(), model_select_window=2,
                     checkpoint_frequency=2, log=lambda *a: None)
    assert tr.saved == ['best_epoch', 'best_epoch', 'epoch0002', 'best_epoch', 'epoch0004',
                        'epoch0004']
    assert [h['epoch'] for h in hist] == [1, 2, 3, 4] and hist[2]['val'] == 4
    cfg = json.load(open(os.path.join(out, 'config.json')))
    assert list(cfg.keys()) == list(train.CONFIG_KEYS) and 'extra' not in cfg
    loss = json.load(open(os.path.join(out, 'loss.json')))
    assert loss[-1]['train'] == 6 and loss[-1]['dataset_val'] == [['synthetic', 1]]
    with pytest.raises(FileExistsError):            # like the reference: never overwrite a run
        train.fit(FakeTrainer([1], [1]), [1], [1], out, CFG, 1, optimizer=object())


def test_fit_without_validation_and_bad_config(tmp_path):
    tr = FakeTrainer([3, 2], [])
    hist = train.fit(tr, [1], None, str(tmp_path / 'r'), CFG, 2, optimizer=object(),
                     log=lambda *a: None)
    # nan moving average never "improves": only the last-epoch checkpoint is written
    assert tr.saved == ['epoch0002'] and hist[0]['val'] != hist[0]['val']
    bad = dict(CFG)
    del bad['emb_dim']
    with pytest.raises(AssertionError):
        train.fit(tr, [1], None, str(tmp_path / 'r2'), bad, 1, optimizer=object())


@pytest.mark.parametrize('dataset', ['generic', 'tennis'])
def test_main_wires_targets_shard_rows_and_loaders(tmp_path, dataset):
    """train.main's data plumbing with the GPU-side constructors replaced: teacher pickles ->
    targets -> split -> shard rows -> pools aligned with their targets; loader lengths; files."""
    import pickle
    import cv2
    import numpy as np
    from vpd_b200 import ingest
    rng = np.random.RandomState(1)
    crop_dir, emb_dir = os.path.join(str(tmp_path), 'crops'), os.path.join(str(tmp_path), 'embs')
    os.makedirs(emb_dir)
    if dataset == 'tennis':
        vdirs = {'match_a/front': ('front__match_a_100_140', 100), 'match_a/back': ('back__match_a_100_140', 100)}
    else:
        vdirs = {'clipA': ('clipA', 0), 'clipB': ('clipB', 0)}
    for vi, (vdir, (stem, start)) in enumerate(vdirs.items()):
        os.makedirs(os.path.join(crop_dir, vdir))
        embs = []
        for f in range(12):
            img = np.full((8, 8, 3), 10 * vi + f, np.uint8)         # pixel value identifies the frame
            cv2.imwrite(os.path.join(crop_dir, vdir, '{}.png'.format(start + f)), img)
            cv2.imwrite(os.path.join(crop_dir, vdir, '{}.flow.png'.format(start + f)), img)
            embs.append((f, np.full((2, 4), 10 * vi + f, np.float32), {'dp_score': 0.9}))
        with open(os.path.join(emb_dir, stem + '.emb.pkl'), 'wb') as fp:
            pickle.dump(embs, fp)
    prefix = os.path.join(str(tmp_path), 'shard')
    ingest.pack_crop_dir(crop_dir, prefix, flow_img='flow', img_dim=8, nested=dataset == 'tennis')
    seen = {}

    def fake_pools(shard, rows):
        return np.asarray(shard.rgb)[rows], np.asarray(shard.flow)[rows], None

    def fake_loader(rgb, flow, mask, teach, length):
        # every pool frame sits next to ITS teacher target
        assert rgb.shape[0] == teach.shape[0] and teach.shape[1:] == (2, 4)
        assert np.array_equal(rgb[:, 0, 0, 0].astype(np.float32), teach[:, 0, 0])
        seen.setdefault('lengths', []).append(length)
        seen.setdefault('sizes', []).append(rgb.shape[0])
        return ['loader', length]

    tr = FakeTrainer([3.0, 2.0], [3.0, 2.5])
    np.random.seed(0)
    hist = train.main(emb_dir, prefix, os.path.join(str(tmp_path), 'run'), CFG['rgb_mean_std'],
                      dataset=dataset, num_epochs=2, batch_size=4, motion=False, target_len=100,
                      checkpoint_frequency=None, model_select_window=1, log=lambda *a: None,
                      _factories={'pools': fake_pools, 'loader': fake_loader,
                                  'trainer': lambda D, flow: (tr, object(), None)})
    assert [h['epoch'] for h in hist] == [1, 2] and seen['lengths'] == [100, 20]
    assert sorted(seen['sizes']) == [5, 19]                          # 24 frames, 80 / 20 split
    with open(os.path.join(str(tmp_path), 'run', 'config.json')) as fp:
        cfg = json.load(fp)
    assert cfg['emb_dim'] == 4 and cfg['use_flow'] is True and cfg['motion'] is False
    assert tr.saved == ['best_epoch', 'best_epoch', 'epoch0002']


def test_pool_loader_raw_yields_the_draw_itself():
    """PoolLoader(raw=True): index batches over the pools (no kernel call, so it runs anywhere);
    flips only with two-row teachers and random_flip, the epoch length of the reference's
    `_TrainDataset`, and the fallback to assembled batches when augmentation / mask noise is on."""
    import torch
    from vpd_b200.train import PoolLoader
    P, B = 10, 4
    rgb = torch.zeros((P, 8, 8, 3), dtype=torch.uint8)
    flow = torch.zeros((P, 8, 8, 3), dtype=torch.uint8)
    teach2 = torch.zeros((P, 2, 6))
    ms = ((0.5, 0.5, 0.5), (0.2, 0.2, 0.2))
    ld = PoolLoader(rgb, flow, teach2, ms, B, 10, seed=1, random_flip=True, raw=True)
    batches = list(ld)
    assert len(batches) == len(ld) == 3 and [b['index'].numel() for b in batches] == [4, 4, 2]
    for b in batches:
        assert b['index'].dtype == torch.int32 and int(b['index'].min()) >= 0 and int(b['index'].max()) < P
        assert b['flip'].dtype == torch.uint8 and b['flip'].shape == b['index'].shape
        assert b['rgb_u8'] is rgb and b['flow_u8'] is flow and b['teacher'] is teach2
        assert b['rgb_mean_std'] == ms
    # one-row teachers have no flipped counterpart; without random_flip nothing is flipped
    assert all(b['flip'] is None for b in PoolLoader(rgb, flow, torch.zeros((P, 6)), ms, B, 8,
                                                     random_flip=True, raw=True))
    assert all(b['flip'] is None for b in PoolLoader(rgb, flow, teach2, ms, B, 8, raw=True))
    # same seed, same draws as the reference-format loader would make
    a = [b['index'].tolist() for b in PoolLoader(rgb, flow, teach2, ms, B, 10, seed=7, raw=True)]
    c = [b['index'].tolist() for b in PoolLoader(rgb, flow, teach2, ms, B, 10, seed=7, raw=True)]
    assert a == c
    assert PoolLoader(rgb, flow, teach2, ms, B, 8, raw=True, augment=True).raw is False
    assert PoolLoader(rgb, flow, teach2, ms, B, 8, raw=True,
                      mask_u8=torch.zeros((P, 8, 8), dtype=torch.uint8)).raw is False
