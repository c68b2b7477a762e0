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
