"""CPU: the keypoint-teacher training oracle (oracle/keypoint_train_ref.py) against the golden
outputs of the unmodified reference `Keypoint_EmbeddingModel.epoch` (losses and first-step
gradients under the dropout masks the reference drew), the parameter containers' initialisation,
and the loaders' zipping logic."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import keypoint_train_ref as T
from vpd_b200 import init, keypoint
from vpd_b200.keypoint_train import FCPoseDecoder, _Arena

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'keypoint_train.npz')
H, BLOCKS, N1, N2, P = 128, 2, 136, 72, 0.2


def _sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().numpy()).tobytes())
    return h.hexdigest()


def initial_state():
    torch.manual_seed(31)
    enc_sd = init.fcresnet_state(39, 32, BLOCKS, H)
    dec = FCPoseDecoder(32, [128, 128], [('h36m', 140)])
    return enc_sd, dec


def unpack_masks(gold, step):
    out = []
    for di, (n, passes) in enumerate(((N1, 3), (N2, 2))):
        bits = np.unpackbits(gold['step{}_masks{}'.format(step, di)])[:passes * 2 * BLOCKS * n * H]
        m = torch.from_numpy(bits.reshape(passes, 2 * BLOCKS, n, H).astype(np.uint8))
        out.append([[m[a, b] for b in range(2 * BLOCKS)] for a in range(passes)])
    return out


def test_initialisation_matches_reference_modules():
    gold = np.load(GOLD)
    enc_sd, dec = initial_state()
    assert _sd_hash(enc_sd) == str(gold['enc_init_sha256'])
    assert _sd_hash(dec.state_dict()) == str(gold['dec_init_sha256'])
    assert dec.fcn_keys == ['fcn.layers.0', 'fcn.layers.2']


def test_oracle_step_matches_reference_epoch():
    gold = np.load(GOLD)
    enc_sd, dec = initial_state()
    b1 = T.synth_batch(N1, 40)
    b2 = T.synth_batch(N2, 50, with_neg=False, with_3d=False)
    masks = unpack_masks(gold, 0)
    res, grads, bufs = T.zipped_step(enc_sd, dec.state_dict(), dec.fcn_keys,
                                     [('h36m', b1), ('pair', b2)], masks, P, BLOCKS)
    n = N1 + N2
    total = sum(v[1] for v in res.values()) / n
    contra = sum(v[0] for v in res.values()) / n
    assert abs(total - float(gold['step0_loss'])) <= 1e-5 * abs(total)
    assert abs(contra - float(gold['step0_contra'])) <= 1e-5 * abs(contra)
    assert abs(res['h36m'][1] / N1 - float(gold['step0_loss_h36m'])) <= 1e-5 * abs(total)
    assert abs(res['pair'][1] / N2 - float(gold['step0_loss_pair'])) <= 1e-5 * abs(total)
    for k in gold.files:
        if k.startswith('grad0/'):
            np.testing.assert_allclose(grads[k[6:]].numpy(), gold[k], rtol=1e-4, atol=1e-7)
    assert int(bufs['layers.2.block.1.num_batches_tracked']) == 5      # 3 + 2 encoder passes
    keep = np.mean([float(m.float().mean()) for d in masks for ps in d for m in ps])
    assert abs(keep - 0.8) < 0.01


def test_arena_views_and_zipper():
    a = _Arena()
    a.add('w', (3, 5), (4, 8))
    a.add('b', (3,))
    a.params = torch.zeros(a.size)
    a.grads = torch.zeros(a.size)
    a.view('w').copy_(torch.arange(15.).view(3, 5))
    full = a.full('w')
    assert full.shape == (4, 8) and float(full[:, 5:].abs().sum()) == 0 and float(full[3].abs().sum()) == 0
    assert torch.equal(a.view('w'), torch.arange(15.).view(3, 5)) and a.full('b').shape == (3,)
    assert a.entries['b'][0] % 4 == 0
    # batch_zipper: the shorter loader skips rounds, every batch is seen exactly once
    np.random.seed(0)
    zipped = list(keypoint.batch_zipper([('a', [1, 2, 3, 4]), ('b', [10, 20])]))
    assert len(zipped) == 4 and [x for z in zipped for n_, x in z if n_ == 'a'] == [1, 2, 3, 4]
    assert [x for z in zipped for n_, x in z if n_ == 'b'] == [10, 20]
    import random
    random.seed(0)
    seen = sorted(x for _, x in keypoint.batch_mulitplexer([('a', [1, 2, 3]), ('b', [10])]))
    assert seen == [1, 2, 3, 10]


class _FakeState:
    def __init__(self, v):
        self.v = torch.tensor([float(v)])

    def state_dict(self):
        return {'v': self.v.clone()}

    def load_state_dict(self, sd):
        self.v = sd['v'].clone()


class _FakeModel:
    """epoch() returns scripted losses; the driver's file logic needs nothing else"""

    def __init__(self, val_losses):
        self.encoder, self.decoders = _FakeState(0), {'3d': _FakeState(0)}
        self.val = list(val_losses)
        self.calls = 0

    def epoch(self, loaders, optimizer=None, weight_3d=1):
        if optimizer is not None:
            self.calls += 1
            self.encoder.v += 1
            return 0.5, 10.0 - self.calls, {'h36m': 10.0 - self.calls}
        return 0.25, self.val[self.calls - 1], {'h36m': self.val[self.calls - 1]}


def test_fit_driver_files_selection_and_resume(tmp_path):
    import json
    from vpd_b200 import keypoint_train as KT
    cfg = {'datasets': [{'name': 'h36m', '3d_pose_shape': [20, 7], 'mean_kp_offset_norms': [1.0]}],
           'num_epochs': 4, 'learning_rate': 1e-4, 'batch_size': 100, 'embedding_dim': 32,
           'encoder_arch': [2, 1024], 'decoder_arch': [2, 512], 'embed_bones': False,
           'augment_camera': True}
    d = os.path.join(str(tmp_path), 'run')
    model, opt = _FakeModel([5.0, 4.0, 4.5, 3.0]), _FakeState(7)
    logs = []
    hist = KT.fit(model, [], [], d, cfg, opt, num_epochs=4, checkpoint_frequency=2, log=logs.append)
    assert [h['epoch'] for h in hist] == [1, 2, 3, 4] and hist[0]['dataset_train'][0] == ('contrast', 0.5)
    assert sum('New best epoch' in l for l in logs) == 3                # epochs 1, 2, 4
    files = sorted(os.listdir(d))
    assert files == ['best_epoch.decoder-3d.pt', 'best_epoch.encoder.pt', 'best_epoch.optimizer.pt',
                     'config.json', 'epoch0002.decoder-3d.pt', 'epoch0002.encoder.pt',
                     'epoch0002.optimizer.pt', 'epoch0004.decoder-3d.pt', 'epoch0004.encoder.pt',
                     'epoch0004.optimizer.pt', 'loss.json']
    with open(os.path.join(d, 'config.json')) as fp:
        assert json.load(fp) == json.loads(json.dumps(cfg))
    assert KT.get_last_checkpoint(d) == 4
    assert float(torch.load(os.path.join(d, 'best_epoch.encoder.pt'))['v']) == 4.0
    # resume: state comes from epoch0004, history is kept, training continues at epoch 5
    model2, opt2 = _FakeModel([9.0] * 6), _FakeState(0)
    model2.calls = 4
    hist2 = KT.fit(model2, [], [], d, cfg, opt2, num_epochs=6, checkpoint_frequency=2, resume=True,
                   log=logs.append)
    assert [h['epoch'] for h in hist2] == [1, 2, 3, 4, 5, 6] and float(opt2.v) == 7.0
    assert float(model2.encoder.v) == 4.0 + 2
    with pytest.raises(AssertionError):
        KT.fit(model, [], [], os.path.join(str(tmp_path), 'x'), {'num_epochs': 1}, opt, num_epochs=1)
