"""CPU: the keypoint-teacher training oracle (oracle/keypoint_train_ref.py) against the golden
outputs of the unmodified reference `Keypoint_EmbeddingModel.epoch` (losses and first-step
gradients under the dropout masks the reference drew), the parameter containers' initialisation,
and the loaders' zipping logic."""
import hashlib
import os

import numpy as np
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
