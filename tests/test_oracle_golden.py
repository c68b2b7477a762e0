"""The oracle restatement vs golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py). CPU only; this is what pins the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import assemble_ref, student_ref
from vpd_b200 import synth


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().numpy()).tobytes()).hexdigest()


def _sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().numpy()).tobytes())
    return h.hexdigest()


@pytest.fixture(scope='module')
def asm(golden_dir):
    return np.load(os.path.join(golden_dir, 'assembly.npz'))


@pytest.fixture(scope='module')
def stu(golden_dir):
    with open(os.path.join(golden_dir, 'student.json')) as fp:
        meta = json.load(fp)
    return np.load(os.path.join(golden_dir, 'student.npz')), meta


RGB_MEAN_STD = {   # vpd_dataset/common.py:14-36
    'tennis': ((0.44157383614877077, 0.47029633580897046, 0.4534017568516162),
               (0.13526736314774856, 0.1208027074415591, 0.1261687563723076)),
    'fs': synth.FS_MEAN_STD,
    'fx': ((0.38402001736617936, 0.34764328219285123, 0.4099846773620623),
           (0.19505844565544309, 0.18984186888162677, 0.1989230425908947)),
    'diving48': ((0.3411329922282787, 0.46349889258964044, 0.5162481674015696),
                 (0.16302619019820488, 0.17092395707914718, 0.19266662199338647)),
    'penn': ((0.43258389316320306, 0.4293850246457961, 0.383481774195889),
             (0.18936336742486998, 0.18502009571154798, 0.18244625387985822)),
    'resnet': ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
}


def test_lut_bit_exact_all_datasets(asm):
    for name, ref_lut in zip(asm['lut_names'], asm['luts']):
        got = assemble_ref.lut(*RGB_MEAN_STD[str(name)])
        assert got.dtype == np.float32
        assert np.array_equal(got.view(np.uint32), ref_lut.view(np.uint32)), name


def test_apply_items_bit_exact(asm):
    rgb, flow = synth.crops(4, seed=11, height=32, width=32)
    got = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD, flip=True)
    assert torch.equal(got, torch.from_numpy(asm['apply_flip']))
    got = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD, flip=False)
    assert torch.equal(got, torch.from_numpy(asm['apply_noflip']))
    got = assemble_ref.apply_batch(rgb.numpy(), None, *synth.FS_MEAN_STD, flip=True)
    assert torch.equal(got, torch.from_numpy(asm['apply_flip_rgbonly']))


def test_apply_full_size_hash(asm):
    rgb, flow = synth.crops(1, seed=13)
    got = assemble_ref.apply_item(rgb[0].numpy(), flow[0].numpy(), *synth.FS_MEAN_STD)
    assert _sha(got) == str(asm['apply128_sha256'])


def test_train_items_bit_exact(asm):
    rgb, flow = synth.crops(4, seed=11, height=32, width=32)
    teach = synth.teacher(4, seed=12, emb_dim=8, motion=True)
    img, emb = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), teach.numpy(),
                                        asm['train_flips'], *synth.FS_MEAN_STD)
    assert torch.equal(img, torch.from_numpy(asm['train_img']))
    assert torch.equal(emb, torch.from_numpy(asm['train_emb']))


def test_motion_targets():
    e = [np.random.RandomState(i).randn(2, 4).astype(np.float32) for i in range(5)]
    vid = [(0, e[0], {'kp_score': 0.9}), (1, e[1], {'kp_score': 0.9}),
           (2, e[2], {'kp_score': 0.1}), (4, e[3], {'kp_score': 0.9}),
           (5, e[4], {'dp_score': 0.7})]
    out = assemble_ref.motion_targets(vid)
    assert [o[0] for o in out] == [1, 5]
    assert out[0][1].shape == (2, 8)
    assert np.array_equal(out[0][1][:, 4:], e[1] - e[0])
    assert np.array_equal(out[1][1][:, :4], e[4])


def test_constructor_bit_identical(stu):
    _, meta = stu
    torch.manual_seed(meta['init_seed'])
    sd = student_ref.init_encoder_state('resnet34', 32, True)
    dsd = student_ref.init_decoder_state(32)
    assert _sd_hash(sd) == meta['encoder_init_sha256']
    assert _sd_hash(dsd) == meta['decoder_init_sha256']
    assert len(sd) == 218
    torch.manual_seed(5)
    sd18 = student_ref.init_encoder_state('resnet18', 26, False)
    assert _sd_hash(sd18) == meta['resnet18_rgb_D26_seed5_sha256']


def test_embed_matches_reference(stu):
    arrays, meta = stu
    torch.manual_seed(0)
    sd = student_ref.init_encoder_state('resnet34', 32, True)
    sd = student_ref.randomize_bn_state(sd, seed=21)
    rgb, flow = synth.crops(4, seed=22)
    x = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD,
                                 flip=True).view(-1, 5, 128, 128)
    got = student_ref.embed(sd, x)
    assert got.shape == (8, 32) and got.dtype == np.float32
    np.testing.assert_allclose(got, arrays['embed_out'], rtol=0, atol=1e-5)


def _curve_batch(meta, rgb, flow, teach, fl, idx, s):
    B = meta['loss_curve']['batch']
    sel = idx[s * B:(s + 1) * B]
    f = fl[s * B:(s + 1) * B]
    return assemble_ref.train_batch(rgb[sel].numpy(), flow[sel].numpy(),
                                    teach[sel].numpy(), f.numpy(), *synth.FS_MEAN_STD)


def test_train_steps_match_reference(stu):
    arrays, meta = stu
    lc = meta['loss_curve']
    B, steps = lc['batch'], lc['steps']
    torch.manual_seed(0)
    sd = student_ref.init_encoder_state('resnet34', 32, True)
    dsd = student_ref.init_decoder_state(32)
    tr = student_ref.OracleTrainer(sd, dsd, lr=lc['lr'])
    rgb, flow = synth.crops(lc['pool'], seed=lc['seeds']['crops'])
    teach = synth.teacher(lc['pool'], seed=lc['seeds']['teacher'], emb_dim=32, motion=True)
    fl = synth.flips(steps * B, seed=lc['seeds']['flips'])
    idx = torch.randint(0, lc['pool'], (steps * B,),
                        generator=torch.Generator().manual_seed(lc['seeds']['index']))
    img, tgt = _curve_batch(meta, rgb, flow, teach, fl, idx, 0)
    loss, grads, out = tr.loss_and_grads(img, tgt, train=True)
    # undo the BN buffer update of this probe pass
    tr2 = student_ref.OracleTrainer(sd, dsd, lr=lc['lr'])
    np.testing.assert_allclose(out.numpy(), arrays['step0_out'], rtol=0, atol=1e-5)
    norms = np.array([g.norm().item() for g in grads])
    np.testing.assert_allclose(norms, arrays['step0_grad_norms'], rtol=2e-4)
    names = student_ref.encoder_param_names('resnet34') + \
        ['decoder.' + n for n in student_ref.DECODER_PARAM_NAMES]
    assert names == meta['param_names']
    gi = names.index('resnet.fc.weight')
    np.testing.assert_allclose(grads[gi].numpy(), arrays['step0_grad_fc'], rtol=1e-3, atol=1e-4)
    # two optimizer steps: losses and updated tensors
    for s in range(2):
        img, tgt = _curve_batch(meta, rgb, flow, teach, fl, idx, s)
        loss = tr2.step(img, tgt) / B
        assert abs(loss - arrays['loss_curve'][s]) <= 1e-4 * abs(arrays['loss_curve'][s])
        np.testing.assert_allclose(tr2.sd['resnet.fc.weight'].detach().numpy(),
                                   arrays['step{}_fc_weight'.format(s)], rtol=0, atol=2e-6)
        np.testing.assert_allclose(tr2.sd['resnet.bn1.running_var'].numpy(),
                                   arrays['step{}_bn1_running_var'.format(s)], rtol=1e-5)
    assert int(tr2.sd['resnet.bn1.num_batches_tracked']) == 2


def test_adamw_numpy_form_matches_torch_form():
    g = torch.Generator().manual_seed(0)
    n = 50021
    p = torch.randn(n, generator=g)
    m = torch.zeros(n); v = torch.zeros(n)
    for t in range(1, 6):
        gr = torch.randn(n, generator=g) * (10.0 ** (t - 3))
        pn, mn, vn = student_ref.adamw_step_numpy(p.numpy(), gr.numpy(), m.numpy(), v.numpy(), t)
        student_ref.adamw_step_torch(p, gr, m, v, t, 5e-4, (0.9, 0.999), 1e-8, 0.01)
        # moments are bit-exact; p inherits torch's not-correctly-rounded vector sqrt
        assert np.array_equal(m.numpy().view(np.int32), mn.view(np.int32))
        assert np.array_equal(v.numpy().view(np.int32), vn.view(np.int32))
        diff = np.abs(p.numpy() - pn)
        assert (diff > 0).mean() < 0.02
        assert diff.max() <= 2.0 ** -22 * max(1.0, float(np.abs(pn).max()))
