"""GPU parity of the keypoint (VIPE*) teacher apply path (vpd_b200/keypoint.py, csrc/mlp.cu +
the implicit-GEMM kernels as 1x1 convolutions) through the C ABI against the oracle
(oracle/keypoint_ref.py, torch CPU fp32) and the golden outputs of the unmodified reference.

Tolerance (floating point, bf16 hidden activations, fp32 accumulation and output): per-pose
cosine >= 0.999 against the fp32 reference and max-abs error <= 2 % of the largest embedding
magnitude (measured values are printed into gpurun_out/keypoint_parity.txt)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import keypoint_ref as K
from vpd_b200 import init, keypoint
from vpd_b200._lib import lib, stream_ptr
from gpu_util import dev, OUT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'keypoint.npz')


def _cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1) + 1e-12)


def _model(in_dim, hidden, blocks, seed):
    torch.manual_seed(seed)
    enc = keypoint.FCResNet(in_dim, 32, blocks, hidden, dropout=0.2)
    sd = K.perturb_bn(enc.state_dict(), seed + 100)
    enc.load_state_dict(sd)
    return keypoint.Keypoint_EmbeddingModel(enc, {}, 'cuda'), sd


@pytest.mark.parametrize('tag,in_dim,joints,hidden,blocks', [('d39', 39, 13, 1024, 2),
                                                            ('d75', 75, 25, 256, 1)])
def test_embed_matches_reference_golden(tag, in_dim, joints, hidden, blocks):
    gold = np.load(GOLD)
    seed = int(gold[tag + '_seed'])
    model, sd = _model(in_dim, hidden, blocks, seed)
    poses = K.synth_poses(96, seed + 200, joints)
    emb = model.embed(poses)
    want = gold[tag + '_emb']
    assert emb.dtype == np.float32 and emb.shape == want.shape
    cos = _cos(emb, want)
    err = np.abs(emb - want).max()
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, 'keypoint_parity.txt'), 'a') as fp:
        fp.write('{}: min cosine {:.6f}, max-abs {:.4e}, max |emb| {:.3f}\n'.format(
            tag, float(cos.min()), float(err), float(np.abs(want).max())))
    assert cos.min() >= 0.999, float(cos.min())
    assert err <= 0.02 * np.abs(want).max(), float(err)
    one = model.embed(poses[3].numpy())                      # [J,3] -> [1,D]
    assert one.shape == (1, 32) and _cos(one, gold[tag + '_emb_one']).min() >= 0.999
    # the nn.Module surface the reference scripts use
    sd2 = model.encoder.state_dict()
    assert list(sd2) == list(sd) and all(torch.equal(sd2[k].cpu(), sd[k]) for k in sd)
    assert sd2['layers.2.block.1.num_batches_tracked'].dtype == torch.int64
    with pytest.raises(NotImplementedError):
        model.encoder.train()(torch.zeros(4, in_dim))


@pytest.mark.parametrize('n', [1, 127, 250, 1000, 70000])
def test_ragged_batches_and_chunking(n):
    model, sd = _model(39, 256, 2, 11)
    poses = K.synth_poses(n, 12)
    emb = model.embed(poses)
    idx = np.unique(np.linspace(0, n - 1, 300).astype(int))
    want = K.embed(sd, poses[idx], 2)
    assert emb.shape == (n, 32) and np.isfinite(emb).all()
    assert _cos(emb[idx], want).min() >= 0.999
    # a row's embedding does not depend on what else is in the batch (eval mode)
    again = model.embed(poses[idx])
    assert np.array_equal(again, emb[idx])


def test_checkpoint_interchange(tmp_path):
    """a reference-format run directory (config.json + best_epoch.encoder.pt) loads"""
    torch.manual_seed(21)
    sd = K.perturb_bn(init.fcresnet_state(39, 32, 2, 128), 22)
    torch.save(sd, os.path.join(tmp_path, 'best_epoch.encoder.pt'))
    with open(os.path.join(tmp_path, 'config.json'), 'w') as fp:
        json.dump({'embedding_dim': 32, 'encoder_arch': [2, 128], 'embed_bones': False}, fp)
    model, bones = keypoint.load_embedding_model(str(tmp_path))
    assert bones is False
    poses = K.synth_poses(40, 23)
    assert _cos(model.embed(poses), K.embed(sd, poses, 2)).min() >= 0.999
    frames = np.repeat(np.arange(20), 2)
    embs = keypoint.embed_video(model, frames, np.full(40, 0.8), np.tile([False, True], 20), poses)
    assert len(embs) == 20 and embs[5][1].shape == (2, 32) and embs[5][1].dtype == np.float32


def test_row_ops_against_torch():
    L, st = lib(), stream_ptr(dev())
    g = torch.Generator().manual_seed(3)
    x = torch.randn((300, 39), generator=g).to(dev())
    out = torch.full((300, 64), 7.0, device=dev(), dtype=torch.bfloat16)
    L.call('vpd_rows_to_bf16', x, out, 300, 39, 64, st)
    assert torch.equal(out[:, :39], x.to(torch.bfloat16)) and float(out[:, 39:].abs().max()) == 0
    a = torch.randn((300, 64), generator=g).to(dev()).to(torch.bfloat16)
    b = torch.randn((300, 64), generator=g).to(dev()).to(torch.bfloat16)
    o = torch.empty_like(a)
    L.call('vpd_axpby_bf16', a, 1.0, b, -1.0, o, a.numel(), st)
    assert torch.equal(o, (a.float() - b.float()).to(torch.bfloat16))
    L.call('vpd_axpby_bf16', a, -0.5, None, 0.0, o, a.numel(), st)
    assert torch.equal(o, (a.float() * -0.5).to(torch.bfloat16))
    C = 200
    gamma, beta, mean, bias = (torch.randn(C, generator=g).to(dev()) for _ in range(4))
    var = (torch.rand(C, generator=g) + 0.1).to(dev())
    scale, shift = torch.empty(C, device=dev()), torch.empty(C, device=dev())
    L.call('vpd_bn_fold', gamma, beta, mean, var, bias, 1e-5, scale, shift, C, st)
    s_ref = gamma / torch.sqrt(var + 1e-5)
    torch.testing.assert_close(scale, s_ref, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(shift, (bias - mean) * s_ref + beta, rtol=1e-5, atol=1e-6)
    h = torch.randn((77, 256), generator=g).to(dev()).to(torch.bfloat16)
    w = (torch.randn((32, 256), generator=g) * 0.1).to(dev())
    bb = torch.randn(32, generator=g).to(dev())
    y = torch.empty((77, 32), device=dev())
    L.call('vpd_linear_rows_f32', h, w, bb, y, 77, 256, 32, st)
    torch.testing.assert_close(y, h.float() @ w.t() + bb, rtol=1e-4, atol=1e-4)
    with pytest.raises(Exception):
        L.call('vpd_linear_rows_f32', h, w, bb, y, 77, 100, 32, st)


def test_apply_pose_dir_end_to_end(tmp_path):
    """apply_vipe_model.main on the device: coco_keypoints.json.gz -> <video>.emb.pkl"""
    import gzip
    import pickle
    from vpd_b200 import keypoint_apply as KA
    torch.manual_seed(41)
    sd = K.perturb_bn(init.fcresnet_state(39, 32, 2, 128), 42)
    model_dir, pose_dir, out_dir = (os.path.join(str(tmp_path), d) for d in ('m', 'p', 'o'))
    os.makedirs(model_dir)
    os.makedirs(pose_dir)
    torch.save(sd, os.path.join(model_dir, 'best_epoch.encoder.pt'))
    with open(os.path.join(model_dir, 'config.json'), 'w') as fp:
        json.dump({'embedding_dim': 32, 'encoder_arch': [2, 128], 'embed_bones': False}, fp)
    g = torch.Generator().manual_seed(43)
    kp = torch.rand((30, 17, 3), generator=g)
    kp[:, :, :2] = kp[:, :, :2] * 200 + 20
    dets = [[f, [[0.9, [0, 0, 1, 1], kp[f].tolist()]]] for f in range(30)]
    with gzip.open(os.path.join(pose_dir, 'clip.json.gz'), 'wt', encoding='ascii') as fp:
        json.dump(dets, fp)
    done = KA.apply_pose_dir(pose_dir, model_dir, out_dir, log=lambda *_: None)
    assert done == [('clip', 30)]
    with open(os.path.join(out_dir, 'clip.emb.pkl'), 'rb') as fp:
        embs = pickle.load(fp)
    assert [e[0] for e in embs] == list(range(30)) and embs[0][1].shape == (2, 32)
    got = np.stack([e[1] for e in embs])                                  # [30, 2, 32]
    for row, fl in ((0, False), (1, True)):
        want = K.embed(sd, KA.normalize_2d_skeletons(kp.numpy(), fl), 2)
        assert _cos(got[:, row], want).min() >= 0.999
