"""GPU parity of the keypoint (VIPE*) teacher TRAINING step (vpd_b200/keypoint_train.py +
csrc/mlp.cu + the implicit-GEMM kernels as 1x1 convolutions), through the C ABI, against the
golden outputs of the unmodified reference `Keypoint_EmbeddingModel.epoch` under the dropout
masks the reference drew, and op by op against torch.

Tolerances (bf16 operands / activations, fp32 accumulation): losses within 1 %, weight
gradients cosine >= 0.98 and norm within 5 % of the reference's fp32 autograd, BatchNorm running
statistics within 2 %, parameters after two AdamW steps within 2.1 * lr * steps."""
import os

import numpy as np
import pytest
import torch

from oracle import keypoint_train_ref as T
from vpd_b200 import keypoint
from vpd_b200._lib import lib, stream_ptr, acc_zeros, acc_from_f64, acc_to_f64
from vpd_b200.keypoint_train import FCPoseDecoder
from gpu_util import dev, OUT
from test_keypoint_train_cpu import GOLD, H, BLOCKS, N1, N2, P, unpack_masks

pytestmark = pytest.mark.gpu
LR = 1e-3


def _log(msg):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, 'keypoint_train_diag.txt'), 'a') as fp:
        fp.write(msg + '\n')


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


def _bf(t):
    return t.to(dev()).to(torch.bfloat16).contiguous()


# ------------------------------------------------------------------------------ op level
@pytest.mark.parametrize('M,C,drop,res', [(136, 128, True, True), (4096, 1024, True, False),
                                          (200, 192, False, False)])
def test_bn1d_fwd_bwd_against_torch(M, C, drop, res):
    L, st = lib(), stream_ptr(dev())
    g = torch.Generator().manual_seed(M + C)
    a = _bf(torch.randn((M, C), generator=g) * 1.5 + 0.3)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev())
    beta = (torch.randn(C, generator=g) * 0.3).to(dev())
    bias = (torch.randn(C, generator=g) * 0.1).to(dev())
    keep = (torch.rand((M, C), generator=g) < 0.8).to(torch.uint8).to(dev()) if drop else None
    r = _bf(torch.randn((M, C), generator=g)) if res else None
    dz = _bf(torch.randn((M, C), generator=g))
    af = a.float()
    stats = acc_from_f64(torch.cat([af.double().sum(0), (af.double() ** 2).sum(0)]))
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    nbt = torch.zeros((), device=dev(), dtype=torch.int64)
    sm, sr = torch.empty(C, device=dev()), torch.empty(C, device=dev())
    out = torch.empty_like(a)
    L.call('vpd_bn1d_fwd', a, stats, gamma, beta, bias, rm, rv, nbt, sm, sr, keep, P if drop else 0.0,
           r, out, M, C, 1, st)
    # torch reference on the same bf16-rounded input
    x = af.clone().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    trm, trv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    y = torch.nn.functional.batch_norm(x + bias, trm, trv, gp, bp, training=True, momentum=0.1)
    y = torch.relu(y)
    if drop:
        y = y * keep.float() / (1 - P)
    if res:
        y = y - r.float()
    torch.testing.assert_close(out.float(), y.detach(), rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(rm, trm, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rv, trv, rtol=1e-3, atol=1e-4)
    assert int(nbt) == 1
    y.backward(dz.float())
    da = torch.empty_like(a)
    dg, db = torch.ones(C, device=dev()), torch.ones(C, device=dev())     # += semantics
    sums = acc_zeros(2 * C, dev())
    L.call('vpd_bn1d_bwd', dz, a, keep, P if drop else 0.0, gamma, beta, sm, sr, sums, da, dg, db,
           M, C, 1, st)
    msg = 'bn1d M={} C={}: da cos {:.5f}, dgamma cos {:.5f}, dbeta cos {:.5f}'.format(
        M, C, _cos(da.float(), x.grad), _cos(dg - 1, gp.grad), _cos(db - 1, bp.grad))
    _log(msg)
    assert _cos(da.float(), x.grad) > 0.995, msg
    torch.testing.assert_close(dg - 1, gp.grad, rtol=2e-2, atol=2e-2 * float(gp.grad.abs().max()))
    torch.testing.assert_close(db - 1, bp.grad, rtol=2e-2, atol=2e-2 * float(bp.grad.abs().max()))


def test_bn1d_row_groups_are_separate_batches():
    """three stacked encoder passes: per-group batch statistics, running statistics updated
    group after group, dgamma / dbeta summed over the groups"""
    L, st = lib(), stream_ptr(dev())
    G, M, C = 3, 136, 128
    g = torch.Generator().manual_seed(77)
    a = _bf(torch.randn((G * M, C), generator=g) * torch.tensor([1.0, 2.0, 0.5]).repeat_interleave(M)[:, None]
            + torch.tensor([0.0, 1.0, -1.0]).repeat_interleave(M)[:, None])
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev())
    beta = (torch.randn(C, generator=g) * 0.3).to(dev())
    keep = (torch.rand((G * M, C), generator=g) < 0.8).to(torch.uint8).to(dev())
    dz = _bf(torch.randn((G * M, C), generator=g))
    stats = acc_zeros(G * 2 * C, dev()) + 7          # the call zeroes it
    L.call('vpd_colstats_bf16', a, stats, M, C, G, st)
    af = a.float().view(G, M, C)
    stf = acc_to_f64(stats).view(G, 2, C)
    torch.testing.assert_close(stf[:, 0].float(), af.sum(1), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(stf[:, 1].float(), (af ** 2).sum(1), rtol=1e-4, atol=1e-2)
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    nbt = torch.zeros((), device=dev(), dtype=torch.int64)
    sm, sr = torch.empty(G * C, device=dev()), torch.empty(G * C, device=dev())
    out = torch.empty_like(a)
    L.call('vpd_bn1d_fwd', a, stats, gamma, beta, None, rm, rv, nbt, sm, sr, keep, P, None, out,
           M, C, G, st)
    x = a.float().clone().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    trm, trv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    ys = []
    for q in range(G):
        y = torch.nn.functional.batch_norm(x[q * M:(q + 1) * M], trm, trv, gp, bp, training=True,
                                           momentum=0.1)
        ys.append(torch.relu(y) * keep[q * M:(q + 1) * M].float() / (1 - P))
    y = torch.cat(ys)
    torch.testing.assert_close(out.float(), y.detach(), rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(rm, trm, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(rv, trv, rtol=1e-3, atol=1e-4)
    assert int(nbt) == G
    y.backward(dz.float())
    da = torch.empty_like(a)
    dg, db = torch.zeros(C, device=dev()), torch.zeros(C, device=dev())
    sums = acc_zeros(G * 2 * C, dev())
    L.call('vpd_bn1d_bwd', dz, a, keep, P, gamma, beta, sm, sr, sums, da, dg, db, M, C, G, st)
    assert _cos(da.float(), x.grad) > 0.995
    torch.testing.assert_close(dg, gp.grad, rtol=2e-2, atol=2e-2 * float(gp.grad.abs().max()))
    torch.testing.assert_close(db, bp.grad, rtol=2e-2, atol=2e-2 * float(bp.grad.abs().max()))


def test_small_ops_against_torch():
    L, st = lib(), stream_ptr(dev())
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn((300, 192), generator=g))
    out = torch.full((192,), 2.0, device=dev())
    L.call('vpd_colsum_bf16', x, out, 300, 192, st)
    torch.testing.assert_close(out - 2, x.float().sum(0), rtol=1e-4, atol=1e-3)
    z = _bf(torch.randn((300, 192), generator=g))
    o = torch.empty_like(x)
    L.call('vpd_relu_mask_bf16', x, z, o, x.numel(), st)
    assert torch.equal(o, torch.where(z > 0, x, torch.zeros_like(x)))
    keep = torch.empty(1 << 20, device=dev(), dtype=torch.uint8)
    L.call('vpd_dropout_mask', keep, keep.numel(), 0.2, 1234, None, 7, st)
    assert set(keep.unique().tolist()) <= {0, 1} and abs(float(keep.float().mean()) - 0.8) < 0.003
    keep2 = torch.empty_like(keep)
    L.call('vpd_dropout_mask', keep2, keep.numel(), 0.2, 1234, None, 8, st)
    assert not torch.equal(keep, keep2)
    add = torch.tensor([5], device=dev(), dtype=torch.int64)         # seed + device-side counter
    L.call('vpd_dropout_mask', keep2, keep.numel(), 0.2, 1229, add, 7, st)
    assert torch.equal(keep, keep2)


@pytest.mark.parametrize('full', [True, False])
def test_vipe_loss_against_torch_autograd(full):
    import torch.nn.functional as F
    L, st = lib(), stream_ptr(dev())
    n, D, Tt, Tpad = 200, 32, 140, 192
    g = torch.Generator().manual_seed(9)
    e1, e2, en = (torch.randn((n, D), generator=g).to(dev()) * s for s in (0.3, 0.3, 0.25))
    e2 = e1 + 0.1 * e2
    e2[5] = e1[5]                                    # zero distance: gradient must be 0, not NaN
    valid = (torch.rand(n, generator=g) > 0.3).float().to(dev())
    p1 = torch.zeros((n, Tpad), device=dev(), dtype=torch.bfloat16)
    p2 = torch.zeros_like(p1)
    p1[:, :Tt] = _bf(torch.randn((n, Tt), generator=g))
    p2[:, :Tt] = _bf(torch.randn((n, Tt), generator=g))
    true3d = torch.randn((n, Tt), generator=g).to(dev())
    gs, w3d = 1.0 / 300, 1.0
    de1, de2, den = (torch.empty((n, D), device=dev()) for _ in range(3))
    dp1, dp2 = torch.empty_like(p1), torch.empty_like(p2)
    sums = torch.zeros(2, device=dev(), dtype=torch.float64)
    if full:
        L.call('vpd_vipe_loss', e1, e2, en, valid, p1, p2, true3d, de1, de2, den, dp1, dp2, sums, n,
               D, Tt, Tpad, w3d, gs, st)
    else:
        L.call('vpd_vipe_loss', e1, e2, None, None, None, None, None, de1, de2, None, None, None,
               sums, n, D, 0, 0, w3d, gs, st)
    a, b, c = (t.clone().requires_grad_(True) for t in (e1, e2, en))
    q1, q2 = p1[:, :Tt].float().requires_grad_(True), p2[:, :Tt].float().requires_grad_(True)
    contra = F.hinge_embedding_loss(torch.norm(a - b, dim=1), torch.ones(n, device=dev()), reduction='sum')
    loss = contra
    if full:
        contra = contra + torch.sum(F.hinge_embedding_loss(
            torch.norm(a - c, dim=1), -torch.ones(n, device=dev()), reduction='none') * valid)
        loss = contra + w3d * (F.mse_loss(q1, true3d, reduction='sum') + F.mse_loss(q2, true3d, reduction='sum'))
    (loss * gs).backward()
    contra, loss = contra.detach(), loss.detach()
    assert abs(float(sums[0]) - float(contra)) <= 1e-5 * abs(float(contra))
    assert abs(float(sums[1]) - float(loss)) <= 1e-4 * abs(float(loss))
    torch.testing.assert_close(de1, a.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(de2, b.grad, rtol=1e-4, atol=1e-7)
    assert torch.isfinite(de1).all() and float(de2[5].abs().max()) == 0.0
    if full:
        torch.testing.assert_close(den, c.grad, rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(dp1[:, :Tt].float(), q1.grad, rtol=1e-2, atol=1e-6)
        torch.testing.assert_close(dp2[:, :Tt].float(), q2.grad, rtol=1e-2, atol=1e-6)
        assert float(dp1[:, Tt:].abs().max()) == 0.0


def test_linear_backward_as_1x1_convolution():
    """the three launches every Linear's backward is made of, on a ragged row count"""
    L, st = lib(), stream_ptr(dev())
    n, cin, cout = 136, 128, 192
    g = torch.Generator().manual_seed(13)
    x, dy = _bf(torch.randn((n, cin), generator=g)), _bf(torch.randn((n, cout), generator=g))
    w = (torch.randn((cout, cin), generator=g) * 0.1).to(dev())
    wt, wtt = (torch.empty(cout * cin, device=dev(), dtype=torch.bfloat16) for _ in range(2))
    L.call('vpd_pack_conv_weight', w, wt, wtt, cout, cin, 1, st)
    y = torch.empty((n, cout), device=dev(), dtype=torch.bfloat16)
    stats = acc_zeros(2 * cout, dev())
    L.call('vpd_conv2d_fwd', x, wt, y, n, 1, 1, cin, cout, 1, 1, 0, None, None, None, 0, stats, st)
    wb = w.to(torch.bfloat16).float()
    ref = x.float() @ wb.t()
    assert _cos(y.float(), ref) > 0.9999
    torch.testing.assert_close(acc_to_f64(stats)[:cout].float(), y.float().sum(0), rtol=1e-3, atol=1e-2)
    dw = torch.zeros((cout, cin), device=dev())
    L.call('vpd_conv2d_wgrad', x, dy, dw, n, 1, 1, cin, cout, 1, 1, 0, st)
    torch.testing.assert_close(dw, dy.float().t() @ x.float(), rtol=1e-3, atol=1e-2)
    dx = torch.empty((n, cin), device=dev(), dtype=torch.bfloat16)
    r = _bf(torch.randn((n, cin), generator=g))
    L.call('vpd_conv2d_dgrad', dy, wtt, dx, n, 1, 1, cin, cout, 1, 1, 0, r, None, None, 0, st)
    assert _cos(dx.float(), dy.float() @ wb + r.float()) > 0.9999


# --------------------------------------------------------------------------- the whole step
def _build():
    torch.manual_seed(31)
    enc = keypoint.FCResNet(39, 32, BLOCKS, H, dropout=P)
    dec = FCPoseDecoder(32, [128, 128], [('h36m', 140)])
    return keypoint.Keypoint_EmbeddingModel(enc, {'3d': dec}, 'cuda'), enc, dec


def test_two_training_steps_match_the_reference():
    gold = np.load(GOLD)
    model, enc, dec = _build()
    opt = model.get_optimizer(LR)
    core = model._core()
    captured = []
    real_step = opt.step

    def step():
        captured.append({k: core.arena.view(k, grad=True).clone() for k in core.arena.entries})
        real_step()
    opt.step = step
    for s in range(2):
        b1 = T.synth_batch(N1, 40 + s)
        b2 = T.synth_batch(N2, 50 + s, with_neg=False, with_3d=False)
        masks = unpack_masks(gold, s)
        dm = iter([[[m.to(dev()).contiguous() for m in ps] for ps in d] for d in masks])
        contra, loss, per = model.epoch([('h36m', [b1]), ('pair', [b2])], optimizer=opt,
                                        weight_3d=1, dropout_masks=dm)
        _log('step {}: contra {:.5f} (ref {:.5f}) loss {:.5f} (ref {:.5f}) h36m {:.5f} ({:.5f}) '
             'pair {:.5f} ({:.5f})'.format(s, contra, float(gold['step%d_contra' % s]), loss,
                                           float(gold['step%d_loss' % s]), per['h36m'],
                                           float(gold['step%d_loss_h36m' % s]), per['pair'],
                                           float(gold['step%d_loss_pair' % s])))
        for got, key in ((contra, 'contra'), (loss, 'loss'), (per['h36m'], 'loss_h36m'),
                         (per['pair'], 'loss_pair')):
            want = float(gold['step{}_{}'.format(s, key)])
            assert abs(got - want) <= 1e-2 * abs(want), (s, key, got, want)
    bad = []
    for k in gold.files:
        if not k.startswith('grad0/'):
            continue
        name = k[6:]
        want = torch.from_numpy(gold[k])
        got = captured[0][name].cpu()
        if '.block.0.bias' in name or '.block.4.bias' in name:
            assert float(got.abs().max()) == 0.0          # bias in front of a batch-stat BN
            continue
        c, ratio = _cos(got, want), float(got.norm() / (want.norm() + 1e-30))
        _log('grad {}: cos {:.4f} norm ratio {:.4f}'.format(name, c, ratio))
        if c < 0.98 or abs(ratio - 1) > 0.05:
            bad.append((name, c, ratio))
    assert not bad, bad
    sd = enc.state_dict()
    for k in gold.files:
        if k.startswith('final/layers'):
            want = torch.from_numpy(gold[k])
            got = sd[k[6:]].cpu()
            if 'num_batches' in k:
                assert int(got) == int(want) == 10
            else:
                torch.testing.assert_close(got, want, rtol=2e-2, atol=2e-3)
    for k, got in (('final/enc.layers.4.weight', sd['layers.4.weight']),
                   ('final/dec.fc_h36m.bias', dec.state_dict()['fc_h36m.bias'])):
        d = (got.cpu() - torch.from_numpy(gold[k])).abs()
        _log('{}: max diff {:.2e}, mean diff {:.2e}'.format(k, float(d.max()), float(d.mean())))
        assert float(d.max()) <= 2.1 * LR * 2 and float(d.mean()) <= 0.3 * LR
    # evaluation epoch (running statistics, no dropout) on the trained weights
    contra, loss, per = model.epoch([('h36m', [T.synth_batch(N1, 70)])])
    _log('eval: contra {:.5f} ({:.5f}) loss {:.5f} ({:.5f})'.format(
        contra, float(gold['eval_contra']), loss, float(gold['eval_loss'])))
    assert abs(loss - float(gold['eval_loss'])) <= 2e-2 * abs(float(gold['eval_loss']))
    # 3-D prediction entry points
    pose = T.synth_batch(8, 71)['pose1']
    emb, pred = model.embed_and_predict3d(pose, 'h36m')
    assert emb.shape == (8, 32) and pred.shape == (8, 140) and np.isfinite(pred).all()


@pytest.mark.parametrize('graphs', [False, True])
def test_training_with_device_dropout_reduces_the_loss(graphs):
    model, enc, dec = _build()
    model._core().use_graphs = graphs          # opt-in CUDA-graph replay of the step
    opt = model.get_optimizer(2e-3)
    data = [T.synth_batch(256, 80 + i) for i in range(4)]
    first = last = None
    for ep in range(6):
        contra, loss, per = model.epoch([('h36m', data)], optimizer=opt)
        first = loss if first is None else first
        last = loss
    _log('device-dropout training (graphs={}): loss {:.4f} -> {:.4f}'.format(graphs, first, last))
    assert np.isfinite(last) and last < 0.95 * first        # measured 29.4 -> 26.4 (ratio 0.90)
    sd = enc.state_dict()
    assert int(sd['layers.2.block.1.num_batches_tracked']) == 6 * 4 * 3
