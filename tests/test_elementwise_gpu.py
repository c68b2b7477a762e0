"""GPU parity of the BN / ReLU / residual / pooling / head kernels against torch fp32
(autograd for the backward kernels), through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

from vpd_b200._lib import lib, stream_ptr, acc_zeros, acc_from_f64
from gpu_util import dev, nhwc_bf16, nchw_f32, rel_err

pytestmark = pytest.mark.gpu


def _stats(y_nchw):
    return acc_from_f64(torch.stack([y_nchw.double().sum((0, 2, 3)),
                                     (y_nchw.double() ** 2).sum((0, 2, 3))]))


def _bn_ref(y, gamma, beta):
    return F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5)


@pytest.mark.parametrize('N,H,W,C,mode', [(4, 32, 32, 64, 'plain'), (3, 16, 16, 128, 'residual'),
                                           (5, 8, 8, 256, 'downsample'), (9, 4, 4, 512, 'residual'),
                                           (2, 12, 20, 64, 'plain')])
def test_bn_act_forward_and_backward(N, H, W, C, mode):
    g = torch.Generator().manual_seed(C + H)
    mk = lambda *s, scale=1.0, shift=0.0: (torch.randn(s, generator=g) * scale + shift)
    y = nhwc_bf16(mk(N, C, H, W, scale=2.0, shift=0.5)).to(dev())
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev()); beta = mk(C, scale=0.2).to(dev())
    rm = mk(C, scale=0.1).to(dev()); rv = (torch.rand(C, generator=g) + 0.5).to(dev())
    rm0, rv0 = rm.clone(), rv.clone()
    nbt = torch.tensor([3], device=dev(), dtype=torch.int64)
    sm = torch.empty(C, device=dev()); sr = torch.empty(C, device=dev())
    M = N * H * W
    yf = nchw_f32(y).requires_grad_(True)
    gam_r = gamma.clone().requires_grad_(True); bet_r = beta.clone().requires_grad_(True)
    ref = _bn_ref(yf, gam_r, bet_r)
    res = res_f = None
    extra = [None] * 8
    if mode == 'residual':
        res = nhwc_bf16(mk(N, C, H, W)).to(dev())
        res_f = nchw_f32(res)
        ref = ref + res_f
    elif mode == 'downsample':
        res = nhwc_bf16(mk(N, C, H, W, scale=1.5, shift=-0.3)).to(dev())
        res_f = nchw_f32(res).requires_grad_(True)
        g2 = (torch.rand(C, generator=g) + 0.5).to(dev()).requires_grad_(True)
        b2 = mk(C, scale=0.2).to(dev()).requires_grad_(True)
        ref = ref + _bn_ref(res_f, g2, b2)
        sm2 = torch.empty(C, device=dev()); sr2 = torch.empty(C, device=dev())
        rm2 = torch.zeros(C, device=dev()); rv2 = torch.ones(C, device=dev())
        nbt2 = torch.zeros(1, device=dev(), dtype=torch.int64)
        extra = [_stats(nchw_f32(res)), g2.detach(), b2.detach(), rm2, rv2, nbt2, sm2, sr2]
    ref = ref.relu()
    z = torch.empty_like(y)
    mask = torch.zeros((M, C // 8), device=dev(), dtype=torch.uint8)
    lib().call('vpd_bn_act_fwd', y, res, z, M, C, 1, _stats(nchw_f32(y)), gamma, beta, rm, rv, nbt,
               sm, sr, *extra, mask, stream_ptr())
    got = nchw_f32(z)
    assert rel_err(got, ref.detach()) < 5e-3
    # the ReLU mask the data-gradient kernels read instead of z: bit j of byte (row, g) is
    # channel 8g + j of the STORED tensor; the stand-alone kernel gives the same bytes
    bits = ((z.view(M, C // 8, 8).float() > 0).to(torch.int32) << torch.arange(8, device=dev())).sum(-1)
    assert torch.equal(mask.to(torch.int32), bits)
    mask2 = torch.zeros_like(mask)
    lib().call('vpd_relu_bitmask', z, mask2, M, C, stream_ptr())
    assert torch.equal(mask2, mask)
    # running statistics like nn.BatchNorm2d (momentum 0.1, unbiased variance)
    yv = nchw_f32(y)
    assert torch.allclose(rm, 0.9 * rm0 + 0.1 * yv.mean((0, 2, 3)), atol=1e-4)
    assert torch.allclose(rv, 0.9 * rv0 + 0.1 * yv.var((0, 2, 3), unbiased=True), rtol=1e-3)
    assert int(nbt) == 4
    assert torch.allclose(sm, yv.mean((0, 2, 3)), atol=1e-4)

    # ---- backward: upstream gradient through ReLU + BN (+ second BN branch)
    dz = nhwc_bf16(mk(N, C, H, W)).to(dev())
    ref.backward(nchw_f32(dz))
    dy = torch.empty_like(y)
    sums = acc_zeros((2, C), dev())
    dgamma = torch.empty(C, device=dev()); dbeta = torch.empty(C, device=dev())
    dmask = torch.empty_like(dz)
    b2args = [None] * 8
    if mode == 'downsample':
        dy2 = torch.empty_like(y); sums2 = torch.zeros_like(sums)
        dg2 = torch.empty(C, device=dev()); db2 = torch.empty(C, device=dev())
        b2args = [res, dy2, extra[1], extra[6], extra[7], sums2, dg2, db2]
    lib().call('vpd_bn_act_bwd', dz, z, dmask, M, C, y, dy, gamma, sm, sr, sums, dgamma, dbeta,
               *b2args, stream_ptr())
    assert rel_err(nchw_f32(dy), yf.grad) < 1.5e-2
    assert rel_err(dgamma, gam_r.grad) < 5e-3 and rel_err(dbeta, bet_r.grad) < 5e-3
    assert rel_err(nchw_f32(dmask), nchw_f32(dz) * (got > 0)) < 1e-6
    if mode == 'downsample':
        assert rel_err(nchw_f32(dy2), res_f.grad) < 1.5e-2
        assert rel_err(dg2, g2.grad) < 5e-3 and rel_err(db2, b2.grad) < 5e-3


@pytest.mark.parametrize('N,H,W', [(3, 64, 64), (2, 16, 32)])
def test_stem_bn_pool_forward_and_backward(N, H, W):
    C = 64
    g = torch.Generator().manual_seed(H)
    y = nhwc_bf16(torch.randn((N, C, H, W), generator=g) * 1.5 + 0.2).to(dev())
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev()); beta = (torch.randn(C, generator=g) * 0.2).to(dev())
    rm = torch.zeros(C, device=dev()); rv = torch.ones(C, device=dev())
    nbt = torch.zeros(1, device=dev(), dtype=torch.int64)
    sm = torch.empty(C, device=dev()); sr = torch.empty(C, device=dev())
    z = torch.empty((N, H // 2, W // 2, C), device=dev(), dtype=torch.bfloat16)
    am = torch.empty((N, H // 2, W // 2, C), device=dev(), dtype=torch.uint8)
    ysel = torch.empty_like(z)
    lib().call('vpd_stem_bn_pool_fwd', y, z, am, ysel, N, H, W, C, _stats(nchw_f32(y)), gamma, beta,
               rm, rv, nbt, sm, sr, stream_ptr())
    # ysel = the pre-BN value the pooled output came from: bn+relu of it reproduces z
    sc = gamma * sr
    zz = torch.relu(ysel.float() * sc + (beta - sm * sc))
    assert rel_err(zz, z.float()) < 5e-3
    yf = nchw_f32(y).requires_grad_(True)
    gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    ref = F.max_pool2d(_bn_ref(yf, gr, br).relu(), 3, 2, 1)
    assert rel_err(nchw_f32(z), ref.detach()) < 5e-3
    assert int(am.max()) <= 8
    dpool = nhwc_bf16(torch.randn((N, C, H // 2, W // 2), generator=g)).to(dev())
    ref.backward(nchw_f32(dpool))
    dy = torch.empty_like(y)
    sums = acc_zeros((2, C), dev())
    dgamma = torch.empty(C, device=dev()); dbeta = torch.empty(C, device=dev())
    lib().call('vpd_stem_bn_pool_bwd', dpool, am, y, ysel, dy, N, H, W, C, gamma, beta, sm, sr, sums,
               dgamma, dbeta, stream_ptr())
    assert rel_err(nchw_f32(dy), yf.grad) < 2e-2
    assert rel_err(dgamma, gr.grad) < 1e-2 and rel_err(dbeta, br.grad) < 1e-2
    # without ysel the reduction gathers the windows from y: bit-identical sums and outputs
    dy2 = torch.empty_like(y)
    sums2 = acc_zeros((2, C), dev())
    dg2 = torch.empty(C, device=dev()); db2 = torch.empty(C, device=dev())
    lib().call('vpd_stem_bn_pool_bwd', dpool, am, y, None, dy2, N, H, W, C, gamma, beta, sm, sr, sums2,
               dg2, db2, stream_ptr())
    assert rel_err(dg2, dgamma) < 1e-5 and rel_err(db2, dbeta) < 1e-5
    assert rel_err(nchw_f32(dy2), nchw_f32(dy)) < 1e-3


@pytest.mark.parametrize('motion,D,B', [(1, 32, 7), (0, 26, 7), (1, 32, 300)])
def test_head_loss_forward_and_backward(motion, D, B):
    # B = 300: the weight-gradient kernel walks the frames in chunks of 128 (two full, one partial)
    HW, Fd, Hd = 16, 512, 128
    T = 2 * D if motion else D
    g = torch.Generator().manual_seed(D)
    z = torch.randn((B, HW, Fd), generator=g).relu().to(torch.bfloat16).to(dev())
    shapes = [(D, Fd), (D,)] + ([(Hd, D), (Hd,), (Hd, Hd), (Hd,), (T, Hd), (T,)] if motion else [])
    ps = [(torch.randn(s, generator=g) * (0.5 / (s[-1] ** 0.5 if len(s) > 1 else 4))).to(dev())
          .requires_grad_(True) for s in shapes]
    params = torch.cat([p.detach().flatten() for p in ps]).contiguous()
    target = torch.randn((B, T), generator=g).to(dev())
    zf = z.float().requires_grad_(True)
    e = F.linear(zf.mean(1), ps[0], ps[1])
    o = e
    if motion:
        o = F.linear(F.relu(F.linear(F.relu(F.linear(e, ps[2], ps[3])), ps[4], ps[5])), ps[6], ps[7])
    loss = F.mse_loss(o, target, reduction='sum')
    loss.backward()
    emb = torch.empty((B, D), device=dev()); out = torch.empty((B, T), device=dev())
    lsum = torch.zeros(1, device=dev(), dtype=torch.float64)
    dz = torch.empty_like(z)
    ws = torch.empty(B * (Fd + 2 * D + 4 * Hd + T), device=dev())
    grads = torch.zeros_like(params)
    lib().call('vpd_head_fwd_bwd', z, B, HW, Fd, D, T, motion, params, target, emb, out, lsum, dz, ws,
               grads, stream_ptr())
    assert torch.allclose(emb, e.detach(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(out, o.detach(), atol=1e-4, rtol=1e-4)
    assert abs(lsum.item() - loss.item()) <= 1e-4 * loss.item()
    assert rel_err(dz.float(), zf.grad) < 5e-3
    ref_g = torch.cat([p.grad.flatten() for p in ps])
    assert rel_err(grads, ref_g) < 1e-4
