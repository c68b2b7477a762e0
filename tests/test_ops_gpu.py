"""GPU parity of the individual CUDA ops (through the C ABI) against the oracle /
a plain torch fp32 reference of the same op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import assemble_ref, student_ref
from vpd_b200 import synth
from vpd_b200._lib import lib, stream_ptr, acc_zeros, acc_from_f64, acc_to_f64
from gpu_util import dev, nhwc_bf16, nchw_f32, rel_err, report

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


MEAN = torch.tensor(synth.FS_MEAN_STD[0], dtype=torch.float32)
STD = torch.tensor(synth.FS_MEAN_STD[1], dtype=torch.float32)


# ------------------------------------------------------------------ K1 assembly
@pytest.mark.parametrize('H,W,B', [(128, 128, 5), (32, 32, 3), (20, 36, 2)])
def test_assemble_apply_bit_exact(H, W, B):
    rgb, flow = synth.crops(B, seed=31, height=H, width=W)
    ref = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD, flip=True)
    out = torch.empty((B, 2, 5, H, W), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), flow.to(dev()), 3, None, None, None, 0, 0,
               MEAN, STD, out, None, B, H, W, 2, stream_ptr())
    assert torch.equal(out.cpu(), ref)
    # no flip, single variant
    ref1 = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD, flip=False)
    out1 = torch.empty((B, 1, 5, H, W), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), flow.to(dev()), 3, None, None, None, 0, 0,
               MEAN, STD, out1, None, B, H, W, 1, stream_ptr())
    assert torch.equal(out1.cpu(), ref1)


def test_assemble_apply_rgb_only_bit_exact():
    rgb, _ = synth.crops(3, seed=32, height=64, width=64)
    ref = assemble_ref.apply_batch(rgb.numpy(), None, *synth.FS_MEAN_STD, flip=True)
    out = torch.empty((3, 2, 3, 64, 64), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), None, 0, None, None, None, 0, 0,
               MEAN, STD, out, None, 3, 64, 64, 2, stream_ptr())
    assert torch.equal(out.cpu(), ref)


def test_assemble_train_bit_exact_with_gather_and_golden(golden_dir):
    import os
    pool, B = 16, 24
    rgb, flow = synth.crops(pool, seed=33)
    teach = synth.teacher(pool, seed=34, emb_dim=32, motion=True)
    flips = synth.flips(B, seed=35)
    idx = torch.randint(0, pool, (B,), generator=torch.Generator().manual_seed(36)).int()
    ref_img, ref_emb = assemble_ref.train_batch(
        rgb[idx.long()].numpy(), flow[idx.long()].numpy(), teach[idx.long()].numpy(),
        flips.numpy(), *synth.FS_MEAN_STD)
    img = torch.empty((B, 1, 5, 128, 128), device=dev(), dtype=torch.float32)
    tgt = torch.empty((B, 64), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), flow.to(dev()), 3, idx.to(dev()),
               flips.to(dev()), teach.to(dev()), 2, 64, MEAN, STD, img, tgt, B, 128, 128, 1,
               stream_ptr())
    assert torch.equal(img.cpu()[:, 0], ref_img)
    assert torch.equal(tgt.cpu(), ref_emb)
    # the reference's own outputs (golden fixture made from GenericDataset.__getitem__)
    g = np.load(os.path.join(golden_dir, 'assembly.npz'))
    rgb, flow = synth.crops(4, seed=11, height=32, width=32)
    teach = synth.teacher(4, seed=12, emb_dim=8, motion=True)
    img = torch.empty((4, 1, 5, 32, 32), device=dev(), dtype=torch.float32)
    tgt = torch.empty((4, 16), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), flow.to(dev()), 3, None,
               torch.from_numpy(g['train_flips']).to(dev()), teach.to(dev()), 2, 16, MEAN, STD,
               img, tgt, 4, 32, 32, 1, stream_ptr())
    assert torch.equal(img.cpu()[:, 0], torch.from_numpy(g['train_img']))
    assert torch.equal(tgt.cpu(), torch.from_numpy(g['train_emb']))
    out = torch.empty((4, 2, 5, 32, 32), device=dev(), dtype=torch.float32)
    lib().call('vpd_assemble_nchw', rgb.to(dev()), flow.to(dev()), 3, None, None, None, 0, 0,
               MEAN, STD, out, None, 4, 32, 32, 2, stream_ptr())
    assert torch.equal(out.cpu(), torch.from_numpy(g['apply_flip']))


def test_assemble_empty_batch_and_bad_width():
    out = torch.empty((0,), device=dev())
    lib().call('vpd_assemble_nchw', out, None, 0, None, None, None, 0, 0, MEAN, STD, out, None,
               0, 128, 128, 1, stream_ptr())
    from vpd_b200._lib import VpdError
    with pytest.raises(VpdError):
        lib().call('vpd_assemble_nchw', out, None, 0, None, None, None, 0, 0, MEAN, STD, out,
                   None, 1, 8, 6, 1, stream_ptr())


def _stem_cells(H, W):
    return (H + 7) // 2, (W + 9) // 4


def _stem_layout_ref(x_nchw):
    """fp32 [B,C,H,W] -> the network input layout like the kernels write it: the image padded by
    3 rows / columns on the top / left with 8 channel slots per pixel, stored space-to-depth
    2 x 4 (common.cuh::stem_pixel_offset): bf16 [B, Hs, Ws, 64], element (a*4 + q)*8 + c of cell
    (i, j) = padded pixel (2i + a, 4j + q), channel c."""
    B, C, H, W = x_nchw.shape
    Hs, Ws = _stem_cells(H, W)
    pad = torch.zeros((B, 2 * Hs, 4 * Ws, 8), dtype=torch.bfloat16)
    pad[:, 3:3 + H, 3:3 + W, :C] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return pad.view(B, Hs, 2, Ws, 4, 8).permute(0, 1, 3, 2, 4, 5).reshape(B, Hs, Ws, 64).contiguous()


def _stem_layout_unpack(cells, H, W):
    """inverse of the cell packing: [B, Hs, Ws, 64] -> padded [B, 2 Hs, 4 Ws, 8]"""
    B, Hs, Ws, _ = cells.shape
    return cells.view(B, Hs, Ws, 2, 4, 8).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * Hs, 4 * Ws, 8)


def test_assemble_stem_layout_bit_exact():
    B = 6
    rgb, flow = synth.crops(B, seed=37)
    flips = synth.flips(B, seed=38)
    teach = synth.teacher(B, seed=39)
    ref_img, ref_emb = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), teach.numpy(),
                                                flips.numpy(), *synth.FS_MEAN_STD)
    out = torch.full((B, 67, 34, 64), 7.0, device=dev(), dtype=torch.bfloat16)
    tgt = torch.empty((B, 64), device=dev())
    lib().call('vpd_assemble_stem', rgb.to(dev()), flow.to(dev()), 3, None, flips.to(dev()),
               teach.to(dev()), 2, 64, MEAN, STD, out, tgt, B, 128, 128, 1, stream_ptr())
    assert torch.equal(out.cpu(), _stem_layout_ref(ref_img))
    assert torch.equal(tgt.cpu(), ref_emb)
    # k = 2 apply variant and the fp32 NCHW -> stem converter
    ref2 = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD, flip=True)
    out2 = torch.empty((B * 2, 67, 34, 64), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_assemble_stem', rgb.to(dev()), flow.to(dev()), 3, None, None, None, 0, 0,
               MEAN, STD, out2, None, B, 128, 128, 2, stream_ptr())
    assert torch.equal(out2.cpu(), _stem_layout_ref(ref2.view(-1, 5, 128, 128)))
    out3 = torch.empty_like(out2)
    lib().call('vpd_nchw_to_stem', ref2.view(-1, 5, 128, 128).to(dev()), out3, B * 2, 5, 128,
               128, stream_ptr())
    assert torch.equal(out3.cpu(), out2.cpu())


# ------------------------------------------------------------------ K5 AdamW
@pytest.mark.parametrize('n', [1, 3, 4099, 1 << 20])
def test_adamw_matches_oracle(n):
    g = torch.Generator().manual_seed(n)
    p = torch.randn(n, generator=g)
    m = torch.zeros(n)
    v = torch.zeros(n)
    dp, dm, dv = p.to(dev()), m.to(dev()), v.to(dev())
    pn, mn, vn = p.numpy(), m.numpy(), v.numpy()
    for t in range(1, 5):
        gr = torch.randn(n, generator=g) * (10.0 ** (t - 2))
        pn, mn, vn = student_ref.adamw_step_numpy(pn, gr.numpy(), mn, vn, t)
        lib().call('vpd_adamw', dp, gr.to(dev()), dm, dv, n, 5e-4, 0.9, 0.999, 1e-8, 0.01, t,
                   1.0, stream_ptr())
    assert np.array_equal(dm.cpu().numpy().view(np.int32), mn.view(np.int32))
    assert np.array_equal(dv.cpu().numpy().view(np.int32), vn.view(np.int32))
    assert np.array_equal(dp.cpu().numpy().view(np.int32), pn.view(np.int32))


def test_adamw_matches_torch_optimizer():
    n = 100003
    g = torch.Generator().manual_seed(5)
    p = torch.randn(n, generator=g)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5e-4)
    dp = p.to(dev())
    dm = torch.zeros(n, device=dev())
    dv = torch.zeros(n, device=dev())
    for t in range(1, 4):
        gr = torch.randn(n, generator=g)
        ref.grad = gr.clone()
        opt.step()
        lib().call('vpd_adamw', dp, gr.to(dev()), dm, dv, n, 5e-4, 0.9, 0.999, 1e-8, 0.01, t,
                   1.0, stream_ptr())
    diff = (dp.cpu() - ref.detach()).abs().max().item()
    assert diff <= 2.0 ** -21, diff


# ------------------------------------------------------------------ K2 convs
def _conv_case(N, H, W, Cin, Cout, k, stride, pad, seed, affine=False, residual=False,
               relu=False, stats=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((N, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    xb = nhwc_bf16(x).to(dev())
    wd = w.to(dev())
    w_tap = torch.empty((k * k, Cout, Cin), device=dev(), dtype=torch.bfloat16)
    wT_tap = torch.empty((k * k, Cin, Cout), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_pack_conv_weight', wd, w_tap, wT_tap, Cout, Cin, k, stream_ptr())
    # the packed mirrors are opaque (pre-tiled, swizzled smem images); as a multiset they
    # are exactly the bf16-rounded weights
    assert torch.equal(w_tap.flatten().float().sort().values.cpu(),
                       w.to(torch.bfloat16).flatten().float().sort().values)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    y = torch.full((N, Ho, Wo, Cout), float('nan'), device=dev(), dtype=torch.bfloat16)
    scale = shift = res = st = None
    if affine:
        scale = (torch.rand(Cout, generator=g) + 0.5).to(dev())
        shift = torch.randn(Cout, generator=g).to(dev())
    if residual:
        res = nhwc_bf16(torch.randn((N, Cout, Ho, Wo), generator=g)).to(dev())
    if stats:
        st = acc_zeros((2, Cout), dev())
    lib().call('vpd_conv2d_fwd', xb, w_tap, y, N, H, W, Cin, Cout, k, stride, pad, scale, shift,
               res, int(relu), st, stream_ptr())
    torch.cuda.synchronize()
    ref = F.conv2d(nchw_f32(xb), w.to(torch.bfloat16).float().to(dev()), stride=stride, padding=pad)
    if affine:
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if residual:
        ref = ref + nchw_f32(res)
    if relu:
        ref = ref.relu()
    got = nchw_f32(y)
    return got, ref, st, (xb, w_tap, wT_tap)


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad
    (2, 32, 32, 64, 64, 3, 1, 1),
    (3, 16, 16, 128, 128, 3, 1, 1),
    (4, 8, 8, 256, 256, 3, 1, 1),
    (9, 4, 4, 512, 512, 3, 1, 1),      # tn = 8, batch tail
    (2, 32, 32, 64, 128, 3, 2, 1),     # strided 3x3
    (2, 32, 32, 64, 128, 1, 2, 0),     # downsample 1x1
    (3, 16, 16, 128, 256, 3, 2, 1),
    (5, 8, 8, 256, 512, 1, 2, 0),
    (1, 24, 40, 64, 64, 3, 1, 1),      # ragged spatial tiles
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d_fwd_matches_fp32(case):
    got, ref, _, _ = _conv_case(*case, seed=1)
    msg = report('conv_fwd{}'.format(case), got, ref)
    assert torch.isfinite(got).all(), msg
    assert rel_err(got, ref) < 6e-3, msg


def test_conv2d_fwd_epilogue_and_stats():
    got, ref, st, _ = _conv_case(8, 16, 16, 128, 128, 3, 1, 1, seed=2, affine=True,
                                 residual=True, relu=True)
    assert rel_err(got, ref) < 6e-3, report('conv_epi', got, ref)
    got, ref, st, _ = _conv_case(8, 32, 32, 64, 64, 3, 1, 1, seed=3, stats=True)
    assert rel_err(got, ref) < 6e-3, report('conv_stats', got, ref)
    s = got.double().sum((0, 2, 3))
    s2 = (got.double() ** 2).sum((0, 2, 3))
    assert torch.allclose(acc_to_f64(st)[0], s, rtol=1e-5, atol=1e-3)
    assert torch.allclose(acc_to_f64(st)[1], s2, rtol=1e-5, atol=1e-3)
    got, ref, st, _ = _conv_case(3, 8, 8, 256, 512, 3, 1, 1, seed=4, stats=True)
    s = got.double().sum((0, 2, 3))
    s2 = (got.double() ** 2).sum((0, 2, 3))
    assert torch.allclose(acc_to_f64(st)[0], s, rtol=1e-5, atol=1e-3)
    assert torch.allclose(acc_to_f64(st)[1], s2, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('N,H,W,Cimg', [(2, 128, 128, 5), (3, 64, 64, 3), (9, 32, 32, 5)])
def test_stem_conv_fwd(N, H, W, Cimg):
    g = torch.Generator().manual_seed(7)
    x = torch.randn((N, Cimg, H, W), generator=g)
    w = torch.randn((64, Cimg, 7, 7), generator=g) * 0.1
    xs = torch.empty((N,) + _stem_cells(H, W) + (64,), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_nchw_to_stem', x.to(dev()), xs, N, Cimg, H, W, stream_ptr())
    ws = torch.empty((20, 64, 64), device=dev(), dtype=torch.bfloat16)   # 8 + 12 class taps
    lib().call('vpd_pack_stem_weight', w.to(dev()), ws, Cimg, stream_ptr())
    y = torch.full((N, H // 2, W // 2, 64), float('nan'), device=dev(), dtype=torch.bfloat16)
    st = acc_zeros((2, 64), dev())
    lib().call('vpd_stem_conv_fwd', xs, ws, y, N, H, W, None, None, 0, st, stream_ptr())
    ref = F.conv2d(x.to(torch.bfloat16).float().to(dev()), w.to(torch.bfloat16).float().to(dev()),
                   stride=2, padding=3)
    got = nchw_f32(y)
    msg = report('stem{}'.format((N, H, W, Cimg)), got, ref)
    assert torch.isfinite(got).all(), msg
    assert rel_err(got, ref) < 6e-3, msg
    assert torch.allclose(acc_to_f64(st)[0], got.double().sum((0, 2, 3)), rtol=1e-5, atol=1e-3)


DGRAD_CASES = [
    (2, 32, 32, 64, 64, 3, 1, 1),
    (3, 16, 16, 128, 128, 3, 1, 1),
    (9, 4, 4, 512, 512, 3, 1, 1),
    (2, 32, 32, 64, 128, 3, 2, 1),
    (3, 16, 16, 128, 256, 3, 2, 1),
    (5, 8, 8, 256, 512, 3, 2, 1),
]


@pytest.mark.parametrize('case', DGRAD_CASES)
@pytest.mark.parametrize('extras', [False, True])
def test_conv2d_dgrad_matches_autograd(case, extras):
    N, H, W, Cin, Cout, k, stride, pad = case
    g = torch.Generator().manual_seed(11)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    w = (torch.randn((Cout, Cin, k, k), generator=g) / (Cout * k * k) ** 0.5)
    dy = nhwc_bf16(torch.randn((N, Cout, Ho, Wo), generator=g)).to(dev())
    w_tap = torch.empty((k * k, Cout, Cin), device=dev(), dtype=torch.bfloat16)
    wT_tap = torch.empty((k * k, Cin, Cout), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_pack_conv_weight', w.to(dev()), w_tap, wT_tap, Cout, Cin, k, stream_ptr())
    wq = w.to(torch.bfloat16).float().to(dev())
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), wq, nchw_f32(dy), stride=stride, padding=pad)
    res = dy_ds = wT_ds = None
    cout_ds = 0
    if extras:
        res = nhwc_bf16(torch.randn((N, Cin, H, W), generator=g)).to(dev())
        ref = ref + nchw_f32(res)
        if stride == 2:
            cout_ds = Cout
            wds = torch.randn((Cout, Cin, 1, 1), generator=g) / Cout ** 0.5
            dy_ds = nhwc_bf16(torch.randn((N, Cout, Ho, Wo), generator=g)).to(dev())
            wT_ds = torch.empty((1, Cin, Cout), device=dev(), dtype=torch.bfloat16)
            lib().call('vpd_pack_conv_weight', wds.to(dev()), None, wT_ds, Cout, Cin, 1, stream_ptr())
            ref = ref + torch.nn.grad.conv2d_input(
                (N, Cin, H, W), wds.to(torch.bfloat16).float().to(dev()), nchw_f32(dy_ds), stride=2)
    dx = torch.full((N, H, W, Cin), float('nan'), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_conv2d_dgrad', dy, wT_tap, dx, N, H, W, Cin, Cout, k, stride, pad, res, dy_ds,
               wT_ds, cout_ds, stream_ptr())
    got = nchw_f32(dx)
    msg = report('dgrad{}{}'.format(case, extras), got, ref)
    assert torch.isfinite(got).all(), msg
    assert rel_err(got, ref) < 6e-3, msg


WGRAD_CASES = [
    (2, 32, 32, 64, 64, 3, 1, 1),
    (16, 32, 32, 64, 64, 3, 1, 1),
    (3, 16, 16, 128, 128, 3, 1, 1),
    (4, 8, 8, 256, 256, 3, 1, 1),
    (9, 4, 4, 512, 512, 3, 1, 1),
    (2, 32, 32, 64, 128, 3, 2, 1),
    (2, 32, 32, 64, 128, 1, 2, 0),
    (3, 16, 16, 128, 256, 3, 2, 1),
    (5, 8, 8, 256, 512, 1, 2, 0),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_conv2d_wgrad_matches_autograd(case):
    N, H, W, Cin, Cout, k, stride, pad = case
    g = torch.Generator().manual_seed(13)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    x = nhwc_bf16(torch.randn((N, Cin, H, W), generator=g)).to(dev())
    dy = nhwc_bf16(torch.randn((N, Cout, Ho, Wo), generator=g)).to(dev())
    dw = torch.zeros((k * k, Cout, Cin), device=dev(), dtype=torch.float32)
    lib().call('vpd_conv2d_wgrad', x, dy, dw, N, H, W, Cin, Cout, k, stride, pad, stream_ptr())
    ref = torch.nn.grad.conv2d_weight(nchw_f32(x), (Cout, Cin, k, k), nchw_f32(dy),
                                      stride=stride, padding=pad)
    got = dw.view(k, k, Cout, Cin).permute(2, 3, 0, 1).contiguous()
    msg = report('wgrad{}'.format(case), got, ref, tol=2e-3)
    assert torch.isfinite(got).all(), msg
    assert rel_err(got, ref) < 2e-3, msg
    # accumulation semantics: a second call doubles the result
    lib().call('vpd_conv2d_wgrad', x, dy, dw, N, H, W, Cin, Cout, k, stride, pad, stream_ptr())
    got2 = dw.view(k, k, Cout, Cin).permute(2, 3, 0, 1)
    assert rel_err(got2, 2 * ref) < 2e-3


@pytest.mark.parametrize('N,H,W,Cimg', [(2, 128, 128, 5), (5, 32, 32, 3)])
def test_stem_conv_wgrad(N, H, W, Cimg):
    g = torch.Generator().manual_seed(17)
    x = torch.randn((N, Cimg, H, W), generator=g)
    xs = torch.empty((N,) + _stem_cells(H, W) + (64,), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_nchw_to_stem', x.to(dev()), xs, N, Cimg, H, W, stream_ptr())
    dy = nhwc_bf16(torch.randn((N, 64, H // 2, W // 2), generator=g)).to(dev())
    dw = torch.zeros((7, 64, 64), device=dev(), dtype=torch.float32)
    lib().call('vpd_stem_conv_wgrad', xs, dy, dw, N, H, W, stream_ptr())
    ref = torch.nn.grad.conv2d_weight(x.to(torch.bfloat16).float().to(dev()), (64, Cimg, 7, 7),
                                      nchw_f32(dy), stride=2, padding=3)
    # dw[kh][co][kw*8+c] -> [co][c][kh][kw]
    full = dw.view(7, 64, 8, 8)                       # kh, co, kw, c
    got = full[:, :, :7, :Cimg].permute(1, 3, 0, 2).contiguous()
    msg = report('stem_wgrad', got, ref, tol=2e-3)
    assert rel_err(got, ref) < 2e-3, msg
    assert full[:, :, 7, :].abs().max().item() == 0.0          # kw == 7 pad never written
    assert full[:, :, :7, Cimg:].abs().max().item() == 0.0      # zero input channels


@pytest.mark.parametrize('case', [(4, 32, 32, 64, 64, 3, 1, 1), (3, 16, 16, 128, 128, 3, 1, 1),
                                  (5, 8, 8, 256, 256, 3, 1, 1), (2, 32, 32, 64, 128, 3, 2, 1),
                                  (9, 4, 4, 512, 512, 3, 1, 1)])
def test_conv2d_dgrad_with_fused_bn_backward_reduction(case):
    N, H, W, Cin, Cout, k, stride, pad = case
    g = torch.Generator().manual_seed(19)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    w = torch.randn((Cout, Cin, k, k), generator=g) / (Cout * k * k) ** 0.5
    dy = nhwc_bf16(torch.randn((N, Cout, Ho, Wo), generator=g)).to(dev())
    wT_tap = torch.empty((k * k, Cin, Cout), device=dev(), dtype=torch.bfloat16)
    lib().call('vpd_pack_conv_weight', w.to(dev()), None, wT_tap, Cout, Cin, k, stream_ptr())
    res = nhwc_bf16(torch.randn((N, Cin, H, W), generator=g)).to(dev())
    z = nhwc_bf16(torch.randn((N, Cin, H, W), generator=g).relu()).to(dev())
    y = nhwc_bf16(torch.randn((N, Cin, H, W), generator=g) * 2 + 0.5).to(dev())
    mean = (torch.randn(Cin, generator=g) * 0.3).to(dev())
    rstd = (torch.rand(Cin, generator=g) + 0.5).to(dev())
    sums = acc_zeros((2, Cin), dev())
    dx = torch.full((N, H, W, Cin), float('nan'), device=dev(), dtype=torch.bfloat16)
    zmask = torch.zeros((N, H, W, Cin // 8), device=dev(), dtype=torch.uint8)
    lib().call('vpd_relu_bitmask', z, zmask, N * H * W, Cin, stream_ptr())
    lib().call('vpd_conv2d_dgrad_bnfused', dy, wT_tap, dx, N, H, W, Cin, Cout, k, stride, pad, res,
               zmask, y, mean, rstd, sums, stream_ptr())
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w.to(torch.bfloat16).float().to(dev()),
                                     nchw_f32(dy), stride=stride, padding=pad) + nchw_f32(res)
    ref = ref * (nchw_f32(z) > 0)
    got = nchw_f32(dx)
    assert rel_err(got, ref) < 6e-3, report('dgrad_fused{}'.format(case), got, ref)
    assert ((got == 0) | (nchw_f32(z) > 0)).all()
    xhat = (nchw_f32(y) - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
    s0 = got.double().sum((0, 2, 3))
    s1 = (got.double() * xhat.double()).sum((0, 2, 3))
    assert torch.allclose(acc_to_f64(sums)[0], s0, rtol=1e-4, atol=1e-2)
    assert torch.allclose(acc_to_f64(sums)[1], s1, rtol=1e-4, atol=1e-2)


# --------------------------------------------------- K1 masked-noise augmentation
def test_assemble_masked_noise_bit_exact_with_given_noise():
    """single_frame.py:179-191 with the noise tensor supplied: img += noise where the mask
    PNG byte is not 0, only on the frames whose coin says so, before the flip."""
    from vpd_b200.assemble import assemble_batch
    B, H, W = 6, 32, 32
    g = torch.Generator().manual_seed(71)
    rgb, flow = synth.crops(B, seed=72, height=H, width=W)
    teach = synth.teacher(B, seed=73, emb_dim=8, motion=True)
    flips = synth.flips(B, seed=74)
    mask = (torch.rand((B, H, W), generator=g) < 0.4).to(torch.uint8) * 255   # 0 = background
    noise = torch.randn((B, 3, H, W), generator=g) * (0.05 ** 0.5)
    on = torch.tensor([1, 0, 1, 1, 0, 1], dtype=torch.uint8)
    ref_img, ref_emb = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), teach.numpy(),
                                                flips.numpy(), *synth.FS_MEAN_STD,
                                                mask_u8=mask.numpy(), noise=noise.numpy(),
                                                noise_on=on.numpy())
    out = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, flip=flips.to(dev()),
                         teacher=teach.to(dev()), mask=mask.to(dev()), noise_on=on.to(dev()),
                         noise=noise.to(dev()))
    assert torch.equal(out['img'].cpu(), ref_img)
    assert torch.equal(out['emb'].cpu(), ref_emb)
    # and it really changed something, only where it may
    plain = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, flip=flips.to(dev()))
    d = (out['img'] - plain['img']).cpu()
    assert d[1].abs().max() == 0 and d[4].abs().max() == 0 and d[:, 3:].abs().max() == 0
    assert d[0, :3].abs().max() > 0


def test_assemble_masked_noise_device_rng_statistics():
    """Device Philox noise: zero on masked-out pixels / switched-off frames / flow planes,
    N(0, sd^2) elsewhere, reproducible per seed, identical in both output layouts."""
    from vpd_b200.assemble import assemble_batch, assemble_stem
    B, H, W = 8, 64, 64
    rgb, flow = synth.crops(B, seed=81, height=H, width=W)
    mask = torch.zeros((B, H, W), dtype=torch.uint8)
    mask[:, :, : W // 2] = 7                                     # left half = person pixels
    args = dict(mask=mask.to(dev()), noise_sd=0.05 ** 0.5)
    base = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD)['img']
    a = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, seed=5, **args)['img']
    b = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, seed=5, **args)['img']
    c = assemble_batch(rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, seed=6, **args)['img']
    assert torch.equal(a, b) and not torch.equal(a, c)
    d = (a - base).cpu()
    assert d[:, :3, :, W // 2:].abs().max() == 0 and d[:, 3:].abs().max() == 0
    nz = d[:, :3, :, : W // 2].flatten().double()
    assert abs(nz.mean().item()) < 4e-3 and abs(nz.std().item() - 0.05 ** 0.5) < 4e-3
    assert abs((nz ** 4).mean().item() / nz.var().item() ** 2 - 3.0) < 0.15   # Gaussian kurtosis
    # the network-layout kernel draws the same noise (then rounds to bf16)
    Hs, Ws = _stem_cells(H, W)
    stem = torch.zeros((B, Hs, Ws, 64), device=dev(), dtype=torch.bfloat16)
    assemble_stem(stem, rgb.to(dev()), flow.to(dev()), synth.FS_MEAN_STD, seed=5, **args)
    got = _stem_layout_unpack(stem, H, W)[:, 3:3 + H, 3:3 + W, :5].permute(0, 3, 1, 2).float()
    assert torch.equal(got, a.to(torch.bfloat16).float())


# ----------------------------------------------------------------- fused SGD
@pytest.mark.parametrize('cfg', [dict(momentum=0.0, weight_decay=0.0),
                                 dict(momentum=0.9, weight_decay=1e-4),
                                 dict(momentum=0.9, weight_decay=5e-4, nesterov=True),
                                 dict(momentum=0.8, dampening=0.1, weight_decay=0.0)])
def test_sgd_matches_torch(cfg):
    n = 4099
    g = torch.Generator().manual_seed(91)
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone().to(dev()))
    opt = torch.optim.SGD([ref_p], lr=0.05, foreach=False, **cfg)
    dp = p0.clone().to(dev())
    buf = torch.zeros(n, device=dev()) if cfg.get('momentum', 0) else None
    for t in range(1, 5):
        grad = torch.randn(n, generator=g).to(dev())
        ref_p.grad = grad.clone()
        opt.step()
        lib().call('vpd_sgd', dp, grad, buf, n, 0.05, cfg.get('momentum', 0.0),
                   cfg.get('dampening', 0.0), cfg.get('weight_decay', 0.0),
                   int(cfg.get('nesterov', False)), int(t == 1), 1.0, stream_ptr())
        # same operations up to fma contraction: a few ulp at most
        assert torch.allclose(dp, ref_p.detach(), rtol=2e-6, atol=2e-7), (cfg, t)
