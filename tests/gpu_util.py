import os

import numpy as np
import torch

from vpd_b200._lib import lib, stream_ptr

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def dev():
    return torch.device('cuda:0')


def nhwc_bf16(x_nchw):
    """fp32 NCHW -> bf16 NHWC contiguous (round to nearest even)."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw_f32(x_nhwc_bf16):
    return x_nhwc_bf16.float().permute(0, 3, 1, 2).contiguous()


def rel_err(got, ref):
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def report(name, got, ref, tol=6e-3):
    """Error summary; dumps a small diagnostic file for offline debugging."""
    d = (got - ref).abs()
    if rel_err(got, ref) < tol and torch.isfinite(got).all():
        return '{}: rel_l2={:.3e}'.format(name, rel_err(got, ref))
    msg = '{}: rel_l2={:.3e} max_abs={:.3e} ref_max={:.3e} frac_bad={:.4f}'.format(
        name, rel_err(got, ref), d.max().item(), ref.abs().max().item(),
        (d > 0.05 * (ref.abs() + 0.05 * ref.abs().max())).float().mean().item())
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, 'diag.log'), 'a') as fp:
        fp.write(msg + '\n')
        if got.dim() == 4:
            bad = (d > 0.05 * (ref.abs() + 0.05 * ref.abs().max()))
            fp.write('  bad per n: {}\n'.format(bad.float().mean((1, 2, 3)).tolist()[:16]))
            fp.write('  bad per c (first 32): {}\n'.format(
                [round(v, 3) for v in bad.float().mean((0, 2, 3)).tolist()[:32]]))
            fp.write('  bad per h: {}\n'.format(
                [round(v, 3) for v in bad.float().mean((0, 1, 3)).tolist()[:32]]))
            fp.write('  bad per w: {}\n'.format(
                [round(v, 3) for v in bad.float().mean((0, 1, 2)).tolist()[:32]]))
            fp.write('  got[0,:4,0,:4]={}\n  ref[0,:4,0,:4]={}\n'.format(
                got[0, :4, 0, :4].tolist(), ref[0, :4, 0, :4].tolist()))
    return msg
