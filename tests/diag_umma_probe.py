"""Diagnostic: UMMA descriptor behaviour with row-shifted starts / odd group strides."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200._lib import lib, stream_ptr
dev = torch.device('cuda:0')
rows = 400
src = torch.randn((rows, 64), generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).to(dev)
def run(row_start, sbo, mode):
    out = torch.zeros((128, 64), device=dev)
    lib().call('vpd_umma_probe', src, rows, row_start, sbo, mode, out, stream_ptr())
    torch.cuda.synchronize()
    m = torch.arange(128, device=dev)
    exp_rows = row_start + (m // 8) * (sbo // 128) + m % 8
    exp = src[exp_rows].float()
    ok = (out == exp).all(dim=1)
    return ok.float().mean().item(), ok[:16].int().tolist()
for sbo in (1024, 2048, 1280, 1152, 2176):
    for row_start in (0, 8, 1, 3, 10, 17):
        for mode in (0, 1):
            frac, first = run(row_start, sbo, mode)
            print('sbo {:5d} row_start {:3d} base_offset_mode {} -> rows correct {:.3f} {}'.format(
                sbo, row_start, mode, frac, first if frac < 1 else ''))
