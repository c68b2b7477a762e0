"""Diagnostic (not a test): per-layer timing of the weight-gradient kernels at batch 256.
    python tests/diag_wgrad.py            (env: VPD_WGRAD_HALO=0|1, VPD_WGRAD_DBG bits, VPD_WGRAD_SPLITS)
Prints us per launch (CUDA events around 20 back-to-back launches) and TFLOP/s per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200._lib import lib, stream_ptr    # noqa: E402

SHAPES = [(256, 32, 32, 64, 64), (256, 16, 16, 128, 128), (256, 8, 8, 256, 256), (256, 4, 4, 512, 512)]


def main():
    dev = torch.device('cuda:0')
    out = []
    for N, H, W, Cin, Cout in SHAPES:
        x = torch.randn((N, H, W, Cin), device=dev).to(torch.bfloat16)
        dy = torch.randn((N, H, W, Cout), device=dev).to(torch.bfloat16)
        dw = torch.zeros((9, Cout, Cin), device=dev)
        call = lambda: lib().call('vpd_conv2d_wgrad', x, dy, dw, N, H, W, Cin, Cout, 3, 1, 1, stream_ptr())
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2.0 * N * H * W * Cin * Cout * 9
        out.append('{}x{}x{} {}->{}: {:.1f} us {:.0f} TF/s'.format(N, H, W, Cin, Cout, us, fl / us / 1e6))
    print('HALO={} DBG={} SPLITS={} | '.format(os.environ.get('VPD_WGRAD_HALO', '1'), os.environ.get('VPD_WGRAD_DBG', '0'),
                                              os.environ.get('VPD_WGRAD_SPLITS', '-')) + ' | '.join(out))


if __name__ == '__main__':
    main()
