"""Diagnostic (not a test): per-CTA phase timing of the generic implicit-GEMM conv kernel.

    python tests/diag_conv_trace.py            # VPD_PAIR=0/1 to compare modes
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200._lib import lib  # noqa: E402

CASES = [  # N, H, W, Cin, Cout (3x3 stride 1, train-mode statistics on)
    (256, 32, 32, 64, 64),
    (256, 16, 16, 128, 128),
    (256, 8, 8, 256, 256),
    (256, 4, 4, 512, 512),
]


def main():
    dev = torch.device('cuda:0')
    s = torch.cuda.current_stream().cuda_stream
    trace = torch.zeros((148, 16), device=dev, dtype=torch.int64)
    use_stats = os.environ.get('DIAG_STATS', '1') == '1'
    for (N, H, W, Cin, Cout) in CASES:
        x = torch.randn((N, H, W, Cin), device=dev).to(torch.bfloat16)
        w = torch.randn((Cout, Cin, 3, 3), device=dev) * 0.02
        w_tap = torch.empty((9, Cout, Cin), device=dev, dtype=torch.bfloat16)
        wT = torch.empty((9, Cin, Cout), device=dev, dtype=torch.bfloat16)
        lib().call('vpd_pack_conv_weight', w, w_tap, wT, Cout, Cin, 3, s)
        y = torch.empty((N, H, W, Cout), device=dev, dtype=torch.bfloat16)
        st = torch.zeros((2, Cout, 2), device=dev, dtype=torch.int64)

        z = torch.randn((N, H, W, Cin), device=dev).to(torch.bfloat16)
        yf = torch.randn((N, H, W, Cin), device=dev).to(torch.bfloat16)
        mean = torch.zeros(Cin, device=dev)
        rstd = torch.ones(Cin, device=dev)
        bs = torch.zeros((2, Cin, 2), device=dev, dtype=torch.int64)
        dgrad = os.environ.get('DIAG_MODE', 'fwd') == 'dgrad'
        alias = os.environ.get('DIAG_ALIAS', '0')
        if alias == '1':      # z and y share one tensor (half the unique bytes)
            yf = z
        elif alias == '2':    # both alias the conv input (already streamed by TMA)
            z = yf = x

        zmask = torch.zeros((N, H, W, Cin // 8), device=dev, dtype=torch.uint8)
        lib().call('vpd_relu_bitmask', z, zmask, N * H * W, Cin, s)

        def run():
            if dgrad:   # Cin == Cout in every case: x doubles as dy, y as dx
                lib().call('vpd_conv2d_dgrad_bnfused', x, wT, y, N, H, W, Cin, Cout, 3, 1, 1, None,
                           zmask, yf, mean, rstd, bs, s)
            else:
                lib().call('vpd_conv2d_fwd', x, w_tap, y, N, H, W, Cin, Cout, 3, 1, 1, None, None,
                           None, 0, st if use_stats else None, s)
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        flops = 2.0 * N * H * W * Cout * Cin * 9
        trace.zero_()
        lib().call('vpd_conv_trace', trace)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        lib().call('vpd_conv_trace', None)
        t = trace.cpu()
        t = t[t[:, 1] != 0]
        if t.shape[0] == 0:   # halo kernels do not trace
            print('case N%d %dx%d %d->%d: %.1f us/launch back-to-back, %.0f TFLOP/s (no trace)'
                  % (N, H, W, Cin, Cout, us, flops / us * 1e-6))
            continue
        g0 = t[:, 0] - t[:, 0].min()
        d = lambda a, b: (t[:, b] - t[:, a]).float()  # noqa: E731
        fmt = lambda v: '%7.0f/%7.0f/%7.0f' % (v.min(), v.median(), v.max())  # noqa: E731
        print('case N%d %dx%d %d->%d: %.1f us/launch back-to-back, %.0f TFLOP/s, %d CTAs traced'
              % (N, H, W, Cin, Cout, us, flops / us * 1e-6, t.shape[0]))
        print('  entry spread (ns)       min/med/max', fmt(g0.float()))
        print('  dependency wait (cyc)              ', fmt(d(1, 2)))
        print('  first operands landed              ', fmt(d(2, 3)))
        print('  mainloop (first full->last issue)  ', fmt(d(3, 4)))
        print('  last issue -> first accum seen     ', fmt(d(4, 5)), '(multi-tile CTAs: negative)')
        print('  epilogue (first accum -> done)     ', fmt(d(5, 6)))
        print('  total entry -> exit                ', fmt(d(1, 7)))
        span = (t[:, 8].max() - t[:, 0].min()).item() / 1e3
        first_exit = (t[:, 8].min() - t[:, 0].min()).item() / 1e3
        late = t[t[:, 0] > t[:, 0].min() + 2000]   # CTAs that entered after the previous kernel drained
        print('  kernel span first entry -> last exit: %.1f us (first exit at %.1f us); '
              'exit spread %.1f us' % (span, first_exit, (t[:, 8].max() - t[:, 8].min()).item() / 1e3))
        if late.shape[0]:
            print('  late CTAs (%d): entry at +%.1f..%.1f us, entry -> exit %.1f us median'
                  % (late.shape[0], (late[:, 0].min() - t[:, 0].min()).item() / 1e3,
                     (late[:, 0].max() - t[:, 0].min()).item() / 1e3,
                     (late[:, 8] - late[:, 0]).float().median().item() / 1e3))


if __name__ == '__main__':
    main()
