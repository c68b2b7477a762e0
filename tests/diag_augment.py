"""Timing of the augmented-batch kernel (K1a) at the training batch size; not a test.
    python tests/diag_augment.py  -> one JSON line (also written to gpurun_out/aug_timing.json)"""
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200 import augment, synth                      # noqa: E402
from vpd_b200.assemble import assemble_batch, assemble_batch_aug   # noqa: E402


def main():
    dev = torch.device('cuda:0')
    B, P, dim = 256, 2048, 128
    rgb, flow = synth.crops(P, seed=1)
    teach = synth.teacher(P, seed=3)
    mask = (torch.rand((P, dim, dim)) > 0.5).to(torch.uint8) * 255
    rgb, flow, teach, mask = rgb.to(dev), flow.to(dev), teach.to(dev), mask.to(dev)
    random.seed(0)
    torch.manual_seed(0)
    params = [augment.draw_batch(B, P, dim, dim, has_mask=torch.ones(P, dtype=torch.bool)).to(dev)
              for _ in range(8)]
    res = {}
    for name, fn in (
            ('aug_full', lambda p: assemble_batch_aug(rgb, flow, synth.FS_MEAN_STD, p, teacher=teach,
                                                      mask=mask, seed=1)),
            ('aug_no_noise', lambda p: assemble_batch_aug(rgb, flow, synth.FS_MEAN_STD, p,
                                                          teacher=teach)),
            ('plain', lambda p: assemble_batch(rgb, flow, synth.FS_MEAN_STD, flip=p.flip,
                                               teacher=teach, index=p.index))):
        for p in params[:3]:
            fn(p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(40):
            fn(params[it % 8])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        by = B * (dim * dim * (3 + 2) + 5 * dim * dim * 4)     # u8 rgb + flow-xy in, fp32 out
        res[name] = {'ms': ms, 'frames_per_s': B / ms * 1e3, 'algorithmic_GBps': by / ms / 1e6}
    print(json.dumps(res))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/aug_timing.json', 'w') as fp:
        json.dump(res, fp)


if __name__ == '__main__':
    main()
