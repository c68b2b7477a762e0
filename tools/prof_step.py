#!/usr/bin/env python3
"""Short workload for ncu captures (B200_PROFILING.md recipe): a few batch-256 training steps
of the student, one apply batch (500 frames x [orig, flipped]) and the reference-layout K1
kernel, eager (no CUDA-graph replay) so that every kernel is its own launch.

    VPD_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:assemble -c 4 \
        -o gpurun_out/r02_k1 python tools/prof_step.py [steps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('VPD_GRAPH', '0')
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer        # noqa: E402
from vpd_b200.assemble import assemble_stem, assemble_batch            # noqa: E402
from vpd_b200._lib import lib                                          # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    B, POOL = 256, 2048
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    enc = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda')
    tr = ModelTrainer(enc, True)
    opt, _ = tr.get_optimizer(5e-4)
    rgb, flow = synth.crops(POOL, seed=1)
    teach = synth.teacher(POOL, seed=3)
    rgb, flow, teach = rgb.to(dev), flow.to(dev), teach.to(dev)
    g = torch.Generator().manual_seed(4)
    tgt = torch.empty((B, 64), device=dev)
    stem = tr.stem_buffer(B, 128, 128)
    for i in range(steps):
        idx = torch.randint(0, POOL, (B,), generator=g).int().to(dev)
        fl = torch.randint(0, 2, (B,), generator=g).to(torch.uint8).to(dev)
        assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD, flip=fl, teacher=teach, index=idx, tgt=tgt)
        tr.train_step_stem(stem, tgt, B, 128, 128, opt)
        assemble_batch(rgb, flow, synth.FS_MEAN_STD, flip=fl, teacher=teach, index=idx)
    torch.cuda.synchronize()
    enc.eval()
    net = enc._native(128, 128, 1000)
    stem_a = lib().call('vpd_net_stem_input', net.handle)
    for i in range(2):
        assemble_stem(stem_a, rgb[i * 500:(i + 1) * 500], flow[i * 500:(i + 1) * 500],
                      synth.FS_MEAN_STD, k=2)
        enc.embed_stem(stem_a, 1000, 128, 128)
    torch.cuda.synchronize()
    print('prof_step done')


if __name__ == '__main__':
    main()
