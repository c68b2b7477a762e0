"""Diagnostic (not a test): a small pass over every kernel family for compute-sanitizer.

    compute-sanitizer --tool memcheck python tests/diag_sanitize.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('VPD_GRAPH', '0')
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer          # noqa: E402
from vpd_b200.assemble import assemble_apply, assemble_stem             # noqa: E402
from vpd_b200._lib import lib                                          # noqa: E402
from vpd_b200.train import PoolLoader                                  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    B, P = 10, 24
    rgb, flow = synth.crops(P, seed=1)
    teach = synth.teacher(P, seed=3)
    rgb, flow, teach = rgb.to(dev), flow.to(dev), teach.to(dev)
    torch.manual_seed(0)
    enc = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda')
    tr = ModelTrainer(enc, True)
    opt, _ = tr.get_optimizer(5e-4)
    for raw in (True, False):          # K1 stem kernel / reference-layout kernel + conversion
        ld = PoolLoader(rgb, flow, teach, synth.FS_MEAN_STD, B, 2 * B, seed=5, random_flip=True, raw=raw)
        print('epoch loss', tr.epoch(ld, optimizer=opt))
    img = assemble_apply(rgb[:B], flow[:B], synth.FS_MEAN_STD, flip=True)   # reference layout, k = 2
    emb = enc.embed(img.reshape(-1, 5, 128, 128))
    print('embed', emb.shape)
    enc.eval()
    net = enc._native(128, 128, 2 * B)                                      # stem layout, k = 2
    stem = lib().call('vpd_net_stem_input', net.handle)
    assemble_stem(stem, rgb[:B], flow[:B], synth.FS_MEAN_STD, k=2)
    print('embed_stem', tuple(enc.embed_stem(stem, 2 * B, 128, 128).shape))
    torch.cuda.synchronize()
    print('sanitize pass done')


if __name__ == '__main__':
    main()
