"""Seeded synthetic inputs of the shapes BASELINE.json names.

CPU torch.Generator streams, so every box with the same torch build produces
identical bytes: uint8 crops ~U{0..255}, flow PNG-style arrays (ch0 = flow-x,
ch1 = flow-y, ch2 = 128 like raft/flow.py:80-84 writes), flip bits
~Bernoulli(.5) and teacher embeddings ~N(0,1) with rows (unflipped, flipped)
as apply_vipe_model.py stores them.
"""
import torch

FS_MEAN_STD = (
    (0.5747710337842444, 0.5644043210903272, 0.6334494151377134),
    (0.21349823115367886, 0.21827191146692457, 0.20393919008463163),
)  # vpd_dataset/common.py:19-22 ('fs')


def crops(n, seed, height=128, width=128):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (n, height, width, 3), generator=g, dtype=torch.uint8)
    flow = torch.randint(0, 256, (n, height, width, 3), generator=g, dtype=torch.uint8)
    flow[..., 2] = 128
    return rgb, flow


def flips(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 2, (n,), generator=g, dtype=torch.uint8)


def teacher(n, seed, emb_dim=32, motion=True):
    g = torch.Generator().manual_seed(seed)
    d = 2 * emb_dim if motion else emb_dim
    return torch.randn((n, 2, d), generator=g, dtype=torch.float32)
