"""Oracle: the keypoint (VIPE*) teacher's eval forward, torch CPU fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates from a state_dict:
  * `FCResNet.forward`        models/module.py:192-204 (Linear + ReLU, blocks, last Linear)
  * `FcResidualBlock.forward` models/module.py:159-177: `x2 = block(x); return x2 - x` with
    block = Linear, BatchNorm1d, ReLU, Dropout, Linear, BatchNorm1d, ReLU, Dropout; in eval
    mode BatchNorm1d uses the running statistics (eps 1e-5) and Dropout is the identity
  * `Keypoint_EmbeddingModel._predict` / `embed` models/keypoint.py:128-160: eval, no_grad,
    `pose.view(n, -1)`, result as a host numpy array.
Pinned by tests/golden/keypoint.npz: `embed()` of the UNMODIFIED reference classes on a seeded
initialisation with perturbed BatchNorm state (oracle/gen_golden.py::gen_keypoint).
"""
import numpy as np
import torch
import torch.nn.functional as F


def perturb_bn(sd, seed):
    """A fresh BatchNorm1d is the identity in eval mode; give every BN non-trivial affine
    parameters and running statistics (deterministic, shared by the generator and the tests)."""
    g = torch.Generator().manual_seed(seed)
    out = type(sd)()
    for k, v in sd.items():
        if k.endswith('running_mean'):
            v = torch.randn(v.shape, generator=g) * 0.3
        elif k.endswith('running_var'):
            v = torch.rand(v.shape, generator=g) * 1.5 + 0.25
        elif '.block.1.' in k or '.block.5.' in k:
            if k.endswith('.weight'):
                v = torch.rand(v.shape, generator=g) + 0.5
            elif k.endswith('.bias'):
                v = torch.randn(v.shape, generator=g) * 0.2
        out[k] = v
    return out


def synth_poses(n, seed, joints=13):
    """normalised 2-D skeletons: (x, y, confidence) per joint, roughly unit scale"""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn((n, joints, 3), generator=g) * 0.5
    p[:, :, 2] = torch.rand((n, joints), generator=g)
    return p


def fcresnet_eval(sd, x, num_blocks):
    x = torch.as_tensor(x, dtype=torch.float32)
    h = F.relu(F.linear(x, sd['layers.0.weight'], sd['layers.0.bias']))
    for i in range(num_blocks):
        p = 'layers.{}.block'.format(2 + i)
        z = h
        for lin, bn in ((0, 1), (4, 5)):
            z = F.linear(z, sd['{}.{}.weight'.format(p, lin)], sd['{}.{}.bias'.format(p, lin)])
            b = '{}.{}'.format(p, bn)
            z = F.batch_norm(z, sd[b + '.running_mean'], sd[b + '.running_var'], sd[b + '.weight'],
                             sd[b + '.bias'], training=False, eps=1e-5)
            z = F.relu(z)
        h = z - h
    last = 'layers.{}'.format(2 + num_blocks)
    return F.linear(h, sd[last + '.weight'], sd[last + '.bias'])


def embed(sd, pose, num_blocks):
    pose = torch.as_tensor(np.asarray(pose), dtype=torch.float32)
    if pose.dim() == 2:
        pose = pose.unsqueeze(0)
    with torch.no_grad():
        return fcresnet_eval(sd, pose.reshape(pose.shape[0], -1), num_blocks).numpy()
