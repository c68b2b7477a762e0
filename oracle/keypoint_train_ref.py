"""Oracle: one training step of the keypoint (VIPE*) teacher, torch CPU fp32 + autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with the dropout keep-masks passed in, what `Keypoint_EmbeddingModel.epoch` does for
one zipped batch (models/keypoint.py:38-126): per dataset batch three weight-sharing encoder
passes (pose1 / pose2 / pose_neg) through `FCResNet` in train mode (models/module.py:159-204:
Linear, BatchNorm1d with batch statistics, ReLU, Dropout, twice, then `x2 - x`),
`F.hinge_embedding_loss(norm(e1 - e2), +1, 'sum')`, `sum(hinge(norm(e1 - en), -1, 'none') *
valid)`, the FC pose decoder (models/module.py:230-246) with `weight_3d * F.mse_loss(..., 'sum')`
for pose1 and pose2; the losses of the zipped batch are summed, divided by the number of samples
and back-propagated (:106-111, models/util.py:50-58).

Dropout: ATen's CPU dropout draws `noise = empty_like(x).bernoulli_(1 - p)` per call and computes
`x * noise / (1 - p)`; `replay_masks` re-draws exactly those tensors from torch's global
generator in call order, which is how the golden generator feeds the unmodified reference and
this restatement the same masks (oracle/gen_golden.py::gen_keypoint_train asserts equality).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def replay_masks(n, hidden, num_blocks, passes, p):
    """the keep masks of `passes` consecutive encoder calls: [pass][2 * num_blocks] bool [n, H]"""
    return [[torch.empty(n, hidden).bernoulli_(1 - p) > 0 for _ in range(2 * num_blocks)]
            for _ in range(passes)]


def synth_batch(n, seed, joints=13, tdim=140, with_neg=True, with_3d=True):
    g = torch.Generator().manual_seed(seed)
    b = OrderedDict()
    b['pose1'] = torch.randn((n, joints, 3), generator=g) * 0.5
    b['pose2'] = b['pose1'] + torch.randn((n, joints, 3), generator=g) * 0.05
    if with_neg:
        b['pose_neg'] = torch.randn((n, joints, 3), generator=g) * 0.5
        b['pose_neg_is_valid'] = (torch.rand(n, generator=g) > 0.25).float()
    if with_3d:
        b['kp_features'] = torch.randn((n, tdim // 7, 7), generator=g) * 0.3
    return b


def encoder_train(sd, x, masks, p, num_blocks, buffers):
    """train-mode FCResNet forward; `buffers` (running stats, num_batches_tracked) are updated in
    place like nn.BatchNorm1d does"""
    h = F.relu(F.linear(x, sd['layers.0.weight'], sd['layers.0.bias']))
    for i in range(num_blocks):
        pre = 'layers.{}.block'.format(2 + i)
        z = h
        for j, (lin, bn) in enumerate(((0, 1), (4, 5))):
            z = F.linear(z, sd['{}.{}.weight'.format(pre, lin)], sd['{}.{}.bias'.format(pre, lin)])
            b = '{}.{}'.format(pre, bn)
            z = F.batch_norm(z, buffers[b + '.running_mean'], buffers[b + '.running_var'],
                             sd[b + '.weight'], sd[b + '.bias'], training=True, momentum=0.1,
                             eps=1e-5)
            buffers[b + '.num_batches_tracked'] += 1
            z = F.relu(z)
            z = z * (masks[2 * i + j].to(z.dtype) / (1 - p))
        h = z - h
    last = 'layers.{}'.format(2 + num_blocks)
    return F.linear(h, sd[last + '.weight'], sd[last + '.bias'])


def decoder_fwd(dsd, emb, target, fcn_keys):
    x = emb
    for k in fcn_keys:
        x = F.relu(F.linear(x, dsd[k + '.weight'], dsd[k + '.bias']))
    return F.linear(x, dsd['fc_{}.weight'.format(target)], dsd['fc_{}.bias'.format(target)])


def zipped_step(enc_sd, dec_sd, fcn_keys, zipped, masks, p, num_blocks, weight_3d=1.0):
    """zipped: [(dataset_name, batch dict)], masks: per dataset batch [pass][2*blocks].
    -> ({name: (contra, loss)}, grads {'enc.<key>' / 'dec.<key>': tensor}, updated buffers)"""
    buf_keys = [k for k in enc_sd if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]
    buffers = {k: enc_sd[k].clone() for k in buf_keys}
    params = {('enc.' + k): v.clone().requires_grad_(True) for k, v in enc_sd.items()
              if k not in buf_keys}
    params.update({('dec.' + k): v.clone().requires_grad_(True) for k, v in dec_sd.items()})
    esd = {k[4:]: v for k, v in params.items() if k.startswith('enc.')}
    dsd = {k[4:]: v for k, v in params.items() if k.startswith('dec.')}
    out = OrderedDict()
    batch_loss, batch_n = 0., 0
    for (name, batch), bm in zip(zipped, masks):
        n = batch['pose1'].shape[0]
        e1 = encoder_train(esd, batch['pose1'].view(n, -1), bm[0], p, num_blocks, buffers)
        contra = 0.
        e2 = None
        k = 1
        if 'pose2' in batch:
            e2 = encoder_train(esd, batch['pose2'].view(n, -1), bm[k], p, num_blocks, buffers)
            k += 1
            contra = contra + F.hinge_embedding_loss(torch.norm(e1 - e2, dim=1),
                                                     torch.ones(n, dtype=torch.int32), reduction='sum')
        if 'pose_neg' in batch:
            en = encoder_train(esd, batch['pose_neg'].view(n, -1), bm[k], p, num_blocks, buffers)
            contra = contra + torch.sum(F.hinge_embedding_loss(
                torch.norm(e1 - en, dim=1), -torch.ones(n, dtype=torch.int32), reduction='none'
            ) * batch['pose_neg_is_valid'])
        loss = 0. + contra
        if 'kp_features' in batch:
            true3d = batch['kp_features'].float()
            loss = loss + weight_3d * F.mse_loss(
                decoder_fwd(dsd, e1, name, fcn_keys).reshape(true3d.shape), true3d, reduction='sum')
            if e2 is not None:
                loss = loss + weight_3d * F.mse_loss(
                    decoder_fwd(dsd, e2, name, fcn_keys).reshape(true3d.shape), true3d, reduction='sum')
        out[name] = (float(torch.as_tensor(contra).detach()), float(loss.detach()))
        batch_loss = batch_loss + loss
        batch_n += n
    (batch_loss / batch_n).backward()
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in params.items()}
    return out, grads, buffers
