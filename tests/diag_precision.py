"""Diagnostic (CPU, not a test): how much storage precision does the student's gradient need?

    python tests/diag_precision.py [batch=32] [train_steps=0]

Runs the fp32 oracle's training step with reduced-precision STORAGE laid over it
(oracle.student_ref.encoder_forward_rounded: weights / stored activations / stored gradients
rounded, fp32 arithmetic) and prints the cosine of selected gradient tensors against plain fp32.
Measured in the build container (batch 32, seeds of SURVEY §8d, reference init):

    storage                      conv1    l1.1.bn2  l2.0.c1  l3.0.c1  l4.1.c1  fc      min
    bf16 everything              0.748    0.721     0.764    0.787    0.882    0.994   0.694
    fp16 everything              0.963    0.968     0.966    0.969    0.983    0.998   0.952
    bf16 weights only            0.893    0.903     0.899    0.908    0.953    0.997   0.836
    bf16 activations only        0.762    0.801     0.777    0.794    0.891    0.994   0.610
    bf16 gradients only          0.9999   1.0000    1.0000   1.0000   1.0000   1.0000  0.9993
    after 60 AdamW steps: bf16   0.902    0.910     0.913    0.935    0.963    0.984   0.870
                          fp16   0.989    0.987     0.990    0.992    0.996    0.998   0.976

i.e. the freshly initialised 34-layer net with batch-statistic BatchNorm turns a 2^-9 relative
storage error into an O(1) change of the early layers' gradient direction; the error shrinks
like sqrt(eps), so 11-bit mantissas (fp16 - the reference's own CUDA autocast precision - or
TF32) still sit at 0.95-0.96 and cosine >= 0.99 on every tensor needs ~14 bits. The hardware
result at batch 256 (tests/test_parity_configs_gpu.py) sits on the bf16 row.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import assemble_ref, student_ref     # noqa: E402
from vpd_b200 import synth                        # noqa: E402

SEL = ['resnet.conv1.weight', 'resnet.layer1.1.bn2.weight', 'resnet.layer2.0.conv1.weight',
       'resnet.layer3.0.conv1.weight', 'resnet.layer4.1.conv1.weight', 'resnet.fc.weight']
CASES = [('bf16 everything', dict()),
         ('fp16 everything', dict(weight_dtype=torch.float16, act_dtype=torch.float16,
                                  grad_dtype=torch.float16)),
         ('bf16 weights only', dict(act_dtype=None, grad_dtype=None)),
         ('bf16 activations only', dict(weight_dtype=None, grad_dtype=None)),
         ('bf16 gradients only', dict(weight_dtype=None, act_dtype=None))]


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm())).item()


def table(otr, img, tgt, cases):
    names = otr.enc_names + ['decoder.' + n for n in student_ref.DECODER_PARAM_NAMES]
    _, ref, _ = otr.loss_and_grads(img, tgt)
    for label, kw in cases:
        _, g = otr.loss_and_grads_rounded(img, tgt, **kw)
        cs = {n: cos(a, b) for n, a, b in zip(names, g, ref)}
        print('{:24s} {}   min {:.4f}'.format(label, ' '.join('{:.4f}'.format(cs[s]) for s in SEL),
                                              min(cs.values())))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rgb, flow = synth.crops(B, seed=1)
    img, tgt = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), synth.teacher(B, seed=3).numpy(),
                                        synth.flips(B, seed=2).numpy(), *synth.FS_MEAN_STD)
    torch.manual_seed(0)
    otr = student_ref.OracleTrainer(student_ref.init_encoder_state('resnet34', 32, True),
                                    student_ref.init_decoder_state(32))
    print('batch {}, reference init'.format(B))
    table(otr, img, tgt, CASES)
    if steps:
        prgb, pflow = synth.crops(512, seed=11)
        pt = synth.teacher(512, seed=13)
        g = torch.Generator().manual_seed(5)
        for _ in range(steps):
            idx = torch.randint(0, 512, (B,), generator=g)
            f = torch.randint(0, 2, (B,), generator=g).to(torch.uint8)
            otr.step(*assemble_ref.train_batch(prgb[idx].numpy(), pflow[idx].numpy(),
                                               pt[idx].numpy(), f.numpy(), *synth.FS_MEAN_STD))
        print('after {} AdamW steps'.format(steps))
        table(otr, img, tgt, CASES[:2])


if __name__ == '__main__':
    main()
