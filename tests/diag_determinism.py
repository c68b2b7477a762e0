"""Diagnostic: is one train step run-to-run deterministic? per-tensor gradient diffs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
from vpd_b200.assemble import assemble_batch
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(1)
m = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda'); t = ModelTrainer(m, True)
r, f = synth.crops(B, seed=41); te = synth.teacher(B, seed=42); fl = synth.flips(B, seed=43)
b = assemble_batch(r.to(dev), f.to(dev), synth.FS_MEAN_STD, flip=fl.to(dev), teacher=te.to(dev))
m._ensure_grads(); m.train()
runs = []
for i in range(3):
    t._loss.zero_(); t._run(b['img'], b['emb'], B, True); torch.cuda.synchronize()
    runs.append((t._loss.item(), m._grads.clone()))
print('losses', [x[0] for x in runs])
table = m._table
for name, arena, off, layout, shape in reversed(table):
    if arena != 0: continue
    import numpy as np
    numel = 7 * 64 * 64 if layout == 2 else int(np.prod(shape))
    a, c = runs[0][1][off:off + numel], runs[1][1][off:off + numel]
    d = (a - c).abs().max().item()
    rel = ((a - c).norm() / (a.norm() + 1e-20)).item()
    print('{:42s} maxdiff {:.3e} rel {:.3e}'.format(name, d, rel))
