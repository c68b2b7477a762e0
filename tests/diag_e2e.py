"""Diagnostic: H2D bandwidth and the e2e epoch() path."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
from vpd_b200.assemble import assemble_batch, assemble_stem
dev = torch.device('cuda:0')
B = 256
x = torch.empty((B, 5, 128, 128), dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device=dev)
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print('H2D pinned 84MB: {:.2f} ms -> {:.1f} GB/s'.format(dt * 1e3, x.numel() * 4 / dt / 1e9))
torch.manual_seed(0)
enc = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda'); tr = ModelTrainer(enc, True)
opt, _ = tr.get_optimizer(5e-4)
rgb, flow = synth.crops(512, seed=1); teach = synth.teacher(512, seed=3)
rgb, flow, teach = rgb.to(dev), flow.to(dev), teach.to(dev)
host = []
for j in range(3):
    idx = torch.randint(0, 512, (B,)).int().to(dev); fl = torch.randint(0, 2, (B,)).to(torch.uint8).to(dev)
    b = assemble_batch(rgb, flow, synth.FS_MEAN_STD, flip=fl, teacher=teach, index=idx)
    host.append({'img': b['img'].cpu().pin_memory(), 'emb': b['emb'].cpu().pin_memory()})
for n in (3, 10, 10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = tr.epoch([host[j % 3] for j in range(n)], optimizer=opt)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('epoch {} steps: {:.2f} ms/step  {:.0f} frames/s loss {:.3f}'.format(n, dt / n * 1e3, B * n / dt, loss))
# fused test numbers
res = []
for fused in (False, True):
    torch.manual_seed(1)
    m = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda'); t = ModelTrainer(m, True); o, _ = t.get_optimizer(5e-4)
    r, f = synth.crops(16, seed=41); te = synth.teacher(16, seed=42); fl = synth.flips(16, seed=43)
    if fused:
        tgt = torch.empty((16, 64), device=dev); ptr = t.stem_buffer(16, 128, 128)
        assemble_stem(ptr, r.to(dev), f.to(dev), synth.FS_MEAN_STD, flip=fl.to(dev), teacher=te.to(dev), tgt=tgt)
        t._loss.zero_(); t.train_step_stem(ptr, tgt, 16, 128, 128, o); loss = t._loss.item() / 16
    else:
        b = assemble_batch(r.to(dev), f.to(dev), synth.FS_MEAN_STD, flip=fl.to(dev), teacher=te.to(dev))
        loss = t.epoch([b], optimizer=o)
    res.append((loss, m.state_dict()['resnet.fc.weight'].cpu(), m._grads.clone()))
d = (res[0][1] - res[1][1]).abs()
print('loss', res[0][0], res[1][0], 'frac>1e-4', (d > 1e-4).float().mean().item(), 'max', d.max().item())
g = (res[0][2] - res[1][2]).abs(); print('grad arena max diff', g.max().item(), 'rel', (g.norm() / res[0][2].norm()).item())
