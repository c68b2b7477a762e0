"""Diagnostic: which forward activation first differs between two identical train steps?"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
from vpd_b200._lib import lib
from vpd_b200.assemble import assemble_batch
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(1)
m = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda'); t = ModelTrainer(m, True)
r, f = synth.crops(B, seed=41); te = synth.teacher(B, seed=42); fl = synth.flips(B, seed=43)
b = assemble_batch(r.to(dev), f.to(dev), synth.FS_MEAN_STD, flip=fl.to(dev), teacher=te.to(dev))
m._ensure_grads(); m.train()
def act(block, which):
    ptr = ctypes.c_void_p(); numel = ctypes.c_int64()
    lib().call('vpd_net_activation', m._net.handle, block, which, B, ctypes.byref(ptr), ctypes.byref(numel))
    if not ptr.value: return None
    x = torch.empty(numel.value, dtype=torch.bfloat16, device='cuda')
    lib().call('vpd_copy_d2d', x, ptr.value, numel.value * 2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return x
snaps = []
for i in range(2):
    t._loss.zero_(); t._run(b['img'], b['emb'], B, True); torch.cuda.synchronize()
    s = {('stem', 0): act(-1, 0), ('stem', 4): act(-1, 4)}
    for blk in range(16):
        for w in range(5):
            s[(blk, w)] = act(blk, w)
    snaps.append(s)
names = {0: 'y1', 1: 'z1', 2: 'y2', 3: 'yds', 4: 'zout'}
for k in snaps[0]:
    a, c = snaps[0][k], snaps[1][k]
    if a is None: continue
    d = (a.float() - c.float())
    nz = (d != 0).float().mean().item()
    print(k, 'frac differing {:.3e} rel {:.3e}'.format(nz, (d.norm() / a.float().norm()).item()))
