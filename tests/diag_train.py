"""Diagnostic (not a test): layer-by-layer comparison of one train step against the oracle."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import assemble_ref, student_ref
from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
from vpd_b200._lib import lib

def cos(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rgb, flow = synth.crops(64, seed=1); teach = synth.teacher(64, seed=3); fl = synth.flips(B, seed=2)
idx = torch.arange(B)
img, tgt = assemble_ref.train_batch(rgb[idx].numpy(), flow[idx].numpy(), teach[idx].numpy(), fl.numpy(), *synth.FS_MEAN_STD)
torch.manual_seed(0)
m = RGBF_EmbeddingModel('resnet34', 32, True, 'cuda'); tr = ModelTrainer(m, True)
m._ensure_grads(); m.train(); tr._loss.zero_()
tr._run(img.cuda(), tgt.cuda(), B, True); torch.cuda.synchronize()
torch.manual_seed(0)
osd = student_ref.init_encoder_state('resnet34', 32, True); odsd = student_ref.init_decoder_state(32)
otr = student_ref.OracleTrainer(osd, odsd)
emb, trace = student_ref.encoder_forward_trace(otr.sd, img, train=True)
out = student_ref.decoder_forward(otr.dsd, emb)
loss = torch.nn.functional.mse_loss(out, tgt, reduction='sum')
inter = [trace['stem.y'], trace['stem.z']] + [t for i in range(16) for t in trace[i] if t is not None]
grads = torch.autograd.grad(loss, otr.params + inter)
pg = grads[:len(otr.params)]
def act(block, which):
    ptr = ctypes.c_void_p(); numel = ctypes.c_int64()
    lib().call('vpd_net_activation', m._net.handle, block, which, B, ctypes.byref(ptr), ctypes.byref(numel))
    t = torch.empty(numel.value, dtype=torch.bfloat16, device='cuda')
    lib().call('vpd_copy_d2d', t, ptr.value, numel.value * 2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return t
def cmp(name, got, ref):
    N, C, H, W = ref.shape
    g = got.view(N, H, W, C).permute(0, 3, 1, 2).float().cpu()
    print('{:28s} cos {:.6f} rel {:.4e}'.format(name, cos(g, ref.detach()), ((g - ref.detach()).norm() / ref.detach().norm()).item()))
cmp('stem.y', act(-1, 0), trace['stem.y']); cmp('stem.z', act(-1, 4), trace['stem.z'])
for i in range(16):
    y1, z1, y2, yds, zo = trace[i]
    cmp('b%d.y1' % i, act(i, 0), y1); cmp('b%d.z1' % i, act(i, 1), z1); cmp('b%d.y2' % i, act(i, 2), y2)
    if yds is not None: cmp('b%d.yds' % i, act(i, 3), yds)
    cmp('b%d.zout' % i, act(i, 4), zo)
print('loss', tr._loss.item(), loss.item())
params, m._params = m._params, m._grads
try: gsd = m._read_state(lambda k: True)
finally: m._params = params
names = student_ref.encoder_param_names('resnet34') + ['decoder.' + n for n in student_ref.DECODER_PARAM_NAMES]
for name, og in reversed(list(zip(names, pg))):
    g = gsd[name].cpu()
    print('{:40s} cos {:.5f} norm {:.4e} ref {:.4e}'.format(name, cos(g, og), g.norm().item(), og.norm().item()))
