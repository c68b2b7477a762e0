import sys, torch
sys.path.insert(0,'/root/repo')
from vpd_b200 import synth
from vpd_b200.assemble import assemble_stem, assemble_batch
dev=torch.device('cuda:0')
rgb,flow=synth.crops(4096,seed=1); teach=synth.teacher(4096,seed=3)
rgb,flow,teach=rgb.to(dev),flow.to(dev),teach.to(dev)
g=torch.Generator().manual_seed(0)
idx=torch.randint(0,4096,(64,256),generator=g).int().to(dev); fl=torch.randint(0,2,(64,256),generator=g).to(torch.uint8).to(dev)
stem=torch.empty((256,67,34,64),device=dev,dtype=torch.bfloat16); tgt=torch.empty((256,64),device=dev)
def t(fn,reps=40):
    for r in range(5): fn(r)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps): fn(r)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1e3
a=t(lambda r: assemble_stem(stem,rgb,flow,synth.FS_MEAN_STD,flip=fl[r%64],teacher=teach,index=idx[r%64],tgt=tgt))
b=t(lambda r: assemble_batch(rgb,flow,synth.FS_MEAN_STD,flip=fl[r%64],teacher=teach,index=idx[r%64]))
print('assemble_stem %.1f us (%.0f GB/s)  assemble_nchw %.1f us (%.0f GB/s)'%(a, 95.6e6/a/1e3, b, 104.86e6/b/1e3))
