"""Diagnostic: does tcgen05.mma.kind::f16 accept fp16 for A and bf16 for B in ONE instruction
(instruction-descriptor a_format = F16, b_format = BF16)? A = 128 random fp16 rows, B = the
bf16 identity: D must come back as A exactly. (Planning input for fp16 activation storage with
bf16 gradients: the weight-gradient GEMM multiplies the two.)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpd_b200._lib import lib, stream_ptr
dev = torch.device('cuda:0')
rows = 128
g = torch.Generator().manual_seed(0)
for name, dtype, mode in (('bf16 x bf16', torch.bfloat16, 0), ('fp16 x bf16', torch.float16, 2)):
    src = (torch.randn((rows, 64), generator=g) * 3).to(dtype).to(dev)
    out = torch.zeros((128, 64), device=dev)
    lib().call('vpd_umma_probe', src, rows, 0, 1024, mode, out, stream_ptr())
    torch.cuda.synchronize()
    print(name, 'exact rows:', (out == src.float()).all(dim=1).float().mean().item(),
          'max abs diff', (out - src.float()).abs().max().item())
