"""One fused attention forward + backward at a given shape (for ncu): python tools/attn_one.py B Nq Nk dk dv"""
import sys
import torch
sys.path.insert(0, '.')
from tartangan_b200 import ops
bt, nq, nk, dk, dv = (int(a) for a in sys.argv[1:6])
q = torch.randn(bt, nq, dk, device='cuda').bfloat16().requires_grad_()
k = torch.randn(bt, nk, dk, device='cuda').bfloat16().requires_grad_()
v = torch.randn(bt, nk, dv, device='cuda').bfloat16().requires_grad_()
go = torch.randn(bt, nq, dv, device='cuda').bfloat16()
for _ in range(3):
    o = ops.FusedAttentionFn.apply(q, k, v)
    torch.autograd.grad(o, (q, k, v), go)
torch.cuda.synchronize()
