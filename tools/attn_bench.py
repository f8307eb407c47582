"""Micro-bench of the attention core at the shapes of the attention configs (C4: cnn '256' + attention,
C5: iqn '512thin' / '512'), per-GPU batch of the 8-GPU runs.  CUDA events, inputs larger than... no: the working set
(q, k, v, o < 30 MB) fits L2 by construction of the layer (it is exp / tensor bound, not HBM bound); an L2 flush
buffer is written between iterations anyway.  One JSON line per (shape, variant)."""
import json
import sys

import torch

sys.path.insert(0, '.')
from tartangan_b200 import ops  # noqa: E402

SHAPES = [('C4/C5thin G 64x64 C64', 16, 4096, 1024, 8, 32), ('C4/C5thin D 32x32 C64', 16, 1024, 256, 8, 32),
          ('C5 512 G 64x64 C128', 8, 4096, 1024, 16, 64), ('C5 512 D 32x32 C128', 8, 1024, 256, 16, 64)]


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, bt, nq, nk, dk, dv in SHAPES:
    q = torch.randn(bt, nq, dk, device='cuda').bfloat16().requires_grad_()
    k = torch.randn(bt, nk, dk, device='cuda').bfloat16().requires_grad_()
    v = torch.randn(bt, nk, dv, device='cuda').bfloat16().requires_grad_()
    go = torch.randn(bt, nq, dv, device='cuda').bfloat16()
    flops_f = 2.0 * bt * nq * nk * (dk + dv)
    flops_b = 2.0 * bt * nq * nk * (2 * dk + 2 * dv + dk)     # S, dP, dV, dK, dQ (as executed: dk padded to 16)
    exps = float(bt) * nq * nk
    for variant in ('fused', 'unfused'):
        ops.state.fused_attention = variant == 'fused'
        with torch.no_grad():
            t_f = timeit(lambda: ops.attention_core(q, k, v))
        o = ops.attention_core(q, k, v)
        t_b = timeit(lambda: torch.autograd.grad(o, (q, k, v), go, retain_graph=True))
        print(json.dumps({'shape': name, 'variant': variant, 'fwd_us': round(t_f, 1), 'bwd_us': round(t_b, 1),
                          'fwd_tflops': round(flops_f / t_f / 1e6, 1), 'bwd_tflops': round(flops_b / t_b / 1e6, 1),
                          'fwd_gexp_per_s': round(exps / t_f / 1e3, 1),
                          'exp_peak_gexp_per_s': round(148 * 16 * 1.965, 1)}))
    ops.state.fused_attention = True
