"""GPU parity of the fused tcgen05 attention kernels (ttg_attn_fwd / ttg_attn_bwd, reference
models/blocks/attention.py:25-34) against plain PyTorch fp32 on the same bf16-rounded inputs, at the shapes of
the attention configs: G side Nq = 64x64, Nk = 32x32; D side Nq = 32x32, Nk = 16x16; C = 64 (dk 8, dv 32) and
C = 128 (dk 16, dv 64).  Tolerances (relative L2): forward 1e-2 (beta is rounded to bf16 before the second MMA),
gradients 2e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _inputs(bt, nq, nk, dk, dv, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(bt, nq, dk, generator=g) * scale).bfloat16()
    k = torch.randn(bt, nk, dk, generator=g).bfloat16()
    v = torch.randn(bt, nk, dv, generator=g).bfloat16()
    go = torch.randn(bt, nq, dv, generator=g).bfloat16()
    return q, k, v, go


def _ref(q, k, v, go):
    q, k, v = (t.double().requires_grad_() for t in (q, k, v))
    s = torch.bmm(q, k.transpose(1, 2))
    o = torch.bmm(F.softmax(s, -1), v)
    gq, gk, gv = torch.autograd.grad(o, (q, k, v), go.double())
    return o, torch.logsumexp(s, -1), gq, gk, gv


SHAPES = [
    (3, 1024, 256, 8, 32),       # D side of '256' / '512thin' (C = 64)
    (2, 4096, 1024, 8, 32),      # G side of '256' / '512thin'
    (2, 1024, 256, 16, 64),      # D side of '512' (C = 128)
    (2, 4096, 1024, 16, 64),     # G side of '512': K / V do not fit with 128-key tiles -> 64-key tiles
    (40, 512, 128, 8, 32),       # more query tiles than SMs: several tiles per CTA, one key tile
    (2, 256, 384, 16, 32),       # odd number of key tiles
]


@pytest.mark.parametrize('shape', SHAPES)
def test_fused_attention_matches_torch(shape):
    from tartangan_b200 import ops
    bt, nq, nk, dk, dv = shape
    q, k, v, go = _inputs(bt, nq, nk, dk, dv, seed=sum(shape))
    o_ref, lse_ref, gq_ref, gk_ref, gv_ref = _ref(q, k, v, go)
    qd, kd, vd = (t.cuda().requires_grad_() for t in (q, k, v))
    assert ops.attention_fused_ok(qd, kd, vd)
    o = ops.FusedAttentionFn.apply(qd, kd, vd)
    assert rel(o, o_ref) < 1e-2
    lse = o.grad_fn.saved_tensors[4]
    assert float((lse.cpu().double() - lse_ref).abs().max()) < 2e-3
    gq, gk, gv = torch.autograd.grad(o, (qd, kd, vd), go.cuda())
    errs = (rel(gq, gq_ref), rel(gk, gk_ref), rel(gv, gv_ref))
    assert max(errs) < 2e-2, errs


def test_fused_attention_large_logits_and_row_sums():
    """Online softmax across key tiles: logits of magnitude ~60 (max subtraction matters), and with v = 1 every
    output row is the sum of beta over keys = 1 (size-independent property)."""
    from tartangan_b200 import ops
    q, k, v, go = _inputs(2, 1024, 1024, 8, 32, seed=3, scale=8.0)
    o_ref, *_ = _ref(q, k, v, go)
    o = ops.FusedAttentionFn.apply(q.cuda(), k.cuda(), v.cuda())
    assert torch.isfinite(o.float()).all()
    assert rel(o, o_ref) < 1e-2
    ones = torch.ones_like(v).cuda()
    o1 = ops.FusedAttentionFn.apply(q.cuda(), k.cuda(), ones).float()
    assert float((o1 - 1).abs().max()) < 1e-2


def test_module_fused_vs_unfused_and_double_backward():
    """SelfAttention2d(64) at 32x32: the fused path equals the bmm / softmax / bmm path, first order and through
    create_graph (the R1 penalty's second backward), in bf16 mode."""
    import copy
    import tartangan_b200 as tb
    from tartangan_b200 import ops
    from tartangan_b200.models.blocks import SelfAttention2d
    tb.set_precision('bf16')
    torch.manual_seed(11)
    m = SelfAttention2d(64).cuda()
    with torch.no_grad():
        m.gamma.fill_(0.7)
        for p in (m.theta.weight, m.phi.weight):
            p.mul_(3.0)
    x0 = torch.randn(2, 64, 32, 32, device='cuda')
    gy = ops.to_internal(torch.randn(2, 64, 32, 32, device='cuda'))
    v = ops.to_internal(torch.randn(2, 64, 32, 32, device='cuda'))
    params = [m.theta.weight, m.phi.weight, m.g.weight, m.o.weight]
    res = {}
    for fused in (True, False):
        ops.state.fused_attention = fused
        try:
            x = x0.clone().requires_grad_()
            y = m(x)
            first = torch.autograd.grad(y, [x] + params, gy, create_graph=True)
            second = torch.autograd.grad(ops.DotFn.apply(ops.to_internal(first[0]), v), params)
            x2 = x0.clone().requires_grad_()
            plain = torch.autograd.grad(m(x2), [x2] + params, gy)          # no create_graph: the fused backward kernel
            res[fused] = (y, first, second, plain)
        finally:
            ops.state.fused_attention = True
    assert rel(res[True][0], res[False][0]) < 1e-2
    for a, b in zip(res[True][1], res[False][1]):
        assert rel(a, b) < 3e-2
    for a, b in zip(res[True][2], res[False][2]):
        assert rel(a, b) < 6e-2
    for a, b in zip(res[True][3], res[False][3]):
        assert rel(a, b) < 3e-2
    for a, b in zip(res[True][3], res[True][1]):
        assert rel(a, b) < 3e-2


def test_unsupported_shape_is_an_error_not_a_fallback():
    from tartangan_b200 import _lib
    q = torch.zeros(1, 64, 8, dtype=torch.bfloat16, device='cuda')
    assert not _lib.lib.ttg_attn_supported(64, 16, 8, 32)
    with pytest.raises(RuntimeError):
        _lib.call('ttg_attn_fwd', q.data_ptr(), q.data_ptr(), q.data_ptr(), q.data_ptr(), q.data_ptr(), 1, 64, 16, 8, 32)
