"""GPU parity of the individual kernels (through the C ABI) against plain PyTorch fp32 on CPU.
fp32 mode: tight tolerance (exact algorithm, fp32 accumulate).  bf16 mode: bf16 tolerance."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {'fp32': 2e-4, 'bf16': 3e-2}


_MODE = ['fp32']


def rel(a, b):
    """fp32: max-abs error relative to the largest reference magnitude.  bf16: relative L2 error
    (a bf16-rounded pre-activation near zero flips its LeakyReLU mask, which moves single
    elements by a factor 5 without being a kernel error)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if _MODE[0] == 'bf16':
        return float((a - b).norm() / b.norm().clamp_min(1e-6))
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.fixture(params=['fp32', 'bf16'])
def mode(request):
    import tartangan_b200 as tb
    tb.set_precision(request.param)
    _MODE[0] = request.param
    yield request.param
    tb.set_precision('bf16')
    _MODE[0] = 'fp32'


def _dev(x, mode):
    from tartangan_b200 import ops
    return ops.to_internal(x.cuda())


@pytest.mark.parametrize('cin,cout,k,hw,n', [(3, 16, 3, 16, 2), (16, 16, 3, 12, 3), (32, 16, 1, 8, 2),
                                             (16, 3, 1, 8, 2), (24, 40, 3, 10, 1), (64, 64, 3, 8, 2)])
def test_conv_fwd_bwd_double(mode, cin, cout, k, hw, n):
    from tartangan_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(n, cin, hw, hw, requires_grad=True)
    w = (torch.randn(cout, cin, k, k) / math.sqrt(cin * k * k)).requires_grad_()
    b = torch.randn(cout, requires_grad=True)
    y = F.conv2d(x, w, b, padding=k // 2)
    gy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy, create_graph=True)
    # second order: d/d(w, gy-path) of <gx, v>
    v = torch.randn_like(x)
    gw2, = torch.autograd.grad((gx * v).sum(), (w,))

    xd = x.detach().cuda().requires_grad_()
    wd = w.detach().cuda().requires_grad_()
    bd = b.detach().cuda().requires_grad_()
    yd = ops.conv2d(ops.to_internal(xd), wd, bd)
    assert rel(yd, y) < TOL[mode]
    gyd = ops.to_internal(gy.cuda())
    gxd, gwd, gbd = torch.autograd.grad(yd, (xd, wd, bd), gyd, create_graph=True)
    assert rel(gxd, gx) < TOL[mode] and rel(gwd, gw) < TOL[mode] and rel(gbd, gb) < TOL[mode]
    gw2d, = torch.autograd.grad(ops.DotFn.apply(ops.to_internal(gxd), ops.to_internal(v.cuda())), (wd,))
    assert rel(gw2d, gw2) < TOL[mode]


@pytest.mark.parametrize('c,hw,n', [(3, 16, 4), (16, 8, 3), (40, 6, 2)])
def test_bn_act_fwd_bwd_double(mode, c, hw, n):
    from tartangan_b200 import ops
    from tartangan_b200.models.layers import BatchNorm2d
    torch.manual_seed(1)
    x = torch.randn(n, c, hw, hw) * 1.5 + 0.3
    bn_ref = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.uniform_(-0.5, 0.5)
    bn = BatchNorm2d(c)
    bn.load_state_dict(bn_ref.state_dict())
    bn = bn.cuda()
    xr = x.clone().requires_grad_()
    y = F.leaky_relu(bn_ref(xr), 0.2)
    gy = torch.randn_like(y)
    gx, ggam, gbet = torch.autograd.grad(y, (xr, bn_ref.weight, bn_ref.bias), gy, create_graph=True)
    v = torch.randn_like(x)
    x2, gam2 = torch.autograd.grad((gx * v).sum(), (xr, bn_ref.weight))

    xd = x.cuda().requires_grad_()
    xi = ops.to_internal(xd)
    yd = ops.bn_act(xi, bn, 0.2)
    assert rel(yd, y) < TOL[mode]
    assert rel(bn.running_mean, bn_ref.running_mean) < 1e-2 and rel(bn.running_var, bn_ref.running_var) < 1e-2
    assert int(bn.num_batches_tracked) == 1
    gxd, ggamd, gbetd = torch.autograd.grad(yd, (xd, bn.weight, bn.bias), ops.to_internal(gy.cuda()), create_graph=True)
    assert rel(gxd, gx) < TOL[mode] * 2 and rel(ggamd, ggam) < TOL[mode] * 2 and rel(gbetd, gbet) < TOL[mode] * 2
    x2d, gam2d = torch.autograd.grad(ops.DotFn.apply(ops.to_internal(gxd), ops.to_internal(v.cuda())), (xd, bn.weight))
    assert rel(x2d, x2) < TOL[mode] * 4 and rel(gam2d, gam2) < TOL[mode] * 4


def test_fp32_wgrad_is_bitwise_repeatable():
    """The fp32 (parity) wgrad path adds its split-K partial sums in a fixed order (ttg_conv2d_wgrad_direct_det): two runs
    on the same inputs give identical bits, like the reference's CPU convolution_backward."""
    import tartangan_b200 as tb
    from tartangan_b200 import ops
    tb.set_precision('fp32')
    try:
        torch.manual_seed(9)
        for n, cin, cout, hw, k in ((8, 16, 16, 32, 3), (4, 128, 64, 16, 3), (16, 3, 32, 32, 1), (2, 24, 40, 10, 3)):
            x = ops.to_internal(torch.randn(n, cin, hw, hw, device='cuda'))
            gy = ops.to_internal(torch.randn(n, cout, hw, hw, device='cuda'))
            runs = [ops.ConvWgradFn.apply(x, gy, k, 0).clone() for _ in range(4)]
            assert all(torch.equal(runs[0], r) for r in runs[1:]), (n, cin, cout, hw, k)
            ref = torch.nn.grad.conv2d_weight(x.float().cpu(), (cout, cin, k, k), gy.float().cpu(), padding=k // 2)
            assert rel(runs[0], ref) < 2e-4
    finally:
        tb.set_precision('bf16')


@pytest.mark.parametrize('act', ['selu', 'elu'])
def test_elu_selu_fwd_bwd_double(mode, act):
    """nn.SELU / nn.ELU (--activation selu|elu, reference trainers/cnn.py:41-45): value, gradient and the second-order
    term the R1 penalty needs (cotangent of x through the backward)."""
    from tartangan_b200 import ops
    torch.manual_seed(3)
    ref_fn, fn = (F.selu, ops.selu) if act == 'selu' else (F.elu, ops.elu)
    for shape in ((2, 16, 8, 8), (3, 5, 6, 6), (7, 33)):
        x = (torch.randn(*shape) * 2).bfloat16().float().requires_grad_()
        y = ref_fn(x)
        gy, v = torch.randn_like(y), torch.randn_like(x)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        gyr = gy.clone().requires_grad_()
        gx_b, = torch.autograd.grad(ref_fn(x), x, gyr, create_graph=True)
        x2, gy2 = torch.autograd.grad((gx_b * v).sum(), (x, gyr))
        xd = x.detach().cuda().requires_grad_()
        four = len(shape) == 4
        xi = ops.to_internal(xd) if four else xd
        cast = (lambda t: ops.to_internal(t.cuda())) if four else (lambda t: t.cuda())
        yd = fn(xi)
        assert rel(yd, y) < TOL[mode]
        gyd = cast(gy).detach().requires_grad_()
        gxd, = torch.autograd.grad(yd, xd, gyd, create_graph=True)
        assert rel(gxd, gx) < TOL[mode]
        probe = ops.DotFn.apply(cast(gxd) if four else gxd, cast(v))
        x2d, gy2d = torch.autograd.grad(probe, (xd, gyd))
        assert rel(x2d, x2) < TOL[mode] * 2 and rel(gy2d, gy2) < TOL[mode] * 2


def test_resample_ops(mode):
    from tartangan_b200 import ops
    torch.manual_seed(2)
    for c, hw in ((3, 16), (16, 8), (8, 4)):
        x = torch.randn(2, c, hw, hw).bfloat16().float().requires_grad_()
        xd = x.detach().cuda().requires_grad_()
        xi = ops.to_internal(xd)
        for ref_fn, fn in ((lambda t: F.avg_pool2d(t, 2), ops.avg_pool2),
                           (lambda t: F.interpolate(t, scale_factor=2, mode='nearest'), ops.upsample2),
                           (lambda t: F.interpolate(t, scale_factor=0.5, mode='bilinear', align_corners=True),
                            ops.bilinear_down),
                           (lambda t: F.max_pool2d(t, 2), ops.max_pool2)):
            y = ref_fn(x)
            g = torch.randn_like(y)
            gx, = torch.autograd.grad(y, x, g)
            yd = fn(xi)
            assert rel(yd, y) < TOL[mode]
            gxd, = torch.autograd.grad(yd, xd, ops.to_internal(g.cuda()))
            assert rel(gxd, gx) < TOL[mode]
        s = ops.spatial_sum(xi)
        assert rel(s, x.sum((2, 3))) < TOL[mode]


def test_iqn_head_and_losses():
    from tartangan_b200 import ops
    torch.manual_seed(3)
    B, C, E, nq = 6, 24, 20, 8
    feats = torch.randn(B, C, requires_grad=True)
    taus = torch.rand(B * nq, 1)
    we = (torch.randn(C, E) * 0.3).requires_grad_()
    be = (torch.randn(C) * 0.1).requires_grad_()
    wo = (torch.randn(1, C) * 0.3).requires_grad_()
    bo = torch.randn(1, requires_grad=True)
    rng = torch.arange(1, E + 1).float()
    emb = torch.tanh(F.linear(torch.cos(taus.repeat(1, E) * math.pi * rng), we, be))
    p_tau = F.linear(feats.repeat(nq, 1) * emb, wo, bo)
    targets = torch.ones(B, 1)
    from oracle.tartan_oracle import quantile_huber
    loss = quantile_huber(p_tau, targets, taus)
    p = p_tau.reshape(nq, -1, 1).mean(0)
    ref_g = torch.autograd.grad(loss + p.sum() * 0.3, (feats, we, be, wo, bo), retain_graph=True)
    # second order through d p / d feats
    gfeat, = torch.autograd.grad(p.sum(), feats, create_graph=True)
    ref2 = torch.autograd.grad((gfeat ** 2).sum(), (we, be, wo))

    d = lambda t: t.detach().cuda().requires_grad_()
    fd, wed, bed, wod, bod = d(feats), d(we), d(be), d(wo), d(bo)
    td = taus.cuda()
    p_tau_d = ops.IqnHeadFn.apply(fd, td, wed, bed, wod, bod, nq)
    assert rel(p_tau_d, p_tau.reshape(-1)) < 1e-4
    loss_d = ops.QuantileHuberFn.apply(p_tau_d, targets.cuda().reshape(-1), td.reshape(-1), nq, 1.0)
    assert rel(loss_d, loss) < 1e-4
    p_d = ops.ColsumFn.apply(p_tau_d.view(nq, B), 1.0 / nq)
    assert rel(p_d, p.reshape(-1)) < 1e-4
    tot = ops.AxpbyFn.apply(loss_d, ops.ColsumFn.apply(p_d.view(B, 1), 1.0).reshape(()), 1.0, 0.3)
    got = torch.autograd.grad(tot, (fd, wed, bed, wod, bod), retain_graph=True)
    for a, b_ in zip(got, ref_g):
        assert rel(a, b_) < 2e-4
    gfeat_d, = torch.autograd.grad(p_d, fd, torch.ones_like(p_d), create_graph=True)
    got2 = torch.autograd.grad(ops.SqsumFn.apply(gfeat_d, 1.0), (wed, bed, wod))
    for a, b_ in zip(got2, ref2):
        assert rel(a, b_) < 2e-4


def test_iqn_head_loss_single_kernel():
    """ops.IqnHeadLossFn: the whole head (embedding, mix, Linear, quantile mean, quantile-Huber loss) as ONE forward
    and ONE backward kernel, against torch autograd on the un-fused formula, incl. the R1 second-order path and
    the no-target call (reference blocks/discriminator.py:164-178, models/iqn.py:91-130)."""
    from tartangan_b200 import ops, _lib
    from oracle.tartan_oracle import quantile_huber
    for B, C, nq in ((6, 24, 8), (33, 128, 8), (5, 256, 64), (9, 16, 3)):
        torch.manual_seed(3 + B)
        E = 20
        feats = torch.randn(B, C, requires_grad=True)
        taus = torch.rand(B * nq, 1)
        we = (torch.randn(C, E) * 0.3).requires_grad_()
        be = (torch.randn(C) * 0.1).requires_grad_()
        wo = (torch.randn(1, C) * 0.3).requires_grad_()
        bo = torch.randn(1, requires_grad=True)
        rng = torch.arange(1, E + 1).float()
        emb = torch.tanh(F.linear(torch.cos(taus.repeat(1, E) * math.pi * rng), we, be))
        p_tau = F.linear(feats.repeat(nq, 1) * emb, wo, bo)
        targets = (torch.rand(B, 1) > 0.5).float()
        loss = quantile_huber(p_tau, targets, taus)
        p = p_tau.reshape(nq, -1, 1).mean(0)
        coef = torch.randn(B, 1)
        ref_g = torch.autograd.grad(1.7 * loss + (p * coef).sum(), (feats, we, be, wo, bo), retain_graph=True)
        gfeat, = torch.autograd.grad(p.sum(), feats, create_graph=True)
        ref2 = torch.autograd.grad((gfeat ** 2).sum(), (we, be, wo))

        d = lambda t: t.detach().cuda().requires_grad_()
        fd, wed, bed, wod, bod = d(feats), d(we), d(be), d(wo), d(bo)
        td = taus.cuda()
        k0 = _lib.Counters.kernels
        p_d, loss_d = ops.IqnHeadLossFn.apply(fd, td, wed, bed, wod, bod, targets.cuda(), nq, 1.0)
        assert _lib.Counters.kernels - k0 == 1
        assert p_d.shape == (B, 1) and rel(p_d, p) < 1e-4 and rel(loss_d, loss) < 1e-4, (B, C, nq)
        k0 = _lib.Counters.kernels
        got = torch.autograd.grad((p_d, loss_d), (fd, wed, bed, wod, bod), (coef.cuda(), torch.tensor(1.7, device='cuda')),
                                  retain_graph=True)
        assert _lib.Counters.kernels - k0 == 1
        for a, b_ in zip(got, ref_g):
            assert rel(a, b_) < 2e-4, (B, C, nq)
        gfeat_d, = torch.autograd.grad(p_d, fd, torch.ones_like(p_d), create_graph=True)
        got2 = torch.autograd.grad(ops.SqsumFn.apply(gfeat_d, 1.0), (wed, bed, wod))
        for a, b_ in zip(got2, ref2):
            assert rel(a, b_) < 2e-4, (B, C, nq)
        p_only = ops.IqnHeadLossFn.apply(fd, td, wed, bed, wod, bod, None, nq, 1.0)
        assert rel(p_only, p) < 1e-4


def test_bce_linear_adam():
    from tartangan_b200 import ops
    from tartangan_b200.optim import FusedAdam
    torch.manual_seed(4)
    x = torch.randn(10, 7, requires_grad=True)
    w = torch.randn(1, 7, requires_grad=True)
    b = torch.randn(1, requires_grad=True)
    y = (torch.rand(10, 1) > 0.5).float()
    loss = F.binary_cross_entropy_with_logits(F.linear(x, w, b), y)
    ref = torch.autograd.grad(loss, (x, w, b))
    d = lambda t: t.detach().cuda().requires_grad_()
    xd, wd, bd = d(x), d(w), d(b)
    loss_d = ops.BceLogitsFn.apply(ops.linear(xd, wd, bd), y.cuda())
    assert rel(loss_d, loss) < 1e-5
    for a, b_ in zip(torch.autograd.grad(loss_d, (xd, wd, bd)), ref):
        assert rel(a, b_) < 1e-4
    # Adam (betas (0, .999)) over three steps vs torch.optim.Adam
    ps = [torch.randn(5, 3), torch.randn(7)]
    ref_p = [p.clone().requires_grad_() for p in ps]
    my_p = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    o_ref = torch.optim.Adam(ref_p, lr=4e-4, betas=(0., 0.999))
    o_my = FusedAdam(my_p, lr=4e-4, betas=(0., 0.999))
    for s in range(3):
        o_my.zero_grad()
        for i, (rp, mp) in enumerate(zip(ref_p, my_p)):
            g = torch.randn_like(rp)
            rp.grad = g.clone()
            mp.grad.copy_(g.cuda())
        o_ref.step()
        o_my.step()
    for rp, mp in zip(ref_p, my_p):
        assert rel(mp, rp) < 1e-5


def test_attention_and_spectral_norm(mode):
    from tartangan_b200 import ops
    from tartangan_b200.models.blocks import SelfAttention2d
    import sys, types
    torch.manual_seed(5)
    m = SelfAttention2d(16)
    with torch.no_grad():
        m.gamma.fill_(0.7)
    x = torch.randn(2, 16, 8, 8).bfloat16().float().requires_grad_()

    def ref_attn(m, x):
        n, c, h, w = x.shape
        th = F.conv2d(x, m.theta.weight).view(n, c // 8, h * w)
        ph = F.max_pool2d(F.conv2d(x, m.phi.weight), 2).view(n, c // 8, h * w // 4)
        g = F.max_pool2d(F.conv2d(x, m.g.weight), 2).view(n, c // 2, h * w // 4)
        beta = F.softmax(torch.bmm(th.transpose(1, 2), ph), -1)
        o = F.conv2d(torch.bmm(g, beta.transpose(1, 2)).view(n, c // 2, h, w), m.o.weight)
        return m.gamma * o + x
    y = ref_attn(m, x)
    gy = torch.randn_like(y)
    params = [m.theta.weight, m.phi.weight, m.g.weight, m.o.weight, m.gamma]
    ref = torch.autograd.grad(y, [x] + params, gy, create_graph=True)
    v = torch.randn_like(x)
    ref2 = torch.autograd.grad((ref[0] * v).sum(), params[:4])
    import copy
    md = copy.deepcopy(m).cuda()
    xd = x.detach().cuda().requires_grad_()
    yd = md(xd)
    assert rel(yd, y) < TOL[mode]
    pd = [md.theta.weight, md.phi.weight, md.g.weight, md.o.weight, md.gamma]
    got = torch.autograd.grad(yd, [xd] + pd, ops.to_internal(gy.cuda()), create_graph=True)
    errs1 = [rel(a, b_) for a, b_ in zip(got, ref)]
    assert max(errs1) < TOL[mode] * 3, errs1
    got2 = torch.autograd.grad(ops.DotFn.apply(ops.to_internal(got[0]), ops.to_internal(v.cuda())), pd[:4])
    errs2 = [rel(a, b_) for a, b_ in zip(got2, ref2)]
    assert max(errs2) < TOL[mode] * 10, (errs1, errs2)
