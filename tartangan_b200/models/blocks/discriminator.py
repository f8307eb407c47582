"""Discriminator blocks — interface of tartangan/models/blocks/discriminator.py.

Same constructors, defaults and state-dict keys.  Every op on these blocks is closed
under differentiation (ops.py) because the R1 penalty back-propagates through the
backward pass of D(real) (models/losses.py:17-30).
"""
import functools

import torch
from torch import nn

from ... import ops
from ..iqn import IQN, iqn_loss
from ..layers import (AvgPool2d, BatchNorm2d, Conv2d, LeakyReLU, Linear, interpolate, native, run_layers)


class DiscriminatorInput(nn.Module):
    """discriminator.py:11-22: 1x1 conv from RGB, no activation."""

    def __init__(self, in_dims, out_dims, conv_factory=Conv2d,
                 activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        self.convs = nn.Sequential(native(conv_factory)(in_dims, out_dims, 1, padding=0, bias=True))

    def forward(self, img):
        return run_layers(self.convs, ops.ensure_internal(img))


class DiscriminatorBlock(nn.Module):
    """Non-residual block (discriminator.py:25-46); BlockModel default, unused by cnn/iqn trainers."""

    def __init__(self, in_dims, out_dims, first_block=False, norm_factory=BatchNorm2d, conv_factory=Conv2d,
                 avg_pool_factory=AvgPool2d, activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        norm_factory, conv_factory = native(norm_factory), native(conv_factory)
        activation_factory, avg_pool_factory = native(activation_factory), native(avg_pool_factory)
        layers = [norm_factory(out_dims), activation_factory(), conv_factory(in_dims, out_dims, 3, padding=1, bias=True),
                  norm_factory(out_dims), activation_factory(), conv_factory(out_dims, out_dims, 3, padding=1, bias=True),
                  avg_pool_factory(2)]
        if first_block:
            layers = layers[2:]
        self.convs = nn.Sequential(*layers)

    def forward(self, x):
        return run_layers(self.convs, ops.ensure_internal(x))


_default_interpolate = functools.partial(interpolate, scale_factor=0.5, mode='bilinear', align_corners=True)


class ResidualDiscriminatorBlock(nn.Module):
    """discriminator.py:49-95: [BN, act,] conv3 -> BN -> act -> conv3 -> avgpool2, plus the
    bilinear(0.5, align_corners=True) skip with a 1x1 projection applied after down-sampling."""

    def __init__(self, in_dims, out_dims, first_block=False, norm_factory=BatchNorm2d, conv_factory=Conv2d,
                 avg_pool_factory=AvgPool2d, activation_factory=functools.partial(LeakyReLU, 0.2),
                 interpolate=_default_interpolate):
        super().__init__()
        norm_factory, conv_factory = native(norm_factory), native(conv_factory)
        activation_factory, avg_pool_factory = native(activation_factory), native(avg_pool_factory)
        layers = [norm_factory(in_dims), activation_factory(), conv_factory(in_dims, out_dims, 3, padding=1, bias=True),
                  norm_factory(out_dims), activation_factory(), conv_factory(out_dims, out_dims, 3, padding=1, bias=True),
                  avg_pool_factory(2)]
        if first_block:
            layers = layers[2:]
        self.convs = nn.Sequential(*layers)
        self.in_dims, self.out_dims = in_dims, out_dims
        self.project_input = None
        if in_dims != out_dims:
            self.project_input = nn.Sequential(conv_factory(in_dims, out_dims, 1))
        self.interpolate = interpolate

    def forward(self, x):
        x = ops.ensure_internal(x)
        layers = list(self.convs)
        fuse_pool = isinstance(layers[-1], AvgPool2d)
        if self.interpolate is _default_interpolate:
            xh, xs = ops.fork_bilinear_down(x)       # (the gradient fan-in of the two branches is one kernel)
        else:
            xs, xh = ops.fork(x)
            xs = self.interpolate(xs)
        if self.project_input is not None:
            with ops.skip_branch(xs) as sb:          # the 1x1 projection runs beside the conv branch
                xs = run_layers(self.project_input, xs)
            h = run_layers(layers[:-1] if fuse_pool else layers, xh)
            xs = sb.join(xs)
        else:
            h = run_layers(layers[:-1] if fuse_pool else layers, xh)
        return ops.avg_pool2_add(h, xs) if fuse_pool else ops.add(xs, h)


class DiscriminatorOutput(nn.Module):
    """discriminator.py:126-146: BN -> act -> sum over H,W -> Linear(C -> out)."""

    def __init__(self, in_dims, out_dims, norm_factory=BatchNorm2d,
                 activation_factory=functools.partial(LeakyReLU, 0.2), output_activation_factory=nn.Identity):
        super().__init__()
        self.activation = nn.Sequential(native(norm_factory)(in_dims), native(activation_factory)())
        self.to_output = nn.Sequential(Linear(in_dims, out_dims), native(output_activation_factory)())

    def forward(self, feats):
        feats = run_layers(self.activation, ops.ensure_internal(feats))
        feats = ops.spatial_sum(feats)
        return run_layers(self.to_output, feats)


class IQNDiscriminatorOutput(nn.Module):
    """discriminator.py:149-178: BN -> act -> sum over H,W -> IQN mix -> Linear(C -> 1);
    returns the mean over quantiles and, when targets are given, the quantile-Huber loss.
    The tau embedding, the mix, the Linear, the quantile mean and the loss are ONE kernel (ops.IqnHeadLossFn)."""

    def __init__(self, in_dims, out_dims, norm_factory=BatchNorm2d,
                 activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        if out_dims != 1:
            raise NotImplementedError('IQNDiscriminatorOutput: only out_dims=1 (as built by IQNDiscriminator)')
        self.activation = nn.Sequential(native(norm_factory)(in_dims), native(activation_factory)())
        self.to_output = nn.Sequential(Linear(in_dims, out_dims))
        self.iqn = IQN(in_dims)
        self.out_dims = out_dims

    def forward(self, feats, targets=None):
        feats = run_layers(self.activation, ops.ensure_internal(feats))
        feats = ops.spatial_sum(feats)                        # (B, C) fp32
        batch = feats.shape[0]
        nq = self.iqn.num_quantiles
        taus = self.iqn.sample_quantiles(batch)               # (nq*B, 1), CPU generator (Appendix B.10)
        emb = self.iqn.quantile_embedding.to_state[0]
        out = self.to_output[0]
        # embedding + mix + Linear(C -> 1) + mean over quantiles + quantile-Huber loss: ONE kernel forward, ONE backward
        if targets is not None:
            assert not targets.requires_grad
            return ops.IqnHeadLossFn.apply(feats, taus, emb.weight, emb.bias, out.weight, out.bias, targets, nq, 1.0)
        return ops.IqnHeadLossFn.apply(feats, taus, emb.weight, emb.bias, out.weight, out.bias, None, nq, 1.0)


class _Unsupported(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f'{type(self).__name__} belongs to the info/text trainers, which are outside '
                                  'the SA-GAN / SA-GAN-IQN training step this package implements')


class DiscriminatorPoolOnlyOutput(_Unsupported):
    pass


class MultiModelDiscriminatorOutput(_Unsupported):
    pass


class LinearOutput(_Unsupported):
    pass


class GaussianParametersOutput(_Unsupported):
    pass
