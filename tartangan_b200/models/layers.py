"""Leaf modules: parameter containers identical to torch.nn's (same init, same
state-dict keys) whose forward runs the sm_100a kernels.  Mirrors the leaf modules
the reference uses (nn.Conv2d / nn.BatchNorm2d / nn.LeakyReLU / nn.Linear /
nn.AvgPool2d / nn.Tanh) plus tartangan/models/layers.py (Interpolate, PixelNorm).
"""
import functools

import torch
from torch import nn

from .. import ops


class Conv2d(nn.Conv2d):
    """nn.Conv2d restricted to what the reference path uses: k in {1,3}, stride 1, padding k//2."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias, **kw)
        k = self.kernel_size[0]
        if (self.kernel_size not in ((1, 1), (3, 3)) or self.stride != (1, 1) or self.padding != (k // 2, k // 2)
                or self.dilation != (1, 1) or self.groups != 1):
            raise NotImplementedError('tartangan_b200.Conv2d supports kernel 1 or 3, stride 1, padding k//2 '
                                      f'(got k={self.kernel_size}, stride={self.stride}, padding={self.padding})')

    def forward(self, x, up=0, out_dtype=None, stats=False):
        return ops.conv2d(ops.ensure_internal(x), self.weight, self.bias, up, out_dtype, stats)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d; blocks fuse it with the following LeakyReLU (ops.bn_act)."""

    def forward(self, x, slope=1.0):
        return ops.bn_act(ops.ensure_internal(x), self, slope)


class LeakyReLU(nn.LeakyReLU):
    def forward(self, x):
        return ops.leaky_relu(x, self.negative_slope)


class ELU(nn.ELU):
    """nn.ELU (--activation elu, reference trainers/cnn.py:44)."""

    def forward(self, x):
        return ops.elu(x, self.alpha)


class SELU(nn.SELU):
    """nn.SELU (--activation selu, reference trainers/cnn.py:43, trainers/iqn.py:44)."""

    def forward(self, x):
        return ops.selu(x)


class AvgPool2d(nn.AvgPool2d):
    def __init__(self, kernel_size=2, **kw):
        super().__init__(kernel_size, **kw)
        if self.kernel_size not in (2, (2, 2)):
            raise NotImplementedError('tartangan_b200.AvgPool2d supports kernel_size=2 only')

    def forward(self, x):
        return ops.avg_pool2(x)


class Linear(nn.Linear):
    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class Tanh(nn.Tanh):
    def forward(self, x):
        return ops.tanh(x)


class Interpolate(nn.Module):
    """tartangan/models/layers.py:6-13; only the two modes the blocks use are implemented."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.args, self.kwargs = args, kwargs

    def forward(self, x):
        return interpolate(x, *self.args, **self.kwargs)


def interpolate(x, size=None, scale_factor=None, mode='nearest', align_corners=None):
    x = ops.ensure_internal(x)
    if size is None and scale_factor == 2 and mode == 'nearest':
        return ops.upsample2(x)
    if size is None and scale_factor == 0.5 and mode == 'bilinear' and align_corners:
        return ops.bilinear_down(x)
    raise NotImplementedError(f'tartangan_b200.interpolate: scale_factor={scale_factor}, mode={mode}, '
                              f'align_corners={align_corners} is not on the reference path')


class PixelNorm(nn.Module):
    """tartangan/models/layers.py:16-22 (imported by trainers/iqn.py, never instantiated there)."""

    def __init__(self, eps=1e-8):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        raise NotImplementedError('PixelNorm is not used by the cnn/iqn trainers; no kernel is provided')


def native(factory):
    """Map a torch.nn factory handed to a block constructor onto the kernel-backed twin."""
    table = {nn.Conv2d: Conv2d, nn.BatchNorm2d: BatchNorm2d, nn.LeakyReLU: LeakyReLU, nn.SELU: SELU, nn.ELU: ELU,
             nn.AvgPool2d: AvgPool2d, nn.Linear: Linear, nn.Tanh: Tanh}
    if isinstance(factory, functools.partial):
        return functools.partial(native(factory.func), *factory.args, **factory.keywords)
    if factory in table:
        return table[factory]
    return factory


def run_layers(layers, x, up_first=False):
    """Run an nn.Sequential-style list, fusing (BatchNorm2d | Identity) + LeakyReLU pairs into one op.

    up_first=True: x is the LOW-resolution input of a block that starts with a nearest x2 upsample.
    Element-wise layers before the first conv run at low resolution (they commute with nearest
    upsampling, batch statistics included) and that conv reads its input through (y>>1, x>>1)
    addressing, so the upsampled tensor is never written."""
    layers = list(layers)
    i = 0
    pending_up = up_first
    while i < len(layers):
        m = layers[i]
        nxt = layers[i + 1] if i + 1 < len(layers) else None
        if isinstance(m, BatchNorm2d) and isinstance(nxt, LeakyReLU):
            x = ops.bn_act(ops.ensure_internal(x), m, nxt.negative_slope, 4 if pending_up else 1)
            i += 2
        elif isinstance(m, BatchNorm2d):
            # followed by ELU / SELU (--activation elu|selu): plain BatchNorm (slope 1), the activation is its own kernel
            x = ops.bn_act(ops.ensure_internal(x), m, 1.0, 4 if pending_up else 1)
            i += 1
        elif isinstance(m, nn.Identity):
            i += 1
        elif isinstance(m, Conv2d):
            # a conv that feeds a train-mode BatchNorm reduces that layer's batch statistics in its epilogue
            feeds_bn = isinstance(nxt, BatchNorm2d) and (nxt.training or nxt.running_mean is None)
            x = m(x, up=1 if pending_up else 0, stats=feeds_bn)
            pending_up = False
            i += 1
        elif pending_up and not isinstance(m, (LeakyReLU, ELU, SELU)):
            x = m(ops.upsample2(x))          # unknown layer type: materialise the upsample first
            pending_up = False
            i += 1
        else:
            x = m(x)
            i += 1
    if pending_up:
        x = ops.upsample2(x)
    return x


class SpectralNormConv2d(Conv2d):
    """Conv2d whose weight is divided by its largest singular value, estimated with one power-iteration
    step per training forward — the semantics of torch.nn.utils.spectral_norm (the hook API that
    tartangan/prep4web.py:33-51 strips): parameters `weight_orig`, buffers `weight_u`, `weight_v`.
    Plugs into the blocks through their conv_factory seam (generator.py:34, discriminator.py:28,52)."""

    def __init__(self, *args, n_power_iterations=1, eps=1e-12, **kw):
        super().__init__(*args, **kw)
        w = self.weight
        del self._parameters['weight']
        self.register_parameter('weight_orig', nn.Parameter(w.data))
        rows, cols = w.shape[0], w[0].numel()
        u = torch.nn.functional.normalize(w.new_empty(rows).normal_(0, 1), dim=0, eps=eps)
        v = torch.nn.functional.normalize(w.new_empty(cols).normal_(0, 1), dim=0, eps=eps)
        self.register_buffer('weight_u', u)
        self.register_buffer('weight_v', v)
        self.n_power_iterations, self.sn_eps = n_power_iterations, eps

    def forward(self, x, up=0, out_dtype=None, stats=False):
        w = ops.SpectralNormFn.apply(self.weight_orig, self.weight_u, self.weight_v,
                                     self.n_power_iterations if self.training else 0, self.sn_eps)
        return ops.conv2d(ops.ensure_internal(x), w, self.bias, up, out_dtype, stats)
