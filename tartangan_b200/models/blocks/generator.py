"""Generator blocks — interface of tartangan/models/blocks/generator.py.

Same constructors, defaults and state-dict keys; the forward passes run the
hand-written kernels.  Fusions relative to the reference module graph: the nearest
x2 upsample (generator.py:58) is never materialised — the leading BatchNorm+LeakyReLU
and the 1x1 skip projection run at low resolution (they commute with nearest
upsampling, batch statistics included), the first 3x3 conv reads its input through
(y>>1, x>>1) addressing and the residual add reads the skip the same way; each
BatchNorm+LeakyReLU pair is one kernel.
"""
import functools
import os

from torch import nn

from ... import ops
from ..layers import (BatchNorm2d, Conv2d, Interpolate, LeakyReLU, Linear, Tanh, native, run_layers)


# TTG_GTAIL_BF16=1: the C -> 3 output conv writes bf16 through the 8-channel TMA staging (-0.2 ms per step).  Off by
# default: rounding the pre-activation to bf16 before tanh lowered the gradient cosine of one mid-generator BatchNorm
# beta from ~0.975 to ~0.968 in 2 of 5 runs of tests/test_gpu_parity_configs.py (bar 0.97); parity comes first.
_FP32_TAIL = os.environ.get('TTG_GTAIL_BF16', '0') != '1'
_RGB_HEAD = os.environ.get('TTG_RGB_HEAD', '1') == '1'       # A/B switch: the fused generator head (ops.RgbHeadFn)


class GeneratorBlock(nn.Module):
    """Non-residual block (generator.py:9-29): BlockModel's default, unused by the cnn/iqn
    trainers.  Like the reference, its first norm is sized with out_dims (Appendix B.14)."""

    def __init__(self, in_dims, out_dims, upsample=True, first_block=False, norm_factory=BatchNorm2d,
                 activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        norm_factory, activation_factory = native(norm_factory), native(activation_factory)
        layers = [norm_factory(out_dims), activation_factory(), Conv2d(in_dims, out_dims, 3, padding=1, bias=True),
                  norm_factory(out_dims), activation_factory(), Conv2d(out_dims, out_dims, 3, padding=1, bias=True)]
        if first_block:
            layers = layers[2:]
        if upsample:
            layers.insert(0, Interpolate(scale_factor=2, mode='nearest'))
        self.convs = nn.Sequential(*layers)

    def forward(self, x):
        return run_layers(self.convs, x)


class ResidualGeneratorBlock(nn.Module):
    """generator.py:32-62: up x2 -> [BN, act,] conv3 -> BN -> act -> conv3, plus (projected) skip."""

    def __init__(self, in_dims, out_dims, upsample=True, first_block=False, norm_factory=BatchNorm2d,
                 conv_factory=Conv2d, activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        norm_factory, conv_factory = native(norm_factory), native(conv_factory)
        activation_factory = native(activation_factory)
        layers = [norm_factory(in_dims), activation_factory(), conv_factory(in_dims, out_dims, 3, padding=1),
                  norm_factory(out_dims), activation_factory(), conv_factory(out_dims, out_dims, 3, padding=1)]
        if first_block:
            layers = layers[2:]
        self.upsample = upsample
        self.project_input = None
        if in_dims != out_dims:
            self.project_input = nn.Sequential(conv_factory(in_dims, out_dims, 1))
        self.convs = nn.Sequential(*layers)

    def forward(self, x):
        x = ops.ensure_internal(x)
        xs, xh = ops.fork(x)
        if self.project_input is not None:
            with ops.skip_branch(xs) as sb:               # beside the conv branch, on the companion stream
                xs = run_layers(self.project_input, xs)  # 1x1 conv commutes with nearest upsampling: run it low-res
            h = run_layers(self.convs, xh, up_first=self.upsample)
            xs = sb.join(xs)
        else:
            h = run_layers(self.convs, xh, up_first=self.upsample)
        return ops.add_up2(h, xs) if self.upsample else ops.add(xs, h)


class GeneratorInputMLP(nn.Module):
    """generator.py:65-80: Linear(latent -> size^2*C) -> act -> view(B, C, size, size)."""

    def __init__(self, latent_dims, output_dims, size=4, norm_factory=nn.BatchNorm1d,
                 activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        self.base_img = nn.Sequential(Linear(latent_dims, size ** 2 * output_dims), native(activation_factory)())
        self.latent_dims, self.output_dims, self.size = latent_dims, output_dims, size

    def forward(self, z):
        img = run_layers(self.base_img, z.float())
        return ops.to_internal(img.view(-1, self.output_dims, self.size, self.size))


class GeneratorInputMLP1d(nn.Module):
    """generator.py:83-98 (text trainer only): constructor kept for surface parity."""

    def __init__(self, latent_dims, output_dims, size=4, norm_factory=nn.BatchNorm1d,
                 activation_factory=functools.partial(LeakyReLU, 0.2)):
        super().__init__()
        self.base = nn.Sequential(Linear(latent_dims, size * output_dims), native(activation_factory)())
        self.latent_dims, self.output_dims, self.size = latent_dims, output_dims, size

    def forward(self, z):
        return run_layers(self.base, z.float()).view(-1, self.output_dims, self.size)


class TiledZGeneratorInput(nn.Module):
    """generator.py:101-112: z tiled over a size x size grid (requires latent_dims == output_dims)."""

    def __init__(self, latent_dims, output_dims, size=4, norm_factory=nn.BatchNorm2d, **_):
        super().__init__()
        self.size = size
        assert latent_dims == output_dims

    def forward(self, z):
        b, c = z.shape
        return ops.SpatialBcastFn.apply(z.float(), (b, c, self.size, self.size), ops.state.act_dtype)


class GeneratorOutput(nn.Module):
    """generator.py:115-129: BN -> act -> conv1x1(C -> data_dims) -> tanh; returns fp32 NCHW."""

    def __init__(self, in_dims, out_dims, norm_factory=BatchNorm2d, conv_factory=Conv2d,
                 activation_factory=functools.partial(LeakyReLU, 0.2), output_activation_factory=Tanh):
        super().__init__()
        self.convs = nn.Sequential(
            native(norm_factory)(in_dims), native(activation_factory)(),
            native(conv_factory)(in_dims, out_dims, 1, padding=0, bias=True),
            native(output_activation_factory)())

    def forward(self, x):
        layers = list(self.convs)
        x = run_layers(layers[:2], ops.ensure_internal(x))
        conv = layers[2]
        if (type(conv) is Conv2d and len(layers) == 4 and isinstance(layers[3], Tanh) and _RGB_HEAD
                and ops.rgb_head_ok(conv, x)):
            # bf16 mode, C -> 3: conv1x1 + tanh + the fp32 NCHW boundary as ONE streaming kernel (fp32 weights and
            # accumulation: no precision is given up, unlike the bf16 tail below)
            return ops.RgbHeadFn.apply(x, conv.weight, conv.bias)
        if (isinstance(conv, Conv2d) and len(layers) == 4 and isinstance(layers[3], Tanh)
                and ops.state.act_dtype == ops.torch.bfloat16 and not _FP32_TAIL):
            # opt-in (see _FP32_TAIL): the C -> 3 conv writes bf16 through the 8-channel TMA staging (45 + 17 us instead of
            # the 180 us fp32 3-channel epilogue), the layout boundary converts to fp32 NCHW and tanh (elementwise, so
            # it commutes with the layout change) runs on the flat fp32 image
            img = ops.from_internal(conv(x))
            return ops.tanh(img.reshape(-1)).view(img.shape)
        x = conv(x, out_dtype=ops.torch.float32) if isinstance(conv, Conv2d) else conv(x)
        x = run_layers(layers[3:], x)
        return ops.from_internal(x)
