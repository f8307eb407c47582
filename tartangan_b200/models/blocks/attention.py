"""SelfAttention2d — interface of tartangan/models/blocks/attention.py (BigGAN-style,
2x2 max-pooled keys/values, learnable gamma initialised to 0)."""
import torch
from torch import nn

from ... import ops
from ..layers import Conv2d


class SelfAttention2d(nn.Module):
    def __init__(self, in_dims, attention_dims=None):
        super().__init__()
        self.in_dims = in_dims
        self.theta = Conv2d(in_dims, in_dims // 8, 1, bias=False)
        self.phi = Conv2d(in_dims, in_dims // 8, 1, bias=False)
        self.g = Conv2d(in_dims, in_dims // 2, 1, bias=False)
        self.o = Conv2d(in_dims // 2, in_dims, 1, bias=False)
        self.gamma = nn.Parameter(torch.tensor(0.), requires_grad=True)

    def forward(self, x, y=None):
        x = ops.ensure_internal(x)
        n, c, h, w = x.shape
        xs, x1 = ops.fork(x)
        x1, x2 = ops.fork(x1)
        x2, x3 = ops.fork(x2)
        # NHWC memory: a feature map is a (positions x channels) matrix per image
        theta = self.theta(x1).permute(0, 2, 3, 1).reshape(n, h * w, c // 8)
        phi = ops.max_pool2(self.phi(x2)).permute(0, 2, 3, 1).reshape(n, h * w // 4, c // 8)
        g = ops.max_pool2(self.g(x3)).permute(0, 2, 3, 1).reshape(n, h * w // 4, c // 2)
        mixed = ops.attention_core(theta, phi, g)                                 # (n, HW, C/2); beta = (n, HW, HW/4)
        mixed = mixed.reshape(n, h, w, c // 2).permute(0, 3, 1, 2)
        o = self.o(mixed)
        return ops.add(ops.ScaleDevFn.apply(o, self.gamma.reshape(1)), xs)
