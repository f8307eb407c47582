"""Model assembly — interface of tartangan/models/pluggan.py (GANConfig, BlockModel,
Generator, Discriminator, IQNDiscriminator, GAN_CONFIGS).  Same factory seam
(input_factory / block_factory / output_factory), same registration order, hence the
same state-dict keys and the same parameter initialisation stream for a given seed.
"""
from collections import namedtuple

from torch import nn

from .blocks import (DiscriminatorBlock, DiscriminatorInput, DiscriminatorOutput, GeneratorBlock,
                     GeneratorOutput, SelfAttention2d, TiledZGeneratorInput)

_GANConfigBase = namedtuple('GANConfig',
                            'base_size, latent_dims, data_dims, blocks, num_blocks_per_scale, attention')


class GANConfig(_GANConfigBase):
    def scale_model(self, scale):
        """pluggan.py:24-28: every width is int(width * scale); blocks becomes a list."""
        return self._replace(blocks=[int(c * scale) for c in self.blocks])


class BlockModel(nn.Module):
    def __init__(self, config, input_factory=None, block_factory=None, output_factory=None):
        super().__init__()
        self.config = config
        self.input_factory = input_factory or self.default_input
        self.block_factory = block_factory or self.default_block
        self.output_factory = output_factory or self.default_output
        self.build()

    def build(self):
        raise NotImplementedError

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        return x

    @property
    def max_size(self):
        return self.config.base_size * 2 ** len(self.config.blocks)


class Generator(BlockModel):
    default_input = TiledZGeneratorInput
    default_block = GeneratorBlock
    default_output = GeneratorOutput

    def build(self):
        cfg = self.config
        width = cfg.blocks[0]
        stack = [self.input_factory(cfg.latent_dims, width, cfg.base_size)]
        for level, out_width in enumerate(cfg.blocks):
            stack.append(self.block_factory(width, out_width, first_block=(level == 0)))
            stack.extend(self.block_factory(out_width, out_width, upsample=False)
                         for _ in range(cfg.num_blocks_per_scale - 1))
            if cfg.attention and level in cfg.attention:
                stack.append(SelfAttention2d(out_width))
            width = out_width
        stack.append(self.output_factory(width, cfg.data_dims))
        self.blocks = nn.Sequential(*stack)


class Discriminator(BlockModel):
    default_input = DiscriminatorInput
    default_block = DiscriminatorBlock
    default_output = DiscriminatorOutput

    def _trunk(self, width, first_flag):
        """Residual down-blocks from the finest level to the coarsest (reversed config.blocks)."""
        cfg = self.config
        stack = []
        levels = list(enumerate(cfg.blocks))[::-1]
        for n, (level, out_width) in enumerate(levels):
            if first_flag:
                stack.append(self.block_factory(width, out_width, first_block=(n == 0)))
            else:
                stack.append(self.block_factory(width, out_width))
            if cfg.attention and level in cfg.attention:
                stack.append(SelfAttention2d(out_width))
            width = out_width
        return stack, width

    def build(self):
        cfg = self.config
        width = cfg.blocks[-1]
        stack = [self.input_factory(cfg.data_dims, width)]
        trunk, width = self._trunk(width, True)
        stack += trunk
        stack.append(self.output_factory(width, 1))
        self.blocks = nn.Sequential(*stack)


class IQNDiscriminator(Discriminator):
    """pluggan.py:114-132: no input conv and no first_block; the head is registered BEFORE the
    trunk, so its parameters come first in parameters()/state_dict()."""
    default_output = DiscriminatorOutput

    def build(self):
        trunk, width = self._trunk(self.config.data_dims, False)
        self.to_output = self.output_factory(width, 1)
        self.blocks = nn.Sequential(*trunk)

    def forward(self, x, targets=None):
        for block in self.blocks:
            x = block(x)
        return self.to_output(x, targets=targets)


def _cfg(latent, blocks, attention=()):
    return GANConfig(base_size=4, latent_dims=latent, data_dims=3, blocks=tuple(blocks),
                     num_blocks_per_scale=1, attention=tuple(attention))


# Named sizes of pluggan.py:199-406 (output side = 4 * 2**len(blocks)).
GAN_CONFIGS = {
    '16': _cfg(100, (64, 32)),
    '32': _cfg(128, (128, 64, 32)),
    '64': _cfg(128, (128, 128, 64, 32)),
    '128': _cfg(256, (128, 128, 64, 32, 16)),
    '128big': _cfg(256, (1024, 1024, 512, 256, 128)),
    '256': _cfg(256, (256, 256, 128, 64, 32, 16)),
    '256big': _cfg(256, (1024, 1024, 512, 256, 128, 64)),
    '512': _cfg(512, (256, 256, 256, 128, 64, 32, 16)),
    '512thin': _cfg(256, (128, 128, 128, 64, 32, 16, 8), (3,)),
    '512thin-test': _cfg(128, (128, 120, 100, 64, 32, 16, 8), (3,)),
    '1024': _cfg(512, (512, 512, 512, 256, 128, 64, 32, 16), (3,)),
    '1024thin': _cfg(256, (256, 256, 256, 128, 64, 32, 16, 8), (3,)),
    'test128': _cfg(64, (64, 32, 16, 8, 4), (3,)),
    'test256': _cfg(256, (200, 180, 128, 64, 32, 16), (3,)),
    # additive entries for BASELINE.json configs 4 and 5 (the reference has no CLI flag for attention)
    '256sa': _cfg(256, (256, 256, 128, 64, 32, 16), (3,)),
    '512sa': _cfg(512, (256, 256, 256, 128, 64, 32, 16), (3,)),
}
