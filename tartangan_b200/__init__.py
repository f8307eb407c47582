"""tartangan_b200 — B200-native (sm_100a) implementation of tartangan's GAN training step.

Drop-in for the path tartangan.models.* / tartangan.trainers.{cnn,iqn}: same module names,
constructors, state-dict keys and CLI; the arithmetic runs in hand-written CUDA kernels
reached through the C ABI in include/ttg_b200.h.  There is no CPU path.
"""
from .ops import set_precision, get_precision  # noqa: F401

__version__ = '0.1.0'


def install_as_tartangan():
    """Alias this package's mirror modules under the reference's import paths so that code
    (and pickled checkpoints) naming ``tartangan.models.pluggan.Generator`` etc. resolve here."""
    import sys
    import types
    from . import models, trainers
    from .models import blocks, iqn, layers, losses, pluggan
    from .trainers import cnn as t_cnn, iqn as t_iqn, trainer as t_trainer, utils as t_utils
    root = types.ModuleType('tartangan')
    root.models, root.trainers = models, trainers
    table = {
        'tartangan': root, 'tartangan.models': models, 'tartangan.models.blocks': blocks,
        'tartangan.models.blocks.generator': blocks.generator,
        'tartangan.models.blocks.discriminator': blocks.discriminator,
        'tartangan.models.blocks.attention': blocks.attention,
        'tartangan.models.iqn': iqn, 'tartangan.models.layers': layers, 'tartangan.models.losses': losses,
        'tartangan.models.pluggan': pluggan, 'tartangan.trainers': trainers,
        'tartangan.trainers.cnn': t_cnn, 'tartangan.trainers.iqn': t_iqn,
        'tartangan.trainers.trainer': t_trainer, 'tartangan.trainers.utils': t_utils,
    }
    for name, mod in table.items():
        sys.modules.setdefault(name, mod)
