from .attention import SelfAttention2d  # noqa
from .discriminator import (  # noqa
    DiscriminatorBlock, DiscriminatorInput, DiscriminatorOutput, DiscriminatorPoolOnlyOutput,
    IQNDiscriminatorOutput, LinearOutput, GaussianParametersOutput, MultiModelDiscriminatorOutput,
    ResidualDiscriminatorBlock,
)
from .generator import (  # noqa
    GeneratorBlock, GeneratorInputMLP, GeneratorInputMLP1d, GeneratorOutput,
    ResidualGeneratorBlock, TiledZGeneratorInput,
)
