"""SA-GAN trainer — interface of tartangan/trainers/cnn.py (CNNTrainer, main)."""
import torch

from .. import ops
from ..models.blocks import DiscriminatorOutput
from ..models.pluggan import Discriminator
from .gan import GanTrainer


class CNNTrainer(GanTrainer):
    discriminator_cls = Discriminator
    d_output_cls = DiscriminatorOutput

    def d_losses(self, real, fake):
        """BCEWithLogits over cat[p_real, p_fake] vs [1.., 0..] (cnn.py:122-131) = the mean of the
        two half-batch means; D(real) and D(fake) stay separate forwards (separate BN statistics)."""
        p_real = self.d(real)
        p_fake = self.d(fake.detach())
        ones = torch.ones_like(p_real)
        l_real = ops.BceLogitsFn.apply(p_real, ones)
        l_fake = ops.BceLogitsFn.apply(p_fake, torch.zeros_like(p_fake))
        return p_real, ops.AxpbyFn.apply(l_real, l_fake, 0.5, 0.5)

    def g_loss(self, fake):
        p = self.d(fake)
        return ops.BceLogitsFn.apply(p, torch.ones_like(p))


def main():
    CNNTrainer.create_from_cli().train()


if __name__ == '__main__':
    main()
