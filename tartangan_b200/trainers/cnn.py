"""SA-GAN trainer — interface of tartangan/trainers/cnn.py (CNNTrainer, main)."""
import torch

from .. import ops
from ..models.blocks import DiscriminatorOutput
from ..models.pluggan import Discriminator
from .gan import GanTrainer


class CNNTrainer(GanTrainer):
    discriminator_cls = Discriminator
    d_output_cls = DiscriminatorOutput

    # BCEWithLogits over cat[p_real, p_fake] vs [1.., 0..] (cnn.py:122-131) = the mean of the two half-batch means;
    # D(real) and D(fake) stay separate forwards (separate BN statistics)
    def d_real(self, real):
        p_real = self.d(real)
        return p_real, ops.BceLogitsFn.apply(p_real, torch.ones_like(p_real))

    def d_fake(self, fake):
        p_fake = self.d(fake.detach())
        return ops.BceLogitsFn.apply(p_fake, torch.zeros_like(p_fake))

    def d_combine(self, l_real, l_fake):
        return ops.AxpbyFn.apply(l_real, l_fake, 0.5, 0.5)

    def g_loss(self, fake):
        p = self.d(fake)
        return ops.BceLogitsFn.apply(p, torch.ones_like(p))


def main():
    CNNTrainer.create_from_cli().train()


if __name__ == '__main__':
    main()
