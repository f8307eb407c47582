"""SA-GAN-IQN trainer — interface of tartangan/trainers/iqn.py (IQNTrainer, main)."""
import torch

from .. import ops
from ..models.blocks import IQNDiscriminatorOutput
from ..models.pluggan import IQNDiscriminator
from .gan import GanTrainer


class IQNTrainer(GanTrainer):
    discriminator_cls = IQNDiscriminator
    d_output_cls = IQNDiscriminatorOutput

    # D returns (mean-over-quantile prediction, quantile-Huber loss) (iqn.py:118-120); d_loss = l_real + l_fake
    def d_real(self, real):
        return self.d(real, targets=torch.ones(len(real), 1, device=self.device))

    def d_fake(self, fake):
        return self.d(fake.detach(), targets=torch.zeros(len(fake), 1, device=self.device))[1]

    def d_combine(self, l_real, l_fake):
        return ops.add(l_real, l_fake)

    def g_loss(self, fake):
        _, loss = self.d(fake, targets=torch.ones(len(fake), 1, device=self.device))
        return loss


def main():
    IQNTrainer.create_from_cli().train()


if __name__ == '__main__':
    main()
