"""SA-GAN-IQN trainer — interface of tartangan/trainers/iqn.py (IQNTrainer, main)."""
import torch

from .. import ops
from ..models.blocks import IQNDiscriminatorOutput
from ..models.pluggan import IQNDiscriminator
from .gan import GanTrainer


class IQNTrainer(GanTrainer):
    discriminator_cls = IQNDiscriminator
    d_output_cls = IQNDiscriminatorOutput

    def d_losses(self, real, fake):
        """D returns (mean-over-quantile prediction, quantile-Huber loss) (iqn.py:118-120)."""
        n = len(real)
        p_real, l_real = self.d(real, targets=torch.ones(n, 1, device=self.device))
        _, l_fake = self.d(fake.detach(), targets=torch.zeros(n, 1, device=self.device))
        return p_real, ops.add(l_real, l_fake)

    def g_loss(self, fake):
        _, loss = self.d(fake, targets=torch.ones(len(fake), 1, device=self.device))
        return loss


def main():
    IQNTrainer.create_from_cli().train()


if __name__ == '__main__':
    main()
