"""Shared body of CNNTrainer / IQNTrainer: model construction from the factory seam and the
D-step / G-step / EMA schedule of reference trainers/cnn.py:29-165 and trainers/iqn.py:30-156.
"""
import functools

import torch
from torch import nn

from .. import ops
from ..models.blocks import (GeneratorInputMLP, GeneratorOutput, ResidualDiscriminatorBlock,
                             ResidualGeneratorBlock, TiledZGeneratorInput)
from ..models.layers import BatchNorm2d, LeakyReLU
from ..models.losses import gradient_penalty
from ..models.pluggan import GAN_CONFIGS, Generator
from ..optim import FlatParams, FusedAdam
from .trainer import Trainer
from .utils import toggle_grad


class GanTrainer(Trainer):
    discriminator_cls = None          # set by subclasses
    d_output_cls = None

    # ------------------------------------------------------------------ construction
    def build_models(self):
        args = self.args
        ops.set_precision(getattr(args, 'precision', 'bf16'))
        cfg = GAN_CONFIGS[args.config]
        if getattr(args, 'attention', None):
            cfg = cfg._replace(attention=tuple(int(i) for i in str(args.attention).split(',') if i != ''))
        self.gan_config = cfg.scale_model(args.model_scale)
        norm_factory = {'id': nn.Identity, 'bn': BatchNorm2d}[args.norm]
        g_input_factory = {'mlp': GeneratorInputMLP, 'tiledz': TiledZGeneratorInput}[args.g_base]
        if args.activation != 'relu':
            raise NotImplementedError(f'--activation {args.activation}: only "relu" (LeakyReLU 0.2, the default) '
                                      'has kernels; selu/elu are not implemented')
        activation_factory = functools.partial(LeakyReLU, 0.2)
        g_factories = dict(
            input_factory=functools.partial(g_input_factory, activation_factory=activation_factory),
            block_factory=functools.partial(ResidualGeneratorBlock, norm_factory=norm_factory,
                                            activation_factory=activation_factory),
            output_factory=functools.partial(GeneratorOutput, norm_factory=norm_factory,
                                             activation_factory=activation_factory))
        self.g = Generator(self.gan_config, **g_factories).to(self.device)
        self.target_g = Generator(self.gan_config, **g_factories).to(self.device)
        self.d = self.discriminator_cls(
            self.gan_config,
            block_factory=functools.partial(ResidualDiscriminatorBlock, norm_factory=norm_factory,
                                            activation_factory=activation_factory),
            output_factory=functools.partial(self.d_output_cls, norm_factory=norm_factory,
                                             activation_factory=activation_factory),
        ).to(self.device)
        nq = getattr(args, 'num_quantiles', None)
        if nq and hasattr(self.d, 'to_output') and hasattr(self.d.to_output, 'iqn'):
            self.d.to_output.iqn.num_quantiles = nq
        self.optimizer_g = FusedAdam(self.g.parameters(), lr=args.lr_g, betas=(0., 0.999))
        self.optimizer_d = FusedAdam(self.d.parameters(), lr=args.lr_d, betas=(0., 0.999))
        self.update_target_generator(1.)       # NB: like the reference this is an EMA step, not a copy (B.1)
        self._target_flat = None
        self.world_size, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world_size, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
            self.broadcast_parameters()

    def broadcast_parameters(self):
        """Data parallel start-up: every rank adopts rank 0's parameters and buffers."""
        for m in (self.g, self.target_g, self.d):
            for t in list(m.parameters()) + list(m.buffers()):
                torch.distributed.broadcast(t.data, 0)

    # ------------------------------------------------------------------ losses (subclass hooks)
    def d_losses(self, real, fake):
        """-> (p_real, d_loss without penalty)"""
        raise NotImplementedError

    def g_loss(self, fake):
        raise NotImplementedError

    # ------------------------------------------------------------------ one optimisation step
    def _reducer(self, optimizer):
        """Bucketed all-reduce over the optimiser's flat gradient buffer (data parallel only)."""
        if self.world_size == 1:
            return None
        flat = optimizer._ensure_flat()
        red = getattr(optimizer, '_ttg_reducer', None)
        if red is None or red.flat_grad is not flat.grad:
            from ..parallel import BucketedAllReduce
            red = BucketedAllReduce(flat.params, flat.offsets, flat.grad, num_buckets=2)
            optimizer._ttg_reducer = red
        return red

    def _backward(self, loss, optimizer):
        """loss.backward() with the gradient exchange overlapped: the loss is scaled by 1/world so the
        summed gradients are the global-batch average."""
        red = self._reducer(optimizer)
        if red is None:
            loss.backward()
            return
        red.begin()
        loss.backward(torch.full_like(loss, 1.0 / self.world_size))
        red.finish()

    def d_step(self, imgs):
        toggle_grad(self.g, False)
        toggle_grad(self.d, True)
        self.optimizer_d.zero_grad()
        with torch.no_grad():
            fake = self.sample_g(len(imgs))
        real = imgs
        if self.args.grad_penalty:
            real = imgs.detach().requires_grad_()
        p_real, d_loss = self.d_losses(real, fake)
        gp = None
        if self.args.grad_penalty:
            gp = gradient_penalty(p_real, real)
            d_loss = ops.AxpbyFn.apply(d_loss, gp, 1.0, float(self.args.grad_penalty))
        self._backward(d_loss, self.optimizer_d)
        self.optimizer_d.step()
        return d_loss, gp

    def g_step(self, imgs):
        toggle_grad(self.g, True)
        toggle_grad(self.d, False)
        self.optimizer_g.zero_grad()
        fake = self.sample_g(len(imgs))
        g_loss = self.g_loss(fake)
        self._backward(g_loss, self.optimizer_g)
        # Adam and the EMA of target_g (update_target_generator) are one kernel
        self.optimizer_g.ema_target = self._flat_target()
        self.optimizer_g.ema_lr = float(self.args.lr_target_g)
        self.optimizer_g.step()
        return g_loss

    def train_batch(self, imgs, as_floats=True):
        imgs = imgs.to(self.device, non_blocking=True)
        self.g.train()
        self.d.train()
        d_loss, gp = self.d_step(imgs)
        g_loss = self.g_step(imgs)
        if not as_floats:
            return dict(g_loss=g_loss.detach(), d_loss=d_loss.detach(), gp=gp)
        gp_val = float(gp.detach()) * self.args.grad_penalty if gp is not None else 0.
        return dict(g_loss=float(g_loss.detach()), d_loss=float(d_loss.detach()), gp=gp_val)

    def _flat_target(self):
        if self._target_flat is None or not self._target_flat.intact():
            self._target_flat = FlatParams(list(self.target_g.parameters()))
        return self._target_flat

    @torch.no_grad()
    def update_target_generator(self, lr=None):
        """target += (g - target) * lr_target_g over parameters only; like the reference the `lr`
        argument is ignored (trainers/cnn.py:158-165, Appendix B.1)."""
        for g_p, t_p in zip(self.g.parameters(), self.target_g.parameters()):
            src, dst = g_p.data.contiguous(), t_p.data
            ops.call('ttg_ema_flat', ops.ptr(dst), ops.ptr(src), dst.numel(), float(self.args.lr_target_g))


def make_trainer(cls, config='64', batch_size=16, gan_config=None, device='cuda', **overrides):
    """Build a trainer without the CLI / filesystem side effects (tests, bench, smoke).
    `gan_config` (a GANConfig) may be given instead of a GAN_CONFIGS key."""
    import argparse
    p = argparse.ArgumentParser()
    cls.add_args_to_parser(p)
    args = p.parse_args(['synthetic', '--batch-size', str(batch_size), '--config', str(config)])
    for k, v in overrides.items():
        if not hasattr(args, k):
            raise AttributeError(f'unknown trainer argument {k}')
        setattr(args, k, v)
    args.device = device
    if gan_config is not None:
        GAN_CONFIGS[f'_custom_{id(gan_config)}'] = gan_config
        args.config = f'_custom_{id(gan_config)}'
    t = cls.__new__(cls)
    t.args, t.steps, t.epoch, t.components, t.run_id = args, 0, 1, [], 'adhoc'
    t.build_models()
    return t
