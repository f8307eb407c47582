"""Write tests/golden/ref_checkpoint/: a checkpoint saved by the UNMODIFIED reference (its
ModelCheckpointComponent.save_checkpoint, components/model_checkpoint.py:32-50: whole pickled objects) after one
CPU training step of a tiny IQN config, plus the plain state dicts the loaded objects must reproduce.
Build container only (needs /root/reference).

    python tools/make_ref_checkpoint.py
"""
import argparse
import contextlib
import io
import os
import sys
import warnings

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tools'))
from oracle.tartan_oracle import tartan_batch  # noqa: E402
from make_golden import _import_reference      # noqa: E402

BLOCKS, LATENT, BATCH = (16, 8, 8), 16, 4


def main():
    pluggan, cnn, iqn = _import_reference()
    from tartangan.trainers.components.model_checkpoint import ModelCheckpointComponent
    out = os.path.join(REPO, 'tests', 'golden', 'ref_checkpoint')
    pluggan.GAN_CONFIGS['ref_ckpt'] = pluggan.GANConfig(base_size=4, latent_dims=LATENT, data_dims=3, attention=(),
                                                        num_blocks_per_scale=1, blocks=BLOCKS)
    cls = iqn.IQNTrainer
    p = argparse.ArgumentParser()
    cls.add_args_to_parser(p)
    for cc in cls.get_component_classes(p.parse_known_args(['/unused'])[0]):
        cc.add_args_to_parser(p)
    args = p.parse_args(['/unused', '--batch-size', str(BATCH), '--config', 'ref_ckpt'])
    args.device = 'cpu'
    t = cls.__new__(cls)
    t.args, t.steps, t.epoch = args, 0, 1
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        t.build_models()
    torch.manual_seed(1000)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t.train_batch(tartan_batch(1234, BATCH, t.g.max_size))     # Adam moments exist
    t.steps = 1
    comp = ModelCheckpointComponent(args)
    comp.trainer = t
    type(t).output_root = property(lambda self: out)               # checkpoints/1/ under tests/golden/ref_checkpoint
    with contextlib.redirect_stdout(io.StringIO()):
        comp.save_checkpoint(1)
    expect = dict(g=t.g.state_dict(), target_g=t.target_g.state_dict(), d=t.d.state_dict(),
                  opt_d=t.optimizer_d.state_dict(), opt_g=t.optimizer_g.state_dict(),
                  blocks=BLOCKS, latent=LATENT, batch=BATCH)
    torch.save(expect, os.path.join(out, 'expected_state.pt'))
    for root, _, files in os.walk(out):
        for f in files:
            print(os.path.join(root, f), os.path.getsize(os.path.join(root, f)))


if __name__ == '__main__':
    main()
