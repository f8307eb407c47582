"""Generate tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Only runs in the build container (needs /root/reference).  The reference is
imported read-only with three import shims for modules that are outside the
arithmetic; nothing from it is copied into the repo,
only the tensors it produces.

    python tools/make_golden.py [case ...]   # rewrites every case (or the named ones)
"""
import argparse
import contextlib
import io
import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.tartan_oracle import tartan_batch  # noqa: E402  (input generator only)


def _import_reference():
    sys.path.insert(0, '/root/reference')
    so = types.ModuleType('smart_open')
    so.open = open
    sys.modules['smart_open'] = so
    b3 = types.ModuleType('boto3')
    b3.resource = b3.client = (lambda *a, **k: None)
    sys.modules['boto3'] = b3
    import tqdm._utils
    tqdm._utils._unicode = str
    import tartangan.models.pluggan as pluggan
    import tartangan.trainers.cnn as cnn
    import tartangan.trainers.iqn as iqn
    return pluggan, cnn, iqn


CASES = {
    # name: (trainer kind, blocks, attention, latent, batch, steps, norm)
    'cnn_tiny': ('cnn', (16, 8, 8), (), 16, 4, 3, 'bn'),
    'iqn_tiny': ('iqn', (16, 8, 8), (), 16, 4, 3, 'bn'),
    'cnn_attn': ('cnn', (16, 16, 8), (1,), 16, 2, 2, 'bn'),
    'iqn_attn': ('iqn', (16, 16, 8), (1,), 16, 2, 2, 'bn'),
    'iqn_nonorm': ('iqn', (16, 8), (), 16, 4, 2, 'id'),
    # round 2: --activation selu|elu (incl. init_params_selu) and --g-base tiledz; optional fields (activation, g_base)
    'cnn_selu': ('cnn', (16, 8, 8), (), 16, 4, 2, 'id', 'selu', 'mlp'),      # BatchNorm + ELU-family kernels: cnn_elu
    'iqn_selu': ('iqn', (16, 8), (), 16, 4, 2, 'id', 'selu', 'mlp'),       # (with 'bn' the selu init zeroes gamma: dead path)
    'cnn_elu': ('cnn', (16, 8, 8), (), 16, 4, 2, 'bn', 'elu', 'mlp'),
    'iqn_tiledz': ('iqn', (16, 8, 8), (), 16, 4, 2, 'bn', 'relu', 'tiledz'),
}


def _clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def make_case(name, pluggan, cnn, iqn):
    kind, blocks, attention, latent, batch, steps, norm = CASES[name][:7]
    activation, g_base = (CASES[name][7:] + ('relu', 'mlp'))[:2] if len(CASES[name]) > 7 else ('relu', 'mlp')
    cfg_key = f'golden_{name}'
    pluggan.GAN_CONFIGS[cfg_key] = pluggan.GANConfig(
        base_size=4, latent_dims=latent, data_dims=3, attention=attention,
        num_blocks_per_scale=1, blocks=blocks)
    mod = cnn if kind == 'cnn' else iqn
    cls = mod.CNNTrainer if kind == 'cnn' else mod.IQNTrainer
    p = argparse.ArgumentParser()
    cls.add_args_to_parser(p)
    for cc in cls.get_component_classes(p.parse_known_args(['/unused'])[0]):
        cc.add_args_to_parser(p)
    args = p.parse_args(['/unused', '--batch-size', str(batch), '--config', cfg_key,
                         '--norm', norm, '--activation', activation, '--g-base', g_base])
    args.device = 'cpu'
    t = cls.__new__(cls)
    t.args, t.steps, t.epoch = args, 0, 1
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        t.build_models()
    if attention:   # gamma initialises to 0 (attention.py:19): make it matter
        with torch.no_grad():
            for m in list(t.g.modules()) + list(t.target_g.modules()) + list(t.d.modules()):
                if hasattr(m, 'gamma'):
                    m.gamma.fill_(0.5)
    size = t.g.max_size
    out = dict(case=name, kind=kind, blocks=blocks, attention=attention, latent=latent,
               batch=batch, steps=steps, norm=norm, size=size,
               **({'activation': activation, 'g_base': g_base} if len(CASES[name]) > 7 else {}),
               init=dict(g=_clone_sd(t.g), target_g=_clone_sd(t.target_g), d=_clone_sd(t.d)))
    # forward-only probes on the initial state (train-mode BN, buffers restored after)
    torch.manual_seed(77)
    z = torch.randn(batch, latent)
    x = tartan_batch(99, batch, size)
    saved = (_clone_sd(t.g), _clone_sd(t.d))
    with torch.no_grad():
        g_out = t.g(z)
        if kind == 'iqn':
            torch.manual_seed(78)
            d_out, d_loss = t.d(x, targets=torch.ones(batch, 1))
            probe = dict(z=z, x=x, g_out=g_out, d_out=d_out, d_loss=d_loss, tau_seed=78)
        else:
            probe = dict(z=z, x=x, g_out=g_out, d_out=t.d(x))
    t.g.load_state_dict(saved[0])
    t.d.load_state_dict(saved[1])
    out['probe'] = probe
    # training steps
    out['imgs'], out['seeds'], out['metrics'] = [], [], []
    for s in range(steps):
        imgs = tartan_batch(1234 + s, batch, size)
        seed = 1000 + s
        torch.manual_seed(seed)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            m = t.train_batch(imgs)
        out['imgs'].append(imgs)
        out['seeds'].append(seed)
        out['metrics'].append(m)
        if s == 0:
            # grads left on the parameters: D grads from the D step (G step runs
            # with D frozen but does not clear them), G grads from the G step.
            out['grads0'] = dict(
                d={k: v.grad.detach().clone() for k, v in t.d.named_parameters() if v.grad is not None},
                g={k: v.grad.detach().clone() for k, v in t.g.named_parameters() if v.grad is not None})
    out['final'] = dict(g=_clone_sd(t.g), target_g=_clone_sd(t.target_g), d=_clone_sd(t.d))
    out['opt_d_exp_avg_sq'] = [st['exp_avg_sq'].clone() for st in t.optimizer_d.state.values()]
    return out


def main():
    pluggan, cnn, iqn = _import_reference()
    os.makedirs(os.path.join(REPO, 'tests', 'golden'), exist_ok=True)
    for name in (sys.argv[1:] or CASES):
        out = make_case(name, pluggan, cnn, iqn)
        path = os.path.join(REPO, 'tests', 'golden', f'{name}.pt')
        torch.save(out, path)
        print(name, out['metrics'], os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
