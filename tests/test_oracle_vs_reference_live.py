"""Pins the CPU oracle - and the command-line surface of the mirror trainers - to the LIVE reference at the NAMED
configurations (build container only: needs /root/reference).

tests/golden/*.pt pin the oracle at tiny custom widths; the GPU parity tests use it as the checker at the reference's
own configs ('32', '64', '128', attention configs).  Here the unmodified reference trainers
(/root/reference/tartangan/trainers/{cnn,iqn}.py: build_models + train_batch) and the oracle start from the same
state dicts, see the same images and the same CPU random stream, and must agree on every loss of two steps and on
every parameter / buffer afterwards (torch CPU fp32 on both sides; measured in the build container: bit-identical
losses and parameters at config '128').  Also pins the oracle's
table of configurations to models/pluggan.py GAN_CONFIGS.  CPU only; skipped where the reference tree is absent.
"""
import argparse
import contextlib
import importlib
import io
import os
import sys
import types
import warnings

import pytest
import torch

from oracle import tartan_oracle as O

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/tartangan'),
                                reason='needs the reference tree (build container only)')


@contextlib.contextmanager
def _reference():
    """The real `tartangan` package, imported with the three shims of SURVEY.md Appendix C, isolated from the
    install_as_tartangan aliases other tests may have left in sys.modules."""
    sys.path.insert(0, '/root/reference')
    added = []
    for name, attrs in (('smart_open', dict(open=open)),
                        ('boto3', dict(resource=lambda *a, **k: None, client=lambda *a, **k: None))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            added.append(name)
    import tqdm._utils
    if not hasattr(tqdm._utils, '_unicode'):
        tqdm._utils._unicode = str
    aside = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'tartangan' or k.startswith('tartangan.')}
    try:
        yield (importlib.import_module('tartangan.models.pluggan'), importlib.import_module('tartangan.trainers.cnn'),
               importlib.import_module('tartangan.trainers.iqn'))
    finally:
        sys.path.remove('/root/reference')
        for k in [k for k in sys.modules if k == 'tartangan' or k.startswith('tartangan.')]:
            del sys.modules[k]
        sys.modules.update(aside)
        for name in added:
            sys.modules.pop(name, None)


def _reference_trainer(kind, cnn, iqn, config, batch, extra=()):
    mod = cnn if kind == 'cnn' else iqn
    cls = mod.CNNTrainer if kind == 'cnn' else mod.IQNTrainer
    p = argparse.ArgumentParser()
    cls.add_args_to_parser(p)
    for cc in cls.get_component_classes(p.parse_known_args(['/unused'])[0]):
        cc.add_args_to_parser(p)
    args = p.parse_args(['/unused', '--batch-size', str(batch), '--config', config] + list(extra))
    args.device = 'cpu'
    t = cls.__new__(cls)
    t.args, t.steps, t.epoch = args, 0, 1
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        t.build_models()
    return t


def _sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def test_config_table_matches_reference():
    with _reference() as (pluggan, _, _):
        ref = dict(pluggan.GAN_CONFIGS)
    for key, spec in O.SPECS.items():
        assert key in ref, key
        c = ref[key]
        assert (c.base_size, c.latent_dims, c.data_dims, tuple(c.blocks), tuple(c.attention)) == tuple(spec), key
        assert c.num_blocks_per_scale == 1, key          # what the oracle's builders assume
    # every configuration BASELINE.json / SURVEY 8d names is in the oracle's table
    for key in ('64', '128', '256', '512', '512thin'):
        assert key in O.SPECS


@pytest.mark.parametrize('kind,config,batch,extra', [
    ('cnn', '64', 2, ()),                       # C1 (cnn '64')
    ('iqn', '64', 2, ()),                       # C2
    ('iqn', '128', 2, ()),                      # C3, the benchmarked configuration
    ('iqn', '32', 3, ('--norm', 'id')),         # the multi-GPU parity worker's configuration, identity norm
    ('cnn', '32', 2, ('--activation', 'elu')),
    ('iqn', '32', 2, ('--g-base', 'tiledz')),
    ('cnn', '256+attention', 2, ()),            # C4: cnn '256' with attention=(3,), gamma set to 0.5
    ('iqn', '512thin+nq64', 2, ()),             # C5: iqn '512thin' (native attention=(3,)), 64 quantiles, gamma 0.5
    ('iqn', '512+attention+nq64', 1, ()),       # C5: iqn '512' with attention=(3,), 64 quantiles
])
def test_oracle_tracks_live_reference(kind, config, batch, extra):
    opts = dict(zip(extra[::2], extra[1::2]))
    config, *variant = config.split('+')
    with _reference() as (pluggan, cnn, iqn):
        spec, nq, steps = O.SPECS[config], O.NUM_QUANTILES, 2
        if 'attention' in variant:
            spec = spec._replace(attention=(3,))
            pluggan.GAN_CONFIGS[config + 'sa'] = pluggan.GAN_CONFIGS[config]._replace(attention=(3,))
            config = config + 'sa'
        t = _reference_trainer(kind, cnn, iqn, config, batch, extra)
        if variant:
            steps = 1                           # (256 / 512 pixel images on host cores)
            with torch.no_grad():               # gamma initialises to 0 (attention.py:19): make the attention matter
                for m in list(t.g.modules()) + list(t.target_g.modules()) + list(t.d.modules()):
                    if hasattr(m, 'gamma'):
                        m.gamma.fill_(0.5)
        if 'nq64' in variant:
            nq = t.d.to_output.iqn.num_quantiles = 64
        size = t.g.max_size
        orc = O.OracleTrainer(kind, spec, _sd(t.g), _sd(t.target_g), _sd(t.d), batch,
                              norm=opts.get('--norm', 'bn'), g_base=opts.get('--g-base', 'mlp'),
                              activation=opts.get('--activation', 'relu'), num_quantiles=nq)
        for step in range(steps):
            imgs = O.tartan_batch(4321 + step, batch, size)
            torch.manual_seed(900 + step)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                ref = t.train_batch(imgs)
            torch.manual_seed(900 + step)
            got = orc.train_batch(imgs)
            for k in ref:
                assert abs(float(got[k]) - float(ref[k])) <= 1e-6 * max(1.0, abs(float(ref[k]))), (step, k, got[k], ref[k])
        for name, mod, state in (('g', t.g, orc.g), ('target_g', t.target_g, orc.target_g), ('d', t.d, orc.d)):
            for k, v in mod.state_dict().items():
                o = state[k].detach()
                if v.dtype.is_floating_point:
                    # Adam(beta1 = 0) moves every element by ~lr * sign(g): rounding-level differences of a gradient
                    # near zero can flip one element by 2 lr; everything else agrees to rounding
                    diff = (o - v).abs()
                    lr = 4e-4 if name == 'd' else 1e-4
                    loose = diff > 1e-5 + 1e-4 * v.abs()
                    assert float(loose.float().mean()) <= 2e-3 and float(diff.max()) <= 4.1 * lr, (name, k, float(diff.max()))
                else:
                    assert torch.equal(o, v), (name, k)


@pytest.mark.parametrize('kind', ['cnn', 'iqn'])
def test_cli_surface_matches_reference(kind):
    """every command-line flag of the reference trainers (trainer.py:271-313 plus the component flags,
    e.g. model_checkpoint.py:110-117 and metrics/fid.py:51-59) exists here with the same default, and a reference
    command line parses to the same values; flags added here are additive (not required)."""
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    ours_cls = CNNTrainer if kind == 'cnn' else IQNTrainer
    ours = argparse.ArgumentParser()
    ours_cls.add_args_to_parser(ours)
    with _reference() as (_, cnn, iqn):
        ref_cls = cnn.CNNTrainer if kind == 'cnn' else iqn.IQNTrainer
        ref = argparse.ArgumentParser()
        ref_cls.add_args_to_parser(ref)
        for cc in ref_cls.get_component_classes(ref.parse_known_args(['/unused', '--fid'])[0]):
            cc.add_args_to_parser(ref)
        ref_actions = [a for a in ref._actions if a.dest != 'help']
        line = ['/data/x.npz', '--batch-size', '32', '--config', '128', '--grad-penalty', '2.5', '--lr-g', '2e-4',
                '--norm', 'id', '--activation', 'selu', '--g-base', 'tiledz', '--checkpoint-freq', '500',
                '--resume-training-step', '1500', '--run-id', 'abc', '--quiet-logs', '--gen-freq', '50']
        ref_ns = vars(ref.parse_args(line))
    ours_by_dest = {a.dest: a for a in ours._actions}
    for a in ref_actions:
        assert a.dest in ours_by_dest, f'missing flag {a.option_strings or a.dest}'
        b = ours_by_dest[a.dest]
        assert set(a.option_strings) <= set(b.option_strings), (a.option_strings, b.option_strings)
        assert a.default == b.default, (a.dest, a.default, b.default)
        assert a.nargs == b.nargs and type(a).__name__ == type(b).__name__, a.dest
    ours_ns = vars(ours.parse_args(line))
    for k, v in ref_ns.items():
        assert ours_ns[k] == v, (k, ours_ns[k], v)
    extra = [a for a in ours._actions if a.dest not in {r.dest for r in ref_actions} and a.dest != 'help']
    assert all(not a.required and a.option_strings for a in extra)        # additive, optional flags only


def _signature(f):
    import functools
    import inspect

    def default(d):
        if d is inspect.Parameter.empty:
            return '<required>'
        if isinstance(d, functools.partial):
            return ('partial', d.func.__name__, d.args, tuple(sorted(d.keywords.items())))
        return getattr(d, '__name__', None) or repr(d)          # classes / functions compare by name
    return [(n, p.kind.name, default(p.default)) for n, p in inspect.signature(f).parameters.items()]


def test_module_constructor_signatures_match_reference():
    """SURVEY 8b, "same constructor signatures and defaults": every class / function of the Python-level contract
    (models.pluggan, models.blocks, models.iqn, models.losses) takes the reference's parameters, in its order, with
    its defaults (factory defaults compared by class name: the mirror's BatchNorm2d / Conv2d are kernel-backed
    subclasses of the torch.nn classes the reference names), and GAN_CONFIGS is the same table."""
    import tartangan_b200.models.blocks as ob
    import tartangan_b200.models.iqn as oi
    import tartangan_b200.models.losses as ol
    import tartangan_b200.models.pluggan as op
    names = {
        'pluggan': ['GANConfig', 'BlockModel', 'Generator', 'Discriminator', 'IQNDiscriminator'],
        'blocks': ['ResidualGeneratorBlock', 'ResidualDiscriminatorBlock', 'GeneratorBlock', 'DiscriminatorBlock',
                   'GeneratorInputMLP', 'TiledZGeneratorInput', 'GeneratorOutput', 'DiscriminatorInput',
                   'DiscriminatorOutput', 'IQNDiscriminatorOutput', 'SelfAttention2d'],
        'iqn': ['IQN', 'CosineQuantileEmbedding', 'iqn_loss'],
        'losses': ['gradient_penalty'],
    }
    ours = dict(pluggan=op, blocks=ob, iqn=oi, losses=ol)
    with _reference() as (rp, _, _):
        ref = dict(pluggan=rp, blocks=importlib.import_module('tartangan.models.blocks'),
                   iqn=importlib.import_module('tartangan.models.iqn'),
                   losses=importlib.import_module('tartangan.models.losses'))
        checked = 0
        for mod, members in names.items():
            for n in members:
                r, o = getattr(ref[mod], n), getattr(ours[mod], n)
                assert _signature(r) == _signature(o), (mod, n, _signature(r), _signature(o))
                if isinstance(r, type) and 'forward' in vars(r):
                    assert _signature(r.forward) == _signature(o.forward), (mod, n, 'forward')
                checked += 1
        assert checked == 20
        assert set(rp.GAN_CONFIGS) <= set(op.GAN_CONFIGS)
        for k, c in rp.GAN_CONFIGS.items():
            assert tuple(c) == tuple(op.GAN_CONFIGS[k]), k


@pytest.mark.parametrize('kind', ['cnn', 'iqn'])
def test_trainer_method_surface_matches_reference(kind):
    """the trainer methods a caller of the reference uses exist here and accept the reference's arguments
    (parameters added here are optional and come last)."""
    from tartangan_b200.trainers import cnn as oc, iqn as oi
    ours = oc.CNNTrainer if kind == 'cnn' else oi.IQNTrainer
    contract = ['create_from_cli', 'add_args_to_parser', 'get_component_classes', 'train', 'build_models', 'train_batch',
                'prepare_dataset', 'sample_z', 'sample_g', 'make_adversarial_batch', 'make_generator_batch',
                'update_target_generator', 'init_params_selu', 'get_state', 'set_state']
    with _reference() as (_, cnn, iqn):
        ref = cnn.CNNTrainer if kind == 'cnn' else iqn.IQNTrainer
        ref_mod = cnn if kind == 'cnn' else iqn
        assert callable(ref_mod.main) and callable((oc if kind == 'cnn' else oi).main)
        for n in contract:
            a, b = _signature(getattr(ref, n)), _signature(getattr(ours, n))
            assert b[:len(a)] == a, (n, a, b)
            assert all(d != '<required>' for _, _, d in b[len(a):]), (n, b)
