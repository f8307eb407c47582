"""Remaining rows of SURVEY.md §8: spectral-norm power iteration (a18), BASELINE config 1 as a 20-step
loss-curve parity run, the attention configs (C4/C5 shapes at reduced batch), checkpoint layout (f-1)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def test_spectral_norm_matches_torch():
    import tartangan_b200 as tb
    from tartangan_b200 import ops
    from tartangan_b200.models.layers import SpectralNormConv2d
    tb.set_precision('fp32')
    torch.manual_seed(0)
    ref = torch.nn.utils.spectral_norm(torch.nn.Conv2d(16, 24, 3, padding=1))
    m = SpectralNormConv2d(16, 24, 3, padding=1)
    with torch.no_grad():
        m.weight_orig.copy_(ref.weight_orig); m.bias.copy_(ref.bias)
        m.weight_u.copy_(ref.weight_u); m.weight_v.copy_(ref.weight_v)
    m = m.cuda()
    x = torch.randn(2, 16, 8, 8)
    for it in range(3):                      # u/v evolve identically over several training forwards
        y = ref(x)
        yd = m(x.cuda())
        assert _rel(yd, y) < 2e-4, it
        assert _rel(m.weight_u, ref.weight_u) < 1e-4 and _rel(m.weight_v, ref.weight_v) < 1e-4
    g = torch.randn_like(y)
    gw, = torch.autograd.grad(y, ref.weight_orig, g)
    gwd, = torch.autograd.grad(yd, m.weight_orig, ops.to_internal(g.cuda()))
    assert _rel(gwd, gw) < 5e-4
    m.eval(); ref.eval()
    assert _rel(m(x.cuda()), ref(x)) < 2e-4
    tb.set_precision('bf16')


@pytest.mark.parametrize('kind,precision', [('cnn', 'fp32'), ('iqn', 'fp32'), ('iqn', 'bf16')])
def test_spectral_norm_training_step_with_r1(kind, precision):
    """--spectral-norm gd: spectral-normalised convs in G and D (conv_factory seam) through two full training steps
    INCLUDING the R1 penalty (the normalised weight is a non-leaf tensor inside the double-backward graph) against
    the oracle, whose spectral-norm conv is pinned to torch.nn.utils.spectral_norm."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GANConfig
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    cfg = GANConfig(base_size=4, latent_dims=32, data_dims=3, blocks=(32, 16, 16), num_blocks_per_scale=1, attention=())
    torch.manual_seed(0)
    t = make_trainer(CNNTrainer if kind == 'cnn' else IQNTrainer, gan_config=cfg, batch_size=8, precision=precision,
                     spectral_norm='gd')
    keys = list(t.d.state_dict())
    assert any(k.endswith('weight_orig') for k in keys) and any(k.endswith('weight_u') for k in keys)
    assert any(k.endswith('weight_orig') for k in t.g.state_dict())
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = O.OracleTrainer(kind, O.Spec(4, 32, 3, (32, 16, 16), ()), cpu(t.g), cpu(t.target_g), cpu(t.d), 8)
    tol = (3e-3, 1e-2) if precision == 'fp32' else (6e-2, 1e-1)
    for s in range(2):
        imgs = O.tartan_batch(60 + s, 8, 32)
        torch.manual_seed(400 + s)
        ref = orc.train_batch(imgs)
        torch.manual_seed(400 + s)
        got = t.train_batch(imgs)
        assert ref['gp'] > 0
        for k in ref:
            assert abs(got[k] - ref[k]) <= tol[min(s, 1)] * max(1.0, abs(ref[k])), (s, k, got[k], ref[k])
        if s == 0 and precision == 'fp32':
            for net, mod in (('d', t.d), ('g', t.g)):
                for k, p in mod.named_parameters():
                    r = orc.last_grads[net].get(k)
                    if r is None or p.grad is None or float(r.abs().max()) < 1e-5:
                        continue
                    assert _rel(p.grad, r) < 2e-2 or (k.endswith('.bias') and float(r.abs().max()) < 1e-4), (net, k)
    for k, v in orc.d.items():       # the power-iteration vectors walked the same path (two D forwards + one per G step)
        if k.endswith('weight_u') or k.endswith('weight_v'):
            assert _rel(t.d.state_dict()[k], v) < (5e-3 if precision == 'fp32' else 5e-2), k


def _sync_from_oracle(t, orc):
    """Copy the oracle's parameters, buffers and Adam moments into the CUDA trainer."""
    with torch.no_grad():
        for mod, sd in ((t.g, orc.g), (t.target_g, orc.target_g), (t.d, orc.d)):
            mod.load_state_dict({k: v.detach() for k, v in sd.items()})
        for opt, oopt, names, sd, mod in ((t.optimizer_g, orc.opt_g, orc.g_params, orc.g, t.g),
                                          (t.optimizer_d, orc.opt_d, orc.d_params, orc.d, t.d)):
            opt._ensure_flat()
            mine = dict(mod.named_parameters())
            for k in names:
                st = oopt.state.get(sd[k])
                if st:
                    opt.state[mine[k]]['exp_avg'].copy_(st['exp_avg'])
                    opt.state[mine[k]]['exp_avg_sq'].copy_(st['exp_avg_sq'])
                    opt._step.fill_(float(st['step']))


def test_config1_cnn64_loss_curve_20_steps():
    """BASELINE.json config 1: trainers.cnn SA-GAN 64x64, batch 16, 20 steps, CUDA (fp32 mode) vs the CPU oracle.

    GAN training with Adam(beta1=0) is chaotic: a last-bit difference in a gradient flips the sign of a
    +-lr parameter step and the two trajectories decorrelate after ~8 steps (the reference itself does this
    between two thread counts).  So parity is asserted two ways: (1) teacher-forced — before every step the
    CUDA trainer adopts the oracle's full state (parameters, BN buffers, Adam moments) and the step's three
    losses must then agree to 2e-3 for all 20 steps; (2) free-running — the first three steps agree to 3 %
    and the curves stay in the same band afterwards."""
    from oracle import tartan_oracle as O
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    # (1) teacher-forced
    torch.manual_seed(0)
    t = make_trainer(CNNTrainer, config='64', batch_size=16, precision='fp32')
    orc = O.OracleTrainer('cnn', O.SPECS['64'], cpu(t.g), cpu(t.target_g), cpu(t.d), 16)
    ref_curve = []
    for s in range(20):
        _sync_from_oracle(t, orc)
        imgs = O.tartan_batch(1234 + s, 16, 64)
        torch.manual_seed(2000 + s)
        ref = orc.train_batch(imgs)
        torch.manual_seed(2000 + s)
        got = t.train_batch(imgs)
        ref_curve.append(ref)
        for k in ref:
            # g_loss is evaluated after the D update of the same step (Adam sign noise): slightly looser
            assert abs(got[k] - ref[k]) <= (4e-3 if k == 'g_loss' else 2e-3) * max(1.0, abs(ref[k])), (s, k, got[k], ref[k])
    # (2) free-running
    torch.manual_seed(0)
    t = make_trainer(CNNTrainer, config='64', batch_size=16, precision='fp32')
    free = []
    for s in range(20):
        torch.manual_seed(2000 + s)
        free.append(t.train_batch(O.tartan_batch(1234 + s, 16, 64)))
    for s in range(3):
        for k in ('d_loss', 'gp', 'g_loss'):
            assert abs(free[s][k] - ref_curve[s][k]) <= 3e-2 * max(1.0, abs(ref_curve[s][k])), (s, k)
    for k in ('d_loss', 'gp'):
        a = sum(m[k] for m in free[10:]) / 10
        b = sum(m[k] for m in ref_curve[10:]) / 10
        assert abs(a - b) <= 0.35 * b, (k, a, b)


@pytest.mark.parametrize('kind,config', [('cnn', '256sa'), ('iqn', '512thin')])
def test_attention_configs_run_and_track_oracle(kind, config):
    """C4 / C5 architectures (self-attention at 64x64 feature maps in G, 32x32 in D) at batch 2, scaled to 1/4
    width so the CPU oracle finishes in seconds; fp32 mode, one step, gamma set to 0.5 so attention matters."""
    from oracle import tartan_oracle as O
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    scale = 0.25 if config == '256sa' else 0.5
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer if kind == 'iqn' else CNNTrainer, config=config, batch_size=2, precision='fp32',
                     model_scale=scale, num_quantiles=(64 if kind == 'iqn' else 8))
    with torch.no_grad():
        for m in list(t.g.modules()) + list(t.d.modules()) + list(t.target_g.modules()):
            if hasattr(m, 'gamma'):
                m.gamma.fill_(0.5)
    cfg = t.gan_config
    spec = O.Spec(4, cfg.latent_dims, 3, tuple(cfg.blocks), tuple(cfg.attention))
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = O.OracleTrainer(kind, spec, cpu(t.g), cpu(t.target_g), cpu(t.d), 2, num_quantiles=(64 if kind == 'iqn' else 8))
    imgs = O.tartan_batch(5, 2, t.g.max_size)
    torch.manual_seed(77)
    ref = orc.train_batch(imgs)
    torch.manual_seed(77)
    got = t.train_batch(imgs)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 5e-3 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


@pytest.mark.parametrize('kind,config,kw', [('cnn', '256sa', {}), ('iqn', '512thin', {'num_quantiles': 64}),
                                            ('iqn', '512', {'num_quantiles': 64, 'attention': '3'})])
def test_attention_configs_full_width_bf16(kind, config, kw):
    """BASELINE configs 4 and 5 at FULL width (256 / 8-channel layers, fused tcgen05 attention at 64x64 and 32x32),
    bf16 mode, batch 2: first-step losses within 5 % of the CPU oracle (fp32) for '256sa' / '512thin' (the
    bf16 first-step bar of DESIGN.md section 6); for every config the fused and the un-fused attention paths agree
    and a CUDA-graph replay of the step reproduces the eager losses' magnitude."""
    from oracle import tartan_oracle as O
    from tartangan_b200 import ops
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    cls = IQNTrainer if kind == 'iqn' else CNNTrainer

    def build(**extra):
        torch.manual_seed(0)
        t = make_trainer(cls, config=config, batch_size=2, precision='bf16', **kw, **extra)
        with torch.no_grad():
            for m in list(t.g.modules()) + list(t.d.modules()) + list(t.target_g.modules()):
                if hasattr(m, 'gamma'):
                    m.gamma.fill_(0.5)
        return t
    res = {}
    for fused in (True, False):
        ops.state.fused_attention = fused
        try:
            t = build()
            imgs = O.tartan_batch(5, 2, t.g.max_size)
            if fused and config != '512':
                cfg = t.gan_config
                spec = O.Spec(4, cfg.latent_dims, 3, tuple(cfg.blocks), tuple(cfg.attention))
                cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
                orc = O.OracleTrainer(kind, spec, cpu(t.g), cpu(t.target_g), cpu(t.d), 2, num_quantiles=kw.get('num_quantiles', 8))
                torch.manual_seed(77)
                res['oracle'] = orc.train_batch(imgs)
            torch.manual_seed(77)
            res[fused] = t.train_batch(imgs)
        finally:
            ops.state.fused_attention = True
    for k in res[True]:
        assert res[True][k] == res[True][k] and abs(res[True][k]) < 1e4, (k, res[True])
        # (g_loss is computed after D's Adam(beta1=0) step: see the note on the graph comparison below; seen 3.06 % once)
        assert abs(res[True][k] - res[False][k]) <= (6e-2 if k == 'g_loss' else 3e-2) * max(1.0, abs(res[False][k])), (k, res[True], res[False])
        if 'oracle' in res:
            assert abs(res[True][k] - res['oracle'][k]) <= 5e-2 * max(1.0, abs(res['oracle'][k])), (k, res[True], res['oracle'])
    tg = build(cuda_graph=True)
    torch.manual_seed(77)
    got = tg.train_batch(O.tartan_batch(5, 2, tg.g.max_size))
    for k in got:
        # d_loss / gp are computed BEFORE the first optimiser update: graph and eager agree tightly.  g_loss is computed
        # after D's Adam(beta1=0) step, whose +-lr moves depend on the sign of near-zero gradients; the bf16 tensor-core
        # wgrad adds its per-CTA partial sums with fp32 reductions in arrival order, so at batch 2 two runs of the SAME
        # code differ by up to ~4 % there (seen: 3.6 % on one box, < 3 % on the others)
        tol = 6e-2 if k == 'g_loss' else 3e-2
        assert abs(got[k] - res[True][k]) <= tol * max(1.0, abs(res[True][k])), (k, got, res[True])


def test_checkpoint_layout_round_trip(tmp_path):
    """components/model_checkpoint.py layout: {output}/{run_id}/checkpoints/{steps}/{g,g_target,d,opt_d,opt_g}.pt + trainer.json"""
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    from oracle import tartan_oracle as O
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer, config='32', batch_size=4, precision='fp32', output=str(tmp_path), model_scale=0.25)
    t.run_id = 'run'
    imgs = O.tartan_batch(1, 4, 32)
    torch.manual_seed(1)
    t.train_batch(imgs)
    t.steps = 7
    t.save_checkpoint()
    root = tmp_path / 'run' / 'checkpoints' / '7'
    assert sorted(os.listdir(root)) == ['d.pt', 'g.pt', 'g_target.pt', 'opt_d.pt', 'opt_g.pt', 'trainer.json']
    assert json.load(open(root / 'trainer.json')) == {'epoch': 1, 'steps': 7}
    import tartangan_b200
    tartangan_b200.install_as_tartangan()                 # the files name the reference's classes (whole objects)
    d_obj = torch.load(root / 'd.pt', weights_only=False)
    assert type(d_obj).__name__ == 'IQNDiscriminator' and not isinstance(d_obj, dict)
    sd = d_obj.state_dict()
    assert list(sd.keys())[0].startswith('to_output.activation.0')       # head first (pluggan.py:126-127)
    opt = torch.load(root / 'opt_d.pt', weights_only=False)
    assert type(opt) is torch.optim.Adam                                  # what the reference pickles (cnn.py:84-85)
    assert set(opt.state_dict()['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    torch.manual_seed(0)
    t2 = make_trainer(IQNTrainer, config='32', batch_size=4, precision='fp32', output=str(tmp_path), model_scale=0.25)
    t2.run_id, t2.steps = 'run', 7
    t2.load_checkpoint()
    for k, v in t.g.state_dict().items():
        assert torch.equal(t2.g.state_dict()[k], v), k
    torch.manual_seed(2)
    a = t.train_batch(imgs)
    torch.manual_seed(2)
    b = t2.train_batch(imgs)
    for k in a:
        assert abs(a[k] - b[k]) <= 5e-3 * max(1.0, abs(a[k])), (k, a[k], b[k])


def test_load_checkpoint_written_by_the_reference(tmp_path):
    """f-1, reference -> here: tests/golden/ref_checkpoint was saved by the UNMODIFIED reference
    (ModelCheckpointComponent.save_checkpoint: whole pickled objects, tools/make_ref_checkpoint.py).  The trainer resumes
    from it (parameters, BN buffers, Adam moments, step counter) and the next step matches the oracle resumed from the
    same state."""
    import shutil
    from conftest import GOLDEN_DIR
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GANConfig
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    exp = torch.load(os.path.join(GOLDEN_DIR, 'ref_checkpoint', 'expected_state.pt'), weights_only=False)
    shutil.copytree(os.path.join(GOLDEN_DIR, 'ref_checkpoint', 'checkpoints'), tmp_path / 'run' / 'checkpoints')
    cfg = GANConfig(base_size=4, latent_dims=exp['latent'], data_dims=3, blocks=tuple(exp['blocks']),
                    num_blocks_per_scale=1, attention=())
    torch.manual_seed(5)
    t = make_trainer(IQNTrainer, gan_config=cfg, batch_size=exp['batch'], precision='fp32', output=str(tmp_path))
    t.run_id, t.steps = 'run', 1
    t.load_checkpoint()
    assert t.steps == 1 and t.epoch == 1
    for net in ('g', 'target_g', 'd'):
        sd = getattr(t, net).state_dict()
        assert list(sd) == list(exp[net])
        for k, v in exp[net].items():
            assert torch.equal(sd[k].cpu(), v), (net, k)
    for name, opt in (('opt_d', t.optimizer_d), ('opt_g', t.optimizer_g)):
        got = opt.state_dict()['state']
        for i, st in exp[name]['state'].items():
            assert torch.equal(got[i]['exp_avg_sq'].cpu(), st['exp_avg_sq']) and float(got[i]['step']) == float(st['step'])
    # the resumed trainer continues like the reference would: one more step vs the oracle carrying the same state
    orc = O.OracleTrainer('iqn', O.Spec(4, exp['latent'], 3, tuple(exp['blocks']), ()), exp['g'], exp['target_g'],
                          exp['d'], exp['batch'])
    orc.opt_d.load_state_dict(exp['opt_d'])
    orc.opt_g.load_state_dict(exp['opt_g'])
    imgs = O.tartan_batch(4321, exp['batch'], 32)
    torch.manual_seed(9)
    ref = orc.train_batch(imgs)
    torch.manual_seed(9)
    got = t.train_batch(imgs)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 3e-3 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    for k, v in orc.g.items():
        if v.is_floating_point() and not k.endswith('.bias') and 'running' not in k:
            d = (t.g.state_dict()[k].cpu() - v).abs()
            assert float(d.max()) <= 3e-4 + 1e-4 * float(v.abs().max()), k      # at most one sign-flipped Adam step (lr 1e-4, bias-corrected)


def test_target_g_sampling_after_graphed_steps():
    """f-2: sample_g(target_g=True) (components/image_sampler.py:24-45 draws its grids from target_g) after CUDA-graph
    steps uses the EMA weights the graphs wrote (not a stale packed copy) and equals the oracle generator run on the
    same state dict; the EMA itself tracks the oracle's target_g."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    cfg = GAN_CONFIGS['32']
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer, config='32', batch_size=8, precision='bf16', cuda_graph=True)
    spec = O.Spec(4, cfg.latent_dims, 3, tuple(cfg.blocks), ())
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = O.OracleTrainer('iqn', spec, cpu(t.g), cpu(t.target_g), cpu(t.d), 8)
    torch.manual_seed(21)
    before = t.sample_g(8, target_g=True).float().cpu()
    for s in range(3):
        imgs = O.tartan_batch(70 + s, 8, 32)
        torch.manual_seed(500 + s)
        orc.train_batch(imgs)
        torch.manual_seed(500 + s)
        t.train_batch(imgs)
    t.target_g.train()
    torch.manual_seed(21)
    after = t.sample_g(8, target_g=True).float().cpu()
    assert after.shape == (8, 3, 32, 32) and float(after.abs().max()) <= 1.0
    assert float((after - before).abs().max()) > 0            # the EMA moved the target generator
    sd = cpu(t.target_g)
    torch.manual_seed(21)
    with torch.no_grad():
        want = O.generator({k: v.clone() for k, v in sd.items()}, spec, torch.randn(8, cfg.latent_dims))
    assert float((after - want).norm() / want.norm()) < 3e-2  # bf16 kernels vs fp32 oracle on the SAME (current) weights
    for k, v in orc.target_g.items():                         # and the EMA state follows the oracle's
        if v.is_floating_point() and 'running' not in k:
            assert float((sd[k] - v).abs().max()) <= 5e-6 + 1e-3 * 4e-4 * 3, k
