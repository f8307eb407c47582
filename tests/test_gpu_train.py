"""GPU parity of the whole training step: CUDA path (through the C ABI) vs the golden vectors
produced by the unmodified reference, and vs the CPU oracle at reference channel widths."""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu


def _trainer(kind, gan_config, batch, precision, norm='bn', **kw):
    from tartangan_b200.trainers.cnn import CNNTrainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.gan import make_trainer
    cls = CNNTrainer if kind == 'cnn' else IQNTrainer
    return make_trainer(cls, gan_config=gan_config, batch_size=batch, precision=precision, norm=norm, **kw)


def _cfg(g):
    from tartangan_b200.models.pluggan import GANConfig
    return GANConfig(base_size=4, latent_dims=g['latent'], data_dims=3, blocks=tuple(g['blocks']),
                     num_blocks_per_scale=1, attention=tuple(g['attention']))


def _close(a, b, tol):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6)) <= tol


def _noise_grad(g, net, name):
    """Conv biases that feed a train-mode BatchNorm have an analytically zero gradient; the
    reference holds ~1e-6 rounding noise there."""
    ref = g['grads0'][net].get(name)
    if ref is None:
        return False
    scale = max(float(v.abs().max()) for v in g['grads0'][net].values())
    return float(ref.abs().max()) < max(2e-3 * scale, 5e-5)


def _params_close(a, b, lr, rare=0.01):
    """Adam with beta1=0 moves every element by ~lr*sign(g) per step, so an element whose gradient
    is rounding noise may differ by 2*lr; require that to be rare and everything else tight.
    `rare`: since round 2 the fp32 wgrad kernel adds its split-K partial sums in a fixed order (bitwise repeatable,
    tests/test_gpu_ops.py::test_fp32_wgrad_is_bitwise_repeatable), so the set of flipped near-zero gradients no longer
    varies from run to run (round 1 allowed 5-8 % for the float-atomic version); 1 % of a tensor is the bound."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    diff = (a - b).abs()
    tight = diff <= 1e-4 * float(b.abs().max().clamp_min(1e-6)) + 0.05 * lr
    ok = float((~tight).float().mean()) <= rare and float(diff.max()) <= 2.5 * lr + 1e-4 * float(b.abs().max())
    if not ok:
        print(f'_params_close: {float((~tight).float().mean()) * 100:.2f} % of {a.numel()} elements loose (allowed '
              f'{rare * 100:.1f} %), max diff {float(diff.max()):.3e} (allowed {2.5 * lr + 1e-4 * float(b.abs().max()):.3e})')
    return ok


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_golden_fp32(case):
    g = load_golden(case)
    torch.manual_seed(0)
    t = _trainer(g['kind'], _cfg(g), g['batch'], 'fp32', norm=g['norm'], activation=g.get('activation', 'relu'),
                 g_base=g.get('g_base', 'mlp'))
    # same seed, same construction order => the reference's initial parameters (incl. init_params_selu, cnn.py:96-105)
    for net, mod in (('g', t.g), ('d', t.d)):
        for k, ref in g['init'][net].items():
            if ref.is_floating_point() and 'gamma' not in k:       # (the attention goldens overwrite gamma with 0.5)
                assert torch.equal(mod.state_dict()[k].cpu(), ref), (net, k)
    t.g.load_state_dict(g['init']['g'])
    t.target_g.load_state_dict(g['init']['target_g'])
    t.d.load_state_dict(g['init']['d'])
    # module-level forward probes
    pr = g['probe']
    gsd = {k: v.clone() for k, v in t.g.state_dict().items()}
    dsd = {k: v.clone() for k, v in t.d.state_dict().items()}
    with torch.no_grad():
        assert _close(t.g(pr['z'].cuda()), pr['g_out'], 2e-4)
        if g['kind'] == 'iqn':
            torch.manual_seed(pr['tau_seed'])
            p, loss = t.d(pr['x'].cuda(), targets=torch.ones(g['batch'], 1, device='cuda'))
            assert _close(p, pr['d_out'], 2e-4) and _close(loss, pr['d_loss'], 2e-4)
        else:
            assert _close(t.d(pr['x'].cuda()), pr['d_out'], 2e-4)
    t.g.load_state_dict(gsd)
    t.d.load_state_dict(dsd)
    for s in range(g['steps']):
        torch.manual_seed(g['seeds'][s])
        m = t.train_batch(g['imgs'][s])
        for k, v in g['metrics'][s].items():
            assert abs(m[k] - v) <= (2e-3 if s == 0 else 6e-3) * max(1.0, abs(v)), (s, k, m[k], v)
        if s == 0:
            for net, mod in (('d', t.d), ('g', t.g)):
                grads = dict(mod.named_parameters())
                for k, ref in g['grads0'][net].items():
                    if _noise_grad(g, net, k):
                        continue            # analytically-zero bias grads: reference holds fp noise
                    assert _close(grads[k].grad, ref, 5e-3), (net, k)
    for net, mod in (('g', t.g), ('target_g', t.target_g), ('d', t.d)):
        sd = mod.state_dict()
        for k, ref in g['final'][net].items():
            if ref.is_floating_point():
                if net != 'target_g' and _noise_grad(g, net, k):
                    continue        # Adam(beta1=0) turns fp noise on zero-gradient biases into +-lr steps
                if 'running_' in k:
                    assert _close(sd[k], ref, 1e-2), (net, k)
                else:
                    assert _params_close(sd[k], ref, lr=4e-4 * g['steps']), (net, k)
            else:
                assert torch.equal(sd[k].cpu(), ref), (net, k)


@pytest.mark.parametrize('kind', ['cnn', 'iqn'])
def test_vs_oracle_reference_widths(kind):
    """config '32' at the reference's real channel widths (128, 64, 32), batch 8, two steps, fp32."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    cfg = GAN_CONFIGS['32']
    torch.manual_seed(0)
    t = _trainer(kind, cfg, 8, 'fp32')
    spec = O.Spec(4, cfg.latent_dims, 3, tuple(cfg.blocks), ())
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = O.OracleTrainer(kind, spec, cpu(t.g), cpu(t.target_g), cpu(t.d), 8)
    for s in range(2):
        imgs = O.tartan_batch(50 + s, 8, 32)
        torch.manual_seed(300 + s)
        ref = orc.train_batch(imgs)
        torch.manual_seed(300 + s)
        got = t.train_batch(imgs)
        for k in ref:
            assert abs(got[k] - ref[k]) <= (3e-3 if s == 0 else 1e-2) * max(1.0, abs(ref[k])), (s, k, got[k], ref[k])
    for k, v in orc.d.items():
        if v.is_floating_point() and not k.endswith('.bias') and 'running' not in k:
            # near-zero gradients may flip their Adam sign against the CPU summation order; with the fixed-order
            # wgrad reduction (round 2) the flipped set is the same in every run
            # (measured: every tensor <= 1 % with the summation order of this commit, one tensor at ~1-4 % with another
            # equally valid order of the same sums; 4 % is the bound, against 8 % while the kernel was not repeatable)
            assert _params_close(t.d.state_dict()[k], v, lr=8e-4, rare=0.04), k


@pytest.mark.parametrize('kind', ['cnn', 'iqn'])
def test_bf16_step_tracks_oracle(kind):
    """bf16 activations: losses of the first step within 5 % of the fp32 oracle."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    cfg = GAN_CONFIGS['32']
    torch.manual_seed(0)
    t = _trainer(kind, cfg, 8, 'bf16')
    spec = O.Spec(4, cfg.latent_dims, 3, tuple(cfg.blocks), ())
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = O.OracleTrainer(kind, spec, cpu(t.g), cpu(t.target_g), cpu(t.d), 8)
    imgs = O.tartan_batch(50, 8, 32)
    torch.manual_seed(300)
    ref = orc.train_batch(imgs)
    torch.manual_seed(300)
    got = t.train_batch(imgs)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 0.08 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


@pytest.mark.parametrize('kind', ['cnn', 'iqn'])
def test_cuda_graph_step_matches_eager(kind):
    """The graph-captured step consumes the same CPU random stream and produces the same losses and
    parameters as the eager step (fp32 mode)."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    cfg = GAN_CONFIGS['32']
    res = []
    for graphed in (False, True):
        torch.manual_seed(0)
        t = _trainer(kind, cfg, 8, 'fp32', cuda_graph=graphed)
        ms = []
        for s in range(3):
            torch.manual_seed(500 + s)
            ms.append(t.train_batch(O.tartan_batch(60 + s, 8, 32)))
        res.append((ms, {k: v.detach().clone() for k, v in t.d.state_dict().items()}))
    # step 0 must agree tightly; later steps only loosely: wgrad sums with fp32 atomics, and Adam with
    # beta1 = 0 turns last-bit gradient noise into +-lr parameter steps (same spread between two eager runs)
    for s, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        tol = 3e-3 if s == 0 else 5e-2
        for k in a:
            assert abs(a[k] - b[k]) <= tol * max(1.0, abs(a[k])), (s, k, a[k], b[k])
    for k, v in res[0][1].items():
        if v.is_floating_point() and 'running' not in k and not k.endswith('.bias'):
            drift = float((res[1][1][k].float() - v.float()).abs().mean())
            assert drift <= 0.25 * 4e-4 * 3, (k, drift)      # far below the 3*lr an element can move
