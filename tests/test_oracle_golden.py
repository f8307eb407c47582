"""Pins the CPU oracle (oracle/tartan_oracle.py) to vectors produced by the
unmodified reference (tools/make_golden.py).  CPU only."""
import torch

from oracle import tartan_oracle as O


def _spec(g):
    return O.Spec(4, g['latent'], 3, tuple(g['blocks']), tuple(g['attention']))


def _close(a, b, tol=2e-5):
    a, b = a.float(), b.float()
    denom = b.abs().max().clamp_min(1e-6)
    return float((a - b).abs().max() / denom) <= tol


def test_forward_probe(golden):
    g = golden
    spec, pr = _spec(g), g['probe']
    gsd = {k: v.clone() for k, v in g['init']['g'].items()}
    dsd = {k: v.clone() for k, v in g['init']['d'].items()}
    O.ACTIVATION = g.get('activation', 'relu')
    try:
        _probe(g, spec, pr, gsd, dsd)
    finally:
        O.ACTIVATION = 'relu'


def _probe(g, spec, pr, gsd, dsd):
    with torch.no_grad():
        img = O.generator(gsd, spec, pr['z'], g['norm'], g.get('g_base', 'mlp'))
        assert _close(img, pr['g_out'])
        if g['kind'] == 'iqn':
            torch.manual_seed(pr['tau_seed'])
            p, loss = O.iqn_discriminator(dsd, spec, pr['x'], torch.ones(g['batch'], 1), g['norm'])
            assert _close(p, pr['d_out']) and _close(loss, pr['d_loss'])
        else:
            assert _close(O.discriminator(dsd, spec, pr['x'], g['norm']), pr['d_out'])


def test_train_steps(golden):
    g = golden
    tr = O.OracleTrainer(g['kind'], _spec(g), g['init']['g'], g['init']['target_g'],
                         g['init']['d'], g['batch'], norm=g['norm'], g_base=g.get('g_base', 'mlp'),
                         activation=g.get('activation', 'relu'))
    for s in range(g['steps']):
        torch.manual_seed(g['seeds'][s])
        m = tr.train_batch(g['imgs'][s])
        for k, v in g['metrics'][s].items():
            assert abs(m[k] - v) <= 2e-5 * max(1.0, abs(v)), (s, k, m[k], v)
        if s == 0:
            for net in ('d', 'g'):
                for k, ref in g['grads0'][net].items():
                    assert _close(tr.last_grads[net][k], ref, 1e-4), (net, k)
    for net, sd in (('g', tr.g), ('target_g', tr.target_g), ('d', tr.d)):
        for k, ref in g['final'][net].items():
            if ref.is_floating_point():
                assert _close(sd[k], ref, 1e-4), (net, k)
            else:
                assert torch.equal(sd[k], ref), (net, k)


def test_tartan_batch_is_deterministic():
    a, b = O.tartan_batch(5, 3, 32), O.tartan_batch(5, 3, 32)
    assert torch.equal(a, b) and a.shape == (3, 3, 32, 32)
    assert float(a.min()) >= -1 and float(a.max()) <= 1


def test_spectral_norm_restatement_matches_torch():
    """The oracle's spectral-norm conv against torch.nn.utils.spectral_norm itself (forward, u / v updates and the
    weight_orig gradient over three training forwards)."""
    torch.manual_seed(4)
    ref = torch.nn.utils.spectral_norm(torch.nn.Conv2d(6, 10, 3, padding=1))
    sd = {'c.' + k: v.detach().clone() for k, v in ref.state_dict().items()}
    sd['c.weight_orig'].requires_grad_()
    for i in range(3):
        x = torch.randn(2, 6, 8, 8)
        y_ref = ref(x)
        y = O._conv(sd, 'c', x, 1)
        assert _close(y, y_ref, 1e-5)
        g = torch.randn_like(y)
        gw_ref, = torch.autograd.grad(y_ref, ref.weight_orig, g)
        gw, = torch.autograd.grad(y, sd['c.weight_orig'], g)
        assert _close(gw, gw_ref, 1e-5)
        assert _close(sd['c.weight_u'], ref.weight_u, 1e-5) and _close(sd['c.weight_v'], ref.weight_v, 1e-5)
