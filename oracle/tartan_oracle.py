"""CPU oracle for the tartangan GAN training step.  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict driven) fp32 restatement, on CPU torch,
of the arithmetic of the reference hot path.  It exists so that the CUDA path
in ``tartangan_b200`` can be checked on a box where ``/root/reference`` does
not exist.  Only ``tests/``, ``__graft_entry__.smoke()``, the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and the
measurement tools under ``tools/`` (golden-vector generation; the same-box
stock-PyTorch bar ``tools/bench_eager_gpu.py``, which runs THIS composition on
``cuda`` through cuDNN / ATen) may import it.  The product package never does.

Parity status: PINNED against outputs of the reference itself, run in the build
container (``tools/make_golden.py`` imports ``/root/reference`` read-only and
writes ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them).
At the reference's NAMED configurations (cnn '64', iqn '64', iqn '128', the
attention configs) ``tests/test_oracle_vs_reference_live.py`` runs the
unmodified reference trainers beside this file in the build container: same
state, images and random stream -> bit-identical losses and parameters.
The reference's own test-suite pins nothing on this path, so
those two are the only pins there are.

Reference locations restated here (all under /root/reference/tartangan):
  models/pluggan.py:18-28     GANConfig / scale_model        -> Spec, scaled()
  models/pluggan.py:63-84     Generator.build                -> generator()
  models/pluggan.py:93-110    Discriminator.build            -> discriminator()
  models/pluggan.py:117-132   IQNDiscriminator.build/forward -> iqn_discriminator()
  models/blocks/generator.py:32-62, 65-80, 115-129           -> _g_block, generator()
  models/blocks/discriminator.py:11-22, 49-95, 126-178       -> _d_block, heads
  models/blocks/attention.py:21-35                           -> _attention()
  models/iqn.py:27-46, 76-108, 111-130                       -> _iqn_head(), quantile_huber()
  models/losses.py:17-30                                     -> r1_penalty()
  trainers/cnn.py:107-165, trainers/iqn.py:104-156           -> OracleTrainer.train_batch()
  trainers/trainer.py:153-176                                -> draw order of z / tau
The arithmetic itself lives in torch (conv2d, batch_norm, interpolate, ...),
exactly as it does for the reference (SURVEY.md §8c "where the arithmetic
really lives"); what is restated is the composition and every quirk listed in
SURVEY.md Appendix B.
"""
import math
from collections import namedtuple

import torch
import torch.nn.functional as F

Spec = namedtuple('Spec', 'base_size latent_dims data_dims blocks attention')

# pluggan.py:199-406 (the table of named sizes); only the numbers matter.
SPECS = {
    '16': Spec(4, 100, 3, (64, 32), ()),
    '32': Spec(4, 128, 3, (128, 64, 32), ()),
    '64': Spec(4, 128, 3, (128, 128, 64, 32), ()),
    '128': Spec(4, 256, 3, (128, 128, 64, 32, 16), ()),
    '128big': Spec(4, 256, 3, (1024, 1024, 512, 256, 128), ()),
    '256': Spec(4, 256, 3, (256, 256, 128, 64, 32, 16), ()),
    '256big': Spec(4, 256, 3, (1024, 1024, 512, 256, 128, 64), ()),
    '512': Spec(4, 512, 3, (256, 256, 256, 128, 64, 32, 16), ()),
    '512thin': Spec(4, 256, 3, (128, 128, 128, 64, 32, 16, 8), (3,)),
    '512thin-test': Spec(4, 128, 3, (128, 120, 100, 64, 32, 16, 8), (3,)),
    '1024': Spec(4, 512, 3, (512, 512, 512, 256, 128, 64, 32, 16), (3,)),
    '1024thin': Spec(4, 256, 3, (256, 256, 256, 128, 64, 32, 16, 8), (3,)),
    'test128': Spec(4, 64, 3, (64, 32, 16, 8, 4), (3,)),
    'test256': Spec(4, 256, 3, (200, 180, 128, 64, 32, 16), (3,)),
}

SLOPE = 0.2          # "relu" in the CLI means LeakyReLU(0.2): trainers/cnn.py:41-45
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
NUM_QUANTILES = 8    # models/iqn.py:78
TAU_DIMS = 20        # models/iqn.py:78 quantile_dims


def scaled(spec, scale):
    """pluggan.py:24-28: channel counts are truncated with int()."""
    return spec._replace(blocks=tuple(int(c * scale) for c in spec.blocks))


# --------------------------------------------------------------------------
# leaf helpers working on a flat {name: tensor} dict ("sd")
# --------------------------------------------------------------------------
def _bn(sd, key, x, norm):
    """Train-mode BatchNorm2d (always train mode on this path, Appendix B.3)."""
    if norm != 'bn':
        return x
    out = F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'],
                       sd[key + '.weight'], sd[key + '.bias'],
                       True, BN_MOMENTUM, BN_EPS)
    sd[key + '.num_batches_tracked'] += 1
    return out


ACTIVATION = 'relu'   # --activation of trainers/cnn.py:41-45: 'relu' = LeakyReLU(0.2), 'selu' = nn.SELU, 'elu' = nn.ELU


def _act(x):
    if ACTIVATION == 'selu':
        return F.selu(x)
    if ACTIVATION == 'elu':
        return F.elu(x)
    return F.leaky_relu(x, SLOPE)


def init_params_selu(sd, names):
    """trainers/cnn.py:96-105: vectors zeroed, everything else N(0, 1/fan_in), in parameter order (CPU generator)."""
    for k in names:
        d = sd[k]
        if d.dim() == 1:
            d.zero_()
        else:
            fan_in = d[0].numel()
            d.normal_(std=math.sqrt(1. / fan_in))


# Optional activation trace for the per-layer parity tests: set ``TRACE`` to a dict and every conv /
# residual-block output of the next forward passes is recorded under its state-dict key (detached).
TRACE = None


def _trace(key, out):
    if TRACE is not None:
        TRACE[key] = out.detach().clone()
    return out


def _sn_weight(sd, key):
    """torch.nn.utils.spectral_norm semantics (the hooks tartangan/prep4web.py:33-51 strips; seam: conv_factory,
    generator.py:34, discriminator.py:28,52): ONE power-iteration step per training forward updating u, v in place,
    sigma = u^T W v with u, v treated as constants.  Pinned against torch itself in tests/test_oracle_golden.py."""
    w, u, v = sd[key + '.weight_orig'], sd[key + '.weight_u'], sd[key + '.weight_v']
    wm = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
        u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
    sigma = torch.dot(u.clone(), torch.mv(wm, v.clone()))
    return w / sigma


def _conv(sd, key, x, pad):
    w = _sn_weight(sd, key) if key + '.weight_orig' in sd else sd[key + '.weight']
    return _trace(key, F.conv2d(x, w, sd.get(key + '.bias'), padding=pad))


def _attention(sd, key, x):
    """attention.py:21-35.  N_q = H*W queries, N_kv = H*W/4 max-pooled keys."""
    n, c, h, w = x.shape
    q = _conv(sd, key + '.theta', x, 0).reshape(n, c // 8, h * w)
    k = F.max_pool2d(_conv(sd, key + '.phi', x, 0), 2).reshape(n, c // 8, h * w // 4)
    v = F.max_pool2d(_conv(sd, key + '.g', x, 0), 2).reshape(n, c // 2, h * w // 4)
    beta = torch.softmax(torch.bmm(q.transpose(1, 2), k), dim=-1)
    mixed = torch.bmm(v, beta.transpose(1, 2)).reshape(n, c // 2, h, w)
    return sd[key + '.gamma'] * _conv(sd, key + '.o', mixed, 0) + x


def _residual_convs(sd, key, x, first, norm):
    """The shared [BN, act,] conv3, BN, act, conv3 stack.  When ``first`` the
    leading BN/act are sliced off and the Sequential is re-indexed from 0
    (generator.py:46-48, discriminator.py:69-71)."""
    if first:
        i_c1, i_bn2, i_c2 = 0, 1, 3
    else:
        x = _act(_bn(sd, key + '.convs.0', x, norm))
        i_c1, i_bn2, i_c2 = 2, 3, 5
    x = _conv(sd, f'{key}.convs.{i_c1}', x, 1)
    x = _act(_bn(sd, f'{key}.convs.{i_bn2}', x, norm))
    return _conv(sd, f'{key}.convs.{i_c2}', x, 1)


def _g_block(sd, key, x, first, norm):
    """generator.py:56-62: upsample first (also for the first block), skip
    projection at the upsampled resolution."""
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    h = _residual_convs(sd, key, x, first, norm)
    if key + '.project_input.0.weight' in sd or key + '.project_input.0.weight_orig' in sd:
        x = _conv(sd, key + '.project_input.0', x, 0)
    return _trace(key, x + h)


def _d_block(sd, key, x, first, norm):
    """discriminator.py:90-95: avg-pooled conv stack + bilinear(0.5,
    align_corners=True) skip, projection after the down-sampling."""
    h = F.avg_pool2d(_residual_convs(sd, key, x, first, norm), 2)
    x = F.interpolate(x, scale_factor=0.5, mode='bilinear', align_corners=True)
    if key + '.project_input.0.weight' in sd or key + '.project_input.0.weight_orig' in sd:
        x = _conv(sd, key + '.project_input.0', x, 0)
    return _trace(key, x + h)


# --------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------
def generator(sd, spec, z, norm='bn', g_base='mlp'):
    c0 = spec.blocks[0]
    if g_base == 'mlp':     # generator.py:65-80
        x = _act(F.linear(z, sd['blocks.0.base_img.0.weight'], sd['blocks.0.base_img.0.bias']))
        x = x.view(-1, c0, spec.base_size, spec.base_size)
    else:                   # generator.py:101-112 (tiled z)
        x = z[..., None, None].repeat(1, 1, spec.base_size, spec.base_size)
    idx = 1
    for i, _ in enumerate(spec.blocks):
        x = _g_block(sd, f'blocks.{idx}', x, i == 0, norm)
        idx += 1
        if spec.attention and i in spec.attention:
            x = _attention(sd, f'blocks.{idx}', x)
            idx += 1
    key = f'blocks.{idx}'   # generator.py:115-129
    x = _act(_bn(sd, key + '.convs.0', x, norm))
    return torch.tanh(_conv(sd, key + '.convs.2', x, 0))


def _d_trunk(sd, spec, x, norm, idx, first_flag):
    for n, (i, _) in enumerate(reversed(list(enumerate(spec.blocks)))):
        x = _d_block(sd, f'blocks.{idx}', x, first_flag and n == 0, norm)
        idx += 1
        if spec.attention and i in spec.attention:
            x = _attention(sd, f'blocks.{idx}', x)
            idx += 1
    return x, idx


def discriminator(sd, spec, x, norm='bn'):
    """SA-GAN D: 1x1 input conv, first residual block without leading BN/act,
    head BN -> act -> sum over HW -> Linear (discriminator.py:141-146)."""
    x = _conv(sd, 'blocks.0.convs.0', x, 0)
    x, idx = _d_trunk(sd, spec, x, norm, 1, True)
    key = f'blocks.{idx}'
    feats = _act(_bn(sd, key + '.activation.0', x, norm)).sum((2, 3))
    return F.linear(feats, sd[key + '.to_output.0.weight'], sd[key + '.to_output.0.bias'])


def quantile_huber(preds, target, taus, k=1.0):
    """iqn.py:111-130: rows are quantile-major (row = q*B + b); sum over the
    quantile axis, mean over the rest; Huber term not divided by k."""
    b = target.shape[0]
    nq = preds.shape[0] // b
    err = target.reshape(1, b, -1) - preds.reshape(nq, b, -1)
    mag = err.abs()
    huber = torch.where(mag <= k, 0.5 * err * err, k * (mag - 0.5 * k))
    weight = (taus.reshape(nq, b, -1) - (err < 0).float()).abs()
    return (weight * huber).sum(0).mean()


def _iqn_head(sd, feats, taus, targets):
    """discriminator.py:164-178 + iqn.py:41-46, 91-103."""
    b = feats.shape[0]
    nq = taus.shape[0] // b
    rng = sd['to_output.iqn.quantile_embedding.embedding_range']
    cosines = torch.cos(taus.repeat(1, rng.numel()) * math.pi * rng)
    emb = torch.tanh(F.linear(cosines,
                              sd['to_output.iqn.quantile_embedding.to_state.0.weight'],
                              sd['to_output.iqn.quantile_embedding.to_state.0.bias']))
    mixed = feats.repeat(nq, 1) * emb
    p_tau = F.linear(mixed, sd['to_output.to_output.0.weight'], sd['to_output.to_output.0.bias'])
    p = p_tau.reshape(nq, -1, 1).mean(0)
    if targets is None:
        return p
    return p, quantile_huber(p_tau, targets, taus.repeat(1, 1))


def iqn_discriminator(sd, spec, x, targets=None, norm='bn', taus=None,
                      num_quantiles=NUM_QUANTILES):
    """SA-GAN-IQN D: no input conv, no first_block (BN over RGB first,
    Appendix B.5).  tau ~ U[0,1) from the CPU generator (iqn.py:105-108)."""
    x, _ = _d_trunk(sd, spec, x, norm, 0, False)
    feats = _act(_bn(sd, 'to_output.activation.0', x, norm)).sum((2, 3))
    if taus is None:
        taus = torch.rand(feats.shape[0] * num_quantiles, 1).to(feats.device)
    return _iqn_head(sd, feats, taus, targets)


def r1_penalty(preds, data):
    """losses.py:17-30."""
    grad, = torch.autograd.grad(preds.sum(), data, create_graph=True, retain_graph=True)
    return grad.pow(2).reshape(data.shape[0], -1).sum(1).mean()


# --------------------------------------------------------------------------
# one optimisation step (trainers/cnn.py:107-165, trainers/iqn.py:104-156)
# --------------------------------------------------------------------------
def _split(state_dict):
    """-> (params requiring grad, everything) sharing storage with nothing."""
    sd = {k: v.detach().clone().float() if v.is_floating_point() else v.detach().clone()
          for k, v in state_dict.items()}
    return sd


def _is_param(name):
    return not (name.endswith('running_mean') or name.endswith('running_var')
                or name.endswith('num_batches_tracked') or name.endswith('embedding_range')
                or name.endswith('weight_u') or name.endswith('weight_v'))


class OracleTrainer:
    """State-dict level twin of CNNTrainer / IQNTrainer.  ``kind`` is 'cnn' or
    'iqn'.  Parameters are held in registration order (the order of the state
    dict), which is also the order Adam and the EMA walk them in."""

    def __init__(self, kind, spec, g_state, target_g_state, d_state, batch_size,
                 lr_g=1e-4, lr_d=4e-4, lr_target_g=1e-3, grad_penalty=5.0,
                 norm='bn', g_base='mlp', num_quantiles=NUM_QUANTILES, activation='relu', device='cpu'):
        self.kind, self.spec, self.batch_size = kind, spec, batch_size
        self.activation = activation
        # device != 'cpu' is used by tools/bench_eager_gpu.py only (the same composition through stock cuDNN / cuBLAS
        # kernels, as a same-box library bar); z / tau are still drawn on the CPU like the reference does
        self.device = torch.device(device)
        g_state, target_g_state, d_state = ({k: v.to(self.device) for k, v in sd.items()}
                                            for sd in (g_state, target_g_state, d_state))
        self.norm, self.g_base, self.nq = norm, g_base, num_quantiles
        self.grad_penalty, self.lr_target_g = grad_penalty, lr_target_g
        self.g, self.target_g, self.d = _split(g_state), _split(target_g_state), _split(d_state)
        self.g_params = [k for k in self.g if _is_param(k)]
        self.d_params = [k for k in self.d if _is_param(k)]
        self.opt_g = torch.optim.Adam([self.g[k] for k in self.g_params], lr=lr_g, betas=(0., 0.999))
        self.opt_d = torch.optim.Adam([self.d[k] for k in self.d_params], lr=lr_d, betas=(0., 0.999))
        self.last_grads = {}

    def _toggle(self, sd, names, on):
        for k in names:
            sd[k].requires_grad_(on)

    def _fake(self, n):
        z = torch.randn(n, self.spec.latent_dims).to(self.device)           # trainer.py:153-156
        return generator(self.g, self.spec, z, self.norm, self.g_base)

    def _d(self, x, targets):
        if self.kind == 'iqn':
            return iqn_discriminator(self.d, self.spec, x, targets, self.norm,
                                     num_quantiles=self.nq)
        return discriminator(self.d, self.spec, x, self.norm)

    def train_batch(self, imgs):
        global ACTIVATION
        prev, ACTIVATION = ACTIVATION, self.activation
        try:
            return self._train_batch(imgs)
        finally:
            ACTIVATION = prev

    def _train_batch(self, imgs):
        b = self.batch_size
        # ---- D step
        self._toggle(self.g, self.g_params, False)
        self._toggle(self.d, self.d_params, True)
        self.opt_d.zero_grad()
        fake = self._fake(len(imgs))
        real = imgs.clone()
        labels = torch.zeros(2 * len(imgs), 1, device=self.device)
        labels[:len(imgs)] = 1
        if self.grad_penalty:
            real.requires_grad_()
        if self.kind == 'iqn':
            p_real, l_real = self._d(real, labels[:b])
            p_fake, l_fake = self._d(fake.detach(), labels[b:])
            d_loss = l_real + l_fake
        else:
            p_real = self._d(real, None)
            p_fake = self._d(fake.detach(), None)
            d_loss = F.binary_cross_entropy_with_logits(torch.cat([p_real, p_fake]), labels)
        gp = 0.
        if self.grad_penalty:
            gp = self.grad_penalty * r1_penalty(p_real, real)
            d_loss = d_loss + gp
        d_loss.backward()
        self.last_grads['d'] = {k: self.d[k].grad.detach().clone() for k in self.d_params
                                if self.d[k].grad is not None}
        self.opt_d.step()
        # ---- G step
        self._toggle(self.g, self.g_params, True)
        self._toggle(self.d, self.d_params, False)
        self.opt_g.zero_grad()
        fake = self._fake(len(imgs))
        ones = torch.ones(len(fake), 1, device=self.device)
        if self.kind == 'iqn':
            _, g_loss = self._d(fake, ones)
        else:
            g_loss = F.binary_cross_entropy_with_logits(self._d(fake, None), ones)
        g_loss.backward()
        self.last_grads['g'] = {k: self.g[k].grad.detach().clone() for k in self.g_params
                                if self.g[k].grad is not None}
        self.opt_g.step()
        # ---- EMA of the target generator: parameters only (Appendix B.1)
        with torch.no_grad():
            for k in self.g_params:
                self.target_g[k].add_((self.g[k] - self.target_g[k]) * self.lr_target_g)
        return dict(g_loss=float(g_loss.detach()), d_loss=float(d_loss.detach()),
                    gp=float(gp.detach()) if torch.is_tensor(gp) else float(gp))


def tartan_batch(seed, batch, size):
    """Deterministic synthetic 'tartan-shaped' RGB batch, fp32 NCHW in [-1,1]
    (SURVEY.md §8d): mirrored random sett, warp/weft, 2/2 twill mask."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(batch, 3, size, size)
    ys, xs = torch.meshgrid(torch.arange(size), torch.arange(size), indexing='ij')
    twill = (((xs + ys) // 2) % 2).float()
    for i in range(batch):
        n = int(torch.randint(3, 9, (1,), generator=g))
        widths = torch.randint(1, max(2, size // 8) + 1, (n,), generator=g)
        colours = torch.rand(n, 3, generator=g)
        sett = torch.repeat_interleave(colours, widths, dim=0)
        sett = torch.cat([sett, sett.flip(0)])
        reps = -(-size // sett.shape[0])
        sett = sett.repeat(reps, 1)[:size]                  # (size, 3)
        warp = sett.t()[:, None, :].expand(3, size, size)
        weft = sett.t()[:, :, None].expand(3, size, size)
        out[i] = (twill * warp + (1 - twill) * weft) * 2 - 1
    return out
