"""Parity at the configurations that are BENCHMARKED (BASELINE.json configs[1] and configs[2]): SA-GAN-IQN
'64' (C2) and '128' (C3, the headline), bf16 tensor-core kernels, eager AND CUDA-graph execution, against
the CPU oracle (fp32).  Reference path: /root/reference/tartangan/trainers/iqn.py:104-147 over
models/pluggan.py:223-249.

Bars (which measured what bf16 operands do to this model):
  * per-layer forward activations (every conv output, every residual-block output): relative L2 <= 2 % for EVERY
    tensor ("worst 1.97 %, median 0.96 %" for bf16 forward activations, App. D row 1).  Measured on B200 (round 2):
    median 0.5-0.7 %, worst 1.0-1.3 %;
  * per-parameter gradient cosine after the D backward and after the G backward (App. D row 2: 11-20 % relative L2,
    i.e. cosine 0.98-0.99): >= 0.97 for every tensor with at least 64 elements, >= 0.90 for the short per-channel
    vectors (BatchNorm gamma / beta, biases: 16 elements at the 128x128 end of the generator, a sum of 4 Mi
    bf16-rounded products each), mean over tensors >= 0.975.  Measured: D min 0.993, mean 0.998; G min 0.984 ('64'),
    mean 0.98-0.99; the one tensor below 0.97 is the 16-element gamma of G's output BatchNorm at '128': a sum of 512 Ki
    strongly cancelling bf16 products per element, 0.950 / 0.955 / 0.969 in three runs of the same code (the
    tensor-core wgrad of the layers behind it sums with fp32 atomics, so the value is not run-to-run repeatable).  Parameters whose
    gradient is analytically zero (conv biases that feed a train-mode BatchNorm) are excluded;
  * 20-step loss curves: teacher-forced (state re-synchronised from the oracle before each step) every loss
    within 1 % (measured: worst 0.16 %); free-running first three steps within 10 % and the step 10-19 means within
    20 % (measured 1.7-8 %; App. D row 3: <= 9-17 % per step, GAN trajectories decorrelate after ~8 steps).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

# (config key, per-GPU batch used here).  C2 = iqn '64' batch 64 as benchmarked; C3 = iqn '128' at batch 32:
# the oracle runs on the host cores (batch 256 would take minutes per step), and batch only changes how many
# tiles the same kernels loop over (the N256 kernels are exercised by bench.py's own smoke check).
CONFIGS = [('64', 64), ('128', 32)]
ACT_WORST, ACT_MEDIAN = 0.02, 0.01
COSINE, COSINE_SHORT, COSINE_MEAN = 0.97, 0.90, 0.975


def _cpu_state(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def _make(config, batch, graph=False, precision='bf16'):
    from oracle import tartan_oracle as O
    from tartangan_b200.trainers.gan import make_trainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer, config=config, batch_size=batch, precision=precision, cuda_graph=graph)
    orc = O.OracleTrainer('iqn', O.SPECS[config], _cpu_state(t.g), _cpu_state(t.target_g), _cpu_state(t.d), batch)
    return t, orc


def _rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if a.shape != b.shape and a.dim() == 4 and a.shape[2] * 2 == b.shape[2]:
        # the generator's 1x1 skip projection runs BEFORE the nearest upsample here (they commute)
        a = a.repeat_interleave(2, 2).repeat_interleave(2, 3)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _cos(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _hook_outputs(model):
    from tartangan_b200.models.layers import Conv2d
    from tartangan_b200.models.blocks import ResidualDiscriminatorBlock, ResidualGeneratorBlock
    acts, handles = {}, []
    for name, m in model.named_modules():
        if isinstance(m, (Conv2d, ResidualGeneratorBlock, ResidualDiscriminatorBlock)):
            handles.append(m.register_forward_hook(
                lambda mod, inp, out, name=name: acts.__setitem__(name, out.detach().float().cpu())))
    return acts, handles


@pytest.mark.parametrize('config,batch', CONFIGS)
def test_forward_activations_bf16(config, batch):
    """Every conv output and residual-block output of G(z) and D(x) against the oracle's, same weights and inputs."""
    from oracle import tartan_oracle as O
    t, orc = _make(config, batch)
    size = t.g.max_size
    z = torch.randn(batch, t.gan_config.latent_dims, generator=torch.Generator().manual_seed(5))
    x = O.tartan_batch(77, batch, size)
    worst = {}
    for net, run_ref, run_cuda in (
            ('g', lambda: O.generator(orc.g, orc.spec, z), lambda: t.g(z.cuda())),
            ('d', lambda: O.iqn_discriminator(orc.d, orc.spec, x, torch.ones(batch, 1)),
             lambda: t.d(x.cuda(), targets=torch.ones(batch, 1, device='cuda')))):
        O.TRACE = {}
        try:
            torch.manual_seed(9)
            with torch.no_grad():
                ref_out = run_ref()
            ref_acts = O.TRACE
        finally:
            O.TRACE = None
        acts, handles = _hook_outputs(getattr(t, net))
        torch.manual_seed(9)
        with torch.no_grad():
            out = run_cuda()
        for h in handles:
            h.remove()
        # the generator's C -> 3 output conv is fused with tanh and the layout boundary (ops.RgbHeadFn): its
        # pre-activation is never stored, the IMAGE below is what is compared for that layer
        head = {f'blocks.{len(t.g.blocks) - 1}.convs.2'} if net == 'g' else set()
        assert set(ref_acts) - set(acts) <= head, (net, sorted(set(ref_acts) - set(acts)))
        errs = {k: _rel_l2(acts[k], v) for k, v in ref_acts.items() if k in acts}
        med = sorted(errs.values())[len(errs) // 2]
        k_worst = max(errs, key=errs.get)
        worst[net] = (k_worst, errs[k_worst], med)
        print(f'[parity] {config} {net}: {len(errs)} tensors, median rel-L2 {med:.4f}, worst {k_worst} {errs[k_worst]:.4f}')
        assert errs[k_worst] <= ACT_WORST and med <= ACT_MEDIAN, (net, k_worst, errs[k_worst], med)
        if net == 'g':
            assert _rel_l2(out, ref_out) <= ACT_WORST
        else:
            assert _rel_l2(out[0], ref_out[0]) <= ACT_WORST, 'p_target'
            assert abs(float(out[1]) - float(ref_out[1])) <= 0.03 * max(1.0, abs(float(ref_out[1]))), 'loss'


def _zero_grad_params(orc):
    """Conv biases feeding a train-mode BatchNorm: analytic gradient 0, the oracle holds rounding noise."""
    skip = set()
    for net in ('d', 'g'):
        grads = orc.last_grads[net]
        scale = max(float(v.abs().max()) for v in grads.values())
        for k, v in grads.items():
            if k.endswith('.bias') and float(v.abs().max()) < max(2e-3 * scale, 5e-5):
                skip.add((net, k))
    return skip


@pytest.mark.parametrize('graph', [False, True], ids=['eager', 'graph'])
@pytest.mark.parametrize('config,batch', CONFIGS)
def test_step_gradients_bf16(config, batch, graph):
    """One full train_batch: losses within 5 %, cosine of EVERY parameter gradient (D after d_loss.backward(),
    G after g_loss.backward()) against the oracle's (bars in the module docstring), in eager and in CUDA-graph execution."""
    from oracle import tartan_oracle as O
    t, orc = _make(config, batch, graph=graph)
    imgs = O.tartan_batch(1234, batch, t.g.max_size)
    torch.manual_seed(300)
    ref = orc.train_batch(imgs)
    torch.manual_seed(300)
    got = t.train_batch(imgs)
    torch.cuda.synchronize()
    for k in ref:
        assert abs(got[k] - ref[k]) <= 0.05 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    skip = _zero_grad_params(orc)
    low = []
    for net, mod in (('d', t.d), ('g', t.g)):
        mine = dict(mod.named_parameters())
        cosines = {}
        for k, gref in orc.last_grads[net].items():
            if (net, k) in skip:
                continue
            assert mine[k].grad is not None, (net, k)
            cosines[k] = _cos(mine[k].grad, gref)
        k_min = min(cosines, key=cosines.get)
        print(f'[parity] {config} {"graph" if graph else "eager"} {net}: {len(cosines)} gradients, '
              f'min cosine {k_min} {cosines[k_min]:.4f}, mean {sum(cosines.values()) / len(cosines):.4f}')
        low += [(net, k, round(c, 4)) for k, c in cosines.items()
                if not c >= (COSINE if mine[k].numel() >= 64 else COSINE_SHORT)]
        assert sum(cosines.values()) / len(cosines) >= COSINE_MEAN, (net, sum(cosines.values()) / len(cosines))
    assert not low, low


def _sync_from_oracle(t, orc):
    """The CUDA trainer adopts the oracle's parameters, buffers and Adam moments (teacher forcing)."""
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
    t.parameters_changed()


@pytest.mark.parametrize('config,batch', CONFIGS)
def test_loss_curve_20_steps_bf16_graph(config, batch):
    """20 steps of the benchmarked execution mode (bf16, CUDA graphs) against the fp32 oracle."""
    from oracle import tartan_oracle as O
    size = int(config)
    # (1) teacher-forced: every step starts from the oracle's state
    t, orc = _make(config, batch, graph=True)
    ref_curve, worst = [], 0.0
    for s in range(20):
        _sync_from_oracle(t, orc)
        imgs = O.tartan_batch(1234 + s, batch, size)
        torch.manual_seed(2000 + s)
        ref = orc.train_batch(imgs)
        torch.manual_seed(2000 + s)
        got = t.train_batch(imgs)
        ref_curve.append(ref)
        for k in ref:
            err = abs(got[k] - ref[k]) / max(1.0, abs(ref[k]))
            worst = max(worst, err)
            assert math.isfinite(got[k]) and err <= 0.01, (s, k, got[k], ref[k])
    print(f'[parity] {config} teacher-forced 20 steps: worst loss deviation {worst:.4f}')
    # (2) free-running from the same initial state
    t2, _ = _make(config, batch, graph=True)
    free = []
    for s in range(20):
        torch.manual_seed(2000 + s)
        free.append(t2.train_batch(O.tartan_batch(1234 + s, batch, size)))
    for s in range(3):
        for k in ('d_loss', 'gp', 'g_loss'):
            assert abs(free[s][k] - ref_curve[s][k]) <= 0.10 * max(1.0, abs(ref_curve[s][k])), (s, k, free[s], ref_curve[s])
    for k in ('d_loss', 'gp'):
        a = sum(m[k] for m in free[10:]) / 10
        b = sum(m[k] for m in ref_curve[10:]) / 10
        print(f'[parity] {config} free-running mean {k} steps 10-19: {a:.4f} vs oracle {b:.4f}')
        assert abs(a - b) <= 0.20 * max(abs(b), 0.05), (k, a, b)


def test_target_g_sampling_after_graph_steps():
    """SURVEY 8 f-2 (components/image_sampler.py:24-45): sample_g(target_g=True) under no_grad after graphed
    steps must use the CURRENT weights (the EMA runs inside the graph, behind Python's back) and match an eager
    trainer that took the same steps."""
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    from tartangan_b200.trainers.gan import make_trainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    outs = []
    for graph in (False, True):
        torch.manual_seed(0)
        t = make_trainer(IQNTrainer, config='32', batch_size=8, precision='fp32', cuda_graph=graph, lr_target_g=0.5)
        samples = []
        for s in range(3):
            torch.manual_seed(40 + s)
            t.train_batch(O.tartan_batch(10 + s, 8, 32))
            torch.manual_seed(99)
            with torch.no_grad():
                samples.append(t.sample_g(4, target_g=True).float().cpu())
                g_now = t.sample_g(4).float().cpu()
            assert samples[-1].shape == (4, 3, 32, 32) and float(samples[-1].abs().max()) <= 1.0
            assert torch.isfinite(g_now).all()
        outs.append(samples)
        assert float((samples[0] - samples[2]).abs().max()) > 1e-3          # the EMA (lr 0.5) visibly moved target_g
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= 2e-2, float((a - b).abs().max())
