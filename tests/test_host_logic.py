"""CPU-only checks of the host side: C-ABI surface, module/state-dict parity with the reference's
initial states (golden), loud failure without CUDA, flat parameter buffers, and the data-parallel
gradient exchange over gloo (world size 2)."""
import functools
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def test_header_and_library_agree():
    from tartangan_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 55
    _lib.lib.load()                     # raises if any declared symbol is missing from the .so
    assert _lib.lib.ttg_version() >= 100
    assert isinstance(_lib.last_error(), str)
    out = subprocess.run(['nm', '-D', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ttg_' in l}
    assert set(protos) <= exported, set(protos) - exported


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: include/ttg_b200.h compiles as strict C99 (no C++ / torch types in the signatures), a
    C program that takes the address of EVERY declared entry point links against the library, and the calls that need
    no device (version, error string) work from C."""
    from tartangan_b200 import _lib
    names = sorted(_lib.parse_header())
    src = tmp_path / 'abi.c'
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "ttg_b200.h"\n'
                   'typedef void (*fn)(void);\n'
                   'static fn table[] = {' + ', '.join(f'(fn){n}' for n in names) + '};\n'
                   'int main(void) {\n'
                   '  size_t i, n = sizeof table / sizeof table[0];\n'
                   '  for (i = 0; i < n; ++i) if (!table[i]) return 2;\n'
                   '  if (ttg_version() < 100) return 3;\n'
                   '  if (!ttg_last_error()) return 4;\n'
                   '  printf("%d symbols\\n", (int)n);\n'
                   '  return 0;\n}\n')
    exe = tmp_path / 'abi'
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror',
                        '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe), '-L', lib_dir, '-lttg_b200',
                        f'-Wl,-rpath,{lib_dir}'], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.strip() == f'{len(names)} symbols'


def _build(kind, g):
    from tartangan_b200.models import pluggan
    from tartangan_b200.models.blocks import (DiscriminatorOutput, GeneratorInputMLP, GeneratorOutput,
                                               IQNDiscriminatorOutput, ResidualDiscriminatorBlock,
                                               ResidualGeneratorBlock, TiledZGeneratorInput)
    from tartangan_b200.models.layers import BatchNorm2d
    from torch import nn
    cfg = pluggan.GANConfig(base_size=4, latent_dims=g['latent'], data_dims=3, blocks=tuple(g['blocks']),
                            num_blocks_per_scale=1, attention=tuple(g['attention']))
    norm = {'bn': BatchNorm2d, 'id': nn.Identity}[g['norm']]
    gf = dict(input_factory=TiledZGeneratorInput if g.get('g_base') == 'tiledz' else GeneratorInputMLP,
              block_factory=functools.partial(ResidualGeneratorBlock, norm_factory=norm),
              output_factory=functools.partial(GeneratorOutput, norm_factory=norm))
    torch.manual_seed(0)
    gen = pluggan.Generator(cfg, **gf)
    tgt = pluggan.Generator(cfg, **gf)
    dcls, ocls = ((pluggan.Discriminator, DiscriminatorOutput) if kind == 'cnn'
                  else (pluggan.IQNDiscriminator, IQNDiscriminatorOutput))
    d = dcls(cfg, block_factory=functools.partial(ResidualDiscriminatorBlock, norm_factory=norm),
             output_factory=functools.partial(ocls, norm_factory=norm))
    return gen, tgt, d


def test_modules_match_reference_initial_state(golden):
    """Same seed -> same parameter tensors under the same state-dict keys as the reference modules."""
    g = golden
    if g.get('activation') == 'selu':
        pytest.skip('init_params_selu re-draws every tensor in the trainer (checked on the GPU: test_golden_fp32)')
    gen, tgt, d = _build(g['kind'], g)
    for name, mod in (('g', gen), ('d', d)):
        ref, sd = g['init'][name], mod.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k, v in ref.items():
            if k.endswith('gamma'):
                continue                       # the golden script sets attention gamma to 0.5 after init
            assert torch.equal(sd[k], v), (name, k)
    assert gen.max_size == g['size']


def test_no_cpu_fallback():
    g = load_golden('cnn_tiny')
    gen, _, d = _build('cnn', g)
    with pytest.raises(RuntimeError, match='CUDA'):
        gen(torch.randn(2, g['latent']))
    with pytest.raises(RuntimeError, match='CUDA'):
        d(torch.randn(2, 3, g['size'], g['size']))
    from tartangan_b200.trainers.utils import set_device_from_args
    import argparse
    with pytest.raises(RuntimeError):
        set_device_from_args(argparse.Namespace(no_cuda=True))


def test_flat_params_views():
    from tartangan_b200.optim import FlatParams
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2, 3))]
    before = [p.detach().clone() for p in ps]
    flat = FlatParams(ps)
    assert flat.intact() and flat.numel % 4 == 0
    for p, b, o in zip(ps, before, flat.offsets):
        assert torch.equal(p.data, b) and o % 4 == 0
        assert p.data_ptr() == flat.data.data_ptr() + 4 * o
    flat.attach_grads()
    ps[1].grad.add_(1.0)
    assert float(flat.grad.sum()) == 7.0


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tartangan_b200.optim import FlatParams
    from tartangan_b200.parallel import BucketedAllReduce, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 6), torch.nn.Tanh(), torch.nn.Linear(6, 6), torch.nn.Tanh(), torch.nn.Linear(6, 6), torch.nn.Linear(6, 1))
    flat = FlatParams(list(model.parameters()))
    red = BucketedAllReduce(flat.params, flat.offsets, flat.grad, num_buckets=2)
    assert len(red.buckets) >= 2 and red.buckets[0][0] == 0 and red.buckets[-1][1] == flat.numel
    assert all(a[1] == b[0] for a, b in zip(red.buckets, red.buckets[1:]))      # contiguous cover
    assert sorted(i for _, _, m in red.buckets for i in m) == list(range(len(flat.params)))
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    xs = x[shard_batch(8, world, rank)]
    flat.attach_grads()
    red.begin()
    loss = model(xs).pow(2).mean()
    loss.backward(torch.full_like(loss, 1.0 / world))
    red.finish()
    if rank == 0:
        torch.save(flat.grad.clone(), out)
    dist.destroy_process_group()


def test_data_parallel_gloo_world2(tmp_path):
    """Two ranks, each on half the batch: exchanged gradients equal the full-batch gradient."""
    out = str(tmp_path / 'grad.pt')
    port = 29500 + os.getpid() % 1000
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    from tartangan_b200.optim import FlatParams
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 6), torch.nn.Tanh(), torch.nn.Linear(6, 6), torch.nn.Tanh(), torch.nn.Linear(6, 6), torch.nn.Linear(6, 1))
    flat = FlatParams(list(model.parameters()))
    flat.attach_grads()
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    model(x).pow(2).mean().backward()
    assert torch.allclose(got, flat.grad, atol=1e-6)


class _DirectLinearFn(torch.autograd.Function):
    """y = x w^T whose backward ADDS the weight gradient to w.grad and returns None for it: the CPU twin of
    ops.direct_param_grads (ttg_conv2d_wgrad_tc_acc ...).  autograd still runs w's AccumulateGrad node (with an
    undefined gradient: no add kernel), so the post-accumulate-grad hook fires once every contribution is in."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        w.grad.add_(gy.t() @ x)
        return gy @ w, None


def _dp_worker_direct(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tartangan_b200.optim import FlatParams
    from tartangan_b200.parallel import BucketedAllReduce, shard_batch
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(6, 6) * 0.3) for _ in range(4)]
    flat = FlatParams(ws)
    red = BucketedAllReduce(flat.params, flat.offsets, flat.grad, num_buckets=2)
    torch.manual_seed(7)
    x = torch.randn(8, 6)[shard_batch(8, world, rank)]
    flat.attach_grads()
    red.begin()
    h = x
    for i, w in enumerate(ws):            # parameters 1 and 2 take the in-place path
        h = torch.tanh(_DirectLinearFn.apply(h, w) if i in (1, 2) else h @ w.t())
    loss = h.pow(2).mean()
    loss.backward(torch.full_like(loss, 1.0 / world))
    assert all(p == 0 for p in red.pending)         # every hook fired: the buckets were launched during backward
    red.finish()
    if rank == 0:
        torch.save(flat.grad.clone(), out)
    dist.destroy_process_group()


def test_data_parallel_with_in_place_parameter_gradients(tmp_path):
    """Data parallel + ops.direct_param_grads: parameters whose gradient is added to .grad by the producing kernel
    (backward returns None) still fire their bucket hooks after the last contribution, so the exchange keeps
    overlapping backward and yields the full-batch gradient (gloo, world 2)."""
    out = str(tmp_path / 'grad.pt')
    port = 29500 + (os.getpid() + 17) % 1000
    mp.spawn(_dp_worker_direct, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(6, 6) * 0.3) for _ in range(4)]
    torch.manual_seed(7)
    h = torch.randn(8, 6)
    for w in ws:
        h = torch.tanh(h @ w.t())
    h.pow(2).mean().backward()
    ref = torch.cat([w.grad.reshape(-1) for w in ws])
    assert got.numel() == ref.numel() and torch.allclose(got, ref, atol=1e-6)


def test_bench_reference_arm_prints_contract_line():
    # launched the way torch.distributed.run launches a worker of an N > 1 job: OMP_NUM_THREADS=1 in the environment,
    # which the CPU arm must override (it is to use every host core; rank 0 alone computes)
    env = dict(os.environ, OMP_NUM_THREADS='1', RANK='0', WORLD_SIZE='2', LOCAL_RANK='0')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['cpu_baseline']['omp_num_threads'] == str(os.cpu_count()) == str(line['cpu_baseline']['cores'])
    assert line['impl'] == 'reference' and line['unit'] == 'images/sec' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['e2e']['h2d_bytes_per_step'] == 0


def test_zero_arena_invariants():
    """ops.ZeroArena (pre-zeroed workspaces of a half-step): every slice handed out is all zero, slices never
    overlap within a period, a reset clears whatever earlier periods dirtied (also when a later period is shorter
    than an earlier one), and a request that does not fit yields None (the caller then memsets an ordinary buffer)."""
    from tartangan_b200 import ops
    a = ops.ZeroArena('cpu', nbytes=4096)
    seen = []
    for period, sizes in enumerate([(100, 700, 8, 1024), (16, 16), (2000, 1000), (5000,), (64,)]):
        a.reset()
        spans = []
        for n in sizes:
            t = a.take(n)
            if n > 4096:
                assert t is None
                continue
            assert t is not None and t.numel() == n and int(t.abs().sum()) == 0, (period, n)
            lo = t.data_ptr() - a.buf.data_ptr()
            assert lo % 256 == 0
            assert all(lo >= hi2 or lo + n <= lo2 for lo2, hi2 in spans)
            spans.append((lo, lo + n))
            t.fill_(255)                    # the kernels leave their partial sums behind
        seen.append(spans)
    # without a reset the bump pointer keeps going: still zero, still disjoint from the current period's slices
    t = a.take(32)
    assert t is not None and int(t.abs().sum()) == 0
    assert a.take(4096) is None


def test_fused_attention_shape_support():
    """ttg_attn_supported: the attention shapes of BASELINE configs 4 and 5 are covered, tiny test shapes and channel
    counts outside C = 64 / 128 are not (they take the bmm / softmax / bmm path, never a silent CPU fallback)."""
    from tartangan_b200 import _lib
    ok = _lib.lib.ttg_attn_supported
    for nq, nk, dk, dv in [(4096, 1024, 8, 32), (1024, 256, 8, 32), (4096, 1024, 16, 64), (1024, 256, 16, 64)]:
        assert ok(nq, nk, dk, dv) == 1
    for nq, nk, dk, dv in [(64, 16, 2, 8), (4096, 1024, 32, 128), (1000, 256, 8, 32), (1024, 192, 8, 32), (1024, 64, 8, 32)]:
        assert ok(nq, nk, dk, dv) == 0
    assert _lib.lib.ttg_attn_bwd_workspace_bytes(2, 1024, 8) == (2 * 1024 * 8 + 2 * 2 * 1024) * 4


def test_reference_written_checkpoint_unpickles_into_mirror_classes():
    """f-1, reference -> here, without a GPU: the whole-object files of tests/golden/ref_checkpoint (saved by the
    unmodified reference) unpickle through install_as_tartangan() into this package's classes and give back the
    reference's state dicts, key order included."""
    from conftest import GOLDEN_DIR
    import tartangan_b200
    tartangan_b200.install_as_tartangan()
    root = os.path.join(GOLDEN_DIR, 'ref_checkpoint')
    exp = torch.load(os.path.join(root, 'expected_state.pt'), weights_only=False)
    for name, key in (('g.pt', 'g'), ('g_target.pt', 'target_g'), ('d.pt', 'd'), ('opt_d.pt', 'opt_d'), ('opt_g.pt', 'opt_g')):
        obj = torch.load(os.path.join(root, 'checkpoints', '1', name), weights_only=False)
        sd = obj.state_dict()
        if key.startswith('opt'):
            assert sd['param_groups'][0]['betas'] == (0.0, 0.999)
            for i, st in exp[key]['state'].items():
                assert torch.equal(sd['state'][i]['exp_avg_sq'], st['exp_avg_sq'])
        else:
            assert type(obj).__module__.startswith('tartangan_b200.models')
            assert list(sd) == list(exp[key]) and all(torch.equal(sd[k], v) for k, v in exp[key].items())


@pytest.mark.skipif(not os.path.isdir('/root/reference/tartangan'), reason='needs the reference tree (build container only)')
def test_reference_loader_reads_our_checkpoints(tmp_path):
    """f-1, here -> reference: files written by checkpoint_compat.save_reference_object are unpickled by the UNMODIFIED
    reference (a separate interpreter with only torch + /root/reference on its path: no tartangan_b200) into its own
    classes; `.state_dict()` (all its loader needs, model_checkpoint.py:66) equals ours and the reference can run the
    loaded generator / discriminator."""
    from tartangan_b200.checkpoint_compat import save_reference_object
    from tartangan_b200.models import pluggan
    from tartangan_b200.models.blocks import (GeneratorInputMLP, GeneratorOutput, IQNDiscriminatorOutput,
                                               ResidualDiscriminatorBlock, ResidualGeneratorBlock)
    from tartangan_b200.optim import FusedAdam
    torch.manual_seed(3)
    cfg = pluggan.GANConfig(base_size=4, latent_dims=16, data_dims=3, blocks=(16, 8, 8), num_blocks_per_scale=1, attention=(1,))
    g = pluggan.Generator(cfg, input_factory=GeneratorInputMLP, block_factory=ResidualGeneratorBlock, output_factory=GeneratorOutput)
    d = pluggan.IQNDiscriminator(cfg, block_factory=ResidualDiscriminatorBlock, output_factory=IQNDiscriminatorOutput)
    opt = FusedAdam(g.parameters(), lr=1e-4, betas=(0., 0.999))
    for p in g.parameters():          # state as after a step (the CUDA step itself is not needed to test the format)
        opt.state[p] = {'step': torch.tensor(3.), 'exp_avg': torch.zeros_like(p), 'exp_avg_sq': torch.rand_like(p)}
    for obj, name in ((g, 'g'), (d, 'd'), (opt, 'opt_g')):
        save_reference_object(obj, str(tmp_path / f'{name}.pt'))
        torch.save(obj.state_dict(), str(tmp_path / f'{name}_sd.pt'))
    assert 'tartangan' not in sys.modules or sys.modules['tartangan'].__dict__.get('models') is not None
    script = f"""
import sys, types
sys.path.insert(0, '/root/reference')
so = types.ModuleType('smart_open'); so.open = open; sys.modules['smart_open'] = so
b3 = types.ModuleType('boto3'); b3.resource = b3.client = (lambda *a, **k: None); sys.modules['boto3'] = b3
import torch
for n in ('g', 'd'):
    m = torch.load(r'{tmp_path}/' + n + '.pt', weights_only=False)      # what model_checkpoint.py:64 does
    assert type(m).__module__ == 'tartangan.models.pluggan', type(m)
    sd, ref = m.state_dict(), torch.load(r'{tmp_path}/' + n + '_sd.pt')
    assert list(sd) == list(ref) and all(torch.equal(sd[k], ref[k]) for k in ref)
    m.train()
    out = m(torch.randn(2, 16)) if n == 'g' else m(torch.randn(2, 3, 32, 32), targets=torch.ones(2, 1))[0]
    assert torch.isfinite(out).all()
o = torch.load(r'{tmp_path}/opt_g.pt', weights_only=False)
assert type(o) is torch.optim.Adam
sd, ref = o.state_dict(), torch.load(r'{tmp_path}/opt_g_sd.pt', weights_only=False)
assert sd['param_groups'][0]['lr'] == 1e-4 and tuple(sd['param_groups'][0]['betas']) == (0.0, 0.999)
assert all(torch.equal(sd['state'][i]['exp_avg_sq'], ref['state'][i]['exp_avg_sq']) for i in ref['state'])
assert 'tartangan_b200' not in sys.modules
print('REFERENCE_LOADED_OK')
"""
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert 'REFERENCE_LOADED_OK' in r.stdout, r.stderr[-2000:]


def test_fid_statistics_match_scipy_and_numpy(tmp_path):
    """f-4 (inception_utils.py:180-247, components/metrics/fid.py): covariance, Newton-Schulz matrix square root, Frechet
    distance and Inception score against scipy / numpy; the accumulation loop with a stand-in feature network."""
    import numpy as np
    from scipy import linalg
    from tartangan_b200.trainers import fid as F_
    rng = np.random.RandomState(0)
    a, b = rng.randn(400, 24), rng.randn(300, 24) * 1.3 + 0.2
    ta, tb = torch.tensor(a), torch.tensor(b)
    assert np.allclose(F_.torch_cov(ta).numpy(), np.cov(a, rowvar=False), atol=1e-10)
    s1, s2 = np.cov(a, rowvar=False), np.cov(b, rowvar=False)
    root = F_.sqrt_newton_schulz(torch.tensor(s1 @ s2)).numpy()
    assert np.allclose(root @ root, s1 @ s2, atol=1e-6)
    want = ((a.mean(0) - b.mean(0)) ** 2).sum() + np.trace(s1) + np.trace(s2) - 2 * np.trace(linalg.sqrtm(s1 @ s2).real)
    got = float(F_.frechet_distance(ta.mean(0), F_.torch_cov(ta), tb.mean(0), F_.torch_cov(tb)))
    assert abs(got - want) <= 1e-5 * abs(want)
    assert abs(float(F_.frechet_distance(ta.mean(0), F_.torch_cov(ta), ta.mean(0), F_.torch_cov(ta)))) < 1e-6
    p = torch.softmax(torch.tensor(rng.randn(200, 10) * 2), 1)
    scores = []
    for i in range(5):
        c = p[i * 40:(i + 1) * 40].numpy()
        scores.append(np.exp(np.mean(np.sum(c * (np.log(c) - np.log(c.mean(0, keepdims=True))), 1))))
    m, sd = F_.inception_score(p, 5)
    assert abs(m - np.mean(scores)) < 1e-9 and abs(sd - np.std(scores)) < 1e-9
    # the loop: a stand-in feature network, moments of one distribution, samples of the same and of a shifted one

    class Net(torch.nn.Module):
        def forward(self, x):
            f = x.flatten(1)[:, :16]
            return f, f[:, :10]
    torch.manual_seed(0)
    ref = torch.randn(4000, 3, 4, 4).clamp(-1, 1)
    pool, _ = F_.accumulate_activations(lambda: ref, Net(), 4000)
    np.savez(tmp_path / 'moments.npz', mu=pool.mean(0).numpy(), sigma=F_.torch_cov(pool).numpy())
    metrics = F_.InceptionMetrics(str(tmp_path / 'moments.npz'), 'cpu', net=Net())
    same = metrics.get(lambda: torch.randn(500, 3, 4, 4).clamp(-1, 1), 4000, num_splits=5)
    shifted = metrics.get(lambda: (torch.randn(500, 3, 4, 4) + 0.5).clamp(-1, 1), 4000, num_splits=5)
    assert same[2] < 0.2 and shifted[2] > 10 * max(same[2], 1e-3)
    with pytest.raises(FileNotFoundError):
        F_.load_inception_net(str(tmp_path / 'missing.pth'))


@pytest.mark.skipif(not os.path.isdir('/root/reference/tartangan'), reason='needs the reference tree (build container only)')
def test_inception_feature_wrapper_matches_reference():
    """f-4: InceptionFeatures on a randomly initialised torchvision Inception-v3 gives the same pool features and logits
    as the reference's WrapInception around the SAME network (inception_utils.py:33-81) - no pretrained weights needed."""
    import types
    from torchvision.models import inception_v3
    from tartangan_b200.trainers.fid import InceptionFeatures
    sys.path.insert(0, '/root/reference')
    for name, mod in (('smart_open', dict(open=open)), ('boto3', dict(resource=lambda *a, **k: None, client=lambda *a, **k: None))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(mod)
            sys.modules[name] = m
    # an earlier test may have aliased `tartangan` to this package (install_as_tartangan): set those entries aside,
    # import the real reference module, then put everything back
    aside = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'tartangan' or k.startswith('tartangan.')}
    try:
        import importlib
        ref = importlib.import_module('tartangan.inception_utils')
    finally:
        sys.path.remove('/root/reference')
        for k in [k for k in sys.modules if k == 'tartangan' or k.startswith('tartangan.')]:
            del sys.modules[k]
        sys.modules.update(aside)
    torch.manual_seed(0)
    net = inception_v3(weights=None, aux_logits=True, transform_input=False, init_weights=True).eval()
    x = torch.rand(2, 3, 64, 64) * 2 - 1
    with torch.no_grad():
        p_ref, l_ref = ref.WrapInception(net)(x)
        p, l = InceptionFeatures(net)(x)
    assert torch.allclose(p, p_ref, atol=1e-5, rtol=1e-4) and torch.allclose(l, l_ref, atol=1e-5, rtol=1e-4)


@pytest.mark.skipif(not os.path.isdir('/root/reference/tartangan'), reason='needs the reference tree (build container only)')
def test_slerp_grid_matches_reference():
    """f-2: the 5x5 interpolation grid of the progress sampler against the reference's utils/slerp.py (loaded from its
    file: numpy + torch only), incl. the parallel-corner (linear) branch; same dtype, values to float32 rounding."""
    import importlib.util
    from tartangan_b200.trainers.sampler import slerp, slerp_grid
    spec = importlib.util.spec_from_file_location('_ref_slerp', '/root/reference/tartangan/utils/slerp.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(3)
    for latent, (nr, nc) in ((64, (5, 5)), (256, (5, 5)), (16, (3, 7))):
        z = torch.randn(4, latent)
        import warnings
        with warnings.catch_warnings():         # (numpy 2 deprecation noise from np.dot on torch tensors)
            warnings.simplefilter('ignore')
            want = ref.slerp_grid(*[t for t in z], nr, nc)          # the reference passes CPU tensors (image_sampler.py:47-53)
        got = slerp_grid(*[t.numpy() for t in z], nr, nc)
        assert got.dtype == want.dtype == torch.float32 and got.shape == want.shape == (nr * nc, latent)
        assert torch.allclose(got, want, atol=2e-6, rtol=1e-5)
        assert torch.equal(got[0], z[0]) and torch.allclose(got[-1], z[3], atol=1e-6)
    v = np.arange(1, 9, dtype=np.float32)
    assert np.allclose(slerp(0.25, v, 2 * v), ref.slerp(0.25, v, 2 * v))         # omega = 0: LERP branch


def test_npz_dataset_matches_the_reference_transform_pipeline(tmp_path):
    """f-3 (host side): NpzImageDataset against the transform stack the reference builds for .npz data
    (trainers/trainer.py:68-74: ToPILImage -> RandomCrop -> ToTensor -> Normalize(0.5, 0.5)) with torchvision itself:
    bit-identical pixels for every byte value (crop size = image size: no draw), and a cropped item is one of the
    windows of the image."""
    from torchvision import transforms
    from tartangan_b200.trainers.trainer import NpzImageDataset
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, size=(6, 16, 16, 3), dtype=np.uint8)
    imgs[0] = np.arange(768, dtype=np.uint32).reshape(16, 16, 3) % 256          # every byte value occurs
    np.savez(tmp_path / 'd.npz', images=imgs)
    ds = NpzImageDataset(str(tmp_path / 'd.npz'), 16)
    ref = transforms.Compose([transforms.ToPILImage(), transforms.RandomCrop(16), transforms.ToTensor(),
                              transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    assert len(ds) == 6
    for i in range(6):
        assert torch.equal(ds[i], ref(imgs[i])), i
    small = NpzImageDataset(str(tmp_path / 'd.npz'), 8)
    full = ref(imgs[2])
    item = small[2]
    assert item.shape == (3, 8, 8)
    assert any(torch.equal(item, full[:, y:y + 8, x:x + 8]) for y in range(9) for x in range(9))
