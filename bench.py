"""Headline benchmark: SA-GAN-IQN G+D training images/sec at 128x128 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one IQNTrainer.train_batch (D step with R1 penalty + G step + EMA) on one batch
of 256 synthetic tartan-shaped 128x128 RGB images per GPU (weak scaling, config '128').
`value`  : device-timed, inputs resident in HBM (images; z / tau staged per step like the
           reference does, trainers/trainer.py:153-156).
`e2e`    : the same step through the public trainer API with HOST batches: pinned H2D copy of
           the images and a D2H read of the three loss floats inside the timed region.
`--impl reference`: the reference's CPU implementation of the step (the oracle port under
           oracle/, torch CPU, all host threads) on a bounded batch of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if '--impl=reference' in sys.argv or ('--impl' in sys.argv and 'reference' in sys.argv):
    # The CPU arm is to use every host core.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers when the
    # variable is unset (the driver launches both arms through it for N > 1); torch.set_num_threads() after the OpenMP
    # runtime has started does not undo that for the autograd worker threads (measured here: 4.6 instead of 20-27
    # images/s on 8 cores).  Only rank 0 computes on this arm, so it takes all cores; set before torch is imported.
    os.environ['OMP_NUM_THREADS'] = os.environ['MKL_NUM_THREADS'] = str(os.cpu_count() or 1)

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = '128'
BATCH = 256
SIZE = 128
# algorithmic work per real image per train step: 4 F_G + 11 F_D, bf16 activations
GFLOP_PER_IMG = 8.862
MB_PER_IMG = 73.6
ROOFLINE_IMG_S = 78290.0


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return p['hbm_gbs'], p.get('bf16_tflops_sustained', p['bf16_tflops']), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [s.strip() for s in out.split(',')]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def _roof_time_ms(r, hbm, tflops):
    """Roofline time of a profiler row: max(algorithmic bytes / HBM peak, flops / tensor peak), in ms."""
    return max(r['bytes'] / (hbm * 1e9), r['flops'] / (tflops * 1e12)) * 1e3


def _roofline(rows, by_name, step_ms_events, step_ms, img_s_per_gpu, hbm, tflops, which):
    """The `roofline` object.  Leads with the STEP (images/s against the mixed roofline of SURVEY.md 8d), then the
    dominant kernel FAMILY (the entry point with the largest share of device time) as a TIME-WEIGHTED fraction over
    all its shapes: frac = sum(roofline time of each launch) / sum(measured time of each launch), where the
    roofline time of a launch is max(algorithmic bytes / HBM peak, flops / tensor peak).  `achieved` is the family's
    algorithmic bytes (or flops) divided by its measured time.  Every family with a cost model is listed the same
    way under `families`; `top_shape` is the single (entry point, shape) with the most device time.
    Per-launch durations: CUDA events around each launch of one EAGER step on the launching stream; the graphed step
    that `value` reports is shorter than their sum (no launch gaps), which `events_vs_graph` states."""
    fams = []
    for f in by_name:
        frows = [r for r in rows if r['name'] == f['name']]
        modelled = [r for r in frows if r['bytes'] > 0]
        t_meas = sum(r['ms'] for r in modelled)
        entry = dict(name=f['name'], calls=f['calls'], ms=f['ms'], share_of_step=f['ms'] / max(step_ms_events, 1e-9))
        if modelled and t_meas > 0:
            t_roof = sum(_roof_time_ms(r, hbm, tflops) for r in modelled)
            nbytes, flops = sum(r['bytes'] for r in modelled), sum(r['flops'] for r in modelled)
            bound = 'tensor' if flops / (tflops * 1e12) > nbytes / (hbm * 1e9) else 'hbm'
            entry.update(frac=t_roof / t_meas, bound=bound, gbs=nbytes / t_meas / 1e6, tflops=flops / t_meas / 1e9)
        fams.append(entry)
    lead = next((f for f in fams if 'frac' in f), None)
    top = max((r for r in rows if r['bytes'] > 0), key=lambda r: r['ms'], default=None)
    roof = dict(step_frac_of_mixed_roofline=img_s_per_gpu / ROOFLINE_IMG_S,
                step_hbm_gbs=img_s_per_gpu * MB_PER_IMG / 1e3, step_tflops=img_s_per_gpu * GFLOP_PER_IMG / 1e3)
    if lead is not None:
        peak, unit = (tflops, 'TFLOP/s') if lead['bound'] == 'tensor' else (hbm, 'GB/s')
        roof.update(bound=lead['bound'], achieved=lead['tflops'] if lead['bound'] == 'tensor' else lead['gbs'],
                    peak=peak, unit=unit, frac=lead['frac'],
                    traffic=None,        # DRAM bytes need ncu; the per-kernel captures are under profiles/ (r2_ncu_*.txt)
                    kernel=lead['name'], share_of_step=lead['share_of_step'], launches_per_step=lead['calls'],
                    frac_definition='time-weighted over every shape of the family: sum(max(bytes/HBM, flops/tensor '
                                    'peak)) / sum(measured time); achieved = family bytes (or flops) / family time')
    roof.update(peak_source=which, peak_hbm_gbs=hbm, peak_bf16_tflops=tflops,
                events_vs_graph=dict(sum_of_eager_launch_events_ms=step_ms_events, graphed_step_ms=step_ms),
                families=[{k: (round(v, 4) if isinstance(v, float) else v) for k, v in f.items()} for f in fams[:10]])
    if top is not None:
        per = top['ms'] / max(top['calls'], 1)
        roof['top_shape'] = dict(kernel=top['name'], shape=top.get('key', ''), launches_per_step=top['calls'],
                                 us_per_launch=per * 1e3, share_of_step=top['ms'] / max(step_ms_events, 1e-9),
                                 algorithmic_bytes_per_launch=top['bytes'] / max(top['calls'], 1),
                                 gbs=top['bytes'] / top['ms'] / 1e6, tflops=top['flops'] / top['ms'] / 1e9,
                                 frac=_roof_time_ms(top, hbm, tflops) / top['ms'])
    return roof


def _oracle_step_time(batch, steps, warmup, threads):
    from oracle import tartan_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    spec = O.SPECS[CONFIG]
    # random-init weights of the architecture, drawn with the same torch initialisers
    from tartangan_b200.models import pluggan
    import functools
    from tartangan_b200.models.blocks import (GeneratorInputMLP, GeneratorOutput, IQNDiscriminatorOutput,
                                               ResidualDiscriminatorBlock, ResidualGeneratorBlock)
    cfg = pluggan.GAN_CONFIGS[CONFIG]
    g = pluggan.Generator(cfg, input_factory=GeneratorInputMLP, block_factory=ResidualGeneratorBlock,
                          output_factory=GeneratorOutput)
    tg = pluggan.Generator(cfg, input_factory=GeneratorInputMLP, block_factory=ResidualGeneratorBlock,
                           output_factory=GeneratorOutput)
    d = pluggan.IQNDiscriminator(cfg, block_factory=ResidualDiscriminatorBlock, output_factory=IQNDiscriminatorOutput)
    tr = O.OracleTrainer('iqn', spec, g.state_dict(), tg.state_dict(), d.state_dict(), batch)
    imgs = O.tartan_batch(1234, batch, SIZE)
    for _ in range(warmup):
        tr.train_batch(imgs)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_batch(imgs)
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    """Reference arm: the CPU implementation of the step (oracle port) with all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 8
    # bounded sample: a CPU step of batch 8 takes ~0.15 s per image, so --steps / --warmup are clamped to 5 / 2 and
    # the batch to 8 images (the line says so: `steps`, `warmup`, `config.note`)
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    sec = _oracle_step_time(batch, steps, warmup, threads)
    val = batch / sec
    line = {
        'impl': 'reference', 'metric': 'SA-GAN-IQN G+D train images/sec at 128x128', 'value': val,
        'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"trainers.iqn SA-GAN-IQN config '{CONFIG}' 128x128, CPU sample batch {batch}",
                   'note': f'CPU arm: oracle port (torch CPU fp32, {threads} threads); requested --steps {args.steps} '
                           f'--warmup {args.warmup} clamped to {steps} / {warmup} and the batch to {batch} images so the '
                           'run stays within minutes; images/sec is per image, so the batch size does not scale it'},
        'cpu_baseline': {'value': val, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
                         'omp_num_threads': os.environ.get('OMP_NUM_THREADS'),
                         'sample': f'{steps} train steps of batch {batch} (of the 256/GPU workload), torch CPU fp32'},
        'e2e': {'value': val, 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=('b200', 'reference'))
    ap.add_argument('--batch', type=int, default=BATCH, help='per-GPU batch (default: the BASELINE workload)')
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--strong', action='store_true',
                    help='strong scaling (secondary row, SURVEY 8d): the global batch stays at --batch, each rank takes 1/N')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager', action='store_true', help='disable CUDA-graph execution of the step')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel event timing table here (JSON)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.strong:
        if args.batch % world:
            raise SystemExit(f'--strong: global batch {args.batch} is not divisible by {world} ranks')
        args.batch //= world                 # per-rank batch; everything below is per rank
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from tartangan_b200 import _lib
    from tartangan_b200.profiler import KernelProfiler
    from tartangan_b200.trainers.gan import make_trainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    from tartangan_b200.trainers.trainer import tartan_batch

    torch.manual_seed(0)                     # model init: same on every rank
    trainer = make_trainer(IQNTrainer, config=CONFIG, batch_size=args.batch, precision=args.precision,
                           cuda_graph=not args.eager)
    torch.manual_seed(1000 + rank)           # z / tau stream
    host = tartan_batch(1234 + rank, args.batch, SIZE).contiguous().pin_memory()      # `batch` DISTINCT images
    dev = host.cuda()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-timed throughput (inputs resident in HBM)
    for _ in range(args.warmup):
        trainer.train_batch(dev, as_floats=False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    k0 = _lib.Counters.kernels
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        trainer.train_batch(dev, as_floats=False)
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # CPU time to enqueue one step
    e1.record()
    barrier()
    launches = _lib.Counters.kernels - k0
    ms = e0.elapsed_time(e1)
    # ---- end-to-end: host batch in, three loss floats out, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m = trainer.train_batch(host, as_floats=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    import math
    if not all(math.isfinite(float(v)) for v in m.values()):        # a fast step that computes garbage is not a result
        raise SystemExit(f'bench: non-finite losses after the timed steps: {m}')
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        t = torch.tensor([ms, e2e_s], device='cuda', dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    # ---- per-kernel device timing of one more step (CUDA events on the launching stream)
    trainer.args.cuda_graph = False          # the per-kernel breakdown needs eager launches
    trainer.train_batch(dev, as_floats=False)
    with KernelProfiler() as prof:
        trainer.train_batch(dev, as_floats=False)
    rows, by_name = prof.summary(), prof.by_name()
    step_ms_prof = sum(r['ms'] for r in rows)

    if rank == 0:
        hbm, tflops, which = _peaks()
        value = args.batch * world * args.steps / (ms / 1e3)
        e2e_val = args.batch * world * args.steps / e2e_s
        roof = _roofline(rows, by_name, step_ms_prof, ms / args.steps, value / world, hbm, tflops, which)
        line = {
            'metric': 'SA-GAN-IQN G+D train images/sec at 128x128', 'value': value, 'unit': 'images/sec',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None,
            'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': f"trainers.iqn SA-GAN-IQN config '{CONFIG}' 128x128, batch {args.batch}/GPU, "
                                   f"R1 penalty 5.0, 8 quantiles", 'global_batch': args.batch * world,
                       'parallelism': f'dp{world}',
                       'l2': 'per-step working set (activations of batch 256 at 128x128, >2 GB) exceeds the 126 MB L2'},
            'e2e': {'value': e2e_val, 'unit': 'images/sec', 'h2d_bytes_per_step':
                    (host.numel() * 4 + 2 * args.batch * 256 * 4 + 3 * args.batch * 8 * 4) * world,
                    'd2h_bytes_per_step': 12 * world},
            'gpu_launches': launches, 'host_issue_ms_per_step': host_issue_ms,
            'last_losses': {k: float(v) for k, v in m.items()},
            'clocks': sampler.result(),
            'roofline': roof,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sec = _oracle_step_time(8, 2, 1, threads)
            line['cpu_baseline'] = {'value': 8 / sec, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
                                    'sample': '2 train steps of batch 8 (of the 256/GPU workload), torch CPU fp32 oracle'}
        if args.profile_out:
            with open(args.profile_out, 'w') as f:
                json.dump({'by_name': by_name, 'rows': rows[:60], 'step_ms_events': step_ms_prof,
                           'ms_per_step': ms / args.steps}, f, indent=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # the step graphs hold captured NCCL collectives: release them before the communicator, never hang on teardown
        from tartangan_b200.parallel import shutdown
        sys.stdout.flush()
        torch.distributed.barrier()
        clean = shutdown([trainer])
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
