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
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    sec = _oracle_step_time(batch, steps, warmup, threads)
    val = batch / sec
    line = {
        'impl': 'reference', 'metric': 'SA-GAN-IQN G+D train images/sec at 128x128', 'value': val,
        'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"trainers.iqn SA-GAN-IQN config '{CONFIG}' 128x128, CPU sample batch {batch}"},
        'cpu_baseline': {'value': val, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
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
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager', action='store_true', help='disable CUDA-graph execution of the step')
    ap.add_argument('--profile-out', default=None, help='write the per-kernel event timing table here (JSON)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
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
    base = tartan_batch(1234 + rank, 32, SIZE)
    host = base.repeat((args.batch + 31) // 32, 1, 1, 1)[:args.batch].contiguous().pin_memory()
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
        # dominant kernel = the (entry point, shape) with the largest share of device time
        fam = by_name[0]
        cands = [r for r in rows if r['name'] == fam['name'] and r['bytes'] > 0] or [fam]
        top = max(cands, key=lambda r: r['ms'])
        per_launch_ms = top['ms'] / max(top['calls'], 1)
        ai = top['flops'] / max(top['bytes'], 1)
        if ai > (tflops * 1e12) / (hbm * 1e9):
            achieved = top['flops'] / (top['ms'] / 1e3) / 1e12
            roof = dict(bound='tensor', achieved=achieved, peak=tflops, unit='TFLOP/s', frac=achieved / tflops)
        else:
            achieved = top['bytes'] / (top['ms'] / 1e3) / 1e9
            roof = dict(bound='hbm', achieved=achieved, peak=hbm, unit='GB/s', frac=achieved / hbm)
        # DRAM traffic per launch from the committed `ncu --set full` capture of this kernel+shape (profiles/)
        ncu_traffic = {('ttg_conv2d_tc', 'N256 128x128 16->16 k3 up0'): 229.6e6,       # profiles/r1_ncu_f_conv16.txt
                       ('ttg_bn_act_bwd', 'M4194304 C16'): 637.6e6}                      # profiles/r1_ncu_f_bn_bwd.txt
        roof.update(traffic=ncu_traffic.get((top['name'], top.get('key', ''))), kernel=top['name'], shape=top.get('key', ''),
                    algorithmic_bytes_per_launch=top['bytes'] / max(top['calls'], 1), us_per_launch=per_launch_ms * 1e3,
                    peak_source=which, launches_per_step=top['calls'],
                    share_of_step=top['ms'] / max(step_ms_prof, 1e-9),
                    family_share_of_step=fam['ms'] / max(step_ms_prof, 1e-9),
                    step_frac_of_mixed_roofline=(value / world) / ROOFLINE_IMG_S,
                    step_hbm_gbs=(value / world) * MB_PER_IMG / 1e3, step_tflops=(value / world) * GFLOP_PER_IMG / 1e3)
        line = {
            'metric': 'SA-GAN-IQN G+D train images/sec at 128x128', 'value': value, 'unit': 'images/sec',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision,
            'data': 'synthetic',
            'config': {'workload': f"trainers.iqn SA-GAN-IQN config '{CONFIG}' 128x128, batch {args.batch}/GPU, "
                                   f"R1 penalty 5.0, 8 quantiles", 'global_batch': args.batch * world,
                       'parallelism': f'dp{world}',
                       'l2': 'per-step working set (activations of batch 256 at 128x128, >2 GB) exceeds the 126 MB L2'},
            'e2e': {'value': e2e_val, 'unit': 'images/sec', 'h2d_bytes_per_step':
                    (host.numel() * 4 + 2 * args.batch * 256 * 4 + 3 * args.batch * 8 * 4) * world,
                    'd2h_bytes_per_step': 12 * world},
            'gpu_launches': launches, 'host_issue_ms_per_step': host_issue_ms,
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
