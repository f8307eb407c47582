"""Secondary bench: device-timed train steps of the other BASELINE.json configs (C2, C4, C5) on ONE GPU at the
per-GPU batch of their 8-GPU runs, with the fused attention kernels on and off.  Not the driver's bench line
(bench.py is); results are kept under profiles/ as evidence for the attention row of SURVEY.md section 8.

    python tools/bench_configs.py [--steps 5] [--warmup 3] [--only cnn256sa]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tartangan_b200 import ops  # noqa: E402
from tartangan_b200.trainers.cnn import CNNTrainer  # noqa: E402
from tartangan_b200.trainers.gan import make_trainer  # noqa: E402
from tartangan_b200.trainers.iqn import IQNTrainer  # noqa: E402
from tartangan_b200.trainers.trainer import tartan_batch  # noqa: E402

CASES = {
    'iqn64': ('C2', IQNTrainer, '64', 64, {}),
    'cnn256sa': ('C4', CNNTrainer, '256sa', 16, {}),
    'iqn512thin': ('C5', IQNTrainer, '512thin', 8, {'num_quantiles': 64}),
    'iqn512sa': ('C5', IQNTrainer, '512', 8, {'num_quantiles': 64, 'attention': '3'}),
}


WORLD = int(os.environ.get('WORLD_SIZE', '1'))
RANK = int(os.environ.get('RANK', '0'))


def barrier():
    if WORLD > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def run(name, steps, warmup, fused, graph=True):
    tag, cls, config, batch, kw = CASES[name]
    ops.state.fused_attention = fused
    torch.manual_seed(0)
    t = make_trainer(cls, config=config, batch_size=batch, precision='bf16', cuda_graph=graph, **kw)
    with torch.no_grad():
        for m in list(t.g.modules()) + list(t.d.modules()):
            if hasattr(m, 'gamma'):
                m.gamma.fill_(0.5)
    size = t.g.max_size
    imgs = tartan_batch(1234 + RANK, batch, size).cuda()
    torch.manual_seed(1000 + RANK)
    for _ in range(warmup):
        out = t.train_batch(imgs, as_floats=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = t.train_batch(imgs, as_floats=False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:                                    # max over ranks
        tm = torch.tensor([ms], device='cuda', dtype=torch.float64)
        torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
        ms = float(tm[0])
    out = {k: (float(v) if v is not None else None) for k, v in out.items()}
    return {'case': name, 'baseline_config': tag, 'config': config, 'size': size, 'batch_per_gpu': batch,
            'n_gpus': WORLD, 'global_batch': batch * WORLD,
            'fused_attention': fused, 'cuda_graph': graph, 'ms_per_step': round(ms, 3), 'images_per_sec': round(batch * WORLD / ms * 1e3, 1),
            'peak_mem_gb': round(torch.cuda.max_memory_allocated() / 2**30, 2), 'losses': out}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--only', default=None)
    ap.add_argument('--eager', action='store_true')
    ap.add_argument('--fused-only', action='store_true')
    a = ap.parse_args()
    if WORLD > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
    for name in ([a.only] if a.only else list(CASES)):
        has_attn = name != 'iqn64'
        for fused in ((True, False) if has_attn and not a.fused_only else (True,)):
            torch.cuda.reset_peak_memory_stats()
            line = run(name, a.steps, a.warmup, fused, not a.eager)
            if RANK == 0:
                print(json.dumps(line), flush=True)
            torch.cuda.empty_cache()
    if WORLD > 1:
        torch.distributed.destroy_process_group()
