"""Same-box LIBRARY bar: the reference's training step composed from stock PyTorch ops — the CPU
oracle's code (oracle/tartan_oracle.py, pinned to the reference) with its tensors on cuda:0, i.e. cuDNN convolutions,
ATen batch-norm / elementwise kernels, autograd double backward — timed with CUDA events at the headline workload
(config '128', batch 256, R1 5.0, 8 quantiles).  This is what a user of the reference gets on one B200 without this
repo; bench.py's number is to be read against it (and against the CPU number).

    python tools/bench_eager_gpu.py [--batch 256] [--steps 10] [--warmup 3] [--bf16] [--channels-last]

--bf16: torch.autocast(bfloat16) around the step (the library's mixed-precision path); default fp32 (TF32 off, as stock).
Prints one JSON line.  Measurement tool: not imported by the product package."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='128')
    ap.add_argument('--bf16', action='store_true')
    ap.add_argument('--tf32', action='store_true')
    args = ap.parse_args()
    from oracle import tartan_oracle as O
    from tartangan_b200.models import pluggan
    from tartangan_b200.models.blocks import (GeneratorInputMLP, GeneratorOutput, IQNDiscriminatorOutput,
                                               ResidualDiscriminatorBlock, ResidualGeneratorBlock)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = args.tf32
    torch.manual_seed(0)
    cfg = pluggan.GAN_CONFIGS[args.config]
    mk = lambda: pluggan.Generator(cfg, input_factory=GeneratorInputMLP, block_factory=ResidualGeneratorBlock,
                                   output_factory=GeneratorOutput)
    g, tg = mk(), mk()
    d = pluggan.IQNDiscriminator(cfg, block_factory=ResidualDiscriminatorBlock, output_factory=IQNDiscriminatorOutput)
    tr = O.OracleTrainer('iqn', O.SPECS[args.config], g.state_dict(), tg.state_dict(), d.state_dict(), args.batch,
                         device='cuda')
    size = cfg.base_size * 2 ** len(cfg.blocks)
    imgs = O.tartan_batch(1234, args.batch, size).cuda()

    def step():
        if args.bf16:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                return tr.train_batch(imgs)
        return tr.train_batch(imgs)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({'impl': 'stock PyTorch eager on cuda (cuDNN / ATen), oracle composition',
                      'metric': 'SA-GAN-IQN G+D train images/sec at 128x128', 'value': args.batch / ms * 1e3,
                      'unit': 'images/sec', 'ms_per_step': ms, 'batch': args.batch, 'steps': args.steps,
                      'dtype': 'bf16 autocast' if args.bf16 else ('tf32' if args.tf32 else 'fp32'),
                      'config': args.config, 'last_metrics': m, 'torch': torch.__version__,
                      'cudnn': torch.backends.cudnn.version()}), flush=True)


if __name__ == '__main__':
    main()
