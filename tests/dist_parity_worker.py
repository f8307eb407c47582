"""Worker of tests/test_gpu_dist.py (run under `python -m torch.distributed.run --nproc-per-node 2`).

N-rank vs sharded-oracle parity of the data-parallel training step: every rank runs the CUDA
trainer (CUDA-graph execution with the NCCL exchange captured inside the graph, or eager with the hook-driven
exchange) on ITS shard of the global batch with its own z / tau stream; rank 0 then replays the same step on the
CPU oracle the way the data-parallel semantics define it -- each shard is a separate forward / backward with LOCAL
BatchNorm statistics (= the reference at the per-rank batch, trainers/iqn.py:104-147), the parameter gradients are
the average over shards, one Adam step -- and compares losses, averaged gradients and updated parameters.
Also checks that all ranks hold identical parameters after the steps (the exchange really is an all-reduce).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class ShardedOracle:
    """The oracle's train_batch with the batch split into `world` shards (local BN statistics, averaged gradients)."""

    def __init__(self, O, kind, spec, g, tg, d, per_rank_batch, world):
        self.O, self.world = O, world
        self.t = O.OracleTrainer(kind, spec, g, tg, d, per_rank_batch)

    def step(self, shards, seeds):
        O, t, w = self.O, self.t, self.world
        b = t.batch_size
        rng = []
        # ---- D step: one forward / backward per shard, gradients accumulate (scaled by 1 / world)
        t._toggle(t.g, t.g_params, False); t._toggle(t.d, t.d_params, True)
        t.opt_d.zero_grad()
        d_losses, gps = [], []
        bn_buffers = {k: v.clone() for k, v in t.d.items() if 'running' in k or 'num_batches' in k}
        for r in range(w):
            torch.manual_seed(seeds[r])
            for k, v in bn_buffers.items():           # every rank starts from the same buffers; rank 0's are kept
                t.d[k].copy_(v)
            fake = t._fake(b)
            real = shards[r].clone().requires_grad_()
            p_real, l_real = t._d(real, torch.ones(b, 1))
            _, l_fake = t._d(fake.detach(), torch.zeros(b, 1))
            gp = t.grad_penalty * O.r1_penalty(p_real, real)
            loss = l_real + l_fake + gp
            (loss / w).backward()
            d_losses.append(float(loss.detach())); gps.append(float(gp.detach()))
            rng.append(torch.get_rng_state())
            if r == 0:
                keep = {k: t.d[k].clone() for k in bn_buffers}
        grads_d = {k: t.d[k].grad.detach().clone() for k in t.d_params if t.d[k].grad is not None}
        t.opt_d.step()
        for k, v in keep.items():
            t.d[k].copy_(v)
        # ---- G step
        t._toggle(t.g, t.g_params, True); t._toggle(t.d, t.d_params, False)
        t.opt_g.zero_grad()
        g_losses = []
        for r in range(w):
            torch.set_rng_state(rng[r])
            fake = t._fake(b)
            _, g_loss = t._d(fake, torch.ones(b, 1))
            (g_loss / w).backward()
            g_losses.append(float(g_loss.detach()))
        grads_g = {k: t.g[k].grad.detach().clone() for k in t.g_params if t.g[k].grad is not None}
        t.opt_g.step()
        with torch.no_grad():
            for k in t.g_params:
                t.target_g[k].add_((t.g[k] - t.target_g[k]) * t.lr_target_g)
        return d_losses, gps, g_losses, grads_d, grads_g


def _cos(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def main():
    precision, graph = sys.argv[1], sys.argv[2] == 'graph'
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    from oracle import tartan_oracle as O
    from tartangan_b200.models.pluggan import GAN_CONFIGS
    from tartangan_b200.trainers.gan import make_trainer
    from tartangan_b200.trainers.iqn import IQNTrainer
    config, b = '32', 8
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer, config=config, batch_size=b, precision=precision, cuda_graph=graph)
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    orc = ShardedOracle(O, 'iqn', O.SPECS[config], cpu(t.g), cpu(t.target_g), cpu(t.d), b, world) if rank == 0 else None
    tol_loss, tol_cos = (3e-3, 0.9995) if precision == 'fp32' else (5e-2, 0.97)
    ok = True
    for step in range(2):
        shards = [O.tartan_batch(100 + step * world + r, b, 32) for r in range(world)]
        seeds = [500 + step * world + r for r in range(world)]
        torch.manual_seed(seeds[rank])
        got = t.train_batch(shards[rank])
        torch.cuda.synchronize()
        mine = torch.tensor([got['d_loss'], got['gp'], got['g_loss']], device='cuda')
        allm = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        if rank == 0:
            d_l, gp_l, g_l, gd, gg = orc.step(shards, seeds)
            for r in range(world):
                for name, ref, val in (('d_loss', d_l[r], float(allm[r][0])), ('gp', gp_l[r], float(allm[r][1])),
                                       ('g_loss', g_l[r], float(allm[r][2]))):
                    # later steps start from Adam(beta1=0) updated weights: sign noise on near-zero gradients
                    tol = tol_loss if step == 0 else 4 * tol_loss
                    if abs(val - ref) > tol * max(1.0, abs(ref)):
                        print(f'MISMATCH step {step} rank {r} {name}: {val} vs oracle {ref}'); ok = False
            if step == 0:
                for net, mod, ref in (('d', t.d, gd), ('g', t.g, gg)):
                    params = dict(mod.named_parameters())
                    scale = max(float(v.abs().max()) for v in ref.values())
                    worst = 1.0
                    for k, v in ref.items():
                        if k.endswith('.bias') and float(v.abs().max()) < max(2e-3 * scale, 5e-5):
                            continue                      # analytically zero gradients (bias before BatchNorm)
                        c = _cos(params[k].grad, v)
                        worst = min(worst, c)
                        if not c >= (tol_cos if v.numel() >= 64 else tol_cos - 0.07):
                            print(f'MISMATCH averaged gradient {net}.{k}: cosine {c:.5f}'); ok = False
                    print(f'[dist parity] {precision} {"graph" if graph else "eager"} world {world}: {net} averaged gradients, '
                          f'worst cosine {worst:.5f}')
    # every rank holds the same parameters after the exchange-driven updates
    flat = torch.cat([p.detach().reshape(-1) for p in list(t.d.parameters()) + list(t.g.parameters())])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(flat, ref))], device='cuda')
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        if float(same) != 1.0:
            print('MISMATCH: ranks hold different parameters after the steps'); ok = False
        print('DIST_PARITY_OK' if ok else 'DIST_PARITY_FAILED', flush=True)
    dist.barrier()
    from tartangan_b200.parallel import shutdown
    clean = shutdown([t])
    if rank == 0:
        print('clean NCCL shutdown' if clean else 'NCCL shutdown timed out (process exits anyway)', flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
