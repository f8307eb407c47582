import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from conftest import load_golden
from tartangan_b200.models.pluggan import GANConfig
from tartangan_b200.trainers.iqn import IQNTrainer
from tartangan_b200.trainers.gan import make_trainer
g = load_golden('iqn_tiledz')
cfg = GANConfig(base_size=4, latent_dims=g['latent'], data_dims=3, blocks=tuple(g['blocks']), num_blocks_per_scale=1, attention=())
for rep in range(4):
    torch.manual_seed(0)
    t = make_trainer(IQNTrainer, gan_config=cfg, batch_size=g['batch'], precision='fp32', norm=g['norm'], g_base='tiledz')
    t.g.load_state_dict(g['init']['g']); t.target_g.load_state_dict(g['init']['target_g']); t.d.load_state_dict(g['init']['d'])
    if rep >= 2:          # the test's probe forward passes, then restore
        pr = g['probe']
        gsd = {k: v.clone() for k, v in t.g.state_dict().items()}
        dsd = {k: v.clone() for k, v in t.d.state_dict().items()}
        with torch.no_grad():
            t.g(pr['z'].cuda())
            torch.manual_seed(pr['tau_seed'])
            t.d(pr['x'].cuda(), targets=torch.ones(g['batch'], 1, device='cuda'))
        t.g.load_state_dict(gsd); t.d.load_state_dict(dsd)
    torch.manual_seed(g['seeds'][0])
    m = t.train_batch(g['imgs'][0])
    p = dict(t.d.named_parameters())
    ref = g['grads0']['d']
    out = []
    for k in ('blocks.0.convs.0.weight', 'blocks.0.convs.0.bias', 'blocks.0.convs.2.weight', 'blocks.0.convs.3.bias'):
        d = (p[k].grad.cpu() - ref[k]).abs().max() / ref[k].abs().max()
        out.append(f'{k.split("blocks.0.")[1]} rel {float(d):.2e}')
    print(f'rep {rep}: losses {m}  ' + ' | '.join(out), 'beta', p['blocks.0.convs.0.bias'].grad.cpu().tolist(), flush=True)
print('ref beta', ref['blocks.0.convs.0.bias'].tolist())
x = g['imgs'][0]
print('real image channel means', x.mean((0, 2, 3)).tolist(), 'distinct values in channel 2:', x[:, 2].unique().numel())
