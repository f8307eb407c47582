import sys, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from conftest import load_golden
import torch.nn.functional as F
import tartangan_b200 as tb
from tartangan_b200 import ops
tb.set_precision('fp32')
x = torch.randn(2, 16, 8, 8)
y = ops.leaky_relu(ops.to_internal(x.cuda()), 0.2)
print('lrelu4d', float((y.cpu() - F.leaky_relu(x, 0.2)).abs().max()))
g = load_golden('iqn_nonorm')
from test_gpu_train import _trainer, _cfg
t = _trainer('iqn', _cfg(g), g['batch'], 'fp32', norm='id')
t.g.load_state_dict(g['init']['g'])
from oracle import tartan_oracle as O
spec = O.Spec(4, g['latent'], 3, tuple(g['blocks']), ())
sd = {k: v.clone() for k, v in g['init']['g'].items()}
z = g['probe']['z']
with torch.no_grad():
    xo = F.leaky_relu(F.linear(z, sd['blocks.0.base_img.0.weight'], sd['blocks.0.base_img.0.bias']), 0.2).view(-1, 16, 4, 4)
    xm = t.g.blocks[0](z.cuda())
    print('mlp', float((xm.cpu() - xo).abs().max()))
    for i in (1, 2):
        xo = O._g_block(sd, f'blocks.{i}', xo, i == 1, 'id')
        xm = t.g.blocks[i](xm)
        print('block', i, float((xm.cpu() - xo).abs().max()), float(xo.abs().max()))
    print(t.g.blocks[1])
with torch.no_grad():
    full_o = O.generator(sd, spec, z, 'id')
    full_m = t.g(z.cuda())
    print('full', float((full_m.cpu() - full_o).abs().max()), float((full_o - g['probe']['g_out']).abs().max()))
    k = 'blocks.3'
    a = F.leaky_relu(xo, 0.2)
    am = ops.leaky_relu(xm, 0.2)
    print('act', float((am.cpu() - a).abs().max()))
    c = F.conv2d(a, sd[k + '.convs.2.weight'], sd[k + '.convs.2.bias'])
    cm = t.g.blocks[3].convs[2](am, out_dtype=torch.float32)
    print('conv', float((cm.cpu() - c).abs().max()), cm.shape, cm.stride())
    tm = ops.tanh(cm)
    print('tanh', float((tm.cpu() - torch.tanh(c)).abs().max()))
    fm = ops.from_internal(tm)
    print('from', float((fm.cpu() - torch.tanh(c)).abs().max()))
    print('out block', float((t.g.blocks[3](xm).cpu() - torch.tanh(c)).abs().max()))
