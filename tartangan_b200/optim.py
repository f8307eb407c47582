"""Flat fused Adam (+EMA) — replaces torch.optim.Adam(betas=(0, 0.999)) and the per-parameter
EMA loop of update_target_generator (reference trainers/cnn.py:84-85,158-165).

Parameters, gradients and moments of one model live in contiguous fp32 buffers; every
``p.data`` / ``p.grad`` / state tensor is a view into them, so one kernel launch updates the
whole model and (for data parallel) one NCCL call reduces all gradients.  The optimiser
state keeps torch.optim.Adam's keys (step, exp_avg, exp_avg_sq) so opt_g.pt / opt_d.pt
checkpoints keep their layout.
"""
import torch

from ._lib import call, ptr


class FlatParams:
    """Re-homes the parameters of a module into one flat buffer (views keep names and shapes)."""

    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        # 4-element alignment per tensor keeps every view 16-byte aligned for vector access
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.data[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None
        self._grad_views = [self.grad[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def intact(self):
        p, o = self.params[0], self.offsets[0]
        return p.data_ptr() == self.data.data_ptr() + 4 * o

    def attach_grads(self):
        """Zero the flat gradient buffer and point every p.grad into it (autograd accumulates in place)."""
        self.grad.zero_()
        for p, g in zip(self.params, self._grad_views):
            p.grad = g


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError('FusedAdam: weight_decay / amsgrad are not used by the reference trainers')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        if len(self.param_groups) != 1:
            raise NotImplementedError('FusedAdam: a single parameter group is supported')
        self.flat = None
        self.ema_target = None     # FlatParams of the target generator (fused EMA)
        self.ema_lr = 0.0

    def _ensure_flat(self):
        if self.flat is None or not self.flat.intact():
            params = self.param_groups[0]['params']
            old = {id(p): self.state.get(p) for p in params}
            self.flat = FlatParams(params)
            n, dev = self.flat.numel, self.flat.data.device
            self._m = torch.zeros(n, dtype=torch.float32, device=dev)
            self._v = torch.zeros(n, dtype=torch.float32, device=dev)
            self._step = torch.zeros(1, dtype=torch.float32, device=dev)
            for p, o in zip(self.flat.params, self.flat.offsets):
                st = old.get(id(p)) or {}
                m = self._m[o:o + p.numel()].view(p.shape)
                v = self._v[o:o + p.numel()].view(p.shape)
                if 'exp_avg' in st:
                    m.copy_(st['exp_avg'])
                    v.copy_(st['exp_avg_sq'])
                    self._step.fill_(float(st['step']))
                self.state[p] = {'step': self._step[0], 'exp_avg': m, 'exp_avg_sq': v}
        return self.flat

    def zero_grad(self, set_to_none=False):
        self._ensure_flat().attach_grads()

    @torch.no_grad()
    def step(self, closure=None):
        flat = self._ensure_flat()
        g = self.param_groups[0]
        ema = None
        if self.ema_target is not None:
            if self.ema_target.numel != flat.numel:
                raise RuntimeError('FusedAdam: EMA target layout differs from the optimised model')
            ema = self.ema_target.data
        call('ttg_adam_flat', ptr(flat.data), ptr(flat.grad), ptr(self._m), ptr(self._v), ptr(ema), flat.numel,
             g['lr'], g['betas'][0], g['betas'][1], g['eps'], self.ema_lr, ptr(self._step))
        for p in flat.params:       # invalidate packed-weight caches (ops._packed)
            p._ttg_epoch = getattr(p, '_ttg_epoch', 0) + 1
        if self.ema_target is not None:      # the fused EMA rewrote the target model's parameters as well
            for p in self.ema_target.params:
                p._ttg_epoch = getattr(p, '_ttg_epoch', 0) + 1
        self.repack()

    def repack(self):
        """Refresh the packed conv-filter images of the optimised model in one launch (ops.PackedModel)."""
        from . import ops
        flat = self._ensure_flat()
        pm = getattr(self, '_packed_model', None)
        if pm is None or pm.flat is not flat:
            pm = self._packed_model = ops.PackedModel(flat)
        pm.repack()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.flat = None            # re-flatten, copying the loaded moments in
        self._ensure_flat()
