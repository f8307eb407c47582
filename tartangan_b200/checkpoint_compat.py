"""Whole-object checkpoints in the reference's format (components/model_checkpoint.py:32-74).

The reference pickles the live objects (``torch.save(self.trainer.g, ...)``) and its loader calls
``torch.load(f).state_dict()`` on them (model_checkpoint.py:66).  A file written here must therefore unpickle in
an environment that has only torch + tartangan, into the REFERENCE's classes:

* our mirror modules (``tartangan_b200.models.pluggan.Generator`` ...) are written under the reference's import
  paths (``tartangan.models.pluggan.Generator`` ...), our kernel-backed leaves (``layers.Conv2d`` ...) as the
  torch.nn classes they derive from;
* only the state every ``nn.Module`` needs travels (parameters, buffers, sub-modules, hook tables, ``training``) plus
  plain scalar attributes; factory partials and packed-weight caches stay behind;
* parameters are compact clones (ours are views into one flat buffer, which torch.save would drag along whole);
* the optimisers are written as genuine ``torch.optim.Adam`` objects carrying our moments.

Loading the other way (a reference-written whole object into this package) needs no code here:
``install_as_tartangan()`` makes the reference's class paths resolve to the mirror classes, and
``Trainer.load_checkpoint`` takes ``.state_dict()`` of whatever it unpickles.
"""
import collections
import contextlib
import copyreg
import functools
import pickle
import sys
import types

import torch
from torch import nn

_PREFIX = 'tartangan_b200.'


def _ref_path(cls):
    """(module, qualname) the reference knows this class by, or None when it has no twin."""
    if cls.__name__ == 'SpectralNormConv2d':
        return None                          # torch's spectral_norm is a hook on nn.Conv2d, not a class: no twin
    mod = cls.__module__
    if mod == _PREFIX + 'models.layers':     # kernel-backed leaves: the torch.nn class they derive from
        for base in cls.__mro__[1:]:
            if base.__module__.startswith('torch.nn.modules') and base is not nn.Module:
                return base.__module__, base.__qualname__
    if mod.startswith(_PREFIX + 'models'):
        return 'tartangan.' + mod[len(_PREFIX):], cls.__qualname__
    return None


class _Namespace(contextlib.AbstractContextManager):
    """While active, `tartangan.*` names resolve to stub classes carrying the reference's (module, qualname), so the
    pickler's "is this global importable" check passes without the reference being installed."""

    def __init__(self):
        self.saved, self.stubs = {}, {}

    def stub(self, module, qualname):
        key = (module, qualname)
        if key not in self.stubs:
            parts = module.split('.')
            for i in range(1, len(parts) + 1):
                name = '.'.join(parts[:i])
                if name not in self.saved:
                    self.saved[name] = sys.modules.get(name)
                    sys.modules[name] = types.ModuleType(name)
            cls = type(qualname, (), {})
            cls.__module__, cls.__qualname__ = module, qualname
            setattr(sys.modules[module], qualname, cls)
            self.stubs[key] = cls
        return self.stubs[key]

    def __exit__(self, *exc):
        for name, old in self.saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
        return False


_PLAIN = (int, float, str, bool, type(None), torch.device, torch.dtype, torch.Size)


def _safe(v):
    """True for values that may travel: scalars, tensors, modules and containers of those (no factories, no classes)."""
    if isinstance(v, _PLAIN) or isinstance(v, (torch.Tensor, nn.Module)):
        return True
    if isinstance(v, (tuple, list, set, frozenset)):
        return all(_safe(i) for i in v)
    if isinstance(v, dict):
        return all(isinstance(k, _PLAIN) and _safe(i) for k, i in v.items())
    return False


def _module_state(m):
    from .models import layers
    state = {}
    for k, v in m.__dict__.items():
        if isinstance(v, functools.partial) and v.func is layers.interpolate:
            # the D blocks' skip resampler (discriminator.py:57): the reference holds a partial of F.interpolate
            state[k] = functools.partial(torch.nn.functional.interpolate, *v.args, **v.keywords)
        elif not k.startswith('_ttg') and _safe(v):
            state[k] = v
    return state


def _make_pickler(ns):
    class RefPickler(pickle.Pickler):
        def reducer_override(self, obj):
            if isinstance(obj, nn.Parameter):
                return torch._utils._rebuild_parameter, (obj.data.clone(), obj.requires_grad, collections.OrderedDict())
            if isinstance(obj, nn.Module) and type(obj).__module__.startswith(_PREFIX):
                path = _ref_path(type(obj))
                if path is None:
                    raise NotImplementedError(f'{type(obj).__name__} has no twin in the reference: save with '
                                              '--checkpoint-format state_dict')
                cls = getattr(sys.modules.get(path[0]), path[1], None) if path[0].startswith('torch.') else None
                if cls is None:
                    cls = ns.stub(*path)
                return copyreg._reconstructor, (cls, object, None), _module_state(obj)
            if isinstance(obj, tuple) and type(obj).__module__.startswith(_PREFIX) and hasattr(obj, '_fields'):
                return ns.stub('tartangan.' + type(obj).__module__[len(_PREFIX):], type(obj).__qualname__), tuple(obj)
            return NotImplemented
    return RefPickler


def save_reference_object(obj, path):
    """torch.save(obj, path) such that the unmodified reference unpickles it into its own classes."""
    if isinstance(obj, torch.optim.Optimizer):
        obj = as_torch_adam(obj)
    with _Namespace() as ns:
        shim = types.ModuleType('ttg_ref_pickle')
        shim.Pickler = _make_pickler(ns)
        for name in ('dumps', 'dump', 'loads', 'load', 'Unpickler', 'HIGHEST_PROTOCOL', 'DEFAULT_PROTOCOL',
                     'PicklingError', 'UnpicklingError'):
            setattr(shim, name, getattr(pickle, name))
        shim.__name__ = 'pickle'
        torch.save(obj, path, pickle_module=shim)


def as_torch_adam(opt):
    """A torch.optim.Adam over compact clones of the optimised parameters with this optimiser's state
    (the object the reference pickles as opt_g.pt / opt_d.pt)."""
    g = opt.param_groups[0]
    params = [nn.Parameter(p.data.clone(), requires_grad=p.requires_grad) for p in g['params']]
    ref = torch.optim.Adam(params, lr=g['lr'], betas=tuple(g['betas']), eps=g['eps'])
    sd = opt.state_dict()
    sd = {'state': {k: {n: (v.detach().clone() if torch.is_tensor(v) else v) for n, v in st.items()}
                    for k, st in sd['state'].items()},
          'param_groups': sd['param_groups']}
    ref.load_state_dict(sd)
    return ref
