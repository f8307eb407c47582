"""IQN head — interface of tartangan/models/iqn.py (CosineQuantileEmbedding, IQN, iqn_loss).

The modules hold the parameters under the reference's names; the arithmetic
(cos(tau*pi*k) -> Linear -> tanh -> multiply -> Linear) runs as one fused kernel
from IQNDiscriminatorOutput (ops.IqnHeadFn).
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from .layers import Linear


class CosineQuantileEmbedding(nn.Module):
    """iqn.py:27-46.  `activation` must be tanh (the fused kernel's embedding activation)."""

    def __init__(self, state_dims, embedding_dims=64, activation=nn.Tanh, norm_factory=nn.BatchNorm1d):
        super().__init__()
        if activation not in (nn.Tanh,) and getattr(activation, '__name__', '') != 'Tanh':
            raise NotImplementedError('CosineQuantileEmbedding: only the tanh activation has a kernel')
        self.embedding_dims = embedding_dims
        self.to_state = nn.Sequential(Linear(embedding_dims, state_dims), nn.Tanh())
        self.register_buffer('embedding_range', torch.arange(1, embedding_dims + 1).float())

    def forward(self, quantiles):
        """Un-fused embedding (rows x state_dims); the training path never materialises this."""
        q = quantiles.float().reshape(-1).to(self.embedding_range.device)
        eye = torch.eye(self.to_state[0].out_features, device=q.device)
        ones = torch.ones(self.to_state[0].out_features, device=q.device)
        rows = [ops.IqnHeadFn.apply(eye[i:i + 1].expand(q.numel(), -1).contiguous(), q,
                                    self.to_state[0].weight, self.to_state[0].bias, ones, None, 1)
                for i in range(eye.shape[0])]
        return torch.stack(rows, dim=1)


class IQN(nn.Module):
    """iqn.py:76-108.  num_quantiles=8, quantile_dims=20, mix='mult'."""

    def __init__(self, feature_dims, quantile_dims=20, num_quantiles=8, mix='mult',
                 quantile_embedding_factory=CosineQuantileEmbedding, norm_factory=nn.BatchNorm1d):
        super().__init__()
        if not mix.startswith('mult'):
            raise NotImplementedError("IQN: only mix='mult' (the reference default) has a kernel")
        self.quantile_embedding = quantile_embedding_factory(feature_dims, quantile_dims, norm_factory=norm_factory)
        self.feature_dims = feature_dims
        self.num_quantiles = num_quantiles
        self.mix = mix
        self._device = None
        self.tau_source = None      # optional callable(n_rows) -> device tensor (static buffers for CUDA graphs)

    def forward(self, x):
        raise NotImplementedError('IQN.forward (materialised x*embedding) is fused into IQNDiscriminatorOutput; '
                                  'call that module instead')

    def sample_quantiles(self, n=1):
        """tau ~ U[0,1) drawn from the CPU generator, then copied (iqn.py:105-108)."""
        if self.tau_source is not None:
            return self.tau_source(n * self.num_quantiles)
        if self._device is None:
            self._device = next(self.parameters()).device
        return torch.rand(n * self.num_quantiles, 1).to(self._device)


def iqn_loss(preds, target, taus, k=1.):
    """iqn.py:111-130: quantile-Huber loss, rows quantile-major, sum over quantiles, mean over batch."""
    assert not target.requires_grad
    batch = target.shape[0]
    if target.numel() != batch:
        raise NotImplementedError('iqn_loss: only output_dims == 1 is implemented')
    nq = preds.shape[0] // batch
    return ops.QuantileHuberFn.apply(preds.reshape(-1), target.reshape(-1).float(), taus.reshape(-1), nq, float(k))
