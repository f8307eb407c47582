"""Losses — interface of tartangan/models/losses.py."""
import torch

from .. import ops


def discriminator_hinge_loss(real, fake):
    raise NotImplementedError('hinge losses are assigned but never called by the reference trainers '
                              '(trainers/cnn.py:86-87,125-126); no kernel is provided')


def generator_hinge_loss(fake):
    raise NotImplementedError('hinge losses are assigned but never called by the reference trainers')


def gradient_penalty(preds, data):
    """R1 penalty (losses.py:17-30): mean over the batch of |d sum(preds) / d data|^2, differentiable.

    The inner backward runs with parameter gradients switched off (only d/d data is needed);
    the graph it records is what the outer d_loss.backward() differentiates again."""
    batch_size = data.size(0)
    ones = torch.ones_like(preds)
    with ops.inputs_only_grads():
        grad_dout, = torch.autograd.grad(outputs=preds, inputs=data, grad_outputs=ones,
                                         create_graph=True, retain_graph=True, only_inputs=True)
    assert grad_dout.size() == data.size()
    return ops.SqsumFn.apply(grad_dout, 1.0 / batch_size)
