"""Interface of tartangan/trainers/utils.py."""
import torch


def set_device_from_args(args):
    """trainers/utils.py:5-11.  There is no CPU path: --no-cuda (or a box without a GPU) is an error."""
    if getattr(args, 'no_cuda', False) or not torch.cuda.is_available():
        raise RuntimeError('tartangan_b200 runs on CUDA (sm_100a) only: --no-cuda / no visible GPU is not supported')
    setattr(args, 'device', 'cuda')


def toggle_grad(model, on_or_off):
    for param in model.parameters():
        param.requires_grad_(on_or_off)
