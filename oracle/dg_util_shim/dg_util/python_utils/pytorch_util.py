"""Shim of dg_util.python_utils.pytorch_util (see package docstring)."""
import numpy as np
import torch
from torch import nn


class BaseModel(nn.Module):
    # vince_model.py:57,62,68 read self.device inside __init__, before .to()
    def __init__(self):
        super().__init__()
        self._device = "cpu"
        self.saves = 0

    @property
    def device(self):
        return self._device

    def to(self, device):
        self._device = device
        return super().to(device)

    def restore(self, checkpoint_dir, saved_variable_prefix=None, new_variable_prefix=None, skip_filter=None):
        return 0


def save(model, path, num_to_keep, iteration):
    return None


class _SingleDevice(nn.Module):
    """`.module`-bearing pass-through (same state_dict keys as nn.DataParallel)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def get_data_parallel(module, gpu_ids):
    # callers use `.module` (end_task_base_solver.py:154).  gpu_ids = ["cpu"] (the CPU arm / golden generation) must
    # stay on the CPU even on a box with GPUs: nn.DataParallel would move the module to cuda:0 there.
    ids = [g for g in (gpu_ids or []) if str(g) != "cpu"]
    if not ids or not torch.cuda.is_available():
        return _SingleDevice(module)
    return nn.DataParallel(module, [torch.device(g).index if not isinstance(g, int) else g for g in ids])


class RemoveDim(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return remove_dim(x, self.dim)


class AttentionPool2D(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("dg_util shim: --use-attention is not on the hot path")


def split_dim(x, dim, d1, d2):
    # vince_model.py:147-151: [N,C,H,W] --(2,3,H/3)--> [N,C,3,H/3,W]
    shape = list(x.shape)
    if dim < 0:
        dim += len(shape)
    new_shape = shape[:dim] + [d1, d2] + shape[dim + 1:]
    return x.reshape(new_shape)


def remove_dim(x, dim):
    # vince_model.py:153-155: [N,3,3,C,h,w],(1,2) -> [9N,C,h,w]; each listed dim merges into its predecessor
    if isinstance(dim, (tuple, list)):
        for d in sorted(dim, reverse=True):
            x = remove_dim(x, d)
        return x
    shape = list(x.shape)
    if dim < 0:
        dim += len(shape)
    new_shape = shape[: dim - 1] + [shape[dim - 1] * shape[dim]] + shape[dim + 1:]
    return x.reshape(new_shape)


def expand_new_dim(x, dim, size):
    # vince_model.py:168: arange(N) -> [N,9]
    x = x.unsqueeze(dim)
    shape = [-1] * x.dim()
    shape[dim] = size
    return x.expand(*shape)


def from_numpy(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)
    return x


def to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return x


def stack_dicts_in_list(list_of_dicts, axis=0, concat=False):
    out = {}
    for key in list_of_dicts[0]:
        vals = [d[key] for d in list_of_dicts]
        if isinstance(vals[0], torch.Tensor):
            out[key] = torch.cat(vals, axis) if concat else torch.stack(vals, axis)
        else:
            out[key] = vals
    return out
