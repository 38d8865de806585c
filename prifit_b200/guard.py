"""Mirror of the reference's src/guard.py (guard_exp :6-11, guard_sqrt :13-18, guard_acos :21-23).

On the hot path these are fused into the kernels' epilogues (csrc/common.cuh: guard_expf); the
standalone functions are kept for API compatibility and act on tensors of any device.
"""
import torch


def guard_exp(x, max_value=75, min_value=-13):
    return torch.exp(torch.clamp(x, max=max_value, min=min_value))


def guard_sqrt(x, minimum=1e-5):
    return torch.sqrt(torch.clamp(x, min=minimum))


def guard_acos(x):
    return torch.acos(torch.clamp(x, min=-1.0, max=1.0))
