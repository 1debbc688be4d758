"""Pair builders with the reference's names (sake/functional.py:7-44).

Inside the fused layer these are computed on the fly in the CUDA kernels and never materialised;
the functions here exist for API parity (scripts / tests that call them directly) and are plain
torch tensor ops on whatever device the input lives on."""
import torch

EPSILON = 1e-5
INF = 1e5


def get_x_minus_xt(x):
    # sake/functional.py:7-8
    return x.unsqueeze(-3) - x.unsqueeze(-2)


def get_x_minus_xt_norm(x_minus_xt, epsilon: float = EPSILON):
    # sake/functional.py:10-19
    return (torch.relu((x_minus_xt ** 2).sum(dim=-1, keepdim=True)) + epsilon) ** 0.5


def get_h_cat_ht(h):
    # sake/functional.py:33-44
    n = h.shape[-2]
    shape = (*h.shape[:-2], n, n, h.shape[-1])
    return torch.cat([h.unsqueeze(-3).expand(shape), h.unsqueeze(-2).expand(shape)], dim=-1)
