"""Helpers with the reference's names (sake/utils.py).  ExpNormalSmearing itself runs inside the
fused edge kernel; `exp_normal_smearing_init` restates its initial parameters (utils.py:49-59)."""
import math

import torch


def coloring(x, mean, std):
    # sake/utils.py:7-8
    return std * x + mean


def exp_normal_smearing_init(num_rbf=50, cutoff_lower=0.0, cutoff_upper=5.0):
    # sake/utils.py:49-59 (PhysNet defaults); computed in fp64 then rounded to fp32 like jnp
    start = math.exp(-cutoff_upper + cutoff_lower)
    means = torch.linspace(start, 1.0, num_rbf, dtype=torch.float64).float()
    betas = torch.full((num_rbf,), (2.0 / num_rbf * (1.0 - start)) ** -2, dtype=torch.float64).float()
    return means, betas


def mae(x, y):
    # sake/utils.py:67-69
    return (x - y).abs().mean()
