"""Seeded synthetic inputs of the BASELINE.json shapes, numpy only."""
import numpy as np


def molecules(seed, B, N, n_species, padded=False, n_min=None):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((B, N, 3)) * 0.62 * N ** (1.0 / 3.0)).astype(np.float32)
    z = rng.integers(0, n_species, (B, N))
    h = np.eye(n_species, dtype=np.float32)[z]
    if not padded:
        return h, x, None, None
    n_real = rng.integers(n_min, N + 1, B)
    am = (np.arange(N)[None, :] < n_real[:, None]).astype(np.float32)
    h = h * am[..., None]
    x = x * am[..., None]
    mask = am[:, :, None] * am[:, None, :]
    return h, x, mask, am
