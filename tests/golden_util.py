"""Load tests/golden/*.npz (written by oracle/gen_golden.py from the reference source)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.basename(p)[:-4]
        if kind is None or n.startswith(kind):
            out.append(n)
    return out


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {"params": {}, "in": {}, "out": {}, "meta": {}}
    for k in z.files:
        tag, rest = k.split(":", 1)
        g[{"param": "params", "in": "in", "out": "out", "meta": "meta"}[tag]][rest] = z[k]
    return g


def unflatten(flat):
    t = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = t
        for q in parts[:-1]:
            d = d.setdefault(q, {})
        d[parts[-1]] = v
    return t
