"""Synthetic CT slabs for benchmarks and tests: no dataset is read anywhere."""
import numpy as np


def synthetic_slab(B, size, cols, seed=1234):
    """CT-like volume in [-248, 202] (HU clipped to [-200,250] minus mean 48, preprocessing.py:15-16,
    train_hybrid.py:35,65) with an ellipsoidal 'liver' (+60) holding spherical 'tumors' (-40), and the
    matching int labels {0,1,2} (every slab contains all three classes, train_hybrid.py:127-132).
    Returns vol (B,size,size,cols,1) float32, labels (B,size,size,cols,1) int16."""
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, cols), np.linspace(-1, 1, size), np.linspace(-1, 1, size),
                             indexing="ij")
    vol = np.zeros((B, size, size, cols, 1), np.float32)
    lab = np.zeros((B, size, size, cols, 1), np.int16)
    for b in range(B):
        base = rng.normal(0, 1, (cols, size // 4 + 1, size // 4 + 1)).astype(np.float32)
        base = np.kron(base, np.ones((1, 4, 4), np.float32))[:, :size, :size]
        field = -80.0 + 50.0 * base
        cy, cx = rng.uniform(-0.2, 0.2, 2)
        liver = ((yy - cy) / 0.6) ** 2 + ((xx - cx) / 0.5) ** 2 + (zz / 1.2) ** 2 < 1.0
        field = field + 60.0 * liver
        l = liver.astype(np.int16)
        for _ in range(3):
            ty, tx, tz = cy + rng.uniform(-0.25, 0.25), cx + rng.uniform(-0.2, 0.2), rng.uniform(-0.5, 0.5)
            tumor = ((yy - ty) ** 2 + (xx - tx) ** 2 + ((zz - tz) * 0.8) ** 2 < 0.18 ** 2) & liver
            field = field - 40.0 * tumor
            l[tumor] = 2
        field = np.clip(field, -248.0, 202.0)
        vol[b, :, :, :, 0] = field.transpose(1, 2, 0)
        lab[b, :, :, :, 0] = l.transpose(1, 2, 0)
    return vol, lab
