"""Shared helpers for the tests: synthetic CT slabs, parameter perturbation, error metrics."""
import numpy as np


class Args(object):
    def __init__(self, b=1, input_size=32, input_cols=8):
        self.b, self.input_size, self.input_cols = b, input_size, input_cols


def perturb_params(model, seed=7):
    """BN/Scale parameters and moving statistics away from their 1/0 defaults so the folded
    affine path is exercised: gamma U[.5,1.5], beta N(0,.1), mean N(0,.1), var U[.5,1.5];
    biases N(0,.05)."""
    rng = np.random.default_rng(seed)
    for p in model.params.order:
        w = p.name.rsplit("/", 1)[1]
        if w == "gamma":
            v = rng.uniform(0.5, 1.5, p.shape)
        elif w == "beta":
            v = rng.normal(0, 0.1, p.shape)
        elif w == "moving_mean":
            v = rng.normal(0, 0.1, p.shape)
        elif w == "moving_variance":
            v = rng.uniform(0.5, 1.5, p.shape)
        elif w == "bias":
            v = rng.normal(0, 0.05, p.shape)
        else:
            continue
        model.params.set_value(p.name, v.astype(np.float32))


from h_denseunet_b200.synthetic import synthetic_slab  # noqa: E402,F401


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
