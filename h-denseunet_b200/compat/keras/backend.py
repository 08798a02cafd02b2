"""The four backend calls the scripts make (train_*.py:18, test.py:17,69, lib/funcs.py:31-32)."""
import numpy as np


def set_image_dim_ordering(order):
    if order != "tf":
        raise ValueError("only channels-last ('tf') is supported")


def image_dim_ordering():
    return "tf"


def clear_session():
    import torch
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def eval(x):          # noqa: A001  (Keras name)
    return np.asarray(x)
