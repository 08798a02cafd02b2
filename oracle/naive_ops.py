"""TEST INFRASTRUCTURE -- direct-loop NumPy fp64 restatements of the TensorFlow-1.x op semantics that the vendored
Keras backend maps to (Keras-2.0.8/keras/backend/tensorflow_backend.py, "KB").  They exist to pin
oracle/hdense_oracle.py (which uses torch.nn.functional) against an independent second statement of each op on tiny
shapes; nothing in the product imports this file.  Channels-last tensors throughout, like the reference.
"""
import numpy as np


def conv_valid(x, w, strides):
    """KB:3158 / KB:3307 tf.nn.convolution, padding VALID: cross-correlation (no kernel flip), kernel (*k, Cin, Cout),
    out = ceil((in - k + 1) / s) per spatial axis.  x (N, *spatial, Cin)."""
    nd = x.ndim - 2
    k = w.shape[:nd]
    out_sp = [-(-(x.shape[1 + i] - k[i] + 1) // strides[i]) for i in range(nd)]
    y = np.zeros((x.shape[0], *out_sp, w.shape[-1]), np.float64)
    for idx in np.ndindex(*out_sp):
        sl = tuple(slice(idx[i] * strides[i], idx[i] * strides[i] + k[i]) for i in range(nd))
        patch = x[(slice(None),) + sl]                               # (N, *k, Cin)
        y[(slice(None),) + idx] = np.tensordot(patch, w, axes=(list(range(1, nd + 2)), list(range(nd + 1))))
    return y


def zero_pad(x, p):
    """KB:2020 tf.pad, symmetric zero padding of every spatial axis (ZeroPadding2D/3D)."""
    nd = x.ndim - 2
    return np.pad(x, [(0, 0)] + [(p, p)] * nd + [(0, 0)])


def conv_same3(x, w):
    """padding='same', stride 1, odd kernels: pads (k-1)/2 on both sides (Keras-2.0.8/keras/utils/conv_utils.py:90-116)."""
    nd = x.ndim - 2
    pads = [(0, 0)] + [((w.shape[i] - 1) // 2,) * 2 for i in range(nd)] + [(0, 0)]
    return conv_valid(np.pad(x, pads), w, (1,) * nd)


def batchnorm_train(x, gamma, beta, eps):
    """KB:1635-1663: tf.nn.moments (population variance) then y = (x - mean) * rsqrt(var + eps) * gamma + beta."""
    ax = tuple(range(x.ndim - 1))
    mean, var = x.mean(axis=ax), x.var(axis=ax)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta, mean, var


def batchnorm_infer(x, gamma, beta, mean, var, eps):
    """KB:1684."""
    return (x - mean) / np.sqrt(var + eps) * gamma + beta


def upsample(x, size):
    """KB:1764-1771 / 1797-1827 == np.repeat per axis (Keras-2.0.8/tests/keras/layers/convolutional_test.py:673-681)."""
    for ax, s in enumerate(size):
        x = np.repeat(x, s, axis=1 + ax)
    return x


def max_pool(x, k, s):
    nd = x.ndim - 2
    out_sp = [(x.shape[1 + i] - k) // s + 1 for i in range(nd)]
    y = np.zeros((x.shape[0], *out_sp, x.shape[-1]), np.float64)
    for idx in np.ndindex(*out_sp):
        sl = tuple(slice(idx[i] * s, idx[i] * s + k) for i in range(nd))
        y[(slice(None),) + idx] = x[(slice(None),) + sl].max(axis=tuple(range(1, nd + 1)))
    return y


def avg_pool(x, k):
    nd = x.ndim - 2
    out_sp = [x.shape[1 + i] // k[i] for i in range(nd)]
    y = np.zeros((x.shape[0], *out_sp, x.shape[-1]), np.float64)
    for idx in np.ndindex(*out_sp):
        sl = tuple(slice(idx[i] * k[i], (idx[i] + 1) * k[i]) for i in range(nd))
        y[(slice(None),) + idx] = x[(slice(None),) + sl].mean(axis=tuple(range(1, nd + 1)))
    return y


def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def weighted_ce(y_true, y_pred, crop):
    """loss.py:5-25 (crop: slices 1:S-1 of the 4th axis) / loss.py:27-46, voxel by voxel."""
    w = (0.78, 0.65, 8.57)
    if crop:
        S = y_pred.shape[3]
        y_pred, y_true = y_pred[:, :, :, 1:S - 1], y_true[:, :, :, 1:S - 1]
    lp = y_pred.reshape(-1, 3)
    yt = np.asarray(y_true, np.float64).reshape(-1)
    tot, cnt = 0.0, 0
    for v in range(lp.shape[0]):
        p = softmax(lp[v])
        for c in range(3):
            if yt[v] == float(c):
                tot += w[c] * np.log(min(max(p[c], 1e-10), 1.0))
                cnt += 1
    return -tot / max(cnt, 1)


def nesterov(p, g, m, lr, mu):
    """Keras-2.0.8/keras/optimizers.py:172-181."""
    v = mu * m - lr * g
    return p + mu * v - lr * g, v
