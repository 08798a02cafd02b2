"""Pins the CPU oracle (oracle/hdense_oracle.py):
  * against the numerically pinned upstream Keras tests that exist for this path -- UpSampling2D/3D == np.repeat
    (Keras-2.0.8/tests/keras/layers/convolutional_test.py:673-681,726-736), ZeroPadding2D/3D: zero border, interior
    untouched (same file :508-628), Add / Concatenate (Keras-2.0.8/tests/keras/layers/merge_test.py:13-30,142-177),
    softmax against the NumPy reference inside Keras-2.0.8/tests/keras/activations_test.py:53-68 and relu (:158-164);
  * against oracle/naive_ops.py, an independent direct-loop NumPy fp64 restatement of each TF op, on tiny shapes;
  * against the committed fixtures in tests/golden/ (made by tests/golden/make_golden.py from the oracle itself: they
    guard against drift, they are NOT reference outputs -- TensorFlow 1.x is not installable here, so conv / BN /
    pool / softmax / loss / SGD parity stays "unpinned" with respect to the reference's own arithmetic).
"""
import json
import os

import numpy as np
import torch

from oracle import hdense_oracle as orc
from oracle import naive_ops as nv

HERE = os.path.dirname(os.path.abspath(__file__))


def _cf(x):
    return torch.as_tensor(np.moveaxis(x, -1, 1).copy())


def _cl(t):
    return np.moveaxis(t.detach().numpy(), 1, -1)


def test_upsampling_is_np_repeat():
    rng = np.random.default_rng(0)
    x2 = rng.normal(size=(2, 5, 7, 3))
    for size in [(2, 2), (2, 1), (1, 2)]:
        assert np.array_equal(_cl(orc.upsample(_cf(x2), size)), np.repeat(np.repeat(x2, size[0], 1), size[1], 2))
    x3 = rng.normal(size=(1, 3, 4, 5, 2))
    for size in [(2, 2, 1), (2, 2, 2)]:
        e = x3
        for ax, s in enumerate(size):
            e = np.repeat(e, s, axis=1 + ax)
        assert np.array_equal(_cl(orc.upsample(_cf(x3), size)), e)


def test_zero_padding_border_and_interior():
    x = np.ones((1, 4, 5, 2))
    y = _cl(orc.zero_pad(_cf(x), 3))
    assert y.shape == (1, 10, 11, 2)
    assert np.all(y[:, :3] == 0) and np.all(y[:, -3:] == 0) and np.all(y[:, :, :3] == 0) and np.all(y[:, :, -3:] == 0)
    assert np.all(y[:, 3:-3, 3:-3] == 1)
    x3 = np.ones((1, 2, 3, 4, 1))
    y3 = _cl(orc.zero_pad(_cf(x3), 1))
    assert y3.shape == (1, 4, 5, 6, 1) and y3.sum() == x3.sum() and np.all(y3[:, 1:-1, 1:-1, 1:-1] == 1)


def test_add_and_concat():
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(2, 4, 4, 3)), rng.normal(size=(2, 4, 4, 3))
    assert np.array_equal((torch.as_tensor(a) + torch.as_tensor(b)).numpy(), a + b)
    assert np.array_equal(torch.cat([_cf(a), _cf(b)], dim=1).numpy(), np.moveaxis(np.concatenate([a, b], -1), -1, 1))


def _ctx(params, training=True):
    return orc.Ctx(params, training, dtype=torch.float64)


def test_softmax_and_relu_known_answers():
    """Keras-2.0.8/tests/keras/activations_test.py: softmax against the test's own NumPy reference on its standard
    values (:53-68, rtol 1e-5), relu is the identity on them (:158-164); plus the sign case the upstream test omits."""
    vals = np.array([[0, 0.1, 0.5, 0.9, 1.0]], dtype=np.float32)          # get_standard_values(), :11-14

    def softmax_ref(values):                                               # :56-59
        m = np.max(values)
        e = np.exp(values - m)
        return e / np.sum(e)

    np.testing.assert_allclose(orc.softmax(vals).numpy(), softmax_ref(vals), rtol=1e-5)
    np.testing.assert_allclose(orc.relu(vals).numpy(), vals, rtol=1e-5)
    assert np.array_equal(orc.relu(np.array([-1.5, -0.0, 2.0], np.float32)).numpy(), np.array([0.0, 0.0, 2.0], np.float32))
    # the loss takes the soft-max over the class axis of every voxel (loss.py:9): rows are independent
    x = np.random.default_rng(0).normal(0, 3, (7, 3)).astype(np.float32)
    got = orc.softmax(x, axis=1).numpy()
    for i in range(7):
        np.testing.assert_allclose(got[i], softmax_ref(x[i]), rtol=1e-5)


def test_conv_against_direct_loops():
    rng = np.random.default_rng(2)
    cases = [((1, 9, 8, 3), (3, 3, 3, 5), (1, 1), "same"), ((2, 13, 12, 3), (7, 7, 3, 4), (2, 2), "stem"),
             ((1, 6, 6, 5), (1, 1, 5, 7), (1, 1), "valid"), ((1, 5, 6, 4, 2), (3, 3, 3, 2, 3), (1, 1, 1), "same"),
             ((1, 8, 8, 6, 2), (7, 7, 7, 2, 3), (2, 2, 2), "stem")]
    for xs, ws, st, kind in cases:
        x, w, b = rng.normal(size=xs), rng.normal(size=ws), rng.normal(size=ws[-1])
        ctx = _ctx({"c/kernel": w, "c/bias": b})
        if kind == "same":
            got = _cl(orc.conv(ctx, _cf(x), "c", padding="same"))
            exp = nv.conv_same3(x, w) + b
        elif kind == "stem":          # ZeroPadding(3) + VALID stride-2 convolution, no bias (hybridnet.py:122-123,208-209)
            got = _cl(orc.conv(ctx, orc.zero_pad(_cf(x), 3), "c", strides=st[0], use_bias=False))
            exp = nv.conv_valid(nv.zero_pad(x, 3), w, st)
        else:
            got = _cl(orc.conv(ctx, _cf(x), "c"))
            exp = nv.conv_valid(x, w, st) + b
        assert got.shape == exp.shape and np.allclose(got, exp, atol=1e-10), (xs, ws, kind)


def test_batchnorm_scale_pool_against_direct_loops():
    rng = np.random.default_rng(3)
    x = rng.normal(2.0, 3.0, size=(2, 6, 6, 4))
    p = {"bn/gamma": rng.uniform(0.5, 1.5, 4), "bn/beta": rng.normal(size=4), "bn/moving_mean": rng.normal(size=4),
         "bn/moving_variance": rng.uniform(0.5, 1.5, 4), "sc/gamma": rng.uniform(0.5, 1.5, 4), "sc/beta": rng.normal(size=4)}
    ctx = _ctx(p, True)
    got = _cl(orc.bn(ctx, _cf(x), "bn", 1.1e-5, True))
    exp, mean, var = nv.batchnorm_train(x, p["bn/gamma"], p["bn/beta"], 1.1e-5)
    assert np.allclose(got, exp, atol=1e-9)
    m, v, mom = ctx.bn_updates["bn"]
    assert np.allclose(m.numpy(), mean) and np.allclose(v.numpy(), var) and mom == 0.99      # biased variance (KNORM:179-185)
    assert np.allclose(orc.moving_average_update(p["bn/moving_mean"], mean, 0.99), p["bn/moving_mean"] * 0.99 + mean * 0.01)
    got = _cl(orc.bn(_ctx(p, False), _cf(x), "bn", 1e-3, True))
    assert np.allclose(got, nv.batchnorm_infer(x, p["bn/gamma"], p["bn/beta"], p["bn/moving_mean"], p["bn/moving_variance"], 1e-3))
    assert np.allclose(_cl(orc.scale(ctx, _cf(x), "sc")), x * p["sc/gamma"] + p["sc/beta"])
    xr = np.maximum(x, 0)
    assert np.allclose(_cl(orc.max_pool(orc.zero_pad(_cf(xr), 1), 3, 2)), nv.max_pool(nv.zero_pad(xr, 1), 3, 2))
    assert np.allclose(_cl(orc.avg_pool(_cf(x), 2)), nv.avg_pool(x, (2, 2)))
    x3 = rng.normal(size=(1, 4, 4, 3, 2))
    assert np.allclose(_cl(orc.avg_pool(_cf(x3), (2, 2, 1))), nv.avg_pool(x3, (2, 2, 1)))


def test_loss_and_sgd_against_direct_loops():
    rng = np.random.default_rng(4)
    logits = rng.normal(0, 3, size=(1, 3, 3, 8, 3))
    logits[0, 0, 0, 2] = [60.0, -60.0, -60.0]                  # exercises the 1e-10 clip (loss.py:11)
    y = rng.integers(0, 3, size=(1, 3, 3, 8)).astype(np.float64)
    y[0, 0, 0, 2] = 1
    y[0, 1, 1, 3] = 7                                            # a label outside {0,1,2} is dropped from both sums
    for crop in (True, False):
        got = float(orc.weighted_crossentropy(torch.as_tensor(y), torch.as_tensor(logits), crop=crop))
        assert abs(got - nv.weighted_ce(y, logits, crop)) < 1e-9
    p, g, m = rng.normal(size=5), rng.normal(size=5), rng.normal(size=5)
    a, b = orc.sgd_nesterov_step(p, g, m, 1e-3, 0.9)
    c, d = nv.nesterov(p, g, m, 1e-3, 0.9)
    assert np.allclose(a, c) and np.allclose(b, d)


def test_slice_triplets_and_window_starts():
    vol = torch.arange(2 * 2 * 2 * 5, dtype=torch.float64).reshape(2, 2, 2, 5)
    t = orc.slice_triplets(vol).reshape(2, 5, 3, 2, 2)
    for s in range(5):
        idx = [max(s - 1, 0), s, min(s + 1, 4)]                # hybridnet.py:385-396
        for k in range(3):
            assert torch.equal(t[:, s, k], vol[:, :, :, idx[k]])
    # lib/funcs.py:12,19-27 with cols = 8: stride 2, tail window clamped to z - 8
    assert orc.window_starts(512, 0, 511, 8) == list(range(0, 506, 2))
    assert orc.window_starts(20, 4, 15, 8) == [0, 2, 4, 6, 8, 10, 12]
    assert orc.window_starts(21, 4, 30, 8)[-1] == 13


def test_golden_fixtures():
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
        gold = json.load(f)
    from golden.make_golden import compute
    now = compute()
    assert set(now) == set(gold)
    for k in gold:
        assert np.allclose(now[k], gold[k], rtol=2e-4, atol=1e-6), k
