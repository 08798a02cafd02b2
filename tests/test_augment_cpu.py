"""Host side of the GPU training-sample pipeline (h_denseunet_b200/augment.py) and the restatement of
the reference's sample function (oracle/augment_oracle.py; train_hybrid.py:40-98, train_2ddense.py:40-69)."""
import numpy as np
import pytest

from oracle import augment_oracle as ao


def test_flip_maps_are_numpys_flips():
    from h_denseunet_b200.augment import flip_map
    rng = np.random.RandomState(0)
    for n in (2, 5, 8):
        crop = rng.randint(0, 1000, size=(n, n, 3))
        for k in range(8):
            m00, m01, m10, m11, o0, o1 = flip_map(k, n, n)
            want = ao.flip(crop, k)
            i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
            got = crop[m00 * i + m01 * j + o0, m10 * i + m11 * j + o1]
            assert np.array_equal(got, want), (n, k)
    # non-square crops: only the pure flips are defined
    crop = rng.randint(0, 1000, size=(4, 6, 2))
    for k in (0, 1, 2, 7):
        m00, m01, m10, m11, o0, o1 = flip_map(k, 4, 6)
        i, j = np.meshgrid(np.arange(4), np.arange(6), indexing="ij")
        assert np.array_equal(crop[m00 * i + m01 * j + o0, m10 * i + m11 * j + o1], ao.flip(crop, k))
    with pytest.raises(ValueError):
        flip_map(3, 4, 6)
    with pytest.raises(ValueError):
        flip_map(8, 4, 4)


def test_draws_follow_the_reference_order():
    from h_denseunet_b200.augment import draw_crop
    lines = ["%d %d %d\n" % (100 + 3 * k, 120 + 2 * k, 30 + k) for k in range(50)]
    mn, mx = np.array([20, 30, 5]), np.array([400, 420, 90])
    for flips, cols in ((True, 8), (False, 3)):
        r1, r2 = np.random.RandomState(7), np.random.RandomState(7)
        for _ in range(20):
            a, b, c, half, k = draw_crop(r1, 224, cols, lines, len(lines), mn, mx, flips)
            a2, b2, c2, hd, hr, k2 = ao.draw_sample_params(r2, 224, cols, lines, len(lines), mn, mx, flips)
            assert (a, b, c, half, half, k) == (a2, b2, c2, hd, hr, k2)
            assert 179 // 2 <= half <= 268 // 2 and mn[0] + half <= a <= mx[0] - half - 1
        assert r1.randint(0, 1 << 30) == r2.randint(0, 1 << 30)        # same number of draws consumed


def _loops_resize(img, out, order):
    """Independent scalar restatement (per output pixel) of the two resize modes."""
    h, w, s = img.shape
    res = np.zeros((out, out, s))
    for y in range(out):
        for x in range(out):
            r = (h / out) * (y + 0.5) - 0.5
            c = (w / out) * (x + 0.5) - 0.5
            if order == 0:
                ri = min(max(int(np.floor(r + 0.5)), 0), h - 1)
                ci = min(max(int(np.floor(c + 0.5)), 0), w - 1)
                res[y, x] = img[ri, ci]
                continue
            r0, c0 = int(np.floor(r)), int(np.floor(c))
            fr = []
            for dr in range(-1, 3):
                f = []
                for dc in range(-1, 3):
                    ri, ci = r0 + dr, c0 + dc
                    f.append(img[ri, ci].astype(np.float64) if 0 <= ri < h and 0 <= ci < w else np.zeros(s))
                fr.append(ao._cubic(c - c0, *f))
            res[y, x] = ao._cubic(r - r0, *fr)
    return res


@pytest.mark.parametrize("n_in,n_out", [(10, 12), (13, 12), (12, 12), (9, 16)])
def test_resize_restatements_agree(n_in, n_out):
    rng = np.random.RandomState(n_in * 31 + n_out)
    img = rng.normal(size=(n_in, n_in, 3)).astype(np.float32) * 50
    lab = rng.randint(0, 3, size=(n_in, n_in, 3))
    assert np.array_equal(ao.resize_nearest_edge(lab, n_out, n_out), _loops_resize(lab, n_out, 0))
    got = ao.resize_cubic_constant(img, n_out, n_out, clip=False)
    assert np.allclose(got, _loops_resize(img, n_out, 3), rtol=0, atol=1e-9)
    clipped = ao.resize_cubic_constant(img, n_out, n_out)
    assert clipped.min() >= img.min() and clipped.max() <= img.max()
    if n_in == n_out:                                       # scale 1: src == dst, Catmull-Rom at offset 0 returns the sample
        assert np.allclose(clipped, img, atol=1e-12)
        assert np.array_equal(ao.resize_nearest_edge(lab, n_out, n_out), lab)


def test_cubic_reproduces_ramps_and_keeps_cval_outside_the_range():
    ramp = np.tile(np.arange(20, dtype=np.float32)[:, None, None], (1, 20, 1)) + 5.0      # strictly positive: 0 outside the range
    out = ao.resize_cubic_constant(ramp, 16, 16)
    src = (20 / 16) * (np.arange(16) + 0.5) - 0.5
    inner = (src >= 1) & (src <= 18)
    assert np.allclose(out[inner][:, 4:12, 0], (src[inner] + 5.0)[:, None], atol=1e-9)     # linear precision in the interior
    assert out.min() >= 5.0                                  # border overshoot / cval mixing is clipped into the range ...
    z = np.zeros((6, 6, 1), np.float32)
    z[2:4, 2:4] = 7.0
    z[:] += 1.0
    far = ao.resize_cubic_constant(z, 6, 6)
    assert np.allclose(far, z)


def test_make_sample_shapes_and_class_rule():
    rng = np.random.RandomState(3)
    img = rng.normal(size=(64, 60, 20)).astype(np.float32) * 100
    seg = rng.randint(0, 3, size=(64, 60, 20)).astype(np.uint8)
    x, y = ao.make_sample(img, seg, 30, 28, 10, 11, 11, 8, 5, 48, 16)
    assert x.shape == (16, 16, 8) and y.shape == (16, 16, 8) and x.dtype == np.float32 and y.dtype == np.int16
    x2, y2 = ao.make_sample(img, seg, 30, 28, 10, 11, 11, 3, 0, 48, 16, two_d=True)
    assert x2.shape == (16, 16, 3) and y2.shape == (16, 16)
    assert ao.batch_has_all_classes(y) and not ao.batch_has_all_classes(np.zeros((4, 4)))
