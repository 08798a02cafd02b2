"""GPU training-sample pipeline (csrc/augment.cu through the C-ABI, augment.py) against the restatement of
train_hybrid.py:40-98 / train_2ddense.py:40-69 (oracle/augment_oracle.py) on the same synthetic volumes and the same draws.
Labels: bit-exact.  Image: double-precision arithmetic on both sides, stored as float32 -> 1 ulp of the value range."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu


def _volume(seed, shape=(96, 90, 40), integral=True):
    rng = np.random.RandomState(seed)
    img = rng.uniform(-200, 250, size=shape)
    img = np.round(img) if integral else img
    seg = np.zeros(shape, np.uint8)
    H, W, S = shape
    seg[H // 5:H * 3 // 4, W // 5:W * 3 // 4, S // 6:S * 5 // 6] = 1
    seg[H * 3 // 8:H // 2, W // 3:W // 2, S // 3:S * 3 // 5] = 2
    liver = np.argwhere(seg >= 1)
    tumor = np.argwhere(seg == 2)
    box = np.concatenate([liver.min(0), liver.max(0)])
    return img.astype(np.float32), seg, liver[::97], tumor[::13], box


def _dataset(dev, compact, shape=(96, 90, 40)):
    from h_denseunet_b200.augment import DeviceVolumes
    dv = DeviceVolumes(dev, compact=compact)
    vols = [_volume(s, shape, integral=True) for s in (1, 2)]
    for v in vols:
        dv.add(*v)
    return dv, vols


@pytest.mark.parametrize("compact", [False, True])
def test_hybrid_samples_match_the_restatement(cuda_dev, compact):
    from h_denseunet_b200.augment import CropGenerator
    dv, vols = _dataset(cuda_dev, compact)
    gen = CropGenerator(dv, batch_size=3, input_size=32, input_cols=8, mean=48, rng=np.random.RandomState(5), liverlist=(),
                        reject_missing_class=False)
    seen = set()
    for _ in range(12):
        batch = next(gen)
        X, Y = batch.host()
        assert X.shape == (3, 32, 32, 8, 1) and Y.shape == (3, 32, 32, 8, 1) and Y.dtype == np.int16
        for n, (count, a, b, c, half, k) in enumerate(batch.params):
            seen.add(k)
            x, y = ao.make_sample(vols[count][0], vols[count][1], a, b, c, half, half, 8, k, 48, 32)
            assert np.array_equal(Y[n, ..., 0], y), (count, a, b, c, half, k)
            assert np.allclose(X[n, ..., 0], x, rtol=0, atol=1e-4), (np.abs(X[n, ..., 0] - x).max(), k)
        hist = [int((Y == v).sum()) for v in (0, 1, 2)]
    assert len(seen) >= 6                                    # the eight flips / rotations were exercised


def test_2d_samples_match_the_restatement(cuda_dev):
    from h_denseunet_b200.augment import CropGenerator
    dv, vols = _dataset(cuda_dev, False)
    gen = CropGenerator(dv, batch_size=2, input_size=32, input_cols=3, mean=48, two_d=True, rng=np.random.RandomState(9), liverlist=())
    for _ in range(6):
        batch = next(gen)
        X, Y = batch.host()
        assert X.shape == (2, 32, 32, 3) and Y.shape == (2, 32, 32, 1)
        assert float(batch.x[..., 3].abs().max()) == 0.0     # the pad channel of the 2-D stem stays zero
        for n, (count, a, b, c, half, k) in enumerate(batch.params):
            assert k == 0
            x, y = ao.make_sample(vols[count][0], vols[count][1], a, b, c, half, half, 3, 0, 48, 32, two_d=True)
            assert np.array_equal(Y[n, ..., 0], y)
            assert np.allclose(X[n], x, rtol=0, atol=1e-4)


def test_clip_keeps_exact_cval_when_zero_is_outside_the_range(cuda_dev):
    """All intensities above the mean: 0 lies outside the crop's range, border taps mix in cval = 0 and get clipped up to the
    minimum -- except samples that are exactly 0 (none here) -- as _clip_warp_output does."""
    from h_denseunet_b200.augment import CropGenerator, DeviceVolumes
    img, seg, liver, tumor, box = _volume(4)
    img = np.abs(img) + 100.0
    dv = DeviceVolumes(cuda_dev)
    dv.add(img, seg, liver, tumor, box)
    gen = CropGenerator(dv, batch_size=1, input_size=40, input_cols=8, mean=48, rng=np.random.RandomState(2), liverlist=(),
                        reject_missing_class=False)
    batch = next(gen)
    X, Y = batch.host()
    count, a, b, c, half, k = batch.params[0]
    x, y = ao.make_sample(img, seg, a, b, c, half, half, 8, k, 48, 40)
    assert x.min() >= 52.0 - 1e-6
    assert np.allclose(X[0, ..., 0], x, rtol=0, atol=1e-4) and np.array_equal(Y[0, ..., 0], y)


def test_class_rejection_and_counts(cuda_dev):
    from h_denseunet_b200.augment import CropGenerator, DeviceVolumes
    img, seg, liver, tumor, box = _volume(6)
    seg_no_tumor = np.where(seg == 2, 1, seg).astype(np.uint8)
    dv = DeviceVolumes(cuda_dev)
    dv.add(img, seg_no_tumor, liver, liver, box)             # volume 0 can never show class 2
    dv.add(img, seg, liver, tumor, box)
    order = iter([0, 0, 1, 1, 1, 1, 1, 1, 1, 1])
    gen = CropGenerator(dv, batch_size=1, input_size=32, input_cols=8, mean=48, rng=np.random.RandomState(1), liverlist=(),
                        choice=lambda idx: next(order))
    batch = next(gen)
    assert gen.rejected >= 2 and batch.params[0][0] == 1     # the two class-2-free batches were discarded (train_hybrid.py:126-131)
    X, Y = batch.host()
    assert batch.counts == [int((Y == v).sum()) for v in (0, 1, 2)] and min(batch.counts) > 0


def test_fit_generator_consumes_device_batches(cuda_dev):
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.augment import CropGenerator
    from util import Args
    dv, vols = _dataset(cuda_dev, False, shape=(200, 190, 24))
    size, cols = 64, 8
    model = hdn.dense_rnn_net(Args(b=1, input_size=size, input_cols=cols), device=str(cuda_dev), precision="mixed")
    model.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    gen = CropGenerator(dv, batch_size=1, input_size=size, input_cols=cols, mean=48, rng=np.random.RandomState(11), liverlist=())
    hist = model.fit_generator(gen, steps_per_epoch=3, epochs=1, verbose=0)
    assert np.isfinite(hist["loss"][0]) and model.h2d_bytes == 0
    # same draws through the host path: the model sees the same batch whichever way it arrives
    model2 = hdn.dense_rnn_net(Args(b=1, input_size=size, input_cols=cols), device=str(cuda_dev), precision="mixed")
    model2.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    model2.set_weights_dict(model.get_weights_dict())
    gen_a = CropGenerator(dv, batch_size=1, input_size=size, input_cols=cols, mean=48, rng=np.random.RandomState(13), liverlist=())
    gen_b = CropGenerator(dv, batch_size=1, input_size=size, input_cols=cols, mean=48, rng=np.random.RandomState(13), liverlist=())
    batch = next(gen_a)
    model._consume_device_batch(model._net(True), batch)
    X, Y = next(gen_b).host()
    net2 = model2._net(True)
    model2._upload(net2, X)
    model2._labels(net2, Y)
    assert torch.equal(list(model._net(True).inputs.values())[0].data, list(net2.inputs.values())[0].data)
    assert torch.equal(model._net(True).loss.labels, net2.loss.labels)
