"""GPU post-processing (csrc/postproc.cu) against scipy.ndimage on the same volumes: bit-exact.
The oracle (oracle/postproc_oracle.py) restates test.py:71-115 with scipy; index / byte work, so the bar is equality."""
import numpy as np
import pytest
import torch
from scipy import ndimage

from oracle import postproc_oracle as po

pytestmark = pytest.mark.gpu


def _blobs(shape, seed, density=0.5, smooth=2.0):
    rng = np.random.default_rng(seed)
    f = ndimage.gaussian_filter(rng.normal(size=shape), smooth)
    return (f > np.quantile(f, 1.0 - density)).astype(np.uint8)


@pytest.mark.parametrize("shape,seed", [((24, 20, 17), 0), ((40, 33, 9), 1), ((7, 5, 3), 2), ((64, 64, 48), 3)])
def test_primitives_match_scipy(cuda_dev, shape, seed):
    from h_denseunet_b200.postprocess import PostProcessor
    pp = PostProcessor(shape, cuda_dev)
    for density in (0.15, 0.5, 0.85):
        x = _blobs(shape, seed * 10 + int(density * 100), density)
        t = torch.from_numpy(x).to(cuda_dev)
        assert np.array_equal(pp.dilate(t).cpu().numpy(), ndimage.binary_dilation(x, iterations=1).astype(np.uint8))
        assert np.array_equal(pp.fill_holes(t).cpu().numpy(), ndimage.binary_fill_holes(x).astype(np.uint8))
        assert np.array_equal(pp.largest_component(t).cpu().numpy(), po.largest_component(x))


def test_largest_component_tie_goes_to_the_first_in_raster_order(cuda_dev):
    from h_denseunet_b200.postprocess import PostProcessor
    x = np.zeros((8, 8, 8), np.uint8)
    x[1:3, 1:3, 1:3] = 1            # 8 voxels
    x[5:7, 5:7, 5:7] = 1            # 8 voxels, later in raster order
    x[0, 7, 7] = 1
    pp = PostProcessor(x.shape, cuda_dev)
    got = pp.largest_component(torch.from_numpy(x).to(cuda_dev)).cpu().numpy()
    exp = np.zeros_like(x)
    exp[1:3, 1:3, 1:3] = 1
    assert np.array_equal(got, exp) and np.array_equal(po.largest_component(x), exp)
    # diagonal (26-connectivity) contact joins components; hole filling uses 6-connectivity for the background
    y = np.zeros((6, 6, 6), np.uint8)
    y[1, 1, 1] = y[2, 2, 2] = y[3, 3, 3] = 1
    y[5, 0, 0] = 1
    got = PostProcessor(y.shape, cuda_dev).largest_component(torch.from_numpy(y).to(cuda_dev)).cpu().numpy()
    assert got.sum() == 3 and np.array_equal(got, po.largest_component(y))


@pytest.mark.parametrize("shape,seed", [((48, 40, 24), 5), ((96, 96, 40), 6)])
def test_full_pipeline_matches_the_reference_steps(cuda_dev, shape, seed):
    """test.py:71-115 end to end on synthetic probability volumes and a synthetic stage-1 liver mask."""
    from h_denseunet_b200.postprocess import postprocess_scores
    rng = np.random.default_rng(seed)
    liver = _blobs(shape, seed, 0.3, 4.0)
    s1 = np.clip(ndimage.gaussian_filter(liver.astype(np.float32), 1.5) + rng.normal(0, 0.15, shape), 0, 1).astype(np.float32)
    tum = _blobs(shape, seed + 100, 0.08, 2.5) & liver
    s2 = np.clip(ndimage.gaussian_filter(tum.astype(np.float32), 1.0) * 1.6 + rng.normal(0, 0.1, shape), 0, 1).astype(np.float32)
    mask = ndimage.binary_dilation(_blobs(shape, seed + 7, 0.35, 4.0), iterations=1).astype(np.uint8)      # test.py:60-62
    got = postprocess_scores(s1.copy(), s2.copy(), mask, 0.5, 0.9, device=cuda_dev)
    exp = po.postprocess_scores(s1.copy(), s2.copy(), mask, 0.5, 0.9)
    assert got.dtype == np.uint8 and set(np.unique(got)) <= {0, 1, 2}
    assert np.array_equal(got, exp), (int((got != exp).sum()), got.size)
    assert (exp == 2).sum() > 0 and (exp == 1).sum() > 0
