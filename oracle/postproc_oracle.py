"""CPU ORACLE of the post-processing step -- TEST INFRASTRUCTURE, NOT PRODUCT (only tests/ may import it).

Restates test.py:71-115 with the libraries that ARE in this image: scipy.ndimage for dilation / hole filling (the
reference's own calls) and ndimage.label with the full 3x3x3 structure in place of skimage.measure.label (absent here;
measure.label's default connectivity in 3-D is the full one and both number components in raster order, so
regionprops' area list and `box.index(max(box))` map to the same component).  Pinned by the reference's call sites only:
the reference holds no test or fixture for this step (parity unpinned beyond the library semantics)."""
import numpy as np
from scipy import ndimage

FULL = np.ones((3, 3, 3), bool)


def largest_component(x):
    """measure.label + regionprops area + box.index(max(box)) + 1 (test.py:83-91)."""
    lab, num = ndimage.label(np.asarray(x) != 0, structure=FULL)
    if num == 0:
        return np.zeros(lab.shape, np.uint8)
    box = [int((lab == i + 1).sum()) for i in range(num)] if num < 64 else list(np.bincount(lab.ravel())[1:])
    label_num = box.index(max(box)) + 1
    return (lab == label_num).astype(np.uint8)


def postprocess_scores(score1, score2, mask, thres_liver=0.5, thres_tumor=0.9):
    result1 = (np.asarray(score1) >= thres_liver).astype(np.uint8)                    # test.py:73-76
    result2 = (np.asarray(score2) >= thres_tumor).astype(np.uint8)
    result1[result2 == 1] = 1                                                         # :77
    liver_res = largest_component(result1)                                            # :81-91
    m = ndimage.binary_dilation(np.asarray(mask) != 0, iterations=1)                  # :94
    liver_labels = ndimage.binary_fill_holes(largest_component(m)).astype(int)        # :95-104
    segmask = ndimage.binary_fill_holes(result2 * liver_labels).astype(np.uint8)      # :107-110
    liver_res = ndimage.binary_fill_holes(liver_res).astype(int)                      # :112
    liver_res[segmask == 1] = 2                                                       # :113
    return liver_res.astype(np.uint8)
