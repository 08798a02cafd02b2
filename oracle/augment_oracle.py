"""TEST INFRASTRUCTURE (see oracle/README): CPU restatement of the reference's training-sample pipeline.

Follows `load_seq_crop_data_masktumor_try` (train_hybrid.py:40-98, train_2ddense.py:40-69): liver/tumour centred crop at a
random scale, mean subtraction, one of 8 flips/rotations (hybrid script only), `skimage.transform.resize` of the label map
(order 0, mode 'edge') and of the image (order 3, mode 'constant', cval 0, clip, preserve_range) to the network size.

PARITY UNPINNED for the two `resize` calls: scikit-image (requirements.txt:63, 0.13.1) is not in this image.  What is
restated here is its documented algorithm for a 3-D array whose third extent does not change: `warp` per channel with the
scaling transform  src = scale * (dst + 0.5) - 0.5  (skimage/transform/_warps.py, `resize` -> `warp` -> `_warp_fast`),
nearest = `round` with edge clamping, order 3 = separable Catmull-Rom cubic convolution on the 4x4 neighbourhood around
floor(src) with out-of-range taps = cval, then `_clip_warp_output` (clip to the input's range; exact cval samples are kept
when cval lies outside that range).  Release 0.13.1's cubic kernel evaluated the same polynomial at a differently scaled
offset (fixed upstream later); the mathematically defined form is used here.  The transform is exact arithmetic here;
skimage estimates it from three corner points (differences ~1e-16, which only matter on exact rounding ties of order 0).
Everything else (crop arithmetic, flips, the np.random call order) is plain numpy and bit-exact by construction.
"""
import numpy as np


def draw_sample_params(rng, input_size, cols, lines, numid, minindex, maxindex, flips=True):
    """The random draws and the crop arithmetic of train_hybrid.py:47-60 (Python-2 integer division), in the reference's
    call order on `rng` (a np.random.RandomState or the np.random module).  `cols` = 3 and flips=False give
    train_2ddense.py:47-58.  Returns (a, b, c, half_d, half_r, flip_num)."""
    scale = rng.uniform(0.8, 1.2)
    deps = int(input_size * scale)
    rows = int(input_size * scale)
    sed = rng.randint(1, numid)
    cen = np.array(str(lines[sed - 1]).split(), dtype=int) if isinstance(lines[sed - 1], (str, bytes)) else np.asarray(lines[sed - 1], dtype=int)
    a = min(max(minindex[0] + deps // 2, cen[0]), maxindex[0] - deps // 2 - 1)
    b = min(max(minindex[1] + rows // 2, cen[1]), maxindex[1] - rows // 2 - 1)
    c = min(max(minindex[2] + cols // 2, cen[2]), maxindex[2] - cols // 2 - 1)
    flip_num = int(rng.randint(0, 8)) if flips else 0
    return int(a), int(b), int(c), deps // 2, rows // 2, flip_num


def flip(arr, flip_num):
    """train_hybrid.py:67-94."""
    if flip_num == 1:
        return np.flipud(arr)
    if flip_num == 2:
        return np.fliplr(arr)
    if flip_num == 3:
        return np.rot90(arr, k=1, axes=(1, 0))
    if flip_num == 4:
        return np.rot90(arr, k=3, axes=(1, 0))
    if flip_num == 5:
        return np.rot90(np.fliplr(arr), k=1, axes=(1, 0))
    if flip_num == 6:
        return np.rot90(np.fliplr(arr), k=3, axes=(1, 0))
    if flip_num == 7:
        return np.fliplr(np.flipud(arr))
    return arr


def _cubic(x, f0, f1, f2, f3):
    return f1 + 0.5 * x * (f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3 + x * (3.0 * (f1 - f2) + f3 - f0)))


def _src_coords(n_in, n_out):
    return (float(n_in) / n_out) * (np.arange(n_out, dtype=np.float64) + 0.5) - 0.5


def resize_nearest_edge(img, out_h, out_w):
    """resize(order=0, mode='edge', preserve_range=True) of an (h, w, s) array to (out_h, out_w, s)."""
    h, w = img.shape[:2]
    r = np.clip(np.floor(_src_coords(h, out_h) + 0.5).astype(np.int64), 0, h - 1)      # C round(): half away from zero (>= -0.5 here)
    c = np.clip(np.floor(_src_coords(w, out_w) + 0.5).astype(np.int64), 0, w - 1)
    return img[r][:, c].astype(np.float64)


def resize_cubic_constant(img, out_h, out_w, cval=0.0, clip=True):
    """resize(order=3, mode='constant', cval=0, clip=True, preserve_range=True) of an (h, w, s) array."""
    img = img.astype(np.float64)
    h, w = img.shape[:2]
    pad = np.full((h + 4, w + 4) + img.shape[2:], cval, dtype=np.float64)
    pad[2:-2, 2:-2] = img

    def taps(n_in, n_out):
        src = _src_coords(n_in, n_out)
        i0 = np.floor(src).astype(np.int64)
        return i0, src - i0

    r0, xr = taps(h, out_h)
    c0, xc = taps(w, out_w)
    xr = xr.reshape((-1, 1) + (1,) * (img.ndim - 2))
    xc = xc.reshape((1, -1) + (1,) * (img.ndim - 2))
    rows = []
    for dr in range(-1, 3):
        line = pad[r0 + dr + 2]                                     # (out_h, w+4, s)
        f = [line[:, c0 + dc + 2] for dc in range(-1, 3)]           # each (out_h, out_w, s)
        rows.append(_cubic(xc, *f))
    out = _cubic(xr, *rows)
    if clip:
        lo, hi = img.min(), img.max()
        keep = None if lo <= cval <= hi else (out == cval)
        out = np.clip(out, lo, hi)
        if keep is not None:
            out[keep] = cval
    return out


def make_sample(img, tumor, a, b, c, half_d, half_r, cols, flip_num, mean, input_size, two_d=False):
    """train_hybrid.py:60-98 (two_d=False) / train_2ddense.py:59-69 (two_d=True) for given draws.
    `img` float32 (H, W, S), `tumor` integer (H, W, S).  Returns (X float32 (size,size,cols), Y int16 (size,size,cols) or
    (size,size) for the 2-D script)."""
    lo_c, hi_c = (c - cols // 2, c + cols // 2 + 1) if two_d else (c - cols // 2, c + cols // 2)
    ci = img[a - half_d:a + half_d, b - half_r:b + half_r, lo_c:hi_c].astype(np.float32).copy()
    ct = tumor[a - half_d:a + half_d, b - half_r:b + half_r, lo_c:hi_c].copy()
    ci -= np.float32(mean)
    ci, ct = flip(ci, flip_num), flip(ct, flip_num)
    y = resize_nearest_edge(ct, input_size, input_size)
    x = resize_cubic_constant(ci, input_size, input_size)
    y = y.astype(np.int16)
    return x.astype(np.float32), (y[:, :, 1] if two_d else y)


def batch_has_all_classes(Y):
    """train_hybrid.py:126-131: a batch missing one of the three classes is discarded and redrawn."""
    return bool(np.sum(Y == 0)) and bool(np.sum(Y == 1)) and bool(np.sum(Y == 2))
