"""Training-sample pipeline with the volumes resident on the GPU.

Reference: `load_seq_crop_data_masktumor_try` + `generate_arrays_from_file` (train_hybrid.py:40-133,
train_2ddense.py:40-126): per sample a liver- or tumour-centred crop at a random scale in [0.8, 1.2), mean subtraction,
one of 8 flips / rotations (hybrid script), cubic / nearest `resize` to the network size, on a 14-thread pool; a hybrid
batch that misses one of the three classes is discarded.  At B200 step times that host work (and the 100 MB upload per
slab) is the bottleneck, so here the training set is uploaded ONCE (`DeviceVolumes`, slice-major), the random draws are
made on the host in the reference's np.random call order, and one kernel per sample (`hdn_aug_sample`, csrc/augment.cu)
gathers, flips, resamples and writes the engine's input layout.  `CropGenerator` is a generator in the reference's sense:
`model.fit_generator(CropGenerator(...), steps_per_epoch, epochs)`; it yields `DeviceBatch` objects that the model
consumes without any host copy.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

LIVERLIST = (32, 34, 38, 41, 47, 87, 89, 91, 105, 106, 114, 115, 119)      # train_hybrid.py:39: volumes sampled liver-only


def flip_map(flip_num, h, w):
    """(m00, m01, m10, m11, o0, o1) with  flipped[i][j] = crop[m00*i + m01*j + o0][m10*i + m11*j + o1]  for the eight cases of
    train_hybrid.py:67-94, derived from numpy's own flipud / fliplr / rot90 applied to index grids (views, O(1))."""
    if flip_num in (3, 4, 5, 6) and h != w:
        raise ValueError("rotations need a square crop, got %dx%d" % (h, w))
    ii = np.broadcast_to(np.arange(h)[:, None], (h, w))
    jj = np.broadcast_to(np.arange(w)[None, :], (h, w))

    def f(a):
        if flip_num == 1:
            return np.flipud(a)
        if flip_num == 2:
            return np.fliplr(a)
        if flip_num == 3:
            return np.rot90(a, k=1, axes=(1, 0))
        if flip_num == 4:
            return np.rot90(a, k=3, axes=(1, 0))
        if flip_num == 5:
            return np.rot90(np.fliplr(a), k=1, axes=(1, 0))
        if flip_num == 6:
            return np.rot90(np.fliplr(a), k=3, axes=(1, 0))
        if flip_num == 7:
            return np.fliplr(np.flipud(a))
        if flip_num == 0:
            return a
        raise ValueError("flip_num must be in 0..7")

    fi, fj = f(ii), f(jj)
    o0, o1 = int(fi[0, 0]), int(fj[0, 0])
    m00 = int(fi[1, 0]) - o0 if fi.shape[0] > 1 else 1
    m01 = int(fi[0, 1]) - o0 if fi.shape[1] > 1 else 0
    m10 = int(fj[1, 0]) - o1 if fj.shape[0] > 1 else 0
    m11 = int(fj[0, 1]) - o1 if fj.shape[1] > 1 else 1
    return m00, m01, m10, m11, o0, o1


def _centre(line):
    if isinstance(line, (str, bytes)):
        return np.array(line.split(), dtype=int)            # np.fromstring(cen, dtype=int, sep=' ')
    return np.asarray(line, dtype=int)


def draw_crop(rng, input_size, cols, lines, numid, minindex, maxindex, flips):
    """train_hybrid.py:47-60,67 / train_2ddense.py:47-58: the draws of one sample in the reference's order; Python-2 `/` on
    ints is floor division.  Returns (a, b, c, half, flip_num): the crop is rows [a-half, a+half), columns [b-half, b+half)."""
    scale = rng.uniform(0.8, 1.2)
    deps = int(input_size * scale)
    sed = rng.randint(1, numid)
    cen = _centre(lines[sed - 1])
    half = deps // 2
    a = min(max(minindex[0] + half, cen[0]), maxindex[0] - half - 1)
    b = min(max(minindex[1] + half, cen[1]), maxindex[1] - half - 1)
    c = min(max(minindex[2] + cols // 2, cen[2]), maxindex[2] - cols // 2 - 1)
    flip_num = int(rng.randint(0, 8)) if flips else 0
    return int(a), int(b), int(c), half, flip_num


class DeviceVolumes:
    """The training set on the device: per volume a slice-major (S, H, W) image (float32, or int16 when the values are
    integral and `compact`) and a uint8 label volume, plus the host-side sampling tables of `load_fast_files`
    (train_hybrid.py:136-181): liver / tumour voxel lists and the liver box grown by 3 voxels."""

    def __init__(self, device="cuda:0", compact=False):
        self.device = torch.device(device)
        self.compact = compact
        self.img, self.seg, self.shape = [], [], []
        self.liverlines, self.tumorlines, self.minindex, self.maxindex = [], [], [], []

    def add(self, img, seg, liver_voxels, tumor_voxels, box):
        """img, seg: (H, W, S) host arrays; liver_voxels / tumor_voxels: the rows of LiverPixels / TumorPixels txt files
        (strings "x y z" or int triples); box: the six numbers of LiverBox/box_i.txt (min xyz, max xyz)."""
        img = np.asarray(img)
        seg = np.asarray(seg)
        if img.shape != seg.shape or img.ndim != 3:
            raise ValueError("image %s and segmentation %s must be 3-D arrays of one shape" % (img.shape, seg.shape))
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(img, dtype=np.float32).transpose(2, 0, 1)))
        if self.compact:
            t16 = t.to(torch.int16)
            if not torch.equal(t16.to(torch.float32), t):
                raise ValueError("compact storage needs integral intensities inside the int16 range")
            t = t16
        self.img.append(t.to(self.device))
        self.seg.append(torch.from_numpy(np.ascontiguousarray(np.asarray(seg).astype(np.uint8).transpose(2, 0, 1))).to(self.device))
        self.shape.append(img.shape)
        box = np.asarray(box, dtype=float).reshape(-1)
        mn, mx = np.array(box[0:3], dtype=int), np.array(box[3:6], dtype=int)
        for k in range(3):                                   # train_hybrid.py:153-158
            mn[k] = max(mn[k] - 3, 0)
            mx[k] = min(img.shape[k], mx[k] + 3)
        self.minindex.append(mn)
        self.maxindex.append(mx)
        self.liverlines.append(list(liver_voxels))
        self.tumorlines.append(list(tumor_voxels))
        return len(self.img) - 1

    def __len__(self):
        return len(self.img)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.img) + sum(t.numel() for t in self.seg)


class DeviceBatch:
    """One batch in the engine's device layout: x (N, S, H, W) fp32 [hybrid / 3-D] or (N, H, W, 4) [2-D, 3 slices as
    channels + the zero pad channel], y (N, S, H, W) / (N, H, W) fp32 class indices; `ready` is recorded on the producing
    stream.  `counts` = class histogram of y (host ints)."""

    is_device_batch = True

    def __init__(self, x, y, ready, counts, params, slot=None):
        self.x, self.y, self.ready, self.counts, self.params = x, y, ready, counts, params
        self._slot = slot

    def release(self, used_event):
        """Called by the consumer after it enqueued its last read of x / y: the producer waits for `used_event` before it
        overwrites this slot."""
        if self._slot is not None:
            self._slot[2] = used_event

    def __len__(self):
        return self.x.shape[0]

    def host(self):
        """(X, Y) as the reference's generator yields them (numpy; hybrid: (N,H,W,S,1) float32 / int16)."""
        self.ready.synchronize()
        if self.x.dim() == 4 and self.x.shape[-1] == 4 and self.y.dim() == 3:
            return self.x[..., :3].cpu().numpy(), self.y.cpu().numpy().astype(np.int16)[..., None]
        return (self.x.permute(0, 2, 3, 1).unsqueeze(-1).cpu().numpy(),
                self.y.permute(0, 2, 3, 1).unsqueeze(-1).cpu().numpy().astype(np.int16))


class CropGenerator:
    """`generate_arrays_from_file` (train_hybrid.py:100-133; two_d=True: train_2ddense.py:70-100) on the device.

    rng: a np.random.RandomState (or the np.random module, which is what the reference draws from); `choice` is the
    volume draw (`np.random.choice(trainidx)` in the hybrid script, `random.choice` in the 2-D one: pass your own).
    Two output slots alternate so that batch k+1 is produced (on its own stream) while step k consumes batch k."""

    def __init__(self, volumes, batch_size, input_size, input_cols, mean, two_d=False, rng=None, trainidx=None, liverlist=LIVERLIST,
                 reject_missing_class=None, choice=None):
        if not len(volumes):
            raise ValueError("no volumes")
        self.v, self.b, self.size, self.cols, self.mean, self.two_d = volumes, int(batch_size), int(input_size), int(input_cols), float(mean), two_d
        self.rng = rng if rng is not None else np.random
        self.trainidx = list(trainidx) if trainidx is not None else list(range(len(volumes)))
        self.liverlist = set(liverlist)
        self.reject = (not two_d) if reject_missing_class is None else reject_missing_class
        self.choice = choice or (lambda idx: self.rng.choice(idx))
        dev = volumes.device
        self.lib = _lib.load()
        self.stream = torch.cuda.Stream(device=dev)
        self.slots = []
        with torch.cuda.stream(self.stream):
            for _ in range(2):
                if two_d:
                    x = torch.zeros((self.b, self.size, self.size, 4), dtype=torch.float32, device=dev)
                    y = torch.zeros((self.b, self.size, self.size), dtype=torch.float32, device=dev)
                else:
                    x = torch.zeros((self.b, self.cols, self.size, self.size), dtype=torch.float32, device=dev)
                    y = torch.zeros((self.b, self.cols, self.size, self.size), dtype=torch.float32, device=dev)
                self.slots.append([x, y, None])
            self.counts = torch.zeros(4, dtype=torch.int32, device=dev)
            self.scratch = torch.zeros(4, dtype=torch.int32, device=dev)
        self.k = 0
        self.rejected = 0

    def __iter__(self):
        return self

    def draw(self):
        """One batch of parameter records, draws in the order of train_hybrid.py:105-118 + :47-67."""
        recs = []
        for _ in range(self.b):
            count = int(self.choice(self.trainidx))
            num = self.rng.randint(0, 6)
            if num < 3 or count in self.liverlist:
                lines = self.v.liverlines[count]
            else:
                lines = self.v.tumorlines[count]
            recs.append((count, lines, len(lines)))
        out = []
        cols = 3 if self.two_d else self.cols
        for count, lines, numid in recs:                      # pool.map keeps the list order (train_hybrid.py:120)
            a, b, c, half, flip_num = draw_crop(self.rng, self.size, cols, lines, numid, self.v.minindex[count], self.v.maxindex[count],
                                                flips=not self.two_d)
            out.append((count, a, b, c, half, flip_num))
        return out

    def _launch(self, n, rec, x, y):
        count, a, b, c, half, flip_num = rec
        H, W, S = self.v.shape[count]
        cols = 3 if self.two_d else self.cols
        c_lo = c - cols // 2
        cs = cols // 2 * 2 + 1 if self.two_d else cols // 2 * 2
        if cs != cols:
            raise ValueError("input_cols must be even for the hybrid pipeline")
        g = _lib.Aug()
        img = self.v.img[count]
        g.vol, g.seg, g.vol_i16 = img.data_ptr(), self.v.seg[count].data_ptr(), 1 if img.dtype == torch.int16 else 0
        g.VS, g.VH, g.VW = S, H, W
        g.a0, g.b0, g.c0 = a - half, b - half, c_lo
        g.ch, g.cw, g.cs = 2 * half, 2 * half, cs
        g.m00, g.m01, g.m10, g.m11, g.o0, g.o1 = flip_map(flip_num, 2 * half, 2 * half)
        g.mean = self.mean
        g.out_h = g.out_w = self.size
        if self.two_d:
            g.xs_s, g.xs_h, g.xs_w = 1, self.size * 4, 4
            g.ys0, g.yns = 1, 1                                # cropp_tumor[:, :, 1] (train_2ddense.py:69)
        else:
            g.xs_s, g.xs_h, g.xs_w = self.size * self.size, self.size, 1
            g.ys0, g.yns = 0, cs
        _lib.check(self.lib.hdn_aug_sample(C.byref(g), x[n].data_ptr(), y[n].data_ptr(), self.counts.data_ptr(), self.scratch.data_ptr(),
                                           self.stream.cuda_stream), "hdn_aug_sample")

    def __next__(self):
        slot = self.slots[self.k % 2]
        x, y, used = slot
        while True:
            recs = self.draw()
            with torch.cuda.stream(self.stream):
                if used is not None:
                    self.stream.wait_event(used)              # the step that consumed this slot two batches ago
                self.counts.zero_()
                for n, rec in enumerate(recs):
                    self._launch(n, rec, x, y)
                ready = torch.cuda.Event()
                ready.record(self.stream)
                counts = self.counts.cpu().tolist()[:3] if self.reject else None
            if self.reject and min(counts) == 0:              # train_hybrid.py:126-131
                self.rejected += 1
                continue
            break
        self.k += 1
        return DeviceBatch(x, y, ready, counts, recs, slot)

    next = __next__
