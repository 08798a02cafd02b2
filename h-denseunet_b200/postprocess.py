"""GPU post-processing of the sliding-window result: test.py:71-115 with the scipy / skimage calls
replaced by libhdn kernels (csrc/postproc.cu).  numpy in / numpy out, volumes in the reference's (H, W, Z) layout."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class PostProcessor(object):
    """Device buffers for one volume shape; every method takes / returns uint8 device tensors of that shape."""

    def __init__(self, shape, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the post-processing kernels have no CPU path")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = tuple(int(s) for s in shape)
        self.n = int(np.prod(self.shape))
        self.ws = torch.empty(8 * self.n + 8, dtype=torch.uint8, device=self.dev)

    @property
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def _new(self):
        return torch.empty(self.shape, dtype=torch.uint8, device=self.dev)

    def threshold(self, score_liver, score_tumor, thres_liver, thres_tumor):
        a, b = self._new(), self._new()
        _lib.check(self.lib.hdn_post_threshold(score_liver.data_ptr(), score_tumor.data_ptr(), a.data_ptr(), b.data_ptr(), self.n,
                                               float(thres_liver), float(thres_tumor), self._st), "hdn_post_threshold")
        return a, b

    def dilate(self, x):
        o = self._new()
        _lib.check(self.lib.hdn_post_dilate(x.data_ptr(), o.data_ptr(), *self.shape, self._st), "hdn_post_dilate")
        return o

    def largest_component(self, x):
        o = self._new()
        _lib.check(self.lib.hdn_post_largest_component(x.data_ptr(), o.data_ptr(), *self.shape, self.ws.data_ptr(), self.ws.numel(), self._st),
                   "hdn_post_largest_component")
        return o

    def fill_holes(self, x):
        o = self._new()
        _lib.check(self.lib.hdn_post_fill_holes(x.data_ptr(), o.data_ptr(), *self.shape, self.ws.data_ptr(), self.ws.numel(), self._st),
                   "hdn_post_fill_holes")
        return o

    def logical_and(self, a, b):
        o = self._new()
        _lib.check(self.lib.hdn_post_and(a.data_ptr(), b.data_ptr(), o.data_ptr(), self.n, self._st), "hdn_post_and")
        return o

    def compose(self, liver, tumor):
        o = self._new()
        _lib.check(self.lib.hdn_post_compose(liver.data_ptr(), tumor.data_ptr(), o.data_ptr(), self.n, self._st), "hdn_post_compose")
        return o


def postprocess_scores(score1, score2, mask, thres_liver=0.5, thres_tumor=0.9, device=None):
    """test.py:71-115.  score1 / score2: the class-1 / class-2 probability volumes predict_tumor_inwindow returns;
    mask: the stage-1 liver mask as test.py holds it at line 63 (labels merged, dilated once).  Returns the uint8
    segmentation (0 background, 1 liver, 2 tumour) that test.py saves."""
    score1 = np.ascontiguousarray(score1, dtype=np.float32)
    score2 = np.ascontiguousarray(score2, dtype=np.float32)
    pp = PostProcessor(score1.shape, device)
    s1 = torch.from_numpy(score1).to(pp.dev)
    s2 = torch.from_numpy(score2).to(pp.dev)
    m = torch.from_numpy(np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)).to(pp.dev)
    result1, result2 = pp.threshold(s1, s2, thres_liver, thres_tumor)                 # :73-77
    liver_res = pp.largest_component(result1)                                         # :81-91
    liver_labels = pp.fill_holes(pp.largest_component(pp.dilate(m)))                  # :94-104
    segmask = pp.fill_holes(pp.logical_and(result2, liver_labels))                    # :107-109
    liver_res = pp.fill_holes(liver_res)                                              # :112
    out = pp.compose(liver_res, segmask)                                              # :113-114
    return out.cpu().numpy()
