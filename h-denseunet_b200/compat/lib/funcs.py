"""Shim for lib/funcs.py: predict_tumor_inwindow (:4-51) on the engine, and the post-processing helpers
get_binary_mask / GeneSeglivertumor (:131-153) on the GPU post-processing kernels (csrc/postproc.cu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _root  # noqa: F401,E402
import numpy as np  # noqa: E402
from h_denseunet_b200 import predict_tumor_inwindow  # noqa: F401,E402


def GeneSeglivertumor(score):
    """lib/funcs.py:138-153: threshold 0.5, largest connected component (26-connectivity)."""
    import torch
    from h_denseunet_b200.postprocess import PostProcessor
    score = np.ascontiguousarray(score, dtype=np.float32)
    pp = PostProcessor(score.shape)
    s = torch.from_numpy(score).to(pp.dev)
    liver, _ = pp.threshold(s, torch.zeros_like(s), 0.5, 2.0)        # second threshold can never fire
    return pp.largest_component(liver).cpu().numpy().astype(np.int64)


def get_binary_mask(score, id):  # noqa: A002  (reference signature, lib/funcs.py:131)
    return np.int16(GeneSeglivertumor(score))
