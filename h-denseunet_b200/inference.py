"""Sliding-window inference (lib/funcs.py:4-51, the test.py path) with the accumulation on the GPU.

Multi-GPU (BASELINE config 5): the windows of one volume are independent, so with one process per GPU
(torch.distributed initialised) the window list is cut into `world` contiguous z-ranges, every rank evaluates its own
range into its own score / count accumulators with NO collective inside the loop, and ONE sum-all-reduce of the
accumulators at the end merges the <= 4 overlapping boundary slices of neighbouring ranges (all other slices are
non-zero on exactly one rank).  Every rank returns the full result, like the single-process call.

2-D slice reuse (opt-in: reuse_2d=True or HDN_WINDOW_REUSE=1): consecutive windows overlap by 75 %, and
in inference every BatchNorm uses moving statistics, so the 2-D network's result for a slice depends only on that
slice's (z-1, z, z+1) triplet.  Interior slices of a window (triplet not clamped, hybridnet.py:388-391) are therefore
identical in every window that contains them as an interior slice; only the two edge slices (clamped triplets,
hybridnet.py:385-387,392-395) are window specific.  With reuse, a window after the first evaluates the 2-D network on
step + 2 slices (the new interior slices and the two edges) instead of all `cols`, moves the still-valid results inside
the hybrid program's own 2-D output buffers and runs the 3-D part only: 8 -> 4 slice evaluations per window at the
reference's 8-slice / stride-2 setting.
"""
import os

import numpy as np
import torch

from . import engine, models


def window_starts(z, mini_z, maxi_z, cols):
    """lib/funcs.py:12,19-27: window start list incl. the clamped tail window (py2 integer division)."""
    step = cols // 4
    right = int(min(z, maxi_z + 10) - cols)
    left = max(0, min(mini_z - 5, right))
    return [z - cols if c > z - cols else c for c in range(left, right + step, step)]


def shard_windows(starts, world, rank):
    """Contiguous z-range of the window list owned by `rank`: ceil(n / world) windows per rank, in z order, so a
    rank touches one z-interval of the volume plus a (cols - step)-slice halo shared with its neighbours."""
    n = len(starts)
    per = (n + world - 1) // world if world > 0 else n
    return starts[rank * per:(rank + 1) * per]


def _dist_info(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(group), dist.get_rank(group)
    return None, 1, 0


def new_slices(cols, delta):
    """Window-relative slice indices whose 2-D result must be computed when the window moved by `delta` slices:
    the two edge slices (clamped triplets) and the `delta` interior slices the previous window did not hold as
    interior slices.  The interior slices 1 .. cols-2-delta are reused from position s + delta."""
    return [0] + list(range(cols - 1 - delta, cols - 1)) + [cols - 1]


class SliceReuse(object):
    """The 2-D network on step + 2 slices, sharing the model's parameters, plus the bookkeeping that splices its
    results into the hybrid program's 2-D output buffers (the 3-D part's inputs, read in place)."""

    def __init__(self, model, net, size, cols):
        assert model.kind == "hybrid" and model.b == 1
        self.model, self.net, self.size, self.cols = model, net, size, cols
        self.R = cols // 4 + 2
        self.net2 = engine.Net(model.params, net.device, False, model.precision, backend=model.backend, dropout=False)
        models.unet2d_net(self.net2, self.R, size, size, model.mode)
        self.net2.compile()
        self.in2 = list(self.net2.inputs.values())[0]
        self.feat_h = net.outputs["feature2d"].view.buf.data         # (cols, 1, H, W, 64)  conv_up4 output (pre bn_up4)
        self.log_h = net.outputs["logits2d"].buf.data                 # (cols, 1, H, W, 3)
        self.feat_2 = self.net2.outputs["feature"].view.buf.data     # (R, 1, H, W, 64)
        self.log_2 = self.net2.outputs["logits"].buf.data
        self.i3d = next(i for i, op in enumerate(net.ops) if isinstance(op, engine.Cat4Op))
        self.prev = None
        self.stage = np.zeros((self.R, size, size, 3), np.float32)
        self.slices_evaluated = 0

    def window(self, imgs, c, box):
        """Evaluate the window starting at volume slice c; returns after the hybrid logits buffer is filled."""
        model, net, cols = self.model, self.net, self.cols
        delta = None if self.prev is None else c - self.prev
        self.prev = c
        model._upload(net, box)
        if delta is None or delta <= 0 or delta + 2 > self.R:
            net.forward()                                  # first window (or an irregular step): everything
            self.slices_evaluated += cols
            return
        keep = cols - 2 - delta                            # interior slices that stay valid, now at positions 1 .. keep
        if keep > 0:
            for t in (self.feat_h, self.log_h):
                t[1:1 + keep].copy_(t[1 + delta:1 + delta + keep].clone())
        todo = new_slices(cols, delta)
        todo = todo + [todo[-1]] * (self.R - len(todo))    # a shorter tail step: pad the batch with a repeat
        size = self.size
        for r, s in enumerate(todo):
            for k in range(3):
                sk = min(max(s - 1 + k, 0), cols - 1)      # hybridnet.py:385-395: triplet clamped to the window
                self.stage[r, :, :, k] = imgs[0:size, 0:size, c + sk]
        model._h2d(self.stage, self.in2.data, three_d=False)
        self.net2.forward()
        for r, s in enumerate(todo):
            self.feat_h[s].copy_(self.feat_2[r])
            self.log_h[s].copy_(self.log_2[r])
        net.accum.zero_()
        for op in net.ops[self.i3d:]:
            op.forward()
        self.slices_evaluated += len(set(todo))


def predict_tumor_inwindow(model, imgs_test, num, mini, maxi, args, group=None, reuse_2d=None, stats=None):
    """Drop-in for lib.funcs.predict_tumor_inwindow: returns (score[..., num-2], score[..., num-1]).
    Windows are evaluated by the engine; soft-max, edge-slice drop and overlap averaging
    (funcs.py:31-48) are hdn_window_accumulate / hdn_window_finalize on the device.
    `stats` (optional dict) receives the number of windows and of 2-D slice evaluations of this rank."""
    assert num == 3, "the reference always calls with num=3 (test.py:68)"
    size, cols = args.input_size, args.input_cols
    x, y, z = imgs_test.shape
    net = model._net(False)
    be, dev = net.be, net.device
    HW = size * size
    score = torch.zeros((z, size, size, 2), dtype=torch.float32, device=dev)
    count = torch.zeros((z,), dtype=torch.int32, device=dev)
    box = np.zeros((model.b, size, size, cols, 1), np.float32)
    logits = net.outputs["logits"]
    dist, world, rank = _dist_info(group)
    starts = window_starts(z, int(mini[2]), int(maxi[2]), cols)
    if reuse_2d is None:
        reuse_2d = os.environ.get("HDN_WINDOW_REUSE", "0") not in ("", "0")
    reuse = SliceReuse(model, net, size, cols) if (reuse_2d and model.kind == "hybrid" and model.b == 1 and cols >= 8) else None
    mine = shard_windows(starts, world, rank)
    for c in mine:
        box[0, :, :, :, 0] = imgs_test[0:size, 0:size, c:c + cols]
        if reuse is not None:
            reuse.window(imgs_test, c, box)
        else:
            model._upload(net, box)
            net.forward()
        be.window_accumulate(logits.buf.data, score, count, cols, HW, c)
    if stats is not None:
        stats["windows"] = len(mine)
        stats["slices_2d"] = reuse.slices_evaluated if reuse is not None else len(mine) * cols
    if world > 1:
        # the one exchange of the whole volume: sum the per-rank accumulators (overlap slices at the range boundaries)
        dist.all_reduce(score, group=group)
        dist.all_reduce(count, group=group)
    be.window_finalize(score, count, z, HW)
    s = score.permute(1, 2, 0, 3).cpu().numpy()          # (z,H,W,2) -> (H,W,z,2)
    out1 = np.zeros((x, y, z), np.float32)
    out2 = np.zeros((x, y, z), np.float32)
    out1[0:size, 0:size] = s[..., 0]
    out2[0:size, 0:size] = s[..., 1]
    return out1, out2
