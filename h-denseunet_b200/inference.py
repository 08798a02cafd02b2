"""Sliding-window inference (lib/funcs.py:4-51, the test.py path) with the accumulation on the GPU.

Multi-GPU (BASELINE config 5): the windows of one volume are independent, so with one process per GPU
(torch.distributed initialised) the window list is cut into `world` contiguous z-ranges, every rank evaluates its own
range into its own score / count accumulators with NO collective inside the loop, and ONE sum-all-reduce of the
accumulators at the end merges the <= 6 overlapping boundary slices of neighbouring ranges (all other slices are
non-zero on exactly one rank).  Every rank returns the full result, like the single-process call.
"""
import numpy as np
import torch


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


def predict_tumor_inwindow(model, imgs_test, num, mini, maxi, args, group=None):
    """Drop-in for lib.funcs.predict_tumor_inwindow: returns (score[..., num-2], score[..., num-1]).
    Windows are evaluated by the engine; soft-max, edge-slice drop and overlap averaging
    (funcs.py:31-48) are hdn_window_accumulate / hdn_window_finalize on the device."""
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
    for c in shard_windows(starts, world, rank):
        box[0, :, :, :, 0] = imgs_test[0:size, 0:size, c:c + cols]
        model._upload(net, box)
        net.forward()
        be.window_accumulate(logits.buf.data, score, count, cols, HW, c)
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
