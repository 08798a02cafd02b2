"""Sliding-window inference (lib/funcs.py:4-51, the test.py path) with the accumulation on the GPU."""
import numpy as np
import torch


def window_starts(z, mini_z, maxi_z, cols):
    """lib/funcs.py:12,19-27: window start list incl. the clamped tail window (py2 integer division)."""
    step = cols // 4
    right = int(min(z, maxi_z + 10) - cols)
    left = max(0, min(mini_z - 5, right))
    return [z - cols if c > z - cols else c for c in range(left, right + step, step)]


def predict_tumor_inwindow(model, imgs_test, num, mini, maxi, args):
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
    for c in window_starts(z, int(mini[2]), int(maxi[2]), cols):
        box[0, :, :, :, 0] = imgs_test[0:size, 0:size, c:c + cols]
        model._upload(net, box)
        net.forward()
        be.window_accumulate(logits.buf.data, score, count, cols, HW, c)
    be.window_finalize(score, count, z, HW)
    s = score.permute(1, 2, 0, 3).cpu().numpy()          # (z,H,W,2) -> (H,W,z,2)
    out1 = np.zeros((x, y, z), np.float32)
    out2 = np.zeros((x, y, z), np.float32)
    out1[0:size, 0:size] = s[..., 0]
    out2[0:size, 0:size] = s[..., 1]
    return out1, out2
