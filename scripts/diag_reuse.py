"""Diagnostic for inference.SliceReuse on the CUDA engine: where do the spliced 2-D results differ from a full evaluation?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import h_denseunet_b200 as hdn  # noqa: E402
from h_denseunet_b200 import inference  # noqa: E402
from util import Args, perturb_params  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
    if os.environ.get("HDN_DIAG_DIRTY"):
        junk = [torch.full((1 << 28,), float("nan"), device="cuda") for _ in range(8)]      # 8 GiB of NaN through the caching allocator
        del junk
    size, cols = 32, 8
    a = Args(b=1, input_size=size, input_cols=cols)
    m = hdn.dense_rnn_net(a, precision=prec)
    perturb_params(m)
    rng = np.random.default_rng(11)
    vol = rng.normal(0, 60, (size, size, 23)).astype(np.float32)
    net = m._net(False)
    box = np.zeros((1, size, size, cols, 1), np.float32)

    def full(c):
        box[0, :, :, :, 0] = vol[:, :, c:c + cols]
        m._upload(net, box)
        net.forward()
        torch.cuda.synchronize()
        return (net.outputs["feature2d"].view.buf.data.clone(), net.outputs["logits2d"].buf.data.clone(),
                net.outputs["logits"].buf.data.clone())

    f0, l0, o0 = full(0)
    f2, l2, o2 = full(2)
    # interior slices of window 2 at positions 1..4 were interior slices 3..6 of window 0
    print("[%s] interior reuse assumption: feat %.3e logits %.3e" % (prec, float((f2[1:5] - f0[3:7]).abs().max()), float((l2[1:5] - l0[3:7]).abs().max())))
    # the 4-slice program on window 2's new slices
    r = inference.SliceReuse(m, net, size, cols)
    todo = inference.new_slices(cols, 2)
    for i, s in enumerate(todo):
        for k in range(3):
            sk = min(max(s - 1 + k, 0), cols - 1)
            r.stage[i, :, :, k] = vol[:, :, 2 + sk]
    m._h2d(r.stage, r.in2.data, three_d=False)
    r.net2.forward()
    torch.cuda.synchronize()
    for i, s in enumerate(todo):
        print("   slice %d: 4-slice program vs full window: feat %.3e logits %.3e" % (
            s, float((r.feat_2[i] - f2[s]).abs().max()), float((r.log_2[i] - l2[s]).abs().max())))
    # input buffers equal?
    in_h = [b for b in net.buffers if b.name == "input2d"][0].data
    for i, s in enumerate(todo):
        print("   slice %d: staged triplet vs hybrid triplet %.3e" % (s, float((r.in2.data[i] - in_h[s]).abs().max())))
    # now the real path
    st0, st1 = {}, {}
    mini, maxi = np.array([0, 0, 4]), np.array([31, 31, 15])
    s1, _ = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a, reuse_2d=False, stats=st0)
    r1, _ = hdn.predict_tumor_inwindow(m, vol, 3, mini, maxi, a, reuse_2d=True, stats=st1)
    d = np.abs(s1 - r1)
    print("   NaN in full result: %d, in reuse result: %d" % (int(np.isnan(s1).sum()), int(np.isnan(r1).sum())))
    d = np.nan_to_num(d, nan=9.0)
    print("   predict_tumor_inwindow: max diff %.3e; per-slice max %s" % (d.max(), np.round(d.max(axis=(0, 1)), 4).tolist()))


if __name__ == "__main__":
    main()
