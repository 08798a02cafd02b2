"""Time (CUDA events) one convolution through the C-ABI at a realistic size; used under ncu for the
per-kernel captures kept in profiles/.  python scripts/prof_conv.py <case> <pass> [reps] [precision 1|2]
(precision 1 = bf16 operands, 2 = bf16x3; a library built with HDN_NVCC_EXTRA=-DHDN_TC_TIMING also prints the per-role
wait / work cycle counters of CTA 0)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from test_gpu_tc import Case  # noqa: E402

CASES = {
    # name: (kwargs, description)
    "3dconv_up4": dict(N=1, D=8, H=512, W=512, cin=96, cout=64, k=(3, 3, 3), ups=((2, 2, 2),), fold=(True,), bias=True, stats=True,
                       src_pad=0, out_pad=0),
    "conv_up4": dict(N=8, D=1, H=512, W=512, cin=96, cout=64, k=(1, 3, 3), ups=((1, 2, 2),), fold=(True,), bias=True, src_pad=0, out_pad=0),
    "fianl_conv": dict(N=1, D=8, H=512, W=512, cin=64, cout=64, k=(3, 3, 3), ups=((1, 1, 1), (1, 1, 1)), fold=(True, True), bias=True,
                       stats=True, src_pad=0, out_pad=0),
    "dense2_x2": dict(N=48, D=1, H=128, W=128, cin=192, cout=48, k=(1, 3, 3), fold=(True,), src_pad=0, out_pad=288),
    "dense2_x1": dict(N=48, D=1, H=128, W=128, cin=240, cout=192, k=(1, 1, 1), fold=(True,), src_pad=144, out_pad=0),
    "dense4_x1": dict(N=48, D=1, H=32, W=32, cin=1200, cout=192, k=(1, 1, 1), fold=(True,), src_pad=912, out_pad=0),
    "dense4_x2": dict(N=48, D=1, H=32, W=32, cin=192, cout=48, k=(1, 3, 3), fold=(True,), src_pad=0, out_pad=2064),
    "3ddense2_x2": dict(N=1, D=12, H=128, W=128, cin=128, cout=32, k=(3, 3, 3), fold=(True,), src_pad=0, out_pad=160),
}


def headline(kw):
    """The same layer at the headline batch: 2 slabs of 48 slices (3-D layers: N=2, D=48; 2-D layers: N=96)."""
    kw = dict(kw)
    if kw["D"] > 1:
        kw["N"], kw["D"] = 2, 48
    else:
        kw["N"] = 96
    return kw


def main():
    name, which = sys.argv[1], sys.argv[2]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    prec = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda:0")
    kw = CASES[name]
    full = os.environ.get("HDN_PROF_FULL", "0") not in ("", "0")
    if full:
        kw = headline(kw)
    c = Case(dev, tc=prec, device_fill=full, **kw)
    op, d, net = c.op, c.op.desc, c.net
    flops = 2.0 * d.out.M * d.Cin * d.Cout * d.k[0] * d.k[1] * d.k[2]
    if which == "fprop":
        fn = lambda: (op._set_prec(0), net.be.conv_fprop(d))
        op.prec = [prec] * 3
    elif which == "wgrad":
        op.prec = [prec] * 3
        fn = lambda: (op._set_prec(2), net.be.conv_wgrad(d, op.w.g, None))
    else:
        from h_denseunet_b200.engine import EpiDesc
        op.prec = [prec] * 3
        epis = []
        for s, b in zip(op.srcs, c.src_bufs):
            v = s.act.view
            S = torch.zeros((2, v.C), dtype=torch.float64, device=dev)
            epis.append(EpiDesc(0, True, dx=v, s=S, center=s.act.fold.mean if s.act.fold is not None else None))
        flops *= len(epis)
        fn = lambda: (op._set_prec(1), net.be.conv_dgrad(d, epis))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print("%s %s: %.3f ms  %.1f TFLOP/s  (M=%d Cin=%d Cout=%d taps=%d)" % (name, which, ms, flops / ms / 1e9, d.out.M, d.Cin, d.Cout,
                                                                          d.k[0] * d.k[1] * d.k[2]))


if __name__ == "__main__":
    main()
