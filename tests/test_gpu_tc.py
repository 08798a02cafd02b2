"""tcgen05 convolution kernels (precision 1: bf16 operands; precision 2: "bf16x3", operands split into a bf16
head and tail, three MMAs per K step; fp32 TMEM accumulation in both) against the fp32 FMA kernels of the same
C-ABI call on identical descriptors: fprop (output, bias, dropout mask, batch statistics), dgrad (dx / du, ReLU
mask, S1/S2 sums, up-sampling reduction, two sources) and wgrad.
Tolerance, precision 1: operands are rounded to bf16 (2^-9 relative), accumulation is fp32, so a correct kernel
sits at ~3e-3 rel-L2 from the fp32 result; the bound used is 1.5e-2.  Precision 2: head + tail carry ~16
significant bits and only the tail x tail product (2^-18) is dropped, so a correct kernel sits at ~1e-5; the
bound is 1e-4 (a missing cross term shows up at ~1e-3).  A layout / descriptor error gives O(1)."""
import numpy as np
import pytest
import torch

from h_denseunet_b200 import engine
from h_denseunet_b200.engine import Act, EpiDesc, Src
from util import rel_l2

pytestmark = pytest.mark.gpu
TOLS = {1: 1.5e-2, 2: 1e-4}
TC_NAME = {1: "bf16", 2: "bf16x3"}
PRECS = [pytest.param(1, id="bf16"), pytest.param(2, id="bf16x3")]


def _rand_fold_params(ps, rng):
    for p in ps.order:
        w = p.name.rsplit("/", 1)[1]
        if w == "gamma":
            v = rng.uniform(0.5, 1.5, p.shape)
        elif w in ("beta", "moving_mean"):
            v = rng.normal(0, 0.3, p.shape)
        elif w == "moving_variance":
            v = rng.uniform(0.5, 1.5, p.shape)
        elif w == "bias":
            v = rng.normal(0, 0.5, p.shape)
        else:
            continue
        ps.set_value(p.name, v.astype(np.float32))


class Case(object):
    """One convolution on a fresh Net: sources (optionally folded BN+ReLU, optionally up-sampled),
    output window inside a wider buffer."""

    def __init__(self, dev, N, D, H, W, cin, cout, k, ups=((1, 1, 1),), fold=(True,), bias=False, stats=False,
                 drop=0.0, src_pad=16, out_pad=32, seed=0, stride=(1, 1, 1), src_c=None, tc=1, device_fill=False):
        rng = np.random.default_rng(seed)
        self.net = net = engine.Net(engine.ParamStore(seed), dev, True, TC_NAME[tc], dropout=drop > 0)
        srcs = []
        self.src_bufs = []
        for i, up in enumerate(ups):
            b = net.buffer("src%d" % i, N, D * stride[0] // up[0], H * stride[1] // up[1], W * stride[2] // up[2],
                           src_c if src_c else cin + src_pad)
            b.requires_grad = True
            self.src_bufs.append(b)
            v = b.view(0 if src_c else src_pad, cin)
            a = net.fold(v, "bn%d" % i, "sc%d" % i, 1e-3, False, 0.99, True, True) if fold[i] else Act(v)
            srcs.append(Src(a, up))
        ob = net.buffer("out", N, D, H, W, cout + out_pad)
        self.out = ob.view(out_pad // 2, cout)
        if stats:
            ob.need_stats[self.out.coff:self.out.coff + cout] = True
        p = tuple(x // 2 for x in k)
        net.conv("cv", srcs, cout, k, s=stride, p=p, bias=bias, out=self.out, drop_rate=drop)
        self.op = [o for o in net.ops if isinstance(o, engine.ConvOp)][0]
        _rand_fold_params(net.params, rng)
        net.compile()
        if device_fill:                       # headline-size tensors: draw on the device (1.6e9 host normals take minutes)
            g = torch.Generator(device=dev).manual_seed(seed + 1)
            for b in self.src_bufs:
                b.data.normal_(generator=g)
                b.ensure_grad().zero_()
            ob.data.zero_()
            ob.ensure_grad().normal_(generator=g)
        else:
            g = torch.Generator(device="cpu").manual_seed(seed + 1)
            for b in self.src_bufs:
                b.data.copy_(torch.randn(b.data.shape, generator=g))
                b.ensure_grad().zero_()
            ob.data.zero_()
            ob.ensure_grad().copy_(torch.randn(ob.data.shape, generator=g))
        for f in net.folds:
            f.forward()
        self.supported = tuple(self.op.prec)
        self.drop = drop

    def fprop(self, prec):
        net, op = self.net, self.op
        net.accum.zero_()
        self.out.buf.data.zero_()
        op.prec = [prec] * 3
        op.forward()
        torch.cuda.synchronize()
        o = self.out
        y = o.buf.data[..., o.coff:o.coff + o.C].cpu().numpy()
        rest = torch.cat([o.buf.data[..., :o.coff].reshape(-1), o.buf.data[..., o.coff + o.C:].reshape(-1)])
        untouched = float(rest.abs().max()) if rest.numel() else 0.0
        st = None
        if o.buf.stats is not None:
            st = o.buf.stats[:, o.coff:o.coff + o.C].cpu().numpy()
        return y, st, untouched

    def dgrad(self, prec, mode, accumulate, with_sums=True):
        net, op, d = self.net, self.op, self.op.desc
        epis, outs = [], []
        for s, b in zip(op.srcs, self.src_bufs):
            v = s.act.view
            S = torch.zeros((2, v.C), dtype=torch.float64, device=net.device) if with_sums else None
            ctr = s.act.fold.mean if s.act.fold is not None else None
            if mode == 0:
                b.grad.fill_(0.25 if accumulate else 7.0)
                e = EpiDesc(0, accumulate, dx=v, s=S, center=ctr)
                outs.append((b.grad, S, v))
            else:
                du = torch.full((v.M, v.C), 0.25 if accumulate else 7.0, dtype=torch.float32, device=net.device)
                e = EpiDesc(1, accumulate, du=du, s=S, center=ctr)
                outs.append((du, S, v))
            epis.append(e)
        d.__dict__.pop("_c_epis", None)
        op.prec = [prec] * 3
        op._set_prec(1)
        net.be.conv_dgrad(d, epis)
        torch.cuda.synchronize()
        res = []
        for t, S, v in outs:
            g = t[..., v.coff:v.coff + v.C] if mode == 0 else t
            res.append((g.cpu().numpy().copy(), S.cpu().numpy().copy() if S is not None else np.ones((2, 1))))
        return res


CASES = {
    "1x1_flat": dict(N=2, D=1, H=24, W=20, cin=96, cout=192, k=(1, 1, 1)),
    "1x1_flat_tail": dict(N=1, D=3, H=7, W=7, cin=248, cout=128, k=(1, 1, 1)),
    "3x3_dense": dict(N=2, D=1, H=20, W=12, cin=192, cout=48, k=(1, 3, 3)),
    "3x3x3_dense": dict(N=1, D=3, H=16, W=8, cin=128, cout=32, k=(3, 3, 3)),
    "3x3_up_skip_bias_stats": dict(N=1, D=1, H=32, W=16, cin=96, cout=96, k=(1, 3, 3), ups=((1, 1, 1), (1, 2, 2)),
                                   fold=(False, True), bias=True, stats=True),
    "3x3x3_up222_bias_stats": dict(N=1, D=4, H=16, W=16, cin=96, cout=64, k=(3, 3, 3), ups=((2, 2, 2),), bias=True, stats=True),
    "3x3x3_up122_504": dict(N=1, D=3, H=8, W=8, cin=504, cout=504, k=(3, 3, 3), ups=((1, 2, 2),), bias=True),
    "3x3_wide": dict(N=1, D=1, H=8, W=8, cin=2208, cout=768, k=(1, 3, 3), ups=((1, 2, 2),), bias=True),
    "1x1_classifier": dict(N=1, D=2, H=16, W=16, cin=64, cout=3, k=(1, 1, 1), bias=True, out_pad=0),
    "stem3d_7x7x7s2": dict(N=1, D=4, H=16, W=16, cin=4, cout=96, k=(7, 7, 7), stride=(2, 2, 2), src_c=4, fold=(False,), stats=True),
    "stem2d_7x7s2": dict(N=2, D=1, H=24, W=16, cin=3, cout=96, k=(1, 7, 7), stride=(1, 2, 2), src_c=4, fold=(False,), stats=True),
    "3x3x3_two_src": dict(N=1, D=4, H=16, W=8, cin=64, cout=64, k=(3, 3, 3), ups=((1, 1, 1), (1, 1, 1)), fold=(True, True),
                          bias=True, stats=True),
}


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", sorted(CASES))
def test_fprop_tc_vs_fp32(cuda_dev, name, prec):
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    c = Case(cuda_dev, tc=prec, **kw)
    assert c.supported[0] == prec, "tcgen05 fprop does not take %s" % name
    TOL = TOLS[prec]
    y0, s0, _ = c.fprop(0)
    y1, s1, untouched = c.fprop(prec)
    assert untouched == 0.0, "wrote outside the output channel window"
    err = rel_l2(y1, y0)
    assert err < TOL, (name, err)
    if s0 is not None:
        assert rel_l2(s1[0], s0[0]) < TOL and rel_l2(s1[1], s0[1]) < TOL


@pytest.mark.parametrize("prec", PRECS)
def test_fprop_tc_dropout_mask_identical(cuda_dev, prec):
    kw = dict(CASES["3x3x3_up222_bias_stats"])
    kw.update(drop=0.3, stats=False, fold=(True,))
    c = Case(cuda_dev, tc=prec, **kw)
    TOL = TOLS[prec]
    c.op.desc.drop_seed = 12345
    for key in ("_c_f", "_c_g"):
        c.op.desc.__dict__.pop(key, None)
    net_seed = c.net.step_seed
    c.net.step_seed = lambda name: 12345
    y0, _, _ = c.fprop(0)
    y1, _, _ = c.fprop(prec)
    c.net.step_seed = net_seed
    assert np.array_equal(y0 == 0, y1 == 0)
    assert abs(float((y0 == 0).mean()) - 0.3) < 0.02
    assert rel_l2(y1, y0) < TOL


DG = ["stem3d_7x7x7s2", "1x1_flat", "1x1_flat_tail", "3x3_dense", "3x3x3_dense", "3x3_up_skip_bias_stats", "3x3x3_up222_bias_stats",
      "3x3x3_up122_504", "3x3x3_two_src"]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", DG)
@pytest.mark.parametrize("mode,accumulate", [(0, False), (0, True), (1, False), (1, True)])
def test_dgrad_tc_vs_fp32(cuda_dev, name, mode, accumulate, prec):
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    c = Case(cuda_dev, tc=prec, **kw)
    assert c.supported[1] == prec, "tcgen05 dgrad does not take %s" % name
    TOL = TOLS[prec]
    stem = name.startswith("stem")        # a stem input has no BN in front: plain dx, no S1/S2 (engine.Planner.epilogue_for)
    if stem and mode == 1:
        pytest.skip("stem inputs never take the du form")
    r0 = c.dgrad(0, mode, accumulate, with_sums=not stem)
    r1 = c.dgrad(prec, mode, accumulate, with_sums=not stem)
    for (g0, S0), (g1, S1) in zip(r0, r1):
        assert rel_l2(g1, g0) < TOL, (name, mode, accumulate, rel_l2(g1, g0))
        assert rel_l2(S1[0], S0[0]) < TOL and rel_l2(S1[1], S0[1]) < TOL


def _wgrad(c, prec):
    net, op, d = c.net, c.op, c.op.desc
    net.params.grads.zero_()
    op.prec = [prec] * 3
    op._set_prec(2)
    net.be.conv_wgrad(d, op.w.g, None if op.bias is None else op.bias.g)
    torch.cuda.synchronize()
    return op.w.g.cpu().numpy().copy(), None if op.bias is None else op.bias.g.cpu().numpy().copy()


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", sorted(CASES))
def test_wgrad_tc_vs_fp32(cuda_dev, name, prec):
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    c = Case(cuda_dev, tc=prec, **kw)
    if c.supported[2] != prec:
        pytest.skip("tcgen05 wgrad does not take %s" % name)
    TOL = TOLS[prec]
    w0, b0 = _wgrad(c, 0)
    w1, b1 = _wgrad(c, prec)
    assert rel_l2(w1, w0) < TOL, (name, rel_l2(w1, w0))
    if b0 is not None:
        assert rel_l2(b1, b0) < 1e-4


# ---- second-generation weight gradient (conv_tc2_wgrad.cu: bf16 pre-pass + TMA tile loads + tcgen05) ---------------
W2 = ["3x3_dense", "3x3x3_dense", "3x3_up_skip_bias_stats", "3x3x3_up222_bias_stats", "3x3x3_up122_504", "3x3_wide",
      "3x3x3_two_src", "1x1_flat", "1x1_flat_tail"]
W2_EXTRA = {
    # ragged grid (H, W not multiples of the 16 x 8 tile), channel tails inside a 128-channel tile / a 64-channel block
    "3x3x3_ragged": dict(N=2, D=3, H=20, W=12, cin=96, cout=64, k=(3, 3, 3), bias=True),
    "3x3_ragged_c200": dict(N=1, D=2, H=24, W=20, cin=200, cout=72, k=(1, 3, 3)),
    # 1x1x1: flat position tiles (2-D tensor map), M not a multiple of 128, a transition-like shape with many channel tiles
    "1x1_flat_m_tail": dict(N=3, D=1, H=13, W=11, cin=1200, cout=192, k=(1, 1, 1)),
    "1x1_up_two_src": dict(N=1, D=2, H=16, W=16, cin=64, cout=40, k=(1, 1, 1), ups=((1, 1, 1), (2, 2, 2)), fold=(False, True), bias=True),
}


def _switch(name, value):
    from h_denseunet_b200 import _lib
    _lib.check(_lib.load().hdn_set_switch(name.encode(), int(value)), "hdn_set_switch")


@pytest.mark.parametrize("layout", [0, 1], ids=["planes", "sw128"])
@pytest.mark.parametrize("name", W2 + sorted(W2_EXTRA))
def test_wgrad_tc2_vs_fp32_and_gen1(cuda_dev, name, layout):
    """tc2 weight gradient against the fp32 FMA kernel (bf16 operand bound) and against the first-generation tcgen05
    kernel: identical bf16 operands, fp32 accumulation in another order -> 1e-4."""
    kw = dict(CASES.get(name) or W2_EXTRA[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    _switch("HDN_WGRAD_TC2", 1)
    _switch("HDN_TC2_LAYOUT", layout)
    try:
        c = Case(cuda_dev, tc=1, **kw)
        assert c.supported[2] == 1
        c.op.desc.precision = 1
        assert c.net.be.conv_tc_workspace(c.op.desc, 2) > 0, "tc2 does not take %s" % name   # gen 1 needs no wgrad scratch
        w0, b0 = _wgrad(c, 0)
        w2, b2 = _wgrad(c, 1)
        _switch("HDN_WGRAD_TC2", 0)
        w1, _ = _wgrad(c, 1)
    finally:
        _switch("HDN_WGRAD_TC2", 1)
        _switch("HDN_TC2_LAYOUT", 0)
    assert rel_l2(w2, w0) < TOLS[1], (name, rel_l2(w2, w0))
    assert rel_l2(w2, w1) < 1e-4, (name, rel_l2(w2, w1))
    if b0 is not None:
        assert rel_l2(b2, b0) < 1e-4


# ---- fprop / dgrad producer forms: HDN_TC_TMA 0 = SIMT producers everywhere, 1 = TMA mode for the 1x3x3 / 3x3x3 layers,
# 2 = TMA mode for the 1x1x1 layers as well (default; the tests above run in it) -----------------------------------------
TMA_CASES = ["1x1_flat", "1x1_flat_tail", "3x3_dense", "3x3x3_dense", "3x3_up_skip_bias_stats", "3x3x3_up222_bias_stats",
             "3x3x3_up122_504", "3x3x3_two_src", "3x3_wide"]


@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", TMA_CASES)
def test_fprop_dgrad_producer_forms(cuda_dev, name, prec, level):
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    _switch("HDN_TC_TMA", level)
    try:
        c = Case(cuda_dev, tc=prec, **kw)
        TOL = TOLS[prec]
        y0, s0, _ = c.fprop(0)
        y1, s1, untouched = c.fprop(prec)
        assert untouched == 0.0
        assert rel_l2(y1, y0) < TOL, (name, rel_l2(y1, y0))
        if s0 is not None:
            assert rel_l2(s1[0], s0[0]) < TOL and rel_l2(s1[1], s0[1]) < TOL
        if name in DG:
            for mode, acc in ((0, False), (0, True), (1, True)):
                r0 = c.dgrad(0, mode, acc)
                r1 = c.dgrad(prec, mode, acc)
                for (g0, S0), (g1, S1) in zip(r0, r1):
                    assert rel_l2(g1, g0) < TOL, (name, mode, acc, rel_l2(g1, g0))
                    assert rel_l2(S1[0], S0[0]) < TOL and rel_l2(S1[1], S0[1]) < TOL
    finally:
        _switch("HDN_TC_TMA", 2)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", TMA_CASES)
def test_fprop_dgrad_tma_swizzle128(cuda_dev, name, prec):
    """HDN_TC_SW128=1: the TMA mode with 128-byte swizzled operand rows (one tile load per stage, [head | tail] interleaved
    operand tensor, swizzled weight rows, K-major SWIZZLE_128B descriptors with sub-row K offsets)."""
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    _switch("HDN_TC_SW128", 1)
    try:
        c = Case(cuda_dev, tc=prec, **kw)
        TOL = TOLS[prec]
        y0, s0, _ = c.fprop(0)
        y1, s1, untouched = c.fprop(prec)
        assert untouched == 0.0
        assert rel_l2(y1, y0) < TOL, (name, rel_l2(y1, y0))
        if s0 is not None:
            assert rel_l2(s1[0], s0[0]) < TOL and rel_l2(s1[1], s0[1]) < TOL
        if name in DG:
            for mode, acc in ((0, False), (0, True), (1, True)):
                r0 = c.dgrad(0, mode, acc)
                r1 = c.dgrad(prec, mode, acc)
                for (g0, S0), (g1, S1) in zip(r0, r1):
                    assert rel_l2(g1, g0) < TOL, (name, mode, acc, rel_l2(g1, g0))
                    assert rel_l2(S1[0], S0[0]) < TOL and rel_l2(S1[1], S0[1]) < TOL
    finally:
        _switch("HDN_TC_SW128", 0)


@pytest.mark.parametrize("fold", [0, 1], ids=["three_mma", "folded"])
@pytest.mark.parametrize("name", TMA_CASES)
def test_fprop_dgrad_x3_forms(cuda_dev, name, fold):
    """Both issue schemes of bf16x3.  HDN_TC_X3FOLD=1 (default): two MMAs per K step -- A_hi x [B_hi | B_lo] with N = 2*BN into a
    double-width accumulator (the halves are added when the epilogue reads TMEM) and A_lo x B_hi; 0: three MMAs."""
    kw = dict(CASES[name])
    kw.setdefault("fold", tuple(True for _ in kw.get("ups", ((1, 1, 1),))))
    _switch("HDN_TC_X3FOLD", fold)
    try:
        c = Case(cuda_dev, tc=2, **kw)
        TOL = TOLS[2]
        y0, s0, _ = c.fprop(0)
        y1, s1, untouched = c.fprop(2)
        assert untouched == 0.0
        assert rel_l2(y1, y0) < TOL, (name, rel_l2(y1, y0))
        if s0 is not None:
            assert rel_l2(s1[0], s0[0]) < TOL and rel_l2(s1[1], s0[1]) < TOL
        if name in DG:
            for mode, acc in ((0, False), (0, True), (1, True)):
                r0 = c.dgrad(0, mode, acc)
                r1 = c.dgrad(2, mode, acc)
                for (g0, S0), (g1, S1) in zip(r0, r1):
                    assert rel_l2(g1, g0) < TOL, (name, mode, acc, rel_l2(g1, g0))
                    assert rel_l2(S1[0], S0[0]) < TOL and rel_l2(S1[1], S0[1]) < TOL
    finally:
        _switch("HDN_TC_X3FOLD", 1)
