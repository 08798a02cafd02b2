"""Launch plans of every convolution of the hybrid net at the headline shape (512x512x48), per pass and precision,
from hdn_conv_tc_plan (host arithmetic only: runs without a GPU).  The program is built small on the CPU reference
backend and its descriptors are rescaled to the full-size grid with fake (aligned) device pointers."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def conv_descriptors(size=512, cols=48, batch=1, small=(64, 8)):
    import h_denseunet_b200 as hdn
    from h_denseunet_b200 import engine, _lib
    from torch_backend import TorchBackend
    from util import Args
    m = hdn.dense_rnn_net(Args(b=1, input_size=small[0], input_cols=small[1]), backend=TorchBackend(), device="cpu", precision="fp32")
    net = m._net(True)
    fh, fd = size // small[0], cols // small[1]
    out = []
    for op in net.ops:
        if not isinstance(op, engine.ConvOp):
            continue
        d = op.desc
        c = _lib.Conv()
        two_d = d.out.D == 1                       # 2-D convs: N = slices, D = 1
        sc_n = (fd * batch) if two_d else batch
        sc_d = 1 if two_d else fd
        c.N, c.D, c.H, c.W = d.out.N * sc_n, d.out.D * sc_d, d.out.H * fh, d.out.W * fh
        c.Cin, c.Cout = d.Cin, d.Cout
        c.kd, c.kh, c.kw = d.k
        c.sd, c.sh, c.sw = d.s
        c.pd, c.ph, c.pw = d.p
        c.nsrc = len(d.srcs)
        for i, s in enumerate(d.srcs):
            v = s.act.view
            has = s.act.fold is not None
            c.src[i] = _lib.Src(_lib.Tensor(0x10000, v.buf.C, v.coff), v.D * sc_d, v.H * fh, v.W * fh, s.up[0], s.up[1], s.up[2],
                                0x20000 if has else 0, 0x30000 if has else 0, 1 if s.act.relu else 0)
        c.w = 0x40000
        c.bias = 0x50000 if d.bias is not None else 0
        c.y = _lib.Tensor(0x60000, d.out.buf.C, d.out.coff)
        c.drop_keep = 1.0
        out.append((op.name, c))
    return out


def plans(precision, **kw):
    from h_denseunet_b200 import _lib
    lib = _lib.load()
    rows = []
    for name, c in conv_descriptors(**kw):
        c.precision = precision
        for ps in range(3):
            o = (C.c_int32 * 16)()
            if not lib.hdn_conv_tc_supported(C.byref(c), ps):
                rows.append((name, ps, None, 0, 0))
                continue
            rc = lib.hdn_conv_tc_plan(C.byref(c), ps, o)
            assert rc == 0, (name, ps, lib.hdn_last_error())
            rows.append((name, ps, list(o), int(lib.hdn_conv_tc_workspace(C.byref(c), ps)), int(c.N) * c.D * c.H * c.W))
    return rows


if __name__ == "__main__":
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    seen = set()
    print("%-22s pass  BN tiles KB/ci CK/CW nsb/G nraw tmem   smem flat   P x3 s2d    work fit  K/BNe NC/CI      ws" % "conv")
    for name, ps, o, ws, _m in plans(prec):
        key = (ps, tuple(o) if o else None)
        if key in seen:
            continue
        seen.add(key)
        if o is None:
            print("%-22s %4d  -- fp32 FMA path" % (name, ps))
        else:
            print("%-22s %4d %3d %5d %5d %5d %5d %4d %4d %6d %4d %3d %2d %3d %7d %3d %6d %5d %7d" % ((name, ps) + tuple(o) + (ws,)))
