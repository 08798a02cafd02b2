"""Forward/backward engine: a static program of fused ops over preallocated NDHWC buffers.

The reference builds a Keras graph of ~830 layers and lets TensorFlow differentiate it
(Keras-2.0.8/keras/engine/training.py:948-967, KB:2310).  Here the builders (models.py) emit a
short program of *fused* ops -- every BatchNorm/Scale/ReLU/ZeroPadding/UpSampling/Add/concat of
the reference disappears into a convolution's load or store indexing -- and each op carries
its own hand-written backward.  All arithmetic happens in libhdn.so (include/hdn.h); this file
only plans buffers, descriptors and launch order.

Layout: every tensor is (N, D, H, W, C) fp32, channels last; 2-D tensors have D == 1 and the
reference's 3-D (N,H,W,S,C) tensors use D = S.  A dense block owns ONE buffer of its final
width; each conv_block writes its growth_rate new channels at a channel offset and reads the
window [0, Cin) -- the reference's tf.concat (merge.py:328-332) never happens.
"""
import ctypes as C
import zlib

import numpy as np
import torch

from . import _lib


# Net(precision=...) -> hdn_conv.precision per pass (fprop, dgrad, wgrad) of the convolutions the tcgen05 path takes
# (include/hdn.h: 1 = bf16 operands, 2 = bf16x3 head + tail split; fp32 accumulation in both).
#   "bf16"   : everything on rounded-to-bf16 operands (fastest; ~1.5e-2 from the fp32 oracle after 200+ layers)
#   "bf16x3" : everything split (fp32-grade results in all three passes)
#   "mixed"  : the activation path -- fprop and dgrad, whose rounding errors compound layer after layer -- split, the
#              weight gradients (a leaf: their rounding error stays in that one tensor, ~3e-3) on plain bf16 operands
TC_PRECISION = {"bf16": (1, 1, 1), "bf16x3": (2, 2, 2), "mixed": (2, 2, 1)}

# HDN_POISON=1 (debug aid): activation / gradient buffers start as NaN instead of uninitialised memory, so that a read of an
# element no kernel has written shows up in the results (fresh CUDA pages are zero: such reads hide in a new process)
import os as _os
POISON = _os.environ.get("HDN_POISON", "0") not in ("", "0")


# ------------------------------------------------------------------------- descriptors
class TView(object):
    """Channel window [coff, coff+C) of a Buffer, optionally re-read with other (N, D)."""

    def __init__(self, buf, coff, C_, N=None, D=None):
        self.buf, self.coff, self.C = buf, coff, C_
        self.N = buf.N if N is None else N
        self.D = buf.D if D is None else D
        assert self.N * self.D == buf.N * buf.D
        self.H, self.W = buf.H, buf.W

    @property
    def M(self):
        return self.N * self.D * self.H * self.W

    @property
    def ldc(self):
        return self.buf.C

    def sub(self, off, C_):
        return TView(self.buf, self.coff + off, C_, self.N, self.D)

    def as_nd(self, N, D):
        return TView(self.buf, self.coff, self.C, N, D)


class Buffer(object):
    def __init__(self, net, name, N, D, H, W, C_):
        self.net, self.name = net, name
        self.N, self.D, self.H, self.W, self.C = N, D, H, W, C_
        self.data = torch.empty((N, D, H, W, C_), dtype=torch.float32, device=net.device)
        if POISON and self.data.device.type == "cuda":
            self.data.fill_(float("nan"))      # HDN_POISON=1: a kernel that consumes a never-written element turns the result into NaN
        self.grad = None
        self.requires_grad = False
        self.need_stats = np.zeros(C_, bool)
        self.stats = None            # double [2, C] slot in the accumulator arena
        self.ginit = np.zeros(C_, bool)

    def view(self, coff=0, C_=None, N=None, D=None):
        return TView(self, coff, self.C - coff if C_ is None else C_, N, D)

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.empty_like(self.data)
            if POISON and self.grad.device.type == "cuda":
                self.grad.fill_(float("nan"))
        return self.grad


class Act(object):
    """Lazy activation: stored tensor + (optional) folded BN/Scale affine + ReLU, applied by the
    consumer on load.  Nothing is materialised."""

    def __init__(self, view, fold=None, relu=False):
        self.view, self.fold, self.relu = view, fold, relu

    @property
    def requires_grad(self):
        return self.view.buf.requires_grad or (self.fold is not None and self.fold.has_trainable)


class Src(object):
    def __init__(self, act, up=(1, 1, 1)):
        if isinstance(act, TView):
            act = Act(act)
        self.act, self.up = act, tuple(up)


class ConvDesc(object):
    """Python mirror of hdn_conv (include/hdn.h)."""

    def __init__(self, **kw):
        self.stat = None            # double tensor [2, Cout] or None
        self.drop_keep, self.drop_seed, self.precision = 1.0, 0, 0
        self.bias = None
        self.__dict__.update(kw)


class EpiDesc(object):
    """Python mirror of hdn_dgrad_epi."""

    def __init__(self, mode, accumulate, dx=None, du=None, s=None, center=None):
        self.mode, self.accumulate, self.dx, self.du, self.s, self.center = mode, accumulate, dx, du, s, center


class PoolDesc(object):
    def __init__(self, kind, out, src, pool_d):
        self.kind, self.out, self.src, self.pool_d = kind, out, src, pool_d


# ------------------------------------------------------------------------- CUDA backend
def _ptr(t):
    return 0 if t is None else t.data_ptr()


class CudaBackend(object):
    """Executes descriptors through the C-ABI.  ctypes structs are built once and cached."""

    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        self.launches = 0
        self.prof = None          # list of (key, flops, start_event, end_event, op name) while profiling
        self._cur = ""

    def _run(self, key, flops, n_launch, fn, *args, nbytes=0.0):
        """Issue one C-ABI call; with profiling on, bracket it with CUDA events on the launch stream.
        prof entries: (class key, algorithmic flops, start event, end event, op name, algorithmic HBM bytes)."""
        self.launches += n_launch
        if self.prof is None:
            return fn(*args)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        self.prof.append((key, flops, s, e, self._cur, nbytes))
        return rc

    @staticmethod
    def _conv_key(d, what):
        return "conv_%s[%s] k%dx%dx%d" % (what, ("simt", "tc", "tc-x3")[d.precision], d.k[0], d.k[1], d.k[2])

    @staticmethod
    def _conv_flops(d):
        return 2.0 * d.out.M * d.Cin * d.Cout * d.k[0] * d.k[1] * d.k[2]

    def _conv_bytes(self, d, which, epis=None):
        """Algorithmic HBM bytes of one convolution pass (fp32 activations read / written once; weights and the
        halo re-reads of neighbouring tiles not counted): the numerator of the HBM roofline of the 1x1 layers."""
        if self.prof is None:
            return 0.0
        m_out = float(d.out.M) * d.Cout
        src = [float(s.act.view.M) * d.Cin for s in d.srcs]
        if which == 1:
            b = m_out                                           # dY
            for n, e in zip(src, epis):
                if e.mode != 2:
                    b += n * (3.0 if e.accumulate else 2.0)     # stored value (ReLU mask, S2) + gradient write (+ read)
            return 4.0 * b
        return 4.0 * (sum(src) + m_out)                         # fprop: sources + y;  wgrad: sources + dY

    @property
    def stream(self):
        return torch.cuda.current_stream().cuda_stream

    # -- struct builders
    @staticmethod
    def _tensor(view, grad=False):
        t = view.buf.grad if grad else view.buf.data
        return _lib.Tensor(t.data_ptr(), view.buf.C, view.coff)

    def _src(self, s):
        v, f = s.act.view, s.act.fold
        return _lib.Src(self._tensor(v), v.D, v.H, v.W, s.up[0], s.up[1], s.up[2],
                        _ptr(f.a if f is not None else None), _ptr(f.b if f is not None else None),
                        1 if s.act.relu else 0)

    def _conv(self, d, y_grad=False):
        key = "_c_g" if y_grad else "_c_f"
        c = d.__dict__.get(key)
        if c is None:
            c = _lib.Conv()
            c.N, c.D, c.H, c.W = d.out.N, d.out.D, d.out.H, d.out.W
            c.Cin, c.Cout = d.Cin, d.Cout
            c.kd, c.kh, c.kw = d.k
            c.sd, c.sh, c.sw = d.s
            c.pd, c.ph, c.pw = d.p
            c.nsrc = len(d.srcs)
            for i, s in enumerate(d.srcs):
                c.src[i] = self._src(s)
            c.w = d.w.data_ptr()
            c.bias = _ptr(d.bias)
            c.y = self._tensor(d.out, grad=y_grad)
            if d.stat is not None and not y_grad:
                c.stat_sum = d.stat[0].data_ptr()
                c.stat_sq = d.stat[1].data_ptr()
            c.drop_keep, c.drop_seed, c.precision = d.drop_keep, d.drop_seed, d.precision
            d.__dict__[key] = c
        ws = getattr(d, "ws", None)
        if ws is not None and c.ws != ws.data_ptr():
            c.ws, c.ws_bytes = ws.data_ptr(), ws.numel() * ws.element_size()
        return c

    def _epi(self, e, Cin):
        c = e.__dict__.get("_c")
        if c is None:
            c = _lib.DgradEpi()
            if e.dx is not None:
                c.dx = self._tensor(e.dx, grad=True)
            c.du = _ptr(e.du)
            c.mode, c.accumulate = e.mode, 1 if e.accumulate else 0
            if e.s is not None:
                c.s1, c.s2 = e.s[0].data_ptr(), e.s[1].data_ptr()
                c.center = _ptr(e.center)
            e.__dict__["_c"] = c
        return c

    # -- ops
    def conv_tc_supported(self, d, which):
        c = self._conv(d)
        c.precision = d.precision          # the cached struct may carry another pass's precision
        return bool(self.lib.hdn_conv_tc_supported(C.byref(c), which))

    def conv_tc_workspace(self, d, which):
        c = self._conv(d)
        c.precision = d.precision
        return int(self.lib.hdn_conv_tc_workspace(C.byref(c), which))

    def conv_fprop(self, d):
        self._cur = d.name
        _lib.check(self._run(self._conv_key(d, "fprop"), self._conv_flops(d), 1, self.lib.hdn_conv_fprop,
                             C.byref(self._conv(d)), self.stream, nbytes=self._conv_bytes(d, 0)), "hdn_conv_fprop " + d.name)

    def conv_dgrad(self, d, epis):
        arr = d.__dict__.get("_c_epis")
        if arr is None:
            arr = (_lib.DgradEpi * 2)()
            for i, e in enumerate(epis):
                arr[i] = self._epi(e, d.Cin)
            d.__dict__["_c_epis"] = arr
        nact = sum(1 for e in epis if e.mode != 2)
        self._cur = d.name
        _lib.check(self._run(self._conv_key(d, "dgrad"), self._conv_flops(d) * nact, nact, self.lib.hdn_conv_dgrad,
                             C.byref(self._conv(d, True)), arr, self.stream, nbytes=self._conv_bytes(d, 1, epis)),
                   "hdn_conv_dgrad " + d.name)

    def conv_wgrad(self, d, dw, dbias):
        self._cur = d.name
        _lib.check(self._run(self._conv_key(d, "wgrad"), self._conv_flops(d), 1 + (dbias is not None),
                             self.lib.hdn_conv_wgrad, C.byref(self._conv(d, True)), dw.data_ptr(), _ptr(dbias),
                             self.stream, nbytes=self._conv_bytes(d, 2)), "hdn_conv_wgrad " + d.name)

    def _pool(self, d, y_grad):
        key = "_c_g" if y_grad else "_c_f"
        c = d.__dict__.get(key)
        if c is None:
            c = _lib.Pool()
            c.kind = d.kind
            c.N, c.D, c.H, c.W, c.C = d.out.N, d.out.D, d.out.H, d.out.W, d.out.C
            c.pool_d = d.pool_d
            c.src = self._src(d.src)
            c.y = self._tensor(d.out, grad=y_grad)
            c.argidx = _ptr(getattr(d, "argidx", None))
            d.__dict__[key] = c
        return c

    def pool_fwd(self, d):
        _lib.check(self._run("pool_fwd", 0.0, 1, self.lib.hdn_pool_fwd, C.byref(self._pool(d, False)), self.stream), "hdn_pool_fwd")

    def pool_bwd(self, d, epi):
        _lib.check(self._run("pool_bwd", 0.0, 1, self.lib.hdn_pool_bwd, C.byref(self._pool(d, True)), C.byref(self._epi(epi, d.out.C)), self.stream),
                   "hdn_pool_bwd")

    def col_stats(self, view, stat):
        _lib.check(self._run("col_stats", 0.0, 1, self.lib.hdn_col_stats, self._tensor(view), view.M, view.C, stat[0].data_ptr(), stat[1].data_ptr(),
                                          self.stream), "hdn_col_stats")

    def bn_fold(self, f, mode):
        c = f.__dict__.get("_c_fold%d" % mode)
        if c is None:
            c = _lib.BnFold()
            c.C, c.mode, c.count = f.C, mode, float(f.view.M)
            if mode == 1:
                c.sum, c.sumsq = f.stat[0].data_ptr(), f.stat[1].data_ptr()
            c.mov_mean, c.mov_var = f.mov_mean.t.data_ptr(), f.mov_var.t.data_ptr()
            c.gamma, c.beta = f.gamma.t.data_ptr(), f.beta.t.data_ptr()
            if f.sgamma is not None:
                c.sgamma, c.sbeta = f.sgamma.t.data_ptr(), f.sbeta.t.data_ptr()
            c.eps, c.momentum = f.eps, f.momentum
            c.a, c.b, c.mean, c.rstd = f.a.data_ptr(), f.b.data_ptr(), f.mean.data_ptr(), f.rstd.data_ptr()
            f.__dict__["_c_fold%d" % mode] = c
        _lib.check(self._run("bn_fold", 0.0, 1, self.lib.hdn_bn_fold, C.byref(c), self.stream), "hdn_bn_fold " + f.name)

    def bn_param_grad(self, f, mode):
        c = f.__dict__.get("_c_grad")
        if c is None:
            c = _lib.BnGrad()
            c.C, c.mode, c.count = f.C, mode, float(f.view.M)
            c.s1, c.s2 = f.S[0].data_ptr(), f.S[1].data_ptr()
            c.mean, c.rstd = f.mean.data_ptr(), f.rstd.data_ptr()
            c.gamma, c.beta = f.gamma.t.data_ptr(), f.beta.t.data_ptr()
            c.sgamma = _ptr(f.sgamma.t if f.sgamma is not None else None)
            c.dgamma = _ptr(f.gamma.g)
            c.dbeta = _ptr(f.beta.g)
            if f.sgamma is not None:
                c.dsgamma, c.dsbeta = _ptr(f.sgamma.g), _ptr(f.sbeta.g)
            if mode == 1:
                c.k0, c.k1, c.k2 = f.k[0].data_ptr(), f.k[1].data_ptr(), f.k[2].data_ptr()
            f.__dict__["_c_grad"] = c
        _lib.check(self._run("bn_param_grad", 0.0, 1, self.lib.hdn_bn_param_grad, C.byref(c), self.stream), "hdn_bn_param_grad " + f.name)

    def bn_bwd_apply(self, f, accumulate):
        v = f.view
        _lib.check(self._run("bn_bwd_apply", 0.0, 1, self.lib.hdn_bn_bwd_apply, f.du.data_ptr(), self._tensor(v), self._tensor(v, True), v.M, v.C,
                                             f.k[0].data_ptr(), f.k[1].data_ptr(), f.k[2].data_ptr(),
                                             f.mean.data_ptr(), 1 if accumulate else 0, self.stream), "hdn_bn_bwd_apply " + f.name)

    def dropout_bwd(self, view, keep, seed):
        _lib.check(self._run("dropout_bwd", 0.0, 1, self.lib.hdn_dropout_bwd, self._tensor(view, True), view.M, view.C, keep, seed, self.stream),
                   "hdn_dropout_bwd")

    def wce_accum(self, logits, labels, N, D, HW, d0, d1, acc):
        _lib.check(self._run("wce_accum", 0.0, 1, self.lib.hdn_wce_accum, logits.data_ptr(), labels.data_ptr(), N, D, HW, d0, d1, acc.data_ptr(),
                                          self.stream), "hdn_wce_accum")

    def wce_grad(self, logits, labels, dlogits, N, D, HW, d0, d1, acc, gscale):
        _lib.check(self._run("wce_grad", 0.0, 1, self.lib.hdn_wce_grad, logits.data_ptr(), labels.data_ptr(), dlogits.data_ptr(), N, D, HW, d0, d1,
                                         acc.data_ptr(), gscale, self.stream), "hdn_wce_grad")

    def triplets(self, vol, out, B, S, HW):
        _lib.check(self._run("triplets", 0.0, 1, self.lib.hdn_triplets, vol.data_ptr(), out.data_ptr(), B, S, HW,
                             out.shape[-1], self.stream), "hdn_triplets")

    def cat4(self, vol, logits, out, M, k):
        _lib.check(self._run("cat4", 0.0, 1, self.lib.hdn_cat4, vol.data_ptr(), logits.data_ptr(), out.data_ptr(), M, k, self.stream), "hdn_cat4")

    def cat4_bwd(self, dout, dlogits, M, k, accumulate):
        _lib.check(self._run("cat4_bwd", 0.0, 1, self.lib.hdn_cat4_bwd, dout.data_ptr(), dlogits.data_ptr(), M, k, 1 if accumulate else 0,
                                         self.stream), "hdn_cat4_bwd")

    def sgd(self, p, g, m, n, lr, mu, gscale):
        _lib.check(self._run("sgd_nesterov", 0.0, 1, self.lib.hdn_sgd_nesterov, p.data_ptr(), g.data_ptr(), m.data_ptr(), n, lr, mu, gscale, self.stream),
                   "hdn_sgd_nesterov")

    def window_accumulate(self, logits, score, count, S, HW, z0):
        _lib.check(self._run("window_accumulate", 0.0, 2, self.lib.hdn_window_accumulate, logits.data_ptr(), score.data_ptr(), count.data_ptr(), S, HW, z0,
                                                  self.stream), "hdn_window_accumulate")

    def window_finalize(self, score, count, Z, HW):
        _lib.check(self._run("window_finalize", 0.0, 1, self.lib.hdn_window_finalize, score.data_ptr(), count.data_ptr(), Z, HW, self.stream),
                   "hdn_window_finalize")


# ------------------------------------------------------------------------- parameters
class Param(object):
    def __init__(self, name, shape, trainable):
        self.name, self.shape, self.trainable = name, tuple(shape), trainable
        self.size = int(np.prod(shape))
        self.offset = None
        self.t = None     # value view
        self.g = None     # gradient view (trainable only)


class ParamStore(object):
    """All parameters of a model in two flat fp32 arenas: trainable (with gradient and momentum
    arenas of the same layout, so the optimizer and the data-parallel reduce are ONE launch) and
    state (frozen weights, BN moving statistics).  Host values live in `host` until realised."""

    def __init__(self, seed=0):
        self.params = {}
        self.order = []
        self.host = {}
        self.rng = np.random.default_rng(seed)
        self.realised = False
        self.train = self.grads = self.moms = self.state = None

    def get(self, name, shape, init, trainable):
        p = self.params.get(name)
        if p is not None:
            assert p.shape == tuple(shape), (name, p.shape, shape)
            return p
        assert not self.realised, "parameter %s requested after the arenas were realised" % name
        p = Param(name, shape, trainable)
        self.params[name] = p
        self.order.append(p)
        self.host[name] = np.ascontiguousarray(init(self.rng, shape), dtype=np.float32)
        return p

    def realise(self, device, alloc=None):
        """Allocate the arenas on `device` and upload host values.  `alloc(n)` may provide the
        trainable/grad arenas (the data-parallel path uses IPC-shareable cudaMalloc memory)."""
        if self.realised:
            return
        for trainable in (True, False):
            off = 0
            for p in self.order:
                if p.trainable == trainable:
                    p.offset = off
                    off += (p.size + 3) // 4 * 4      # 16-byte aligned tensors
            if trainable:
                self.n_train = off
            else:
                self.n_state = off
        mk = alloc or (lambda n: torch.zeros(max(n, 4), dtype=torch.float32, device=device))
        self.train = mk(self.n_train)
        self.grads = mk(self.n_train)
        self.moms = torch.zeros(max(self.n_train, 4), dtype=torch.float32, device=device)
        self.state = torch.zeros(max(self.n_state, 4), dtype=torch.float32, device=device)
        self.train.zero_()
        self.grads.zero_()
        for p in self.order:
            arena = self.train if p.trainable else self.state
            p.t = arena[p.offset:p.offset + p.size].view(p.shape)
            if p.trainable:
                p.g = self.grads[p.offset:p.offset + p.size].view(p.shape)
            p.t.copy_(torch.from_numpy(self.host[p.name]))
        self.realised = True

    def set_value(self, name, value):
        p = self.params[name]
        v = np.ascontiguousarray(value, dtype=np.float32).reshape(p.shape)
        self.host[name] = v
        if self.realised:
            p.t.copy_(torch.from_numpy(v))

    def get_value(self, name):
        p = self.params[name]
        if self.realised:
            return p.t.detach().cpu().numpy().copy()
        return self.host[name].copy()

    def get_grad(self, name):
        p = self.params[name]
        return None if p.g is None else p.g.detach().cpu().numpy().copy()


# ------------------------------------------------------------------------- initialisers
def glorot_uniform(rng, shape):
    """Keras default kernel_initializer (KCONV:92, initializers.py VarianceScaling fan_avg uniform)."""
    rf = int(np.prod(shape[:-2]))
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)


def random_normal(rng, shape):
    """kernel_initializer='normal' -> RandomNormal(0, 0.05) (densenet.py:71, initializers.py:72)."""
    return rng.normal(0.0, 0.05, size=shape)


def zeros(rng, shape):
    return np.zeros(shape, np.float32)


def ones(rng, shape):
    return np.ones(shape, np.float32)


# ------------------------------------------------------------------------- ops
class Fold(object):
    """BatchNormalization (+ Scale) folded to a per-channel (a, b); an op of the program.

    forward : hdn_bn_fold   (training statistics come from the producers' epilogues)
    backward: hdn_bn_param_grad from the consumers' S1/S2 sums, then (training mode) the
              second pass hdn_bn_bwd_apply that turns the consumers' du into dx."""

    def __init__(self, net, name, view, bn_name, scale_name, eps, learn, momentum, bn_trainable, scale_trainable):
        ps, C_ = net.params, view.C
        self.net, self.name, self.view, self.C = net, name, view, C_
        self.eps, self.momentum = eps, momentum
        self.train_mode = bool(learn and net.training)
        self.gamma = ps.get(bn_name + "/gamma", (C_,), ones, bn_trainable)
        self.beta = ps.get(bn_name + "/beta", (C_,), zeros, bn_trainable)
        self.mov_mean = ps.get(bn_name + "/moving_mean", (C_,), zeros, False)
        self.mov_var = ps.get(bn_name + "/moving_variance", (C_,), ones, False)
        if scale_name is not None:
            self.sgamma = ps.get(scale_name + "/gamma", (C_,), ones, scale_trainable)
            self.sbeta = ps.get(scale_name + "/beta", (C_,), zeros, scale_trainable)
        else:
            self.sgamma = self.sbeta = None
        self.has_trainable = net.training and (bn_trainable or (scale_name is not None and scale_trainable))
        dev = net.device
        self.a = torch.empty(C_, dtype=torch.float32, device=dev)
        self.b = torch.empty_like(self.a)
        self.mean = torch.empty_like(self.a)
        self.rstd = torch.empty_like(self.a)
        self.k = None
        self.S = None
        self.stat = None
        self.du = None
        self.du_init = False
        self.n_consumers_bwd = 0
        if self.train_mode:
            view.buf.need_stats[view.coff:view.coff + C_] = True

    # reserve accumulator slots
    def reserve(self):
        net = self.net
        if self.train_mode:
            self.stat = (self.view.buf.stats[0][self.view.coff:self.view.coff + self.C],
                         self.view.buf.stats[1][self.view.coff:self.view.coff + self.C])

    def forward(self):
        self.net.be.bn_fold(self, 1 if self.train_mode else 0)

    def plan_backward(self, planner):
        self.active_bwd = self.n_consumers_bwd > 0 and (self.has_trainable or self.train_mode)
        if not self.active_bwd:
            return
        if self.train_mode:
            dev = self.net.device
            self.k = [torch.empty(self.C, dtype=torch.float32, device=dev) for _ in range(3)]
            self.apply = self.view.buf.requires_grad
            if self.apply:
                self.apply_acc = planner.claim(self.view)
            planner.scratch_free(self)

    def backward(self):
        if not self.active_bwd:
            return
        be = self.net.be
        be.bn_param_grad(self, 1 if self.train_mode else 0)
        if self.train_mode and self.apply:
            be.bn_bwd_apply(self, self.apply_acc)


class ConvOp(object):
    def __init__(self, net, name, srcs, w, bias, out, k, s, p, drop_rate=0.0):
        self.net, self.name, self.srcs, self.w, self.bias, self.out = net, name, srcs, w, bias, out
        self.k, self.s, self.p = k, s, p
        self.drop_rate = drop_rate if net.training and net.dropout else 0.0
        self.desc = None

    def reserve(self):
        pass

    def compile(self):
        net, o = self.net, self.out
        st = None
        if o.buf.need_stats[o.coff:o.coff + o.C].any():
            st = (o.buf.stats[0][o.coff:o.coff + o.C], o.buf.stats[1][o.coff:o.coff + o.C])
        self.desc = ConvDesc(name=self.name, out=o, Cin=self.srcs[0].act.view.C, Cout=o.C, k=self.k, s=self.s, p=self.p,
                             srcs=self.srcs, w=self.w.t, bias=None if self.bias is None else self.bias.t, stat=st,
                             drop_keep=1.0 - self.drop_rate, drop_seed=0, precision=0)
        self.prec = [0, 0, 0]
        if net.precision in TC_PRECISION:
            for i in range(3):
                self.desc.precision = TC_PRECISION[net.precision][i]
                self.prec[i] = self.desc.precision if net.be.conv_tc_supported(self.desc, i) else 0
                if self.prec[i]:
                    net.ws_need = max(net.ws_need, net.be.conv_tc_workspace(self.desc, i))
            self.desc.precision = 0
        net.report.append((self.name, tuple(self.prec)))

    def _set_prec(self, which):
        d = self.desc
        d.precision = self.prec[which]
        for key in ("_c_f", "_c_g"):
            c = d.__dict__.get(key)
            if c is not None and c.precision != d.precision:
                c.precision = d.precision

    def forward(self):
        d = self.desc
        if self.drop_rate:
            d.drop_seed = self.net.step_seed(self.name)
            c = d.__dict__.get("_c_f")
            if c is not None:
                c.drop_seed = d.drop_seed
        self._set_prec(0)
        self.net.be.conv_fprop(d)

    def plan_backward(self, planner):
        self.do_w = self.w.trainable and self.net.training
        self.do_d = [s.act.requires_grad for s in self.srcs]
        self.active_bwd = self.out.buf.requires_grad and (self.do_w or any(self.do_d))
        if not self.active_bwd:
            return
        self.epis = []
        if any(self.do_d):
            for s, need in zip(self.srcs, self.do_d):
                self.epis.append(planner.epilogue_for(s.act) if need else EpiDesc(2, False))

    def backward(self):
        if not self.active_bwd:
            return
        be, d = self.net.be, self.desc
        if self.drop_rate:
            be.dropout_bwd(self.out, 1.0 - self.drop_rate, d.drop_seed)
        if self.do_w:
            self._set_prec(2)
            be.conv_wgrad(d, self.w.g, None if self.bias is None or not self.bias.trainable else self.bias.g)
        if self.epis:
            self._set_prec(1)
            be.conv_dgrad(d, self.epis)


class PoolOp(object):
    def __init__(self, net, kind, src, out, pool_d):
        self.net, self.kind, self.src, self.out, self.pool_d = net, kind, src, out, pool_d
        self.name = "pool"

    def reserve(self):
        pass

    def compile(self):
        o = self.out
        self.desc = PoolDesc(self.kind, o, self.src, self.pool_d)
        if self.kind == 0 and self.net.training and self.net.device.type == "cuda":
            # arg-max taps saved by the forward pass for the backward pass (1 byte per pooled element)
            self.desc.argidx = torch.zeros(o.M * o.C, dtype=torch.uint8, device=self.net.device)
        self.stat = None
        if o.buf.need_stats[o.coff:o.coff + o.C].any():
            self.stat = (o.buf.stats[0][o.coff:o.coff + o.C], o.buf.stats[1][o.coff:o.coff + o.C])

    def forward(self):
        self.net.be.pool_fwd(self.desc)
        if self.stat is not None:
            self.net.be.col_stats(self.out, self.stat)

    def plan_backward(self, planner):
        self.active_bwd = self.out.buf.requires_grad and self.src.act.requires_grad
        if self.active_bwd:
            self.epi = planner.epilogue_for(self.src.act)

    def backward(self):
        if self.active_bwd:
            self.net.be.pool_bwd(self.desc, self.epi)


class TripletOp(object):
    """hybridnet.py:385-396: volume (B,S,H,W,1) -> slice triplets (B*S,1,H,W,3)."""

    def __init__(self, net, vol, out):
        self.net, self.vol, self.out, self.name = net, vol, out, "triplets"

    def reserve(self):
        pass

    def compile(self):
        pass

    def forward(self):
        v = self.vol
        self.net.be.triplets(v.buf.data, self.out.buf.data, v.N, v.D, v.H * v.W)

    def plan_backward(self, planner):
        pass

    def backward(self):
        pass


class Cat4Op(object):
    """hybridnet.py:409-411: 3-D net input = concat([CT, 250 * logits2d]) on the channel axis."""

    def __init__(self, net, vol, logits, out, k=250.0):
        self.net, self.vol, self.logits, self.out, self.kk, self.name = net, vol, logits, out, k, "cat4"

    def reserve(self):
        pass

    def compile(self):
        pass

    def forward(self):
        self.net.be.cat4(self.vol.buf.data, self.logits.buf.data, self.out.buf.data, self.out.M, self.kk)

    def plan_backward(self, planner):
        self.active_bwd = self.out.buf.requires_grad and self.logits.buf.requires_grad
        if self.active_bwd:
            self.acc = planner.claim(self.logits)

    def backward(self):
        if self.active_bwd:
            self.net.be.cat4_bwd(self.out.buf.grad, self.logits.buf.grad, self.out.M, self.kk, self.acc)


class LossOp(object):
    """loss.py:5-46 weighted cross entropy, forward + dlogits."""

    def __init__(self, net, logits, crop):
        self.net, self.logits, self.crop, self.name = net, logits, crop, "wce"
        b = logits.buf
        assert logits.coff == 0 and logits.C == 3 and b.C == 3, "loss needs a contiguous 3-channel logits buffer"
        self.labels = torch.zeros((logits.N, logits.D, logits.H, logits.W), dtype=torch.float32, device=net.device)
        self.gscale = 1.0

    def reserve(self):
        self.acc = self.net.accum_alloc(2)

    def compile(self):
        v = self.logits
        self.d0, self.d1 = (1, v.D - 1) if self.crop else (0, v.D)

    def forward(self):
        v = self.logits
        self.net.be.wce_accum(v.buf.data, self.labels, v.N, v.D, v.H * v.W, self.d0, self.d1, self.acc)

    def plan_backward(self, planner):
        acc = planner.claim(self.logits)
        assert not acc

    def backward(self):
        v = self.logits
        self.net.be.wce_grad(v.buf.data, self.labels, v.buf.grad, v.N, v.D, v.H * v.W, self.d0, self.d1, self.acc,
                             self.gscale)

    def value(self):
        a = self.acc.cpu().numpy()
        return float(-a[0] / max(a[1], 1.0))


# ------------------------------------------------------------------------- planner
class Planner(object):
    """Walks the program in reverse once and fixes, per gradient write, whether it is the first
    write to its channel window (overwrite) or a later one (accumulate); also places the
    training-mode BN `du` temporaries in a shared scratch arena (first-fit, static)."""

    def __init__(self, net):
        self.net = net
        self.free = []        # (offset, size)
        self.top = 0
        self.live = {}

    def claim(self, view):
        b = view.buf
        b.ensure_grad()
        m = b.ginit[view.coff:view.coff + view.C]
        if m.all():
            return True
        if m.any():
            raise RuntimeError("gradient window of %s partially initialised" % b.name)
        b.ginit[view.coff:view.coff + view.C] = True
        return False

    def scratch_alloc(self, fold, n):
        for i, (o, s) in enumerate(self.free):
            if s >= n:
                self.free[i] = (o + n, s - n)
                self.live[fold] = (o, n)
                return o
        o = self.top
        self.top += n
        self.live[fold] = (o, n)
        return o

    def scratch_free(self, fold):
        if fold in self.live:
            self.free.append(self.live.pop(fold))

    def epilogue_for(self, act):
        f, v = act.fold, act.view
        if f is None:
            return EpiDesc(0, self.claim(v), dx=v)
        f.n_consumers_bwd += 1
        s = None
        if f.train_mode or f.has_trainable:
            if f.S is None:
                f.S = self.net.accum_alloc(2 * f.C).view(2, f.C)
            s = f.S
        if f.train_mode:
            if not hasattr(f, "du_off"):
                f.du_off = self.scratch_alloc(f, v.M * v.C)
                f.du_first = True
            acc = not f.du_first
            f.du_first = False
            return EpiDesc(1, acc, du=("scratch", f), s=s, center=f.mean)
        return EpiDesc(0, self.claim(v), dx=v, s=s, center=f.mean)


# ------------------------------------------------------------------------- the net
class Net(object):
    """A compiled forward(/backward) program for one model at one input shape and one mode."""

    def __init__(self, params, device, training, precision="fp32", backend=None, dropout=False):
        self.params, self.device, self.training = params, torch.device(device), training
        self.precision, self.dropout = precision, dropout
        self.be = backend if backend is not None else CudaBackend()
        self.ops, self.buffers, self.folds = [], [], []
        self.report = []
        self._accum_req = []
        self.accum = None
        self.compiled = False
        self.step = 0                 # optimizer steps taken (Model restores it from a checkpoint)
        self.seed_salt = 0            # per run / per replica (Model sets it): replicas must not draw identical masks
        self.inputs = {}
        self.outputs = {}
        self.loss = None
        self.ws_need = 0
        self.ws = None

    # -- graph construction
    def buffer(self, name, N, D, H, W, C_):
        b = Buffer(self, name, N, D, H, W, C_)
        self.buffers.append(b)
        return b

    def input(self, name, N, D, H, W, C_):
        b = self.buffer(name, N, D, H, W, C_)
        b.data.zero_()              # padding channels of an input are never written afterwards
        self.inputs[name] = b
        return b.view()

    def conv(self, name, srcs, cout, k, s=(1, 1, 1), p=(0, 0, 0), bias=True, out=None, init=glorot_uniform,
             trainable=True, drop_rate=0.0):
        srcs = [x if isinstance(x, Src) else Src(x) for x in srcs]
        v0 = srcs[0].act.view
        cin = v0.C
        vdims = [v0.D * srcs[0].up[0], v0.H * srcs[0].up[1], v0.W * srcs[0].up[2]]
        od = [(vdims[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
        for i, x in enumerate(srcs[1:], 1):       # the Add in front of a convolution: same channels, batch and (up-sampled) grid
            v = x.act.view
            if (v.C, v.N, v.D * x.up[0], v.H * x.up[1], v.W * x.up[2]) != (cin, v0.N, vdims[0], vdims[1], vdims[2]):
                raise ValueError("conv %s: source %d is (N=%d, C=%d, grid %s) but source 0 is (N=%d, C=%d, grid %s)" % (
                    name, i, v.N, v.C, (v.D * x.up[0], v.H * x.up[1], v.W * x.up[2]), v0.N, cin, tuple(vdims)))
        if out is None:
            out = self.buffer(name, v0.N, od[0], od[1], od[2], cout).view()
        assert (out.N, out.D, out.H, out.W, out.C) == (v0.N, od[0], od[1], od[2], cout), (name, out.N, out.D, out.H,
                                                                                          out.W, out.C, od)
        w = self.params.get(name + "/kernel", tuple(k) + (cin, cout), init, trainable)
        b = self.params.get(name + "/bias", (cout,), zeros, trainable) if bias else None
        op = ConvOp(self, name, srcs, w, b, out, tuple(k), tuple(s), tuple(p), drop_rate)
        self.ops.append(op)
        if self.training and (trainable or any(x.act.requires_grad for x in srcs)):
            out.buf.requires_grad = True
        return out

    def fold(self, view, bn_name, scale_name=None, eps=1e-3, learn=True, momentum=0.99, bn_trainable=True,
             scale_trainable=True, relu=True):
        f = Fold(self, bn_name, view, bn_name, scale_name, eps, learn, momentum, bn_trainable, scale_trainable)
        self.ops.append(f)
        self.folds.append(f)
        return Act(view, f, relu)

    def maxpool(self, act, out, pool_d):
        src = Src(act)
        self.ops.append(PoolOp(self, 0, src, out, 1 if pool_d else 0))
        if self.training and act.requires_grad:
            out.buf.requires_grad = True
        return out

    def avgpool(self, view, out):
        src = Src(Act(view))
        self.ops.append(PoolOp(self, 1, src, out, 0))
        if self.training and view.buf.requires_grad:
            out.buf.requires_grad = True
        return out

    def triplets(self, vol, out):
        self.ops.append(TripletOp(self, vol, out))

    def cat4(self, vol, logits, out):
        self.ops.append(Cat4Op(self, vol, logits, out))
        if self.training and logits.buf.requires_grad:
            out.buf.requires_grad = True

    def set_loss(self, logits, crop):
        self.loss = LossOp(self, logits, crop)
        self.ops.append(self.loss)

    # -- compilation
    def accum_alloc(self, n):
        if self.accum is not None:
            o = self._accum_top
            assert o + n <= self.accum.numel(), "accumulator arena exhausted"
            self._accum_top += n
            return self.accum[o:o + n]
        raise RuntimeError("accum_alloc before compile")

    def step_seed(self, name):
        return ((zlib.crc32(name.encode()) & 0xFFFFFF) * 1000003 + self.step + 1 + self.seed_salt) & 0xFFFFFFFFFFFF

    def compile(self):
        assert not self.compiled
        self.params.realise(self.device)
        # one double arena: batch statistics, S1/S2 sums, loss accumulators -- zeroed once per step
        need = 8
        for b in self.buffers:
            if b.need_stats.any():
                need += 2 * b.C
        for f in self.folds:
            need += 2 * f.C
        self.accum = torch.zeros(need, dtype=torch.float64, device=self.device)
        self._accum_top = 0
        for b in self.buffers:
            if b.need_stats.any():
                b.stats = self.accum_alloc(2 * b.C).view(2, b.C)
        for op in self.ops:
            op.reserve()
        for op in self.ops:
            if not isinstance(op, Fold):
                op.compile()
        if self.ws_need:
            # one scratch for the packed bf16 weights of whichever convolution is running (stream-ordered reuse)
            self.ws = torch.empty(self.ws_need, dtype=torch.uint8, device=self.device)
            for op in self.ops:
                if isinstance(op, ConvOp):
                    op.desc.ws = self.ws
        if self.training and self.loss is not None:
            planner = Planner(self)
            for op in reversed(self.ops):
                op.plan_backward(planner)
            self.scratch = torch.empty(max(planner.top, 4), dtype=torch.float32, device=self.device)
            if POISON and self.device.type == "cuda":
                self.scratch.fill_(float("nan"))
            for op in self.ops:
                for e in getattr(op, "epis", []) + ([op.epi] if hasattr(op, "epi") else []):
                    if isinstance(e.du, tuple):
                        f = e.du[1]
                        if f.du is None:
                            f.du = self.scratch[f.du_off:f.du_off + f.view.M * f.view.C]
                        e.du = f.du
        self.compiled = True

    # -- execution
    def forward(self):
        self.accum.zero_()
        for op in self.ops:
            op.forward()

    def backward(self):
        self.params.grads.zero_()
        for op in reversed(self.ops):
            op.backward()
        self.step += 1

    def memory_bytes(self):
        n = 0
        for b in self.buffers:
            n += b.data.numel() * 4 + (b.grad.numel() * 4 if b.grad is not None else 0)
        return n
