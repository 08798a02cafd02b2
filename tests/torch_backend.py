"""TEST INFRASTRUCTURE ONLY -- a plain-PyTorch fp32 implementation of every engine backend call.

Two uses, both inside tests/:
  * `-m "not gpu"` tests drive the engine's host logic (program construction, backward planning,
    accumulate flags, scratch placement, data-parallel sharding) on CPU tensors against the oracle.
  * `-m gpu` kernel tests run the SAME descriptor through libhdn.so and through this file and
    compare ("numerics tests for a CUDA kernel compare it against a plain PyTorch fp32 reference
    of the same op").
The product package never imports this module; without libhdn.so the product raises.
"""
import numpy as np
import torch
import torch.nn.functional as F

CLASS_W = (0.78, 0.65, 8.57)


def _win(view, grad=False):
    t = view.buf.grad if grad else view.buf.data
    return t.view(view.N, view.D, view.H, view.W, view.buf.C)[..., view.coff:view.coff + view.C]


def _pro(src, x):
    f = src.act.fold
    if f is not None:
        x = x * f.a + f.b
    if src.act.relu:
        x = torch.relu(x)
    return x


def _up(x, up):
    for ax, u in enumerate(up):
        if u != 1:
            x = x.repeat_interleave(u, dim=1 + ax)
    return x


def _cf(x):
    return x.permute(0, 4, 1, 2, 3)


def _cl(x):
    return x.permute(0, 2, 3, 4, 1)


def _conv_w(d):
    return d.w.permute(4, 3, 0, 1, 2)


def drop_scale(seed, n, keep):
    """hdn_drop_scale(seed, idx, keep) for idx = 0..n-1 (dense element index of the [M, C] output), restated from
    csrc/hdn_common.cuh: a splitmix-style 64-bit hash, its top 24 bits as a uniform in [0, 1), scale = 1/keep where u < keep."""
    with np.errstate(over="ignore"):
        x = np.uint64(seed) * np.uint64(0x100000001B3) + np.arange(n, dtype=np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = ((x >> np.uint64(32)).astype(np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return torch.from_numpy(np.where(u < np.float32(keep), np.float32(1.0) / np.float32(keep), np.float32(0.0)).astype(np.float32))


class TorchBackend(object):
    name = "torch-reference"

    def __init__(self):
        self.launches = 0
        self.masks = []           # 0/1 dropout masks of the forward pass, in call order, engine layout (N, D, H, W, C)

    def conv_tc_supported(self, d, which):
        return False

    def _A(self, d):
        a = None
        for s in d.srcs:
            v = _up(_pro(s, _win(s.act.view)), s.up)
            a = v if a is None else a + v
        return a

    def conv_fprop(self, d):
        y = F.conv3d(_cf(self._A(d)), _conv_w(d), d.bias, stride=d.s, padding=d.p)
        y = _cl(y)
        o = d.out
        y = y[:, :o.D, :o.H, :o.W]
        if d.drop_keep < 1.0:                # tf.nn.dropout (KB:2888): the engine's stateless hash mask, restated (hdn_common.cuh)
            scale = drop_scale(d.drop_seed, y.numel(), d.drop_keep).view(y.shape)
            self.masks.append((scale > 0).to(torch.float32))
            y = y * scale
        _win(o).copy_(y)
        if d.stat is not None:
            yd = y.double().reshape(-1, o.C)
            d.stat[0].add_(yd.sum(0))
            d.stat[1].add_((yd * yd).sum(0))

    def conv_dgrad(self, d, epis):
        dy = _cf(_win(d.out, grad=True)).contiguous()
        A = _cf(self._A(d)).detach().clone().requires_grad_(True)
        y = F.conv3d(A, _conv_w(d), None, stride=d.s, padding=d.p)
        o = d.out
        y = y[:, :, :o.D, :o.H, :o.W]
        (dA,) = torch.autograd.grad(y, A, dy)
        dA = _cl(dA)
        for s, e in zip(d.srcs, epis):
            if e.mode == 2:
                continue
            dz = dA
            v = s.act.view
            if s.up != (1, 1, 1):
                dz = dz.reshape(v.N, v.D, s.up[0], v.H, s.up[1], v.W, s.up[2], v.C).sum(dim=(2, 4, 6))
            self._epilogue(s, e, dz)

    def _epilogue(self, s, e, dz):
        v, f = s.act.view, s.act.fold
        x = _win(v)
        du = dz
        if s.act.relu:
            u = x * f.a + f.b if f is not None else x
            du = torch.where(u > 0, dz, torch.zeros_like(dz))
        if e.s is not None:
            e.s[0].add_(du.double().reshape(-1, v.C).sum(0))
            xc = x.double() - (e.center.double() if e.center is not None else 0.0)
            e.s[1].add_((du.double() * xc).reshape(-1, v.C).sum(0))
        if e.mode == 0:
            g = du * f.a if f is not None else du
            tgt = _win(e.dx, grad=True)
            if e.accumulate:
                tgt.add_(g)
            else:
                tgt.copy_(g)
        else:
            tgt = e.du.view(v.M, v.C)
            if e.accumulate:
                tgt.add_(du.reshape(v.M, v.C))
            else:
                tgt.copy_(du.reshape(v.M, v.C))

    def conv_wgrad(self, d, dw, dbias):
        dy = _cf(_win(d.out, grad=True)).contiguous()
        w = _conv_w(d).detach().clone().requires_grad_(True)
        y = F.conv3d(_cf(self._A(d)), w, None, stride=d.s, padding=d.p)
        o = d.out
        y = y[:, :, :o.D, :o.H, :o.W]
        (g,) = torch.autograd.grad(y, w, dy)
        dw.add_(g.permute(2, 3, 4, 1, 0))
        if dbias is not None:
            dbias.add_(dy.sum(dim=(0, 2, 3, 4)))

    def _pool_fwd_val(self, d, x):
        if d.kind == 0:
            pad = (1, 1, 1, 1, 1, 1) if d.pool_d else (1, 1, 1, 1, 0, 0)
            k = (3, 3, 3) if d.pool_d else (1, 3, 3)
            st = (2, 2, 2) if d.pool_d else (1, 2, 2)
            return F.max_pool3d(F.pad(x, pad), k, st)
        return F.avg_pool3d(x, (1, 2, 2), (1, 2, 2))

    def pool_fwd(self, d):
        x = _cf(_pro(d.src, _win(d.src.act.view)))
        _win(d.out).copy_(_cl(self._pool_fwd_val(d, x)))

    def pool_bwd(self, d, e):
        x = _cf(_pro(d.src, _win(d.src.act.view))).detach().clone().requires_grad_(True)
        y = self._pool_fwd_val(d, x)
        (dz,) = torch.autograd.grad(y, x, _cf(_win(d.out, grad=True)).contiguous())
        self._epilogue(d.src, e, _cl(dz))

    def col_stats(self, view, stat):
        y = _win(view).double().reshape(-1, view.C)
        stat[0].add_(y.sum(0))
        stat[1].add_((y * y).sum(0))

    def bn_fold(self, f, mode):
        if mode == 1:
            M = float(f.view.M)
            mean = (f.stat[0] / M)
            var = (f.stat[1] / M - mean * mean).clamp_min(0).float()
            mean = mean.float()
            f.mov_mean.t.sub_((f.mov_mean.t - mean) * (1.0 - f.momentum))
            f.mov_var.t.sub_((f.mov_var.t - var) * (1.0 - f.momentum))
        else:
            mean, var = f.mov_mean.t, f.mov_var.t
        rstd = torch.rsqrt(var + f.eps)
        a = f.gamma.t * rstd
        b = f.beta.t - mean * a
        if f.sgamma is not None:
            a, b = f.sgamma.t * a, f.sgamma.t * b + f.sbeta.t
        f.a.copy_(a)
        f.b.copy_(b)
        f.mean.copy_(mean)
        f.rstd.copy_(rstd)

    def bn_param_grad(self, f, mode):
        S1, S2 = f.S[0], f.S[1]
        mean, rstd = f.mean.double(), f.rstd.double()
        gam, bet = f.gamma.t.double(), f.beta.t.double()
        gs = f.sgamma.t.double() if f.sgamma is not None else torch.ones_like(gam)
        Sx = rstd * S2
        if f.sgamma is not None and f.sgamma.g is not None:
            f.sbeta.g.add_(S1.float())
            f.sgamma.g.add_((gam * Sx + bet * S1).float())
        if f.gamma.g is not None:
            f.beta.g.add_((gs * S1).float())
            f.gamma.g.add_((gs * Sx).float())
        if mode == 1:
            M = float(f.view.M)
            G = gam * rstd * gs
            f.k[0].copy_(G.float())
            f.k[1].copy_((-G * rstd * Sx / M).float())
            f.k[2].copy_((-G * S1 / M).float())

    def bn_bwd_apply(self, f, accumulate):
        v = f.view
        x = _win(v).reshape(v.M, v.C)
        g = f.k[0] * f.du.view(v.M, v.C) + f.k[1] * (x - f.mean) + f.k[2]
        tgt = _win(v, grad=True)
        g = g.view(tgt.shape)
        if accumulate:
            tgt.add_(g)
        else:
            tgt.copy_(g)

    def dropout_bwd(self, view, keep, seed):
        g = _win(view, grad=True)
        g.mul_(drop_scale(seed, g.numel(), keep).view(g.shape))

    @staticmethod
    def _wce_parts(logits, labels, N, D, HW, d0, d1):
        lg = logits.view(N, D, HW, 3)
        lab = labels.view(N, D, HW)
        dmask = torch.zeros(D, dtype=torch.bool, device=lg.device)
        dmask[d0:d1] = True
        p = torch.softmax(lg, dim=-1)
        valid = ((lab == 0) | (lab == 1) | (lab == 2)) & dmask.view(1, D, 1)
        y = lab.clamp(0, 2).long()
        w = torch.tensor(CLASS_W, dtype=lg.dtype, device=lg.device)[y]
        py = p.gather(-1, y.unsqueeze(-1)).squeeze(-1)
        return p, valid, y, w, py

    def wce_accum(self, logits, labels, N, D, HW, d0, d1, acc):
        p, valid, y, w, py = self._wce_parts(logits, labels, N, D, HW, d0, d1)
        lp = torch.log(py.clamp(1e-10, 1.0))
        acc[0] += (w * lp)[valid].double().sum()
        acc[1] += valid.double().sum()

    def wce_grad(self, logits, labels, dlogits, N, D, HW, d0, d1, acc, gscale):
        p, valid, y, w, py = self._wce_parts(logits, labels, N, D, HW, d0, d1)
        cnt = max(float(acc[1]), 1.0)
        oh = F.one_hot(y, 3).to(p.dtype)
        g = (w / cnt * gscale).unsqueeze(-1) * (p - oh)
        g = g * (valid & (py >= 1e-10)).unsqueeze(-1)
        dlogits.view(N, D, HW, 3).copy_(g)

    def triplets(self, vol, out, B, S, HW):
        v = vol.view(B, S, HW)
        idx = torch.arange(S, device=vol.device)
        tri = torch.stack([(idx - 1).clamp(0, S - 1), idx, (idx + 1).clamp(0, S - 1)], dim=1)
        o = out.view(B, S, HW, -1)
        o.zero_()
        o[..., :3].copy_(v[:, tri].permute(0, 1, 3, 2))

    def cat4(self, vol, logits, out, M, k):
        o = out.view(M, 4)
        o[:, 0] = vol.view(M)
        o[:, 1:] = k * logits.view(M, 3)

    def cat4_bwd(self, dout, dlogits, M, k, accumulate):
        g = k * dout.view(M, 4)[:, 1:]
        if accumulate:
            dlogits.view(M, 3).add_(g)
        else:
            dlogits.view(M, 3).copy_(g)

    def sgd(self, p, g, m, n, lr, mu, gscale):
        gg = g[:n] * gscale
        v = mu * m[:n] - lr * gg
        m[:n].copy_(v)
        p[:n].add_(mu * v - lr * gg)

    def window_accumulate(self, logits, score, count, S, HW, z0):
        p = torch.softmax(logits.view(S, HW, 3), dim=-1)[1:S - 1]
        score.view(-1, HW, 2)[z0 + 1:z0 + S - 1] += p[..., 1:]
        count[z0 + 1:z0 + S - 1] += 1

    def window_finalize(self, score, count, Z, HW):
        score.view(Z, HW, 2).div_((count.float() + 1e-4).view(Z, 1, 1))
