// HBM-bound kernels around the convolutions: pooling, BatchNorm fold / backward, dropout
// backward, the fused weighted cross-entropy, hybrid glue, Nesterov SGD, sliding-window
// accumulation.  All are grid-stride, channel-fastest (coalesced on the NDHWC layout).
#include <stdlib.h>
#include "hdn_common.cuh"

namespace {

constexpr int ET = 256;
inline unsigned grid_for(int64_t n, int per_thread = 1) {
  int64_t b = hdn_cdiv(n, (int64_t)ET * per_thread);
  const int64_t cap = 148 * 32;   // persistent-ish: <= 32 CTAs per SM worth of blocks
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// --------------------------------------------------------------------------- pooling
__global__ void __launch_bounds__(ET) maxpool_fwd(const hdn_pool p, const int64_t total) {
  const hdn_src& s = p.src;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % p.C);
    int64_t m = i / p.C;
    int n, od, oh, ow;
    hdn_decode(m, p.D, p.H, p.W, n, od, oh, ow);
    const int d_lo = p.pool_d ? 2 * od - 1 : od, d_hi = p.pool_d ? 2 * od + 1 : od;
    float best = -INFINITY;
    int arg = 0, idx = 0;
    for (int d = d_lo; d <= d_hi; ++d)
      for (int h = 2 * oh - 1; h <= 2 * oh + 1; ++h)
        for (int w = 2 * ow - 1; w <= 2 * ow + 1; ++w, ++idx) {
          float v = 0.f;   // zero padding (ZeroPadding + VALID max-pool)
          if (d >= 0 && d < s.D && h >= 0 && h < s.H && w >= 0 && w < s.W)
            v = hdn_prologue(s, __ldg(s.t.p + hdn_src_off(s, n, d, h, w) + c), c);
          if (v > best) { best = v; arg = idx; }     // first maximum in scan order (d, h, w)
        }
    ((float*)p.y.p)[m * p.y.ldc + p.y.coff + c] = best;
    if (p.argidx) p.argidx[i] = (unsigned char)arg;
  }
}

__device__ __forceinline__ void epi_store(const hdn_dgrad_epi& e, int64_t m, int c, int C, float a, float du) {
  if (e.mode == 0) {
    float* q = (float*)e.dx.p + m * e.dx.ldc + e.dx.coff + c;
    float g = a * du;
    *q = e.accumulate ? (*q + g) : g;
  } else {
    float* q = e.du + m * C + c;
    *q = e.accumulate ? (*q + du) : du;
  }
}

// S1/S2 per-channel partials live in shared-memory bins (float), flushed once per block
// with double atomics.
__device__ __forceinline__ void bins_init(float* bins, int C) {
  for (int c = threadIdx.x; c < 2 * C; c += ET) bins[c] = 0.f;
  __syncthreads();
}
__device__ __forceinline__ void bins_flush(const float* bins, int C, double* s1, double* s2) {
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += ET) {
    if (bins[c] != 0.f) atomicAdd(s1 + c, (double)bins[c]);
    if (bins[C + c] != 0.f) atomicAdd(s2 + c, (double)bins[C + c]);
  }
}

__global__ void __launch_bounds__(ET) maxpool_bwd(const hdn_pool p, const hdn_dgrad_epi e, const int64_t total) {
  extern __shared__ float bins[];
  const hdn_src& s = p.src;
  if (e.s1) bins_init(bins, p.C);
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % p.C);
    int64_t m = i / p.C;
    int n, id, ih, iw;
    hdn_decode(m, s.D, s.H, s.W, n, id, ih, iw);
    const float x = __ldg(s.t.p + m * s.t.ldc + s.t.coff + c);
    const float a = s.pa ? __ldg(s.pa + c) : 1.f, b = s.pb ? __ldg(s.pb + c) : 0.f;
    const float u = fmaf(a, x, b);
    float dz = 0.f;
    // candidate windows o with 2o-1 <= i <= 2o+1  <=>  o in [i/2, (i+1)/2]
    const int odl = p.pool_d ? id / 2 : id, odh = p.pool_d ? (id + 1) / 2 : id;
    const int ohl = ih / 2, ohh = (ih + 1) / 2, owl = iw / 2, owh = (iw + 1) / 2;
    for (int od = odl; od <= odh; ++od) {
      if (od >= p.D) continue;
      for (int oh = ohl; oh <= ohh; ++oh) {
        if (oh >= p.H) continue;
        for (int ow = owl; ow <= owh; ++ow) {
          if (ow >= p.W) continue;
          // recompute the window's first arg-max (scan order d,h,w; padding counts as 0)
          const int dl = p.pool_d ? 2 * od - 1 : od, dh = p.pool_d ? 2 * od + 1 : od;
          float best = -INFINITY;
          bool mine = false;
          for (int d = dl; d <= dh; ++d)
            for (int h = 2 * oh - 1; h <= 2 * oh + 1; ++h)
              for (int w = 2 * ow - 1; w <= 2 * ow + 1; ++w) {
                float v = 0.f;
                if (d >= 0 && d < s.D && h >= 0 && h < s.H && w >= 0 && w < s.W)
                  v = hdn_prologue(s, __ldg(s.t.p + hdn_src_off(s, n, d, h, w) + c), c);
                if (v > best) { best = v; mine = (d == id && h == ih && w == iw); }
              }
          if (mine) dz += __ldg(p.y.p + ((((int64_t)n * p.D + od) * p.H + oh) * p.W + ow) * p.y.ldc + p.y.coff + c);
        }
      }
    }
    float du = (s.relu && !(u > 0.f)) ? 0.f : dz;
    if (e.s1 && du != 0.f) { atomicAdd(&bins[c], du); atomicAdd(&bins[p.C + c], du * (x - (e.center ? __ldg(e.center + c) : 0.f))); }
    epi_store(e, m, c, p.C, a, du);
  }
  if (e.s1) bins_flush(bins, p.C, e.s1, e.s2);
}

// backward with the arg-max taps saved by the forward pass: an input element receives dY of every window
// whose saved tap points at it (<= 2 windows per pooled axis)
__global__ void __launch_bounds__(ET) maxpool_bwd_idx(const hdn_pool p, const hdn_dgrad_epi e, const int64_t total) {
  extern __shared__ float bins[];
  const hdn_src& s = p.src;
  if (e.s1) bins_init(bins, p.C);
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % p.C);
    int64_t m = i / p.C;
    int n, id, ih, iw;
    hdn_decode(m, s.D, s.H, s.W, n, id, ih, iw);
    float dz = 0.f;
    const int odl = p.pool_d ? id / 2 : id, odh = p.pool_d ? (id + 1) / 2 : id;
    for (int od = odl; od <= odh; ++od) {
      if (od >= p.D) continue;
      const int td = p.pool_d ? id - (2 * od - 1) : 0;
      for (int oh = ih / 2; oh <= (ih + 1) / 2; ++oh) {
        if (oh >= p.H) continue;
        const int th = ih - (2 * oh - 1);
        for (int ow = iw / 2; ow <= (iw + 1) / 2; ++ow) {
          if (ow >= p.W) continue;
          const int tw = iw - (2 * ow - 1);
          const int64_t mo = (((int64_t)n * p.D + od) * p.H + oh) * p.W + ow;
          if (p.argidx[mo * p.C + c] == (unsigned char)((td * 3 + th) * 3 + tw)) dz += __ldg(p.y.p + mo * p.y.ldc + p.y.coff + c);
        }
      }
    }
    const float x = __ldg(s.t.p + m * s.t.ldc + s.t.coff + c);
    const float a = s.pa ? __ldg(s.pa + c) : 1.f, b = s.pb ? __ldg(s.pb + c) : 0.f;
    float du = (s.relu && !(fmaf(a, x, b) > 0.f)) ? 0.f : dz;
    if (e.s1 && du != 0.f) { atomicAdd(&bins[c], du); atomicAdd(&bins[p.C + c], du * (x - (e.center ? __ldg(e.center + c) : 0.f))); }
    epi_store(e, m, c, p.C, a, du);
  }
  if (e.s1) bins_flush(bins, p.C, e.s1, e.s2);
}

// Same backward, four channels per thread: a thread keeps one channel quad for its whole life (quad = tid % (C/4)), so
// S1/S2 are plain register sums flushed once, positions are decoded with 32-bit arithmetic once per quad instead of
// five 64-bit divisions per element, and every access is a 16-byte (or 4-byte arg-max) vector.  Needs C % 4 == 0,
// 16-byte aligned windows and M < 2^31 (checked by the caller).
__global__ void __launch_bounds__(ET) maxpool_bwd_idx_v4(const hdn_pool p, const hdn_dgrad_epi e, const unsigned M_in) {
  extern __shared__ float bins[];
  const hdn_src& s = p.src;
  if (e.s1) bins_init(bins, p.C);
  const unsigned cq = (unsigned)p.C >> 2, ppb = ET / cq;               // quads per position, positions per block pass
  const unsigned quad = threadIdx.x % cq, pl = threadIdx.x / cq;
  const int c = (int)quad * 4;
  float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f), ctr = b;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < ppb) {
    if (s.pa) a = __ldg(reinterpret_cast<const float4*>(s.pa + c));
    if (s.pb) b = __ldg(reinterpret_cast<const float4*>(s.pb + c));
    if (e.s1 && e.center) ctr = __ldg(reinterpret_cast<const float4*>(e.center + c));
    const unsigned sD = (unsigned)s.D, sH = (unsigned)s.H, sW = (unsigned)s.W;
    for (unsigned m = blockIdx.x * ppb + pl; m < M_in; m += gridDim.x * ppb) {
      unsigned r = m;
      const int iw = (int)(r % sW); r /= sW;
      const int ih = (int)(r % sH); r /= sH;
      const int id = (int)(r % sD);
      const int n = (int)(r / sD);
      float dz[4] = {0.f, 0.f, 0.f, 0.f};
      const int odl = p.pool_d ? id / 2 : id, odh = p.pool_d ? (id + 1) / 2 : id;
      for (int od = odl; od <= odh; ++od) {
        if (od >= p.D) continue;
        const int td = p.pool_d ? id - (2 * od - 1) : 0;
        for (int oh = ih / 2; oh <= (ih + 1) / 2; ++oh) {
          if (oh >= p.H) continue;
          const int th = ih - (2 * oh - 1);
          for (int ow = iw / 2; ow <= (iw + 1) / 2; ++ow) {
            if (ow >= p.W) continue;
            const int tw = iw - (2 * ow - 1);
            const unsigned char tap = (unsigned char)((td * 3 + th) * 3 + tw);
            const int64_t mo = (((int64_t)n * p.D + od) * p.H + oh) * p.W + ow;
            const uchar4 t4 = *reinterpret_cast<const uchar4*>(p.argidx + mo * p.C + c);
            if (t4.x == tap || t4.y == tap || t4.z == tap || t4.w == tap) {
              const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.y.p + mo * p.y.ldc + p.y.coff + c));
              if (t4.x == tap) dz[0] += y4.x;
              if (t4.y == tap) dz[1] += y4.y;
              if (t4.z == tap) dz[2] += y4.z;
              if (t4.w == tap) dz[3] += y4.w;
            }
          }
        }
      }
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(s.t.p + (int64_t)m * s.t.ldc + s.t.coff + c));
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
      const float cs[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
      float du[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        du[k] = (s.relu && !(fmaf(as[k], xs[k], bs[k]) > 0.f)) ? 0.f : dz[k];
        s1[k] += du[k];
        s2[k] += du[k] * (xs[k] - cs[k]);
      }
      if (e.mode == 0) {
        float4* q = reinterpret_cast<float4*>((float*)e.dx.p + (int64_t)m * e.dx.ldc + e.dx.coff + c);
        float4 g = make_float4(as[0] * du[0], as[1] * du[1], as[2] * du[2], as[3] * du[3]);
        if (e.accumulate) { const float4 o = *q; g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
        *q = g;
      } else {
        float4* q = reinterpret_cast<float4*>(e.du + (int64_t)m * p.C + c);
        float4 g = make_float4(du[0], du[1], du[2], du[3]);
        if (e.accumulate) { const float4 o = *q; g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
        *q = g;
      }
    }
    if (e.s1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (s1[k] != 0.f) atomicAdd(&bins[c + k], s1[k]);
        if (s2[k] != 0.f) atomicAdd(&bins[p.C + c + k], s2[k]);
      }
    }
  }
  if (e.s1) bins_flush(bins, p.C, e.s1, e.s2);
}

__global__ void __launch_bounds__(ET) avgpool_fwd(const hdn_pool p, const int64_t total) {
  const hdn_src& s = p.src;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % p.C);
    int64_t m = i / p.C;
    int n, od, oh, ow;
    hdn_decode(m, p.D, p.H, p.W, n, od, oh, ow);
    float acc = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw)
        acc += hdn_prologue(s, __ldg(s.t.p + hdn_src_off(s, n, od, 2 * oh + dh, 2 * ow + dw) + c), c);
    ((float*)p.y.p)[m * p.y.ldc + p.y.coff + c] = 0.25f * acc;
  }
}

__global__ void __launch_bounds__(ET) avgpool_bwd(const hdn_pool p, const hdn_dgrad_epi e, const int64_t total) {
  extern __shared__ float bins[];
  const hdn_src& s = p.src;
  if (e.s1) bins_init(bins, p.C);
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % p.C);
    int64_t m = i / p.C;
    int n, id, ih, iw;
    hdn_decode(m, s.D, s.H, s.W, n, id, ih, iw);
    const int oh = ih >> 1, ow = iw >> 1;
    float dz = 0.f;
    if (oh < p.H && ow < p.W)
      dz = 0.25f * __ldg(p.y.p + ((((int64_t)n * p.D + id) * p.H + oh) * p.W + ow) * p.y.ldc + p.y.coff + c);
    const float x = __ldg(s.t.p + m * s.t.ldc + s.t.coff + c);
    const float a = s.pa ? __ldg(s.pa + c) : 1.f, b = s.pb ? __ldg(s.pb + c) : 0.f;
    float du = (s.relu && !(fmaf(a, x, b) > 0.f)) ? 0.f : dz;
    if (e.s1 && du != 0.f) { atomicAdd(&bins[c], du); atomicAdd(&bins[p.C + c], du * (x - (e.center ? __ldg(e.center + c) : 0.f))); }
    epi_store(e, m, c, p.C, a, du);
  }
  if (e.s1) bins_flush(bins, p.C, e.s1, e.s2);
}

// --------------------------------------------------------------------------- batch norm
__global__ void bn_fold_kernel(const hdn_bn_fold_t f) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.C) return;
  float mean, var;
  if (f.mode == 1) {
    double mu = f.sum[c] / f.count;
    double v = f.sumsq[c] / f.count - mu * mu;
    if (v < 0.0) v = 0.0;
    mean = (float)mu; var = (float)v;
    if (f.mov_mean) {
      f.mov_mean[c] -= (f.mov_mean[c] - mean) * (1.0f - f.momentum);
      f.mov_var[c] -= (f.mov_var[c] - var) * (1.0f - f.momentum);
    }
  } else {
    mean = f.mov_mean[c]; var = f.mov_var[c];
  }
  float rstd = rsqrtf(var + f.eps);
  // full-precision reciprocal sqrt (rsqrtf is 2 ulp; refine once)
  rstd = rstd * (1.5f - 0.5f * (var + f.eps) * rstd * rstd);
  float g = f.gamma ? f.gamma[c] : 1.f, bt = f.beta ? f.beta[c] : 0.f;
  float a = g * rstd, b = bt - mean * a;
  if (f.sgamma) { float gs = f.sgamma[c]; a = gs * a; b = gs * b + (f.sbeta ? f.sbeta[c] : 0.f); }
  f.a[c] = a; f.b[c] = b;
  if (f.mean) f.mean[c] = mean;
  if (f.rstd) f.rstd[c] = rstd;
}

__global__ void bn_param_grad_kernel(const hdn_bn_grad_t g) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.C) return;
  const double S1 = g.s1[c], S2 = g.s2[c];
  const double mean = g.mean[c], rstd = g.rstd[c];
  const double gam = g.gamma ? g.gamma[c] : 1.0, bet = g.beta ? g.beta[c] : 0.0;
  const double gs = g.sgamma ? g.sgamma[c] : 1.0;
  const double Sx = rstd * S2;                      // sum du * xhat   (S2 = sum du*(x-mean))
  (void)mean;
  if (g.dsbeta) g.dsbeta[c] += (float)S1;
  if (g.dsgamma) g.dsgamma[c] += (float)(gam * Sx + bet * S1);
  if (g.dbeta) g.dbeta[c] += (float)(gs * S1);
  if (g.dgamma) g.dgamma[c] += (float)(gs * Sx);
  if (g.mode == 1 && g.k0) {
    const double G = gam * rstd * gs, M = g.count;
    g.k0[c] = (float)G;
    g.k1[c] = (float)(-G * rstd * Sx / M);
    g.k2[c] = (float)(-G * S1 / M);
  }
}

__global__ void __launch_bounds__(ET) bn_bwd_apply_kernel(const float* __restrict__ du, hdn_tensor x, hdn_tensor dx,
                                                           int64_t total, int C, const float* k0, const float* k1,
                                                           const float* k2, const float* mean, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % C);
    int64_t m = i / C;
    float v = fmaf(k0[c], du[i], fmaf(k1[c], __ldg(x.p + m * x.ldc + x.coff + c) - mean[c], k2[c]));
    float* q = (float*)dx.p + m * dx.ldc + dx.coff + c;
    *q = accumulate ? (*q + v) : v;
  }
}

__global__ void __launch_bounds__(ET) dropout_bwd_kernel(hdn_tensor g, int64_t total, int C, float keep, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int c = (int)(i % C);
    int64_t m = i / C;
    float* q = (float*)g.p + m * g.ldc + g.coff + c;
    *q *= hdn_drop_scale(seed, (uint64_t)i, keep);
  }
}

// --------------------------------------------------------------------------- loss
__device__ __forceinline__ void softmax3(const float* l, float* p) {
  float m = fmaxf(l[0], fmaxf(l[1], l[2]));
  float e0 = expf(l[0] - m), e1 = expf(l[1] - m), e2 = expf(l[2] - m);
  float inv = 1.0f / (e0 + e1 + e2);
  p[0] = e0 * inv; p[1] = e1 * inv; p[2] = e2 * inv;
}
__device__ __forceinline__ int label_class(float y) { return (y == 0.f) ? 0 : (y == 1.f) ? 1 : (y == 2.f) ? 2 : -1; }
__constant__ float kClassW[3] = {0.78f, 0.65f, 8.57f};   // loss.py:23

__global__ void __launch_bounds__(ET) wce_accum_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                        int64_t total, int D, int64_t HW, int d0, int d1, double* acc) {
  double s = 0.0, n = 0.0;
  for (int64_t v = blockIdx.x * (int64_t)ET + threadIdx.x; v < total; v += (int64_t)gridDim.x * ET) {
    int d = (int)((v / HW) % D);
    if (d < d0 || d >= d1) continue;
    int y = label_class(labels[v]);
    if (y < 0) continue;
    float l[3] = {logits[3 * v], logits[3 * v + 1], logits[3 * v + 2]}, p[3];
    softmax3(l, p);
    float py = fminf(fmaxf(p[y], 1e-10f), 1.0f);
    s += (double)(kClassW[y] * logf(py));
    n += 1.0;
  }
  s = warp_sum_d(s); n = warp_sum_d(n);
  __shared__ double rs[ET / 32], rn[ET / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { rs[w] = s; rn[w] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tn = 0;
    for (int i = 0; i < ET / 32; ++i) { ts += rs[i]; tn += rn[i]; }
    atomicAdd(acc, ts); atomicAdd(acc + 1, tn);
  }
}

__global__ void __launch_bounds__(ET) wce_grad_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                       float* __restrict__ dl, int64_t total, int D, int64_t HW, int d0,
                                                       int d1, const double* acc, float gscale) {
  const float invn = (float)(1.0 / fmax(acc[1], 1.0)) * gscale;
  for (int64_t v = blockIdx.x * (int64_t)ET + threadIdx.x; v < total; v += (int64_t)gridDim.x * ET) {
    int d = (int)((v / HW) % D);
    int y = label_class(labels[v]);
    float g[3] = {0.f, 0.f, 0.f};
    if (d >= d0 && d < d1 && y >= 0) {
      float l[3] = {logits[3 * v], logits[3 * v + 1], logits[3 * v + 2]}, p[3];
      softmax3(l, p);
      if (p[y] >= 1e-10f) {   // gradient of clip_by_value is zero outside [1e-10, 1]
        float k = kClassW[y] * invn;
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = k * (p[c] - (c == y ? 1.f : 0.f));
      }
    }
    dl[3 * v] = g[0]; dl[3 * v + 1] = g[1]; dl[3 * v + 2] = g[2];
  }
}

// --------------------------------------------------------------------------- hybrid glue
__global__ void __launch_bounds__(ET) triplets_kernel(const float* __restrict__ vol, float* __restrict__ out, int B, int S,
                                                       int64_t HW, int ldc) {
  const int64_t total = (int64_t)B * S * HW;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int64_t hw = i % HW;
    int s = (int)((i / HW) % S), b = (int)(i / (HW * S));
    const float* base = vol + (int64_t)b * S * HW + hw;
    int sm = s > 0 ? s - 1 : 0, sp = s < S - 1 ? s + 1 : S - 1;
    float* o = out + i * ldc;
    o[0] = base[(int64_t)sm * HW];
    o[1] = base[(int64_t)s * HW];
    o[2] = base[(int64_t)sp * HW];
    for (int c = 3; c < ldc; ++c) o[c] = 0.f;      // padding channels (16-byte pixels for the tensor-core stem)
  }
}
__global__ void __launch_bounds__(ET) cat4_kernel(const float* __restrict__ vol, const float* __restrict__ lg,
                                                   float4* __restrict__ out, int64_t M, float k) {
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < M; i += (int64_t)gridDim.x * ET)
    out[i] = make_float4(vol[i], k * lg[3 * i], k * lg[3 * i + 1], k * lg[3 * i + 2]);
}
__global__ void __launch_bounds__(ET) cat4_bwd_kernel(const float4* __restrict__ dout, float* __restrict__ dl, int64_t M,
                                                       float k, int accumulate) {
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < M; i += (int64_t)gridDim.x * ET) {
    float4 g = dout[i];
    if (accumulate) { dl[3 * i] += k * g.y; dl[3 * i + 1] += k * g.z; dl[3 * i + 2] += k * g.w; }
    else { dl[3 * i] = k * g.y; dl[3 * i + 1] = k * g.z; dl[3 * i + 2] = k * g.w; }
  }
}

// --------------------------------------------------------------------------- optimizer
__global__ void __launch_bounds__(ET) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  int64_t n, float lr, float mu, float gs) {
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < n; i += (int64_t)gridDim.x * ET) {
    float gi = g[i] * gs;
    float v = mu * m[i] - lr * gi;
    m[i] = v;
    p[i] = p[i] + mu * v - lr * gi;
  }
}

struct PeerTable { float* p[16]; const float* g[16]; };
// rank-owned shard [lo,hi): pull the peers' gradients over NVLink, Nesterov update, push params.
__global__ void __launch_bounds__(ET) dp_reduce_sgd_kernel(const PeerTable t, float* __restrict__ m, int world, int rank,
                                                            int64_t lo, int64_t hi, float lr, float mu, float gs) {
  // 4-wide when the shard start is 16-byte aligned on every arena (the host guarantees lo % 4 == 0)
  const int64_t n4 = (hi - lo) / 4;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < n4; i += (int64_t)gridDim.x * ET) {
    const int64_t e = lo + 4 * i;
    float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      float4 gi = *reinterpret_cast<const float4*>(t.g[r] + e);
      gsum.x += gi.x; gsum.y += gi.y; gsum.z += gi.z; gsum.w += gi.w;
    }
    float4 mm = *reinterpret_cast<float4*>(m + e);
    float4 pp = *reinterpret_cast<const float4*>(t.p[rank] + e);
    float gx = gsum.x * gs, gy = gsum.y * gs, gz = gsum.z * gs, gw = gsum.w * gs;
    mm.x = mu * mm.x - lr * gx; mm.y = mu * mm.y - lr * gy; mm.z = mu * mm.z - lr * gz; mm.w = mu * mm.w - lr * gw;
    pp.x += mu * mm.x - lr * gx; pp.y += mu * mm.y - lr * gy; pp.z += mu * mm.z - lr * gz; pp.w += mu * mm.w - lr * gw;
    *reinterpret_cast<float4*>(m + e) = mm;
    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(t.p[r] + e) = pp;
  }
  // tail
  for (int64_t e = lo + 4 * n4 + blockIdx.x * (int64_t)ET + threadIdx.x; e < hi; e += (int64_t)gridDim.x * ET) {
    float gsum = 0.f;
    for (int r = 0; r < world; ++r) gsum += t.g[r][e];
    gsum *= gs;
    float v = mu * m[e] - lr * gsum;
    m[e] = v;
    float pn = t.p[rank][e] + mu * v - lr * gsum;
    for (int r = 0; r < world; ++r) t.p[r][e] = pn;
  }
}

// Device-side step flags of the data-parallel exchange (peer-mapped, one uint32 slot per (owner, writer) pair): the
// gradient / parameter arenas are handed over between the per-GPU processes on the compute stream itself, without a
// host barrier.  signal: after a system-scope fence (everything this GPU's earlier kernels wrote is visible to its
// peers) store `value` into slot [rank] of every peer's flag row.  wait: spin until all `world` slots of the local row
// have reached `value`.
struct FlagTable { unsigned int* f[16]; };
__global__ void dp_signal_kernel(const FlagTable t, int world, int rank, unsigned int value) {
  if (threadIdx.x < world) {
    __threadfence_system();
    volatile unsigned int* q = t.f[threadIdx.x] + rank;
    *q = value;
    __threadfence_system();
  }
}
__global__ void dp_wait_kernel(const unsigned int* flags, int world, unsigned int value) {
  if (threadIdx.x < world) {
    const volatile unsigned int* q = flags + threadIdx.x;
    while ((int)(*q - value) < 0) __nanosleep(200);     // wrap-safe "slot < value"
    __threadfence_system();
  }
}

// --------------------------------------------------------------------------- host layout -> device layout
// Reference volumes arrive as (N,H,W,S) (channel 1 dropped); the engine's layout is (N,S,H,W).  One pass over the
// staged tensor, reads along W coalesced per slice through a 32x33 shared tile over (W, S).  T = float or short
// (int16 label maps are converted to the fp32 the loss kernels read, train_hybrid.py:127-132).
template <typename T>
__global__ void __launch_bounds__(256) nhws_to_nshw_kernel(const T* __restrict__ in, float* __restrict__ out, int H, int W, int S) {
  __shared__ float tile[32][33];
  const int64_t nh = blockIdx.z;                      // n * H + h
  const int n = (int)(nh / H), h = (int)(nh % H);
  const int w0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {                   // rows = w, fastest = s (contiguous in the source)
    const int w = w0 + i, sidx = s0 + tx;
    if (w < W && sidx < S) tile[i][tx] = (float)in[(nh * W + w) * S + sidx];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {                   // rows = s, fastest = w (contiguous in the destination)
    const int sidx = s0 + i, w = w0 + tx;
    if (w < W && sidx < S) out[(((int64_t)n * S + sidx) * H + h) * W + w] = tile[tx][i];
  }
}

// --------------------------------------------------------------------------- sliding window
__global__ void __launch_bounds__(ET) window_acc_kernel(const float* __restrict__ logits, float* __restrict__ score,
                                                         int S, int64_t HW, int z0) {
  const int64_t total = (int64_t)(S - 2) * HW;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    int64_t hw = i % HW;
    int s = (int)(i / HW) + 1;
    const float* l = logits + ((int64_t)s * HW + hw) * 3;
    float lv[3] = {l[0], l[1], l[2]}, p[3];
    softmax3(lv, p);
    float* q = score + ((int64_t)(z0 + s) * HW + hw) * 2;
    q[0] += p[1]; q[1] += p[2];
  }
}
__global__ void window_cnt_kernel(int* count, int S, int z0) {
  int s = threadIdx.x + 1;
  if (s < S - 1) count[z0 + s] += 1;
}
__global__ void __launch_bounds__(ET) window_fin_kernel(float* __restrict__ score, const int* __restrict__ count, int Z,
                                                         int64_t HW) {
  const int64_t total = (int64_t)Z * HW;
  for (int64_t i = blockIdx.x * (int64_t)ET + threadIdx.x; i < total; i += (int64_t)gridDim.x * ET) {
    float inv = 1.0f / ((float)count[i / HW] + 1e-4f);
    score[2 * i] *= inv; score[2 * i + 1] *= inv;
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

// HDN_POOL_FAST=0/1: four-channels-per-thread max-pool backward (read once per process; default 1, validated on B200 in
// round 1: profiles/r01c_*)
static int hdn_pool_fast() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDN_POOL_FAST");
    v = e ? (atoi(e) != 0 ? 1 : 0) : 1;
  }
  return v;
}
static bool aligned16(const hdn_tensor& t) {
  return t.ldc % 4 == 0 && t.coff % 4 == 0 && (reinterpret_cast<uintptr_t>(t.p) & 15) == 0;
}

extern "C" int hdn_pool_fwd(const hdn_pool* p, void* stream) {
  HDN_CHECK_ARG(p && p->src.t.p && p->y.p, "pool_fwd: null pointer");
  HDN_CHECK_ARG(p->kind == 0 || p->kind == 1, "pool: kind must be 0 (max) or 1 (avg)");
  const int64_t total = (int64_t)p->N * p->D * p->H * p->W * p->C;
  if (p->kind == 0) HDN_LAUNCHED(1), maxpool_fwd<<<grid_for(total), ET, 0, ST>>>(*p, total);
  else HDN_LAUNCHED(1), avgpool_fwd<<<grid_for(total), ET, 0, ST>>>(*p, total);
  HDN_CHECK_LAUNCH("pool_fwd");
  return HDN_OK;
}
extern "C" int hdn_pool_bwd(const hdn_pool* p, const hdn_dgrad_epi* e, void* stream) {
  HDN_CHECK_ARG(p && e && p->src.t.p && p->y.p, "pool_bwd: null pointer");
  const int64_t total = (int64_t)p->N * p->src.D * p->src.H * p->src.W * p->C;
  HDN_CHECK_ARG(p->C <= 4096, "pool_bwd: C > 4096 unsupported");
  const size_t sm = e->s1 ? 2 * (size_t)p->C * sizeof(float) : 0;
  const int64_t M_in = (int64_t)p->N * p->src.D * p->src.H * p->src.W;
  const bool v4 = hdn_pool_fast() && p->kind == 0 && p->argidx && p->C % 4 == 0 && p->C / 4 <= ET && M_in < (1ll << 31) &&
                  aligned16(p->src.t) && aligned16(p->y) &&
                  (e->mode == 0 ? aligned16(e->dx) : (reinterpret_cast<uintptr_t>(e->du) & 15) == 0) &&
                  (!p->src.pa || (reinterpret_cast<uintptr_t>(p->src.pa) & 15) == 0) &&
                  (!p->src.pb || (reinterpret_cast<uintptr_t>(p->src.pb) & 15) == 0) &&
                  (!e->center || (reinterpret_cast<uintptr_t>(e->center) & 15) == 0) &&
                  (reinterpret_cast<uintptr_t>(p->argidx) & 3) == 0;
  if (v4) {
    const unsigned ppb = ET / (unsigned)(p->C / 4);
    int64_t blocks = hdn_cdiv(M_in, (int64_t)ppb);
    if (blocks > 148 * 16) blocks = 148 * 16;
    HDN_LAUNCHED(1), maxpool_bwd_idx_v4<<<(unsigned)blocks, ET, sm, ST>>>(*p, *e, (unsigned)M_in);
  } else if (p->kind == 0 && p->argidx) HDN_LAUNCHED(1), maxpool_bwd_idx<<<grid_for(total), ET, sm, ST>>>(*p, *e, total);
  else if (p->kind == 0) HDN_LAUNCHED(1), maxpool_bwd<<<grid_for(total), ET, sm, ST>>>(*p, *e, total);
  else HDN_LAUNCHED(1), avgpool_bwd<<<grid_for(total), ET, sm, ST>>>(*p, *e, total);
  HDN_CHECK_LAUNCH("pool_bwd");
  return HDN_OK;
}
extern "C" int hdn_bn_fold(const hdn_bn_fold_t* f, void* stream) {
  HDN_CHECK_ARG(f && f->C > 0 && f->a && f->b, "bn_fold: bad descriptor");
  HDN_CHECK_ARG(f->mode == 0 || (f->sum && f->sumsq && f->count > 0), "bn_fold: training mode needs statistics");
  HDN_CHECK_ARG(f->mode == 1 || (f->mov_mean && f->mov_var), "bn_fold: inference mode needs moving statistics");
  HDN_LAUNCHED(1), bn_fold_kernel<<<(unsigned)hdn_cdiv(f->C, 128), 128, 0, ST>>>(*f);
  HDN_CHECK_LAUNCH("bn_fold");
  return HDN_OK;
}
extern "C" int hdn_bn_param_grad(const hdn_bn_grad_t* g, void* stream) {
  HDN_CHECK_ARG(g && g->C > 0 && g->s1 && g->s2 && g->mean && g->rstd, "bn_param_grad: bad descriptor");
  HDN_LAUNCHED(1), bn_param_grad_kernel<<<(unsigned)hdn_cdiv(g->C, 128), 128, 0, ST>>>(*g);
  HDN_CHECK_LAUNCH("bn_param_grad");
  return HDN_OK;
}
extern "C" int hdn_bn_bwd_apply(const float* du, hdn_tensor x, hdn_tensor dx, int64_t M, int C, const float* k0,
                                const float* k1, const float* k2, const float* mean, int accumulate, void* stream) {
  HDN_CHECK_ARG(du && x.p && dx.p && k0 && k1 && k2 && mean && M > 0 && C > 0, "bn_bwd_apply: bad arguments");
  HDN_LAUNCHED(1), bn_bwd_apply_kernel<<<grid_for(M * C, 4), ET, 0, ST>>>(du, x, dx, M * C, C, k0, k1, k2, mean, accumulate);
  HDN_CHECK_LAUNCH("bn_bwd_apply");
  return HDN_OK;
}
extern "C" int hdn_dropout_bwd(hdn_tensor g, int64_t M, int C, float keep, uint64_t seed, void* stream) {
  HDN_CHECK_ARG(g.p && M > 0 && C > 0 && keep > 0.f && keep <= 1.f, "dropout_bwd: bad arguments");
  HDN_LAUNCHED(1), dropout_bwd_kernel<<<grid_for(M * C, 4), ET, 0, ST>>>(g, M * C, C, keep, seed);
  HDN_CHECK_LAUNCH("dropout_bwd");
  return HDN_OK;
}
extern "C" int hdn_wce_accum(const float* logits, const float* labels, int64_t N, int D, int64_t HW, int d0, int d1,
                             double* acc, void* stream) {
  HDN_CHECK_ARG(logits && labels && acc && N > 0 && D > 0 && HW > 0, "wce_accum: bad arguments");
  const int64_t total = N * D * HW;
  HDN_LAUNCHED(1), wce_accum_kernel<<<grid_for(total, 4), ET, 0, ST>>>(logits, labels, total, D, HW, d0, d1, acc);
  HDN_CHECK_LAUNCH("wce_accum");
  return HDN_OK;
}
extern "C" int hdn_wce_grad(const float* logits, const float* labels, float* dlogits, int64_t N, int D, int64_t HW,
                            int d0, int d1, const double* acc, float gscale, void* stream) {
  HDN_CHECK_ARG(logits && labels && dlogits && acc, "wce_grad: null pointer");
  const int64_t total = N * D * HW;
  HDN_LAUNCHED(1), wce_grad_kernel<<<grid_for(total, 4), ET, 0, ST>>>(logits, labels, dlogits, total, D, HW, d0, d1, acc, gscale);
  HDN_CHECK_LAUNCH("wce_grad");
  return HDN_OK;
}
extern "C" int hdn_triplets(const float* vol, float* out, int B, int S, int64_t HW, int ldc, void* stream) {
  HDN_CHECK_ARG(vol && out && B > 0 && S > 0 && HW > 0 && ldc >= 3, "triplets: bad arguments");
  HDN_LAUNCHED(1), triplets_kernel<<<grid_for((int64_t)B * S * HW, 2), ET, 0, ST>>>(vol, out, B, S, HW, ldc);
  HDN_CHECK_LAUNCH("triplets");
  return HDN_OK;
}
extern "C" int hdn_cat4(const float* vol, const float* logits, float* out, int64_t M, float k, void* stream) {
  HDN_CHECK_ARG(vol && logits && out && M > 0, "cat4: bad arguments");
  HDN_LAUNCHED(1), cat4_kernel<<<grid_for(M, 2), ET, 0, ST>>>(vol, logits, (float4*)out, M, k);
  HDN_CHECK_LAUNCH("cat4");
  return HDN_OK;
}
extern "C" int hdn_cat4_bwd(const float* dout, float* dlogits, int64_t M, float k, int accumulate, void* stream) {
  HDN_CHECK_ARG(dout && dlogits && M > 0, "cat4_bwd: bad arguments");
  HDN_LAUNCHED(1), cat4_bwd_kernel<<<grid_for(M, 2), ET, 0, ST>>>((const float4*)dout, dlogits, M, k, accumulate);
  HDN_CHECK_LAUNCH("cat4_bwd");
  return HDN_OK;
}
extern "C" int hdn_sgd_nesterov(float* p, const float* g, float* m, int64_t n, float lr, float mu, float gscale,
                                void* stream) {
  HDN_CHECK_ARG(p && g && m && n > 0, "sgd: bad arguments");
  HDN_LAUNCHED(1), sgd_kernel<<<grid_for(n, 4), ET, 0, ST>>>(p, g, m, n, lr, mu, gscale);
  HDN_CHECK_LAUNCH("sgd");
  return HDN_OK;
}
extern "C" int hdn_dp_reduce_sgd(float* const* peer_p, const float* const* peer_g, float* m_local, int world, int rank,
                                 int64_t lo, int64_t hi, float lr, float mu, float gscale, void* stream) {
  HDN_CHECK_ARG(peer_p && peer_g && m_local && world >= 1 && world <= 16 && rank >= 0 && rank < world,
                "dp_reduce_sgd: bad arguments");
  HDN_CHECK_ARG(lo % 4 == 0 && hi >= lo, "dp_reduce_sgd: shard start must be a multiple of 4");
  if (hi == lo) return HDN_OK;
  PeerTable t;
  for (int r = 0; r < world; ++r) { t.p[r] = peer_p[r]; t.g[r] = peer_g[r]; }
  HDN_LAUNCHED(1), dp_reduce_sgd_kernel<<<grid_for(hi - lo, 8), ET, 0, ST>>>(t, m_local, world, rank, lo, hi, lr, mu, gscale);
  HDN_CHECK_LAUNCH("dp_reduce_sgd");
  return HDN_OK;
}
extern "C" int hdn_window_accumulate(const float* logits, float* score, int* count, int S, int64_t HW, int z0,
                                     void* stream) {
  HDN_CHECK_ARG(logits && score && count && S > 2 && S <= 1024 && HW > 0, "window_accumulate: bad arguments");
  HDN_LAUNCHED(1), window_acc_kernel<<<grid_for((int64_t)(S - 2) * HW, 2), ET, 0, ST>>>(logits, score, S, HW, z0);
  HDN_LAUNCHED(1), window_cnt_kernel<<<1, 1024, 0, ST>>>(count, S, z0);
  HDN_CHECK_LAUNCH("window_accumulate");
  return HDN_OK;
}
extern "C" int hdn_window_finalize(float* score, const int* count, int Z, int64_t HW, void* stream) {
  HDN_CHECK_ARG(score && count && Z > 0 && HW > 0, "window_finalize: bad arguments");
  HDN_LAUNCHED(1), window_fin_kernel<<<grid_for((int64_t)Z * HW, 2), ET, 0, ST>>>(score, count, Z, HW);
  HDN_CHECK_LAUNCH("window_finalize");
  return HDN_OK;
}

extern "C" int hdn_dp_signal(unsigned int* const* peer_flags, int world, int rank, unsigned int value, void* stream) {
  HDN_CHECK_ARG(peer_flags && world >= 1 && world <= 16 && rank >= 0 && rank < world, "dp_signal: bad arguments");
  FlagTable t;
  for (int r = 0; r < world; ++r) t.f[r] = peer_flags[r];
  HDN_LAUNCHED(1), dp_signal_kernel<<<1, 32, 0, ST>>>(t, world, rank, value);
  HDN_CHECK_LAUNCH("dp_signal");
  return HDN_OK;
}
extern "C" int hdn_dp_wait(const unsigned int* flags, int world, unsigned int value, void* stream) {
  HDN_CHECK_ARG(flags && world >= 1 && world <= 16, "dp_wait: bad arguments");
  HDN_LAUNCHED(1), dp_wait_kernel<<<1, 32, 0, ST>>>(flags, world, value);
  HDN_CHECK_LAUNCH("dp_wait");
  return HDN_OK;
}
extern "C" int hdn_layout_nhws_to_nshw(const void* in, float* out, int N, int H, int W, int S, int is_int16, void* stream) {
  HDN_CHECK_ARG(in && out && N > 0 && H > 0 && W > 0 && S > 0 && (int64_t)N * H < 65536, "layout: bad arguments");
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((S + 31) / 32), (unsigned)(N * H));
  if (is_int16) HDN_LAUNCHED(1), nhws_to_nshw_kernel<short><<<grid, 256, 0, ST>>>(reinterpret_cast<const short*>(in), out, H, W, S);
  else HDN_LAUNCHED(1), nhws_to_nshw_kernel<float><<<grid, 256, 0, ST>>>(reinterpret_cast<const float*>(in), out, H, W, S);
  HDN_CHECK_LAUNCH("layout_nhws_to_nshw");
  return HDN_OK;
}
