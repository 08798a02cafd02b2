// fp32 FMA implicit-GEMM convolution family (precision == 0): the parity path and the
// fallback for shapes the tcgen05 path does not take (Cin = 3/4 stems, Cout = 3 classifiers).
// fprop / dgrad / wgrad all share the same gather: output position x tap -> (virtually
// up-sampled, zero-padded) source position, with the BN->Scale->ReLU prologue applied on load.
#include "hdn_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Geo {
  int Dv, Hv, Wv;   // virtual (up-sampled) input dims
};

__device__ __forceinline__ float load_a(const hdn_conv& c, int n, int vd, int vh, int vw, int ci) {
  float v;
  {
    const hdn_src& s = c.src[0];
    int64_t off = hdn_src_off(s, n, vd, vh, vw);
    v = hdn_prologue(s, __ldg(s.t.p + off + ci), ci);
  }
  if (c.nsrc == 2) {
    const hdn_src& s = c.src[1];
    int64_t off = hdn_src_off(s, n, vd, vh, vw);
    v += hdn_prologue(s, __ldg(s.t.p + off + ci), ci);
  }
  return v;
}

// ------------------------------------------------------------------------------ fprop
__global__ void __launch_bounds__(NT) conv_fprop_simt(const hdn_conv c, const int64_t M) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_sum[BN], s_sq[BN];
  const int tid = threadIdx.x;
  const int arow = tid >> 2, akc = (tid & 3) * 4;
  const int brow = tid >> 4, bcol = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Dv = c.src[0].D * c.src[0].ud, Hv = c.src[0].H * c.src[0].uh, Wv = c.src[0].W * c.src[0].uw;

  int an = 0, ad = 0, ah = 0, aw = 0;
  const bool arow_ok = (m0 + arow) < M;
  if (arow_ok) hdn_decode(m0 + arow, c.D, c.H, c.W, an, ad, ah, aw);
  const int vd0 = ad * c.sd - c.pd, vh0 = ah * c.sh - c.ph, vw0 = aw * c.sw - c.pw;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = c.kd * c.kh * c.kw;
  for (int tap = 0; tap < ntaps; ++tap) {
    const int tw = tap % c.kw, th = (tap / c.kw) % c.kh, td = tap / (c.kw * c.kh);
    const int vd = vd0 + td, vh = vh0 + th, vw = vw0 + tw;
    const bool inb = arow_ok && vd >= 0 && vd < Dv && vh >= 0 && vh < Hv && vw >= 0 && vw < Wv;
    for (int c0 = 0; c0 < c.Cin; c0 += BK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int ci = c0 + akc + j;
        float v = 0.f;
        if (inb && ci < c.Cin) v = load_a(c, an, vd, vh, vw, ci);
        As[akc + j][arow] = v;
      }
      {
        int ci = c0 + brow;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int co = n0 + bcol + j;
          float v = 0.f;
          if (ci < c.Cin && co < c.Cout) v = __ldg(c.w + ((int64_t)tap * c.Cin + ci) * c.Cout + co);
          Bs[brow][bcol + j] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const bool do_stats = c.stat_sum != nullptr;
  if (do_stats) {
    if (tid < BN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();
  }
  float csum[4] = {0, 0, 0, 0}, csq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= c.Cout) continue;
      float v = acc[i][j] + (c.bias ? __ldg(c.bias + co) : 0.f);
      if (c.drop_keep < 1.0f) v *= hdn_drop_scale(c.drop_seed, (uint64_t)m * c.Cout + co, c.drop_keep);
      ((float*)c.y.p)[m * c.y.ldc + c.y.coff + co] = v;
      csum[j] += v; csq[j] += v * v;
    }
  }
  if (do_stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_sum[tx * 4 + j], csum[j]);
      atomicAdd(&s_sq[tx * 4 + j], csq[j]);
    }
    __syncthreads();
    if (tid < BN && n0 + tid < c.Cout) {
      atomicAdd(c.stat_sum + n0 + tid, (double)s_sum[tid]);
      atomicAdd(c.stat_sq + n0 + tid, (double)s_sq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------ dgrad
// One launch per source.  Rows = stored source positions, cols = Cin, K = (upsample sub-pos, tap, Cout).
__global__ void __launch_bounds__(NT) conv_dgrad_simt(const hdn_conv c, const hdn_dgrad_epi e,
                                                       const int si, const int64_t Msrc) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_s1[BN], s_s2[BN];
  const hdn_src& s = c.src[si];
  const int tid = threadIdx.x;
  const int arow = tid >> 2, akc = (tid & 3) * 4;
  const int brow = tid >> 4, bcol = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  int an = 0, ad = 0, ah = 0, aw = 0;
  const bool arow_ok = (m0 + arow) < Msrc;
  if (arow_ok) hdn_decode(m0 + arow, s.D, s.H, s.W, an, ad, ah, aw);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = c.kd * c.kh * c.kw;
  const int nsub = s.ud * s.uh * s.uw;
  for (int sub = 0; sub < nsub; ++sub) {
    const int uw_ = sub % s.uw, uh_ = (sub / s.uw) % s.uh, ud_ = sub / (s.uw * s.uh);
    const int qd = ad * s.ud + ud_ + c.pd, qh = ah * s.uh + uh_ + c.ph, qw = aw * s.uw + uw_ + c.pw;
    for (int tap = 0; tap < ntaps; ++tap) {
      const int tw = tap % c.kw, th = (tap / c.kw) % c.kh, td = tap / (c.kw * c.kh);
      const int nd = qd - td, nh = qh - th, nw = qw - tw;
      bool ok = arow_ok && nd >= 0 && nh >= 0 && nw >= 0 && (nd % c.sd == 0) && (nh % c.sh == 0) &&
                (nw % c.sw == 0);
      const int od = nd / c.sd, oh = nh / c.sh, ow = nw / c.sw;
      ok = ok && od < c.D && oh < c.H && ow < c.W;
      const int64_t yoff = ((((int64_t)an * c.D + od) * c.H + oh) * c.W + ow) * c.y.ldc + c.y.coff;
      for (int c0 = 0; c0 < c.Cout; c0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int co = c0 + akc + j;
          As[akc + j][arow] = (ok && co < c.Cout) ? __ldg(c.y.p + yoff + co) : 0.f;
        }
        {
          int co = c0 + brow;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int ci = n0 + bcol + j;
            float v = 0.f;
            if (co < c.Cout && ci < c.Cin) v = __ldg(c.w + ((int64_t)tap * c.Cin + ci) * c.Cout + co);
            Bs[brow][bcol + j] = v;
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  const bool do_s = e.s1 != nullptr;
  if (do_s) {
    if (tid < BN) { s_s1[tid] = 0.f; s_s2[tid] = 0.f; }
    __syncthreads();
  }
  float p1[4] = {0, 0, 0, 0}, p2[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= Msrc) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = n0 + tx * 4 + j;
      if (ci >= c.Cin) continue;
      float x = __ldg(s.t.p + m * s.t.ldc + s.t.coff + ci);
      float a = s.pa ? __ldg(s.pa + ci) : 1.f;
      float b = s.pb ? __ldg(s.pb + ci) : 0.f;
      float du = acc[i][j];
      if (s.relu && !(fmaf(a, x, b) > 0.f)) du = 0.f;
      p1[j] += du; p2[j] += du * (x - (e.center ? __ldg(e.center + ci) : 0.f));
      if (e.mode == 0) {
        float* q = (float*)e.dx.p + m * e.dx.ldc + e.dx.coff + ci;
        float g = a * du;
        *q = e.accumulate ? (*q + g) : g;
      } else {
        float* q = e.du + m * c.Cin + ci;
        *q = e.accumulate ? (*q + du) : du;
      }
    }
  }
  if (do_s) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_s1[tx * 4 + j], p1[j]);
      atomicAdd(&s_s2[tx * 4 + j], p2[j]);
    }
    __syncthreads();
    if (tid < BN && n0 + tid < c.Cin) {
      atomicAdd(e.s1 + n0 + tid, (double)s_s1[tid]);
      atomicAdd(e.s2 + n0 + tid, (double)s_s2[tid]);
    }
  }
}

// ------------------------------------------------------------------------------ wgrad
// grid: x = (ci tile, co tile), y = split over output positions, z = tap.
__global__ void __launch_bounds__(NT) conv_wgrad_simt(const hdn_conv c, float* __restrict__ dw,
                                                       const int64_t M, const int64_t rows_per_split) {
  __shared__ float As[BK][BM + 4];   // [pos][ci]
  __shared__ float Bs[BK][BN + 4];   // [pos][co]
  const int tid = threadIdx.x;
  const int lrow = tid >> 4, lcol = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const int ntn = (c.Cout + BN - 1) / BN;
  const int ci0 = (blockIdx.x / ntn) * BM, co0 = (blockIdx.x % ntn) * BN;
  const int tap = blockIdx.z;
  const int tw = tap % c.kw, th = (tap / c.kw) % c.kh, td = tap / (c.kw * c.kh);
  const int Dv = c.src[0].D * c.src[0].ud, Hv = c.src[0].H * c.src[0].uh, Wv = c.src[0].W * c.src[0].uw;
  const int64_t mb = (int64_t)blockIdx.y * rows_per_split;
  const int64_t me = min(M, mb + rows_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = mb; k0 < me; k0 += BK) {
    const int64_t m = k0 + lrow;
    const bool mok = m < me;
    int n = 0, od = 0, oh = 0, ow = 0;
    if (mok) hdn_decode(m, c.D, c.H, c.W, n, od, oh, ow);
    const int vd = od * c.sd - c.pd + td, vh = oh * c.sh - c.ph + th, vw = ow * c.sw - c.pw + tw;
    const bool inb = mok && vd >= 0 && vd < Dv && vh >= 0 && vh < Hv && vw >= 0 && vw < Wv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + lcol + j;
      As[lrow][lcol + j] = (inb && ci < c.Cin) ? load_a(c, n, vd, vh, vw, ci) : 0.f;
      int co = co0 + lcol + j;
      Bs[lrow][lcol + j] = (mok && co < c.Cout) ? __ldg(c.y.p + m * c.y.ldc + c.y.coff + co) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ci = ci0 + ty * 4 + i;
    if (ci >= c.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tx * 4 + j;
      if (co >= c.Cout) continue;
      atomicAdd(dw + ((int64_t)tap * c.Cin + ci) * c.Cout + co, acc[i][j]);
    }
  }
}

// wgrad of a 1x1x1 convolution with very few output channels (the 3-class classifiers, hybridnet.py:260,419):
// HBM-bound column reduction dw[ci][co] += sum_pos A[pos][ci] * dY[pos][co]; thread = (input channel, position lane).
__global__ void __launch_bounds__(256) conv_wgrad_small_n(const hdn_conv c, float* __restrict__ dw, const int64_t M,
                                                          const int lanes) {
  __shared__ float red[256 * 4];
  const hdn_src& s = c.src[0];
  const int ci = threadIdx.x % c.Cin, pl = threadIdx.x / c.Cin;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    const float a = s.pa ? __ldg(s.pa + ci) : 1.f, b = s.pb ? __ldg(s.pb + ci) : 0.f;
    for (int64_t m = (int64_t)blockIdx.x * lanes + pl; m < M; m += (int64_t)gridDim.x * lanes) {
      float x = fmaf(a, __ldg(s.t.p + m * s.t.ldc + s.t.coff + ci), b);
      if (s.relu) x = fmaxf(x, 0.f);
      const float* dy = c.y.p + m * c.y.ldc + c.y.coff;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < c.Cout) acc[j] = fmaf(x, __ldg(dy + j), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = acc[j];
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += red[(l * c.Cin + ci) * 4 + j];
    for (int j = 0; j < c.Cout; ++j) atomicAdd(dw + (int64_t)ci * c.Cout + j, acc[j]);
  }
}

// Classifier-shaped 1x1x1 convolutions (Cout <= 4: dense167classifer / 2d3dclassifer, hybridnet.py:260,419), vector forms.
// Both passes are HBM-bound column work over M positions x Cin channels; thread = (position lane, channel quad), so a
// position's Cin channels are read as contiguous 16-byte pieces and every thread keeps UNR positions in flight.
// wgrad: dw[ci][co] += sum_pos max(a*x+b,0)[pos][ci] * dY[pos][co]
template <int UNR>
__global__ void __launch_bounds__(256) conv_wgrad_small_n4(const hdn_conv c, float* __restrict__ dw, const int64_t M, const int nq) {
  __shared__ float red[256 * 16];
  const hdn_src& s = c.src[0];
  const int q = threadIdx.x % nq, pl = threadIdx.x / nq, lanes = 256 / nq;
  const int ci = q * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (pl < lanes) {
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s.pa) a = __ldg(reinterpret_cast<const float4*>(s.pa + ci));
    if (s.pb) b = __ldg(reinterpret_cast<const float4*>(s.pb + ci));
    const int64_t stride = (int64_t)gridDim.x * lanes;
    for (int64_t m0 = (int64_t)blockIdx.x * lanes + pl; m0 < M; m0 += stride * UNR) {
      float4 x[UNR];
      float dy[UNR][4];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t m = m0 + u * stride;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) dy[u][j] = 0.f;
        if (m < M) {
          x[u] = __ldg(reinterpret_cast<const float4*>(s.t.p + m * s.t.ldc + s.t.coff + ci));
          const float* d = c.y.p + m * c.y.ldc + c.y.coff;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < c.Cout) dy[u][j] = __ldg(d + j);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (m0 + u * stride >= M) break;
        float v[4] = {fmaf(a.x, x[u].x, b.x), fmaf(a.y, x[u].y, b.y), fmaf(a.z, x[u].z, b.z), fmaf(a.w, x[u].w, b.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (s.relu) v[i] = fmaxf(v[i], 0.f);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(v[i], dy[u][j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x * 16 + i * 4 + j] = acc[i][j];
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += red[(l * nq + q) * 16 + i * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < c.Cout; ++j) atomicAdd(dw + (int64_t)(ci + i) * c.Cout + j, acc[i][j]);
  }
}

// dgrad: dz[pos][ci] = sum_co dY[pos][co] * w[ci][co], then the epilogue of hdn_dgrad_epi (ReLU mask from the stored value,
// S1 / S2 sums, dx (+)= a*du or du (+)= du).  The quad's 4 x Cout weights and its (a, b, center) stay in registers.
template <int UNR>
__global__ void __launch_bounds__(256) conv_dgrad_small_n4(const hdn_conv c, const hdn_dgrad_epi e, const int64_t M, const int nq) {
  __shared__ float red[256 * 8];
  const hdn_src& s = c.src[0];
  const int q = threadIdx.x % nq, pl = threadIdx.x / nq, lanes = 256 / nq;
  const int ci = q * 4;
  float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f), ctr = b;
    if (s.pa) a = __ldg(reinterpret_cast<const float4*>(s.pa + ci));
    if (s.pb) b = __ldg(reinterpret_cast<const float4*>(s.pb + ci));
    if (e.s1 && e.center) ctr = __ldg(reinterpret_cast<const float4*>(e.center + ci));
    float w[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[i][j] = j < c.Cout ? __ldg(c.w + (int64_t)(ci + i) * c.Cout + j) : 0.f;
    const int64_t stride = (int64_t)gridDim.x * lanes;
    for (int64_t m0 = (int64_t)blockIdx.x * lanes + pl; m0 < M; m0 += stride * UNR) {
      float4 x[UNR], old[UNR];
      float dy[UNR][4];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t m = m0 + u * stride;
        x[u] = old[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) dy[u][j] = 0.f;
        if (m < M) {
          x[u] = __ldg(reinterpret_cast<const float4*>(s.t.p + m * s.t.ldc + s.t.coff + ci));
          if (e.accumulate)
            old[u] = e.mode == 0 ? *reinterpret_cast<const float4*>(e.dx.p + m * e.dx.ldc + e.dx.coff + ci)
                                 : *reinterpret_cast<const float4*>(e.du + m * c.Cin + ci);
          const float* d = c.y.p + m * c.y.ldc + c.y.coff;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < c.Cout) dy[u][j] = __ldg(d + j);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t m = m0 + u * stride;
        if (m >= M) break;
        const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
        const float as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w}, cs[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
        float g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float du = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) du = fmaf(dy[u][j], w[i][j], du);
          if (s.relu && !(fmaf(as[i], xs[i], bs[i]) > 0.f)) du = 0.f;
          p1[i] += du; p2[i] += du * (xs[i] - cs[i]);
          g[i] = e.mode == 0 ? as[i] * du : du;
        }
        float4 o = make_float4(g[0] + old[u].x, g[1] + old[u].y, g[2] + old[u].z, g[3] + old[u].w);
        if (e.mode == 0) *reinterpret_cast<float4*>(const_cast<float*>(e.dx.p) + m * e.dx.ldc + e.dx.coff + ci) = o;
        else *reinterpret_cast<float4*>(e.du + m * c.Cin + ci) = o;
      }
    }
  }
  if (e.s1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[threadIdx.x * 8 + i] = p1[i]; red[threadIdx.x * 8 + 4 + i] = p2[i]; }
    __syncthreads();
    if (pl == 0) {
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int i = 0; i < 4; ++i) { p1[i] += red[(l * nq + q) * 8 + i]; p2[i] += red[(l * nq + q) * 8 + 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) { atomicAdd(e.s1 + ci + i, (double)p1[i]); atomicAdd(e.s2 + ci + i, (double)p2[i]); }
    }
  }
}

static bool small_n_shape(const hdn_conv* c) {
  const hdn_src& s = c->src[0];
  return c->kd * c->kh * c->kw == 1 && c->Cout <= 4 && c->Cin % 4 == 0 && c->Cin <= 256 && 256 % (c->Cin / 4) == 0 && c->nsrc == 1 &&
         s.ud == 1 && s.uh == 1 && s.uw == 1 && c->sd == 1 && c->sh == 1 && c->sw == 1 && s.t.ldc % 4 == 0 && s.t.coff % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(s.t.p) & 15) == 0;
}

// out[c] += sum_m y[m][c]   (bias gradient)
__global__ void __launch_bounds__(256) colsum_kernel(hdn_tensor y, int64_t M, int C, float* out,
                                                     int64_t rows_per_block) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int cidx = blockIdx.y * 32 + lane;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_block, me = min(M, mb + rows_per_block);
  float s = 0.f;
  if (cidx < C)
    for (int64_t m = mb + wy; m < me; m += 8) s += __ldg(y.p + m * y.ldc + y.coff + cidx);
  red[wy][lane] = s;
  __syncthreads();
  if (wy == 0 && cidx < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    atomicAdd(out + cidx, t);
  }
}

__global__ void __launch_bounds__(256) colstats_kernel(hdn_tensor y, int64_t M, int C, double* sum, double* sq,
                                                       int64_t rows_per_block) {
  __shared__ float r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int cidx = blockIdx.y * 32 + lane;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_block, me = min(M, mb + rows_per_block);
  float s = 0.f, q = 0.f;
  if (cidx < C)
    for (int64_t m = mb + wy; m < me; m += 8) { float v = __ldg(y.p + m * y.ldc + y.coff + cidx); s += v; q += v * v; }
  r1[wy][lane] = s; r2[wy][lane] = q;
  __syncthreads();
  if (wy == 0 && cidx < C) {
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t += r1[i][lane]; u += r2[i][lane]; }
    atomicAdd(sum + cidx, (double)t);
    atomicAdd(sq + cidx, (double)u);
  }
}

}  // namespace

extern "C" int hdn_col_stats(hdn_tensor y, int64_t M, int C, double* sum, double* sq, void* stream) {
  HDN_CHECK_ARG(y.p && sum && sq && M > 0 && C > 0, "col_stats: bad arguments");
  int64_t rpb = 1024;
  dim3 grid((unsigned)hdn_cdiv(M, rpb), (unsigned)hdn_cdiv(C, 32));
  HDN_LAUNCHED(1), colstats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, M, C, sum, sq, rpb);
  HDN_CHECK_LAUNCH("col_stats");
  return HDN_OK;
}

int hdn_validate_conv(const hdn_conv* c) {
  HDN_CHECK_ARG(c != nullptr, "conv: null descriptor");
  HDN_CHECK_ARG(c->N > 0 && c->D > 0 && c->H > 0 && c->W > 0 && c->Cin > 0 && c->Cout > 0,
                "conv: non-positive dims N=%d D=%d H=%d W=%d Cin=%d Cout=%d", c->N, c->D, c->H, c->W,
                c->Cin, c->Cout);
  HDN_CHECK_ARG(c->kd > 0 && c->kh > 0 && c->kw > 0 && c->sd > 0 && c->sh > 0 && c->sw > 0,
                "conv: bad kernel/stride");
  HDN_CHECK_ARG(c->nsrc == 1 || c->nsrc == 2, "conv: nsrc must be 1 or 2 (got %d)", c->nsrc);
  HDN_CHECK_ARG(c->w != nullptr && c->y.p != nullptr, "conv: null weight/output pointer");
  const int Dv = c->src[0].D * c->src[0].ud, Hv = c->src[0].H * c->src[0].uh, Wv = c->src[0].W * c->src[0].uw;
  for (int i = 0; i < c->nsrc; ++i) {
    const hdn_src& s = c->src[i];
    HDN_CHECK_ARG(s.t.p != nullptr, "conv: src[%d] null", i);
    HDN_CHECK_ARG((s.ud == 1 || s.ud == 2) && (s.uh == 1 || s.uh == 2) && (s.uw == 1 || s.uw == 2),
                  "conv: upsample factors must be 1 or 2");
    HDN_CHECK_ARG(s.D * s.ud == Dv && s.H * s.uh == Hv && s.W * s.uw == Wv,
                  "conv: sources disagree on the virtual input size");
    HDN_CHECK_ARG(s.t.ldc >= s.t.coff + c->Cin, "conv: src[%d] channel window exceeds ldc", i);
  }
  // output grid must match geometry
  HDN_CHECK_ARG((Dv + 2 * c->pd - c->kd) / c->sd + 1 >= c->D && (Hv + 2 * c->ph - c->kh) / c->sh + 1 >= c->H &&
                    (Wv + 2 * c->pw - c->kw) / c->sw + 1 >= c->W,
                "conv: output grid larger than the geometry allows");
  HDN_CHECK_ARG(c->y.ldc >= c->y.coff + c->Cout, "conv: y channel window exceeds ldc");
  HDN_CHECK_ARG(c->drop_keep > 0.f && c->drop_keep <= 1.f, "conv: drop_keep must be in (0,1]");
  HDN_CHECK_ARG(c->precision >= 0 && c->precision <= 2, "conv: precision must be 0 (fp32), 1 (bf16) or 2 (bf16x3), got %d", c->precision);
  return HDN_OK;
}

int hdn_conv_fprop_simt(const hdn_conv* c, cudaStream_t st) {
  const int64_t M = (int64_t)c->N * c->D * c->H * c->W;
  dim3 grid((unsigned)hdn_cdiv(M, BM), (unsigned)hdn_cdiv(c->Cout, BN));
  HDN_LAUNCHED(1), conv_fprop_simt<<<grid, NT, 0, st>>>(*c, M);
  HDN_CHECK_LAUNCH("conv_fprop_simt");
  return HDN_OK;
}

int hdn_conv_dgrad_simt(const hdn_conv* c, const hdn_dgrad_epi* epi, cudaStream_t st) {
  if (small_n_shape(c) && epi[0].mode != 2 &&
      (epi[0].mode == 1 || (epi[0].dx.ldc % 4 == 0 && epi[0].dx.coff % 4 == 0 && (reinterpret_cast<uintptr_t>(epi[0].dx.p) & 15) == 0))) {
    const int64_t M = (int64_t)c->N * c->D * c->H * c->W;
    const int nq = c->Cin / 4, lanes = 256 / nq;
    int64_t blocks = hdn_cdiv(M, (int64_t)lanes * 4 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    HDN_LAUNCHED(1), conv_dgrad_small_n4<4><<<(unsigned)blocks, 256, 0, st>>>(*c, epi[0], M, nq);
    HDN_CHECK_LAUNCH("conv_dgrad_small_n4");
    return HDN_OK;
  }
  for (int si = 0; si < c->nsrc; ++si) {
    if (epi[si].mode == 2) continue;
    const hdn_src& s = c->src[si];
    const int64_t Ms = (int64_t)c->N * s.D * s.H * s.W;
    dim3 grid((unsigned)hdn_cdiv(Ms, BM), (unsigned)hdn_cdiv(c->Cin, BN));
    HDN_LAUNCHED(1), conv_dgrad_simt<<<grid, NT, 0, st>>>(*c, epi[si], si, Ms);
    HDN_CHECK_LAUNCH("conv_dgrad_simt");
  }
  return HDN_OK;
}

int hdn_colsum(hdn_tensor y, int64_t M, int C, float* out, cudaStream_t st) {
  int64_t rpb = 4096;
  dim3 grid((unsigned)hdn_cdiv(M, rpb), (unsigned)hdn_cdiv(C, 32));
  HDN_LAUNCHED(1), colsum_kernel<<<grid, 256, 0, st>>>(y, M, C, out, rpb);
  HDN_CHECK_LAUNCH("colsum");
  return HDN_OK;
}

int hdn_conv_wgrad_simt(const hdn_conv* c, float* dw, cudaStream_t st) {
  const int64_t M = (int64_t)c->N * c->D * c->H * c->W;
  if (small_n_shape(c)) {
    const int nq = c->Cin / 4, lanes = 256 / nq;
    int64_t blocks = hdn_cdiv(M, (int64_t)lanes * 4 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    HDN_LAUNCHED(1), conv_wgrad_small_n4<4><<<(unsigned)blocks, 256, 0, st>>>(*c, dw, M, nq);
    HDN_CHECK_LAUNCH("conv_wgrad_small_n4");
    return HDN_OK;
  }
  if (c->kd * c->kh * c->kw == 1 && c->Cout <= 4 && c->Cin <= 256 && c->nsrc == 1 && c->src[0].ud == 1 && c->src[0].uh == 1 &&
      c->src[0].uw == 1 && c->sd == 1 && c->sh == 1 && c->sw == 1) {
    const int lanes = 256 / c->Cin;
    int64_t blocks = hdn_cdiv(M, (int64_t)lanes * 64);
    if (blocks > 148 * 8) blocks = 148 * 8;
    HDN_LAUNCHED(1), conv_wgrad_small_n<<<(unsigned)blocks, 256, 0, st>>>(*c, dw, M, lanes);
    HDN_CHECK_LAUNCH("conv_wgrad_small_n");
    return HDN_OK;
  }
  const int taps = c->kd * c->kh * c->kw;
  const int64_t tiles = hdn_cdiv(c->Cin, BM) * hdn_cdiv(c->Cout, BN);
  // aim for ~8 waves of 148 SMs, at least 256 rows per split
  int64_t splits = hdn_cdiv(148 * 8, tiles * taps);
  splits = splits < 1 ? 1 : splits;
  int64_t rps = hdn_cdiv(M, splits);
  if (rps < 256) rps = 256;
  rps = hdn_cdiv(rps, BK) * BK;
  splits = hdn_cdiv(M, rps);
  if (splits > 65535) { rps = hdn_cdiv(hdn_cdiv(M, 65535), BK) * BK; splits = hdn_cdiv(M, rps); }
  dim3 grid((unsigned)tiles, (unsigned)splits, (unsigned)taps);
  HDN_LAUNCHED(1), conv_wgrad_simt<<<grid, NT, 0, st>>>(*c, dw, M, rps);
  HDN_CHECK_LAUNCH("conv_wgrad_simt");
  return HDN_OK;
}
