// Training-sample pipeline on the GPU: what `load_seq_crop_data_masktumor_try` does on 14 host
// threads per batch (train_hybrid.py:40-98, train_2ddense.py:40-69) -- crop around a liver / tumour voxel at a random
// scale, subtract the mean, flip / rotate, skimage `resize` to the network size (labels: order 0, mode 'edge'; image:
// order 3, mode 'constant', cval 0, clip, preserve_range) -- as one gather kernel over a volume that stays resident in
// HBM (the 131 training volumes are ~35 GB as fp32), writing straight into the engine's (S, H, W) input layout.
// The random draws stay on the host (augment.py) in the reference's np.random call order.
//
// resize: a 3-D array whose third extent is unchanged goes through `warp` per slice with src = scale*(dst+0.5)-0.5
// (skimage/transform/_warps.py); order 0 = round + edge clamp; order 3 = separable Catmull-Rom on the 4x4 taps around
// floor(src), taps outside the crop = cval; the result is clipped to the crop's value range (_clip_warp_output), exact
// cval samples surviving when cval is outside that range.  Arithmetic in double like `_warp_fast`.
// HBM bound: 16 taps per output, served by L1/L2 (neighbouring outputs share 12 of them).
#include "hdn_common.cuh"
#include <limits.h>

namespace {

constexpr int AT = 256;

__device__ __forceinline__ int fkey(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ float crop_value(const hdn_aug& a, int i, int j, int s) {   // un-flipped crop coordinates
  const int64_t off = ((int64_t)(a.c0 + s) * a.VH + (a.a0 + i)) * a.VW + (a.b0 + j);
  const float v = a.vol_i16 ? (float)reinterpret_cast<const short*>(a.vol)[off] : reinterpret_cast<const float*>(a.vol)[off];
  return v - a.mean;                                     // float32 subtraction, as `cropp_img -= args.mean` on a float32 array
}

__global__ void aug_init_kernel(int* mm) {
  mm[0] = INT_MAX;
  mm[1] = INT_MIN;
}

__global__ void __launch_bounds__(AT) aug_minmax_kernel(hdn_aug a, int* __restrict__ mm) {
  const int64_t n = (int64_t)a.ch * a.cw * a.cs;
  int lo = INT_MAX, hi = INT_MIN;
  for (int64_t t = blockIdx.x * (int64_t)AT + threadIdx.x; t < n; t += (int64_t)gridDim.x * AT) {
    const int j = (int)(t % a.cw), i = (int)((t / a.cw) % a.ch), s = (int)(t / ((int64_t)a.cw * a.ch));
    const int k = fkey(crop_value(a, i, j, s));
    lo = min(lo, k);
    hi = max(hi, k);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}

__device__ __forceinline__ double cubic(double x, double f0, double f1, double f2, double f3) {
  return f1 + 0.5 * x * (f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3 + x * (3.0 * (f1 - f2) + f3 - f0)));
}

__global__ void __launch_bounds__(AT) aug_sample_kernel(hdn_aug a, float* __restrict__ X, float* __restrict__ Y, int* __restrict__ counts,
                                                        const int* __restrict__ mm) {
  const int fh = (a.m00 != 0) ? a.ch : a.cw, fw = (a.m00 != 0) ? a.cw : a.ch;      // extent of the flipped crop
  const double rs = (double)fh / a.out_h, cs = (double)fw / a.out_w;
  const double lo = (double)fkey_inv(mm[0]), hi = (double)fkey_inv(mm[1]);
  const bool keep_cval = !(lo <= 0.0 && 0.0 <= hi);
  const int64_t n = (int64_t)a.cs * a.out_h * a.out_w;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int64_t t = blockIdx.x * (int64_t)AT + threadIdx.x; t < n; t += (int64_t)gridDim.x * AT) {
    const int x = (int)(t % a.out_w), y = (int)((t / a.out_w) % a.out_h), s = (int)(t / ((int64_t)a.out_w * a.out_h));
    const double r = rs * (y + 0.5) - 0.5, c = cs * (x + 0.5) - 0.5;
    if (Y != nullptr && s >= a.ys0 && s < a.ys0 + a.yns) {
      int ri = (int)floor(r + 0.5), ci = (int)floor(c + 0.5);
      ri = min(max(ri, 0), fh - 1);
      ci = min(max(ci, 0), fw - 1);
      const int i = a.m00 * ri + a.m01 * ci + a.o0, j = a.m10 * ri + a.m11 * ci + a.o1;
      const int64_t off = ((int64_t)(a.c0 + s) * a.VH + (a.a0 + i)) * a.VW + (a.b0 + j);
      const int lab = a.seg[off];
      Y[((int64_t)(s - a.ys0) * a.out_h + y) * a.out_w + x] = (float)lab;
      c0 += lab == 0;
      c1 += lab == 1;
      c2 += lab == 2;
    }
    const int r0 = (int)floor(r), q0 = (int)floor(c);
    const double xr = r - r0, xc = c - q0;
    double fr[4];
#pragma unroll
    for (int dr = 0; dr < 4; ++dr) {
      const int ri = r0 + dr - 1;
      double f[4];
#pragma unroll
      for (int dc = 0; dc < 4; ++dc) {
        const int ci = q0 + dc - 1;
        double v = 0.0;
        if (ri >= 0 && ri < fh && ci >= 0 && ci < fw)
          v = (double)crop_value(a, a.m00 * ri + a.m01 * ci + a.o0, a.m10 * ri + a.m11 * ci + a.o1, s);
        f[dc] = v;
      }
      fr[dr] = cubic(xc, f[0], f[1], f[2], f[3]);
    }
    double o = cubic(xr, fr[0], fr[1], fr[2], fr[3]);
    if (!(keep_cval && o == 0.0)) o = fmin(fmax(o, lo), hi);
    X[s * a.xs_s + y * a.xs_h + x * a.xs_w] = (float)o;
  }
  if (counts != nullptr) {
    for (int o = 16; o; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (c0) atomicAdd(counts, c0);
      if (c1) atomicAdd(counts + 1, c1);
      if (c2) atomicAdd(counts + 2, c2);
    }
  }
}

inline unsigned agrid(int64_t n) {
  int64_t b = (n + AT - 1) / AT;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int hdn_aug_sample(const hdn_aug* a, float* x, float* y, int32_t* counts, void* scratch, void* stream) {
  HDN_CHECK_ARG(a && a->vol && x && scratch, "aug_sample: null pointer");
  HDN_CHECK_ARG(y == nullptr || a->seg != nullptr, "aug_sample: label output without a label volume");
  HDN_CHECK_ARG(a->ch > 0 && a->cw > 0 && a->cs > 0 && a->out_h > 0 && a->out_w > 0, "aug_sample: empty crop or output");
  HDN_CHECK_ARG(a->a0 >= 0 && a->b0 >= 0 && a->c0 >= 0 && a->a0 + a->ch <= a->VH && a->b0 + a->cw <= a->VW && a->c0 + a->cs <= a->VS,
                "aug_sample: crop [%d,%d)x[%d,%d)x[%d,%d) outside the volume %dx%dx%d", a->a0, a->a0 + a->ch, a->b0, a->b0 + a->cw, a->c0,
                a->c0 + a->cs, a->VH, a->VW, a->VS);
  const bool straight = a->m00 != 0 && a->m11 != 0 && a->m01 == 0 && a->m10 == 0;
  const bool rotated = a->m01 != 0 && a->m10 != 0 && a->m00 == 0 && a->m11 == 0;
  HDN_CHECK_ARG(straight || rotated, "aug_sample: the flip matrix must be a signed permutation");
  HDN_CHECK_ARG(abs(a->m00) <= 1 && abs(a->m01) <= 1 && abs(a->m10) <= 1 && abs(a->m11) <= 1, "aug_sample: the flip matrix must be a signed permutation");
  HDN_CHECK_ARG(y == nullptr || (a->ys0 >= 0 && a->yns > 0 && a->ys0 + a->yns <= a->cs), "aug_sample: label slices outside the crop");
  cudaStream_t st = (cudaStream_t)stream;
  int* mm = reinterpret_cast<int*>(scratch);
  HDN_LAUNCHED(1), aug_init_kernel<<<1, 1, 0, st>>>(mm);
  HDN_LAUNCHED(1), aug_minmax_kernel<<<agrid((int64_t)a->ch * a->cw * a->cs), AT, 0, st>>>(*a, mm);
  HDN_LAUNCHED(1), aug_sample_kernel<<<agrid((int64_t)a->cs * a->out_h * a->out_w), AT, 0, st>>>(*a, x, y, counts, mm);
  HDN_CHECK_LAUNCH("aug_sample");
  return HDN_OK;
}
