// Shared device/host helpers for libhdn (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hdn.h"

void hdn_set_error(const char* fmt, ...);
// kernels launched by this thread since the library was loaded (hdn_launch_count(); bench.py reports the count of the timed region)
extern thread_local long long hdn_tl_launches;
#define HDN_LAUNCHED(n) (hdn_tl_launches += (n))

#define HDN_CHECK_ARG(cond, ...)                          \
  do {                                                    \
    if (!(cond)) {                                        \
      hdn_set_error(__VA_ARGS__);                         \
      return HDN_ERR_ARG;                                 \
    }                                                     \
  } while (0)

#define HDN_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      hdn_set_error("%s: %s", name, cudaGetErrorString(e__));                    \
      return HDN_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

static inline int64_t hdn_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- dropout mask: stateless hash of (seed, dense element index) -> scale (0 or 1/keep) ----
__host__ __device__ __forceinline__ uint32_t hdn_hash64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  return (uint32_t)(x >> 32);
}
__host__ __device__ __forceinline__ float hdn_drop_scale(uint64_t seed, uint64_t idx, float keep) {
  float u = (float)(hdn_hash64(seed * 0x100000001B3ull + idx) >> 8) * (1.0f / 16777216.0f);
  return (u < keep) ? (1.0f / keep) : 0.0f;
}

// ---- A-operand source access ----------------------------------------------------------------
// Virtual (up-sampled) coordinates -> element offset of channel 0 of the window.
__device__ __forceinline__ int64_t hdn_src_off(const hdn_src& s, int n, int vd, int vh, int vw) {
  int d = (s.ud == 2) ? (vd >> 1) : vd;
  int h = (s.uh == 2) ? (vh >> 1) : vh;
  int w = (s.uw == 2) ? (vw >> 1) : vw;
  return ((((int64_t)n * s.D + d) * s.H + h) * s.W + w) * (int64_t)s.t.ldc + s.t.coff;
}
// prologue: BatchNorm->Scale->ReLU folded to max(a*x+b, 0)
__device__ __forceinline__ float hdn_prologue(const hdn_src& s, float x, int c) {
  float a = s.pa ? __ldg(s.pa + c) : 1.0f;
  float b = s.pb ? __ldg(s.pb + c) : 0.0f;
  float u = fmaf(a, x, b);
  return s.relu ? fmaxf(u, 0.0f) : u;
}

// decode a flat output position m -> (n, d, h, w)
__device__ __forceinline__ void hdn_decode(int64_t m, int D, int H, int W, int& n, int& d, int& h, int& w) {
  w = (int)(m % W); m /= W;
  h = (int)(m % H); m /= H;
  d = (int)(m % D); n = (int)(m / D);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
