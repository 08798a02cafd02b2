// Post-processing of the sliding-window probability volumes on the GPU: what test.py:71-115 and
// lib/funcs.py:138-153 do with scipy.ndimage / skimage.measure on the host --
//   thresholds (0.5 liver, 0.9 tumour; test.py:73-77), 6-neighbour binary dilation (ndimage.binary_dilation, :62,:95),
//   largest connected component (measure.label + regionprops area, :83-91, :96-103: full 26-connectivity in 3-D),
//   hole filling (ndimage.binary_fill_holes, :104,:109,:112: background components under 6-connectivity that do not
//   reach the volume border become foreground).
// Byte / index work, HBM bound: one uint8 per voxel, one int32 label per voxel of scratch.
//
// Connected components: label-equivalence union-find.  Every foreground voxel starts as its own root (label = linear
// index); each voxel unites itself with its foreground "forward" neighbours (13 of 26, or 3 of 6) using atomicMin on
// the root, so a component's final root is its SMALLEST linear index.  Volumes are C-order (X, Y, Z) as the reference
// holds them, so "smallest linear index" is the raster order in which measure.label numbers components: the largest
// component with ties broken towards the first one -- `box.index(max(box))`, test.py:90 -- is max over (size, -root).
#include "hdn_common.cuh"

namespace {

constexpr int PT = 256;
inline unsigned pgrid(int64_t n) {
  int64_t b = (n + PT - 1) / PT;
  return (unsigned)(b > 148 * 64 ? 148 * 64 : (b < 1 ? 1 : b));
}

__global__ void __launch_bounds__(PT) threshold_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                                       unsigned char* __restrict__ r1, unsigned char* __restrict__ r2, int64_t n,
                                                       float t1, float t2) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    const unsigned char b = s2[i] >= t2 ? 1 : 0;
    r2[i] = b;
    r1[i] = (s1[i] >= t1 || b) ? 1 : 0;                  // result1[result2 == 1] = 1 (test.py:77)
  }
}

__global__ void __launch_bounds__(PT) dilate6_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int X,
                                                     int Y, int Z) {
  const int64_t n = (int64_t)X * Y * Z;
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
    unsigned char v = in[i];
    if (!v) {
      if (z > 0 && in[i - 1]) v = 1;
      else if (z + 1 < Z && in[i + 1]) v = 1;
      else if (y > 0 && in[i - Z]) v = 1;
      else if (y + 1 < Y && in[i + Z]) v = 1;
      else if (x > 0 && in[i - (int64_t)Z * Y]) v = 1;
      else if (x + 1 < X && in[i + (int64_t)Z * Y]) v = 1;
    }
    out[i] = v ? 1 : 0;
  }
}

__device__ __forceinline__ int uf_find(const int* L, int i) {
  const volatile int* V = L;                             // other threads lower labels concurrently (atomicMin): always re-read
  int p = V[i];
  while (p != i) { i = p; p = V[i]; }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }        // a < b: hang the larger root under the smaller one
    const int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;                                             // somebody re-rooted b in between: retry from there
  }
}

// fg(i) = (in[i] != 0) ^ invert
__global__ void __launch_bounds__(PT) ccl_init_kernel(const unsigned char* __restrict__ in, int* __restrict__ L, int64_t n, int invert) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT)
    L[i] = ((in[i] != 0) != (invert != 0)) ? (int)i : -1;
}

__global__ void __launch_bounds__(PT) ccl_merge_kernel(int* __restrict__ L, int X, int Y, int Z, int conn26) {
  const int64_t n = (int64_t)X * Y * Z;
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    if (L[i] < 0) continue;
    const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
    // forward half of the neighbourhood: (dx, dy, dz) lexicographically > (0, 0, 0)
    for (int dx = 0; dx <= 1; ++dx)
      for (int dy = (dx ? -1 : 0); dy <= 1; ++dy)
        for (int dz = ((dx || dy) ? -1 : 1); dz <= 1; ++dz) {
          if (!conn26 && (dx != 0) + (dy != 0) + (dz != 0) != 1) continue;
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx >= X || yy < 0 || yy >= Y || zz < 0 || zz >= Z) continue;
          const int64_t j = ((int64_t)xx * Y + yy) * Z + zz;
          if (L[j] >= 0) uf_union(L, (int)i, (int)j);
        }
  }
}

__global__ void __launch_bounds__(PT) ccl_compress_kernel(int* __restrict__ L, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT)
    if (L[i] >= 0) L[i] = uf_find(L, (int)i);
}

// sizes live in a second int32 plane indexed by root
__global__ void __launch_bounds__(PT) ccl_count_kernel(const int* __restrict__ L, int* __restrict__ size, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT)
    if (L[i] >= 0) atomicAdd(&size[L[i]], 1);
}
__global__ void __launch_bounds__(PT) ccl_argmax_kernel(const int* __restrict__ L, const int* __restrict__ size, int64_t n,
                                                        unsigned long long* best) {
  unsigned long long loc = 0ull;
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT)
    if (L[i] == (int)i) {                                // a root
      const unsigned long long key = ((unsigned long long)(unsigned)size[i] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
      loc = key > loc ? key : loc;
    }
  if (loc) atomicMax(best, loc);
}
__global__ void __launch_bounds__(PT) ccl_select_kernel(const int* __restrict__ L, unsigned char* __restrict__ out, int64_t n,
                                                        const unsigned long long* best) {
  const unsigned long long b = *best;
  const int root = b ? (int)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull)) : -2;
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) out[i] = (L[i] == root) ? 1 : 0;
}

// hole filling: background roots that own a border voxel are "outside"
__global__ void __launch_bounds__(PT) border_mark_kernel(const int* __restrict__ L, int* __restrict__ flag, int X, int Y, int Z) {
  const int64_t n = (int64_t)X * Y * Z;
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    if (L[i] < 0) continue;
    const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
    if (x == 0 || y == 0 || z == 0 || x == X - 1 || y == Y - 1 || z == Z - 1) flag[L[i]] = 1;
  }
}
__global__ void __launch_bounds__(PT) fill_select_kernel(const int* __restrict__ L, const int* __restrict__ flag,
                                                         unsigned char* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT)
    out[i] = (L[i] < 0 || !flag[L[i]]) ? 1 : 0;          // foreground, or background that never reaches the border
}

// liver_res = fill(largest(result1)); liver_res[Segmask == 1] = 2  (test.py:106-114), Segmask already filled
__global__ void __launch_bounds__(PT) compose_kernel(const unsigned char* __restrict__ liver, const unsigned char* __restrict__ tumor,
                                                     unsigned char* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) out[i] = tumor[i] ? 2 : (liver[i] ? 1 : 0);
}
__global__ void __launch_bounds__(PT) and_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                                 unsigned char* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) out[i] = (a[i] && b[i]) ? 1 : 0;
}

int ccl(const unsigned char* in, int* L, int X, int Y, int Z, int conn26, int invert, cudaStream_t st) {
  const int64_t n = (int64_t)X * Y * Z;
  HDN_LAUNCHED(1), ccl_init_kernel<<<pgrid(n), PT, 0, st>>>(in, L, n, invert);
  HDN_LAUNCHED(1), ccl_merge_kernel<<<pgrid(n), PT, 0, st>>>(L, X, Y, Z, conn26);
  HDN_LAUNCHED(1), ccl_compress_kernel<<<pgrid(n), PT, 0, st>>>(L, n);
  HDN_CHECK_LAUNCH("post ccl");
  return HDN_OK;
}

}  // namespace

#define PST ((cudaStream_t)stream)
#define POST_ARGS_OK(X, Y, Z) ((X) > 0 && (Y) > 0 && (Z) > 0 && (int64_t)(X) * (Y) * (Z) < (1ll << 31))

extern "C" int hdn_post_threshold(const float* score_liver, const float* score_tumor, unsigned char* liver, unsigned char* tumor,
                                  int64_t n, float thres_liver, float thres_tumor, void* stream) {
  HDN_CHECK_ARG(score_liver && score_tumor && liver && tumor && n > 0, "post_threshold: bad arguments");
  HDN_LAUNCHED(1), threshold_kernel<<<pgrid(n), PT, 0, PST>>>(score_liver, score_tumor, liver, tumor, n, thres_liver, thres_tumor);
  HDN_CHECK_LAUNCH("post_threshold");
  return HDN_OK;
}

extern "C" int hdn_post_dilate(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* stream) {
  HDN_CHECK_ARG(in && out && in != out && POST_ARGS_OK(X, Y, Z), "post_dilate: bad arguments");
  HDN_LAUNCHED(1), dilate6_kernel<<<pgrid((int64_t)X * Y * Z), PT, 0, PST>>>(in, out, X, Y, Z);
  HDN_CHECK_LAUNCH("post_dilate");
  return HDN_OK;
}

// ws: 2 * X*Y*Z int32 + 8 bytes (labels, sizes, best key)
extern "C" int hdn_post_largest_component(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* ws, int64_t ws_bytes,
                                          void* stream) {
  HDN_CHECK_ARG(in && out && ws && POST_ARGS_OK(X, Y, Z), "post_largest_component: bad arguments");
  const int64_t n = (int64_t)X * Y * Z;
  HDN_CHECK_ARG(ws_bytes >= 8 * n + 8, "post_largest_component: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)(8 * n + 8));
  int* L = reinterpret_cast<int*>(ws);
  int* size = L + n;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(size + n);
  int rc = ccl(in, L, X, Y, Z, 1, 0, PST);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(size, 0, (size_t)n * 4 + 8, PST);
  if (e != cudaSuccess) { hdn_set_error("post_largest_component: memset: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  HDN_LAUNCHED(1), ccl_count_kernel<<<pgrid(n), PT, 0, PST>>>(L, size, n);
  HDN_LAUNCHED(1), ccl_argmax_kernel<<<pgrid(n), PT, 0, PST>>>(L, size, n, best);
  HDN_LAUNCHED(1), ccl_select_kernel<<<pgrid(n), PT, 0, PST>>>(L, out, n, best);
  HDN_CHECK_LAUNCH("post_largest_component");
  return HDN_OK;
}

extern "C" int hdn_post_fill_holes(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* ws, int64_t ws_bytes,
                                   void* stream) {
  HDN_CHECK_ARG(in && out && ws && POST_ARGS_OK(X, Y, Z), "post_fill_holes: bad arguments");
  const int64_t n = (int64_t)X * Y * Z;
  HDN_CHECK_ARG(ws_bytes >= 8 * n + 8, "post_fill_holes: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)(8 * n + 8));
  int* L = reinterpret_cast<int*>(ws);
  int* flag = L + n;
  int rc = ccl(in, L, X, Y, Z, 0, 1, PST);               // background components, 6-connectivity
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(flag, 0, (size_t)n * 4, PST);
  if (e != cudaSuccess) { hdn_set_error("post_fill_holes: memset: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  HDN_LAUNCHED(1), border_mark_kernel<<<pgrid(n), PT, 0, PST>>>(L, flag, X, Y, Z);
  HDN_LAUNCHED(1), fill_select_kernel<<<pgrid(n), PT, 0, PST>>>(L, flag, out, n);
  HDN_CHECK_LAUNCH("post_fill_holes");
  return HDN_OK;
}

extern "C" int hdn_post_and(const unsigned char* a, const unsigned char* b, unsigned char* out, int64_t n, void* stream) {
  HDN_CHECK_ARG(a && b && out && n > 0, "post_and: bad arguments");
  HDN_LAUNCHED(1), and_kernel<<<pgrid(n), PT, 0, PST>>>(a, b, out, n);
  HDN_CHECK_LAUNCH("post_and");
  return HDN_OK;
}

extern "C" int hdn_post_compose(const unsigned char* liver, const unsigned char* tumor, unsigned char* out, int64_t n, void* stream) {
  HDN_CHECK_ARG(liver && tumor && out && n > 0, "post_compose: bad arguments");
  HDN_LAUNCHED(1), compose_kernel<<<pgrid(n), PT, 0, PST>>>(liver, tumor, out, n);
  HDN_CHECK_LAUNCH("post_compose");
  return HDN_OK;
}
