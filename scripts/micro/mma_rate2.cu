// Second issue-rate experiment: which feature of the convolution kernel's issue loop slows tcgen05.mma down from the
// 48 cycles an isolated M=128 x N=64 x K=16 SS MMA takes (mma_rate.cu) to the ~88 measured there?  Features are enabled one
// at a time (bit mask):
//   1  all MMAs accumulate into ONE accumulator (the kernel's K loop) instead of two alternating ones
//   2  tcgen05.commit to an mbarrier after every 6 MMAs (one per filter tap), never waited on
//   4  tcgen05.fence::after_thread_sync + an mbarrier try_wait (already complete) before every 6 MMAs
//   8  a second warp streams 8 KB bulk copies global -> shared (a 104 KB ring) as fast as they complete
//  16  the four epilogue-like warps hammer shared memory (st.shared / ld.shared transposes)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../h-denseunet_b200/csrc -o mma_rate2 mma_rate2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "tc_common.cuh"

struct Args { int N, feat, R; const uint8_t* gsrc; };

__global__ void __launch_bounds__(256, 1) mma_rate2_kernel(Args a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar_tap[2], bar_ready, bar_copy[13];
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::mbar_init(&bar, 1); tc::mbar_init(&bar_tap[0], 1); tc::mbar_init(&bar_tap[1], 1); tc::mbar_init(&bar_ready, 1);
    for (int i = 0; i < 13; ++i) tc::mbar_init(&bar_copy[i], 1);
    stop = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tbase, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tbase;
  if (tid == 0) tc::mbar_arrive(&bar_ready);               // a barrier whose phase 0 is complete: try_wait returns at once
  __syncthreads();
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, a.N, 0, 0);
    const uint32_t sA = tc::smem_u32(smem), sB = tc::smem_u32(smem + 48 * 1024);
    const uint64_t ad0 = tc::make_smem_desc(0, 184u * 16u, 128u), bd0 = tc::make_smem_desc(0, (uint32_t)a.N * 16u, 128u);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      __syncwarp();
      t0 = clock64();
      uint64_t ad[6], bd[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        ad[u] = ad0 | (uint64_t)(((sA + (uint32_t)(u % 3) * 8192u + (uint32_t)u * 16u) >> 4) & 0x3FFF);
        bd[u] = bd0 | (uint64_t)(((sB + (uint32_t)(u % 3) * (uint32_t)a.N * 64u) >> 4) & 0x3FFF);
      }
      const uint32_t d0 = tm, d1 = (a.feat & 1) ? tm : tm + (uint32_t)a.N;
      for (int i = 0; i < a.R; i += 6) {
        if (a.feat & 4) { tc::mbar_wait(&bar_ready, 0); tc::tc_fence_after(); }
        if (tc::elect_one_sync()) {
          tc::umma_bf16(d0, ad[0], bd[0], idesc, 1u); tc::umma_bf16(d1, ad[1], bd[1], idesc, 1u);
          tc::umma_bf16(d0, ad[2], bd[2], idesc, 1u); tc::umma_bf16(d1, ad[3], bd[3], idesc, 1u);
          tc::umma_bf16(d0, ad[4], bd[4], idesc, 1u); tc::umma_bf16(d1, ad[5], bd[5], idesc, 1u);
          if (a.feat & 2) tc::umma_commit(&bar_tap[(i / 6) & 1]);
        }
        __syncwarp();
      }
      if (tc::elect_one_sync()) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, (uint32_t)pass & 1u);
      t1 = clock64();
    }
    if (tid == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp == 1) {
    if (a.feat & 8) {                                        // weight-loader-like stream into a 13 x 8 KB ring at 96 KB
      uint32_t ph = 0;
      int s = 0;
      long long n = 0;
      while (!stop) {
        if (tc::elect_one_sync()) {
          tc::mbar_arrive_expect_tx(&bar_copy[s], 8192u);
          tc::bulk_g2s(smem + 96 * 1024 + s * 8192, a.gsrc + ((n * 148 + blockIdx.x) % 4096) * 8192, 8192u, &bar_copy[s]);
        }
        __syncwarp();
        tc::mbar_wait(&bar_copy[s], ph);
        ++n;
        if (++s == 13) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    if (a.feat & 16) {
      float* t = reinterpret_cast<float*>(smem + 200 * 1024 - (8 - warp) * 4352 - 4352);
      float acc = 0.f;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 32; ++i) t[lane * 33 + i] = acc + i;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += t[i * 33 + lane];
        __syncwarp();
      }
      if (acc == 12345.f) out[1000] = 1;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int R = 6 * 1024;
  long long* d;
  uint8_t* g;
  cudaMalloc(&d, 2048 * sizeof(long long));
  cudaMalloc(&g, 4096ull * 8192);
  cudaMemset(g, 0, 4096ull * 8192);
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("grid %d, %d MMAs (M=128, K=16, bf16, SS) per CTA in groups of 6; cycles per MMA (median over CTAs)\n", grid, R);
  const int feats[] = {0, 1, 2, 4, 8, 16, 3, 7, 15, 31};
  for (int N : {64, 128})
    for (int f : feats) {
      Args a{N, f, R, g};
      mma_rate2_kernel<<<grid, 256, 220 * 1024>>>(a, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d feat=%d: %s\n", N, f, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("N=%3d features=%2d (%s%s%s%s%s): %7.1f  (min %.1f max %.1f)\n", N, f, f & 1 ? "one-acc " : "", f & 2 ? "commit/6 " : "", f & 4 ? "wait+fence/6 " : "",
             f & 8 ? "bulk-copies " : "", f & 16 ? "smem-traffic " : "", (double)h[grid / 2] / R, (double)h[0] / R, (double)h[grid - 1] / R);
      fflush(stdout);
    }
  return 0;
}
