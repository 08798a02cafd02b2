// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16) as a function of N, of the
// shared-memory operand layout (SWIZZLE_NONE chunk planes / SWIZZLE_128B rows), of the A operand's home (shared memory or
// tensor memory) and of a 16-byte shift of the A start address (what the patch-shifted implicit GEMM does per tap).
// One CTA per SM, one elected thread issues R back-to-back MMAs into two alternating accumulators; cycles per MMA from
// clock64 around the loop (commit + mbarrier wait included once).  Operands are zeros: the datapath does not care.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../h-denseunet_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Args { int N, layout, ts, shift, R, nbuf, mn; };   // mn: both operands MN-major (the weight-gradient kernel's form)

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Args a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&tbase, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tbase;
  if (warp == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, a.N, a.mn, a.mn);
    const uint32_t sA = tc::smem_u32(smem), sB = tc::smem_u32(smem + 96 * 1024);
    const uint64_t sw_bit = (uint64_t)2 << 61;
    // layout 0: K-major chunk planes, A plane = 180 px * 16 B (patch-like), B plane = N * 16 B;  layout 2: 128-byte swizzled rows
    // MN-major, no swizzle: core matrix = 8 K rows x 16 B (8 MN elements), K groups 128 B apart (LBO), MN groups 2 KB apart (SBO)
    const uint64_t ad0 = a.mn ? tc::make_smem_desc(0, 128u, 2048u)
                              : a.layout ? (tc::make_smem_desc(0, 16u, 1024u) | sw_bit) : tc::make_smem_desc(0, 184u * 16u, 128u);
    const uint64_t bd0 = a.mn ? tc::make_smem_desc(0, 128u, 2048u)
                              : a.layout ? (tc::make_smem_desc(0, 16u, 1024u) | sw_bit) : tc::make_smem_desc(0, (uint32_t)a.N * 16u, 128u);
    const uint32_t a_stride = a.mn ? 0u : a.layout ? 16384u : 8192u;     // bytes between the rotating A buffers
    const uint32_t b_stride = a.mn ? 0u : a.layout ? (uint32_t)a.N * 128u : (uint32_t)a.N * 64u;
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {                   // pass 0 warms up
      __syncwarp();
      t0 = clock64();
      if (tc::elect_one_sync()) {
        // descriptors are loop invariants: the loop body is four MMAs and a counter (an earlier version of this benchmark
        // computed them per iteration with integer divisions and measured its own address arithmetic: ~105 cycles per MMA)
        uint64_t ad[4], bd[4];
        uint32_t ta[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t k = (uint32_t)(u % a.nbuf);
          const uint32_t sh = a.shift ? (uint32_t)((u * 7 + 3) % 19) * (a.layout ? 128u : 16u) : 0u;
          ad[u] = ad0 | (uint64_t)(((sA + k * a_stride + sh) >> 4) & 0x3FFF);
          bd[u] = bd0 | (uint64_t)(((sB + (k % (a.layout ? 2u : 4u)) * b_stride) >> 4) & 0x3FFF);
          ta[u] = tm + 480u + (k & 3u) * 8u;
        }
        const uint32_t d0 = tm, d1 = tm + ((a.ts && a.N > 192) ? 0u : (uint32_t)a.N);
        if (a.ts) {
          for (int i = 0; i < a.R; i += 4) {
            umma_ts(d0, ta[0], bd[0], idesc, 1u); umma_ts(d1, ta[1], bd[1], idesc, 1u);
            umma_ts(d0, ta[2], bd[2], idesc, 1u); umma_ts(d1, ta[3], bd[3], idesc, 1u);
          }
        } else {
          for (int i = 0; i < a.R; i += 4) {
            tc::umma_bf16(d0, ad[0], bd[0], idesc, 1u); tc::umma_bf16(d1, ad[1], bd[1], idesc, 1u);
            tc::umma_bf16(d0, ad[2], bd[2], idesc, 1u); tc::umma_bf16(d1, ad[3], bd[3], idesc, 1u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, (uint32_t)pass & 1u);
      t1 = clock64();
    }
    if (tid == 0) out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int R = 4096;
  const bool quick = argc > 2;                              // second argument: only the base rows + the MN-major rows
  long long* d;
  cudaMalloc(&d, 1024 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("grid %d, %d MMAs (M=128, K=16, bf16) per CTA; cycles per MMA: min / median / max over CTAs; floor = N/2\n", grid, R);
  const int Ns[] = {32, 48, 64, 96, 128, 192, 256};
  for (int ts = 0; ts < 2; ++ts)
    for (int layout = 0; layout <= 2; layout += 2)
      for (int shift = 0; shift < 2; ++shift)
        for (int nbuf = 1; nbuf <= 4; nbuf += 3)
          for (int N : Ns) {
            if (ts && (layout || shift)) continue;
            if (quick && (layout || nbuf > 1 || shift)) continue;
            Args a{N, layout, ts, shift, R, nbuf, 0};
            mma_rate_kernel<<<grid, 128, 200 * 1024>>>(a, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d layout=%d ts=%d: %s\n", N, layout, ts, cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            std::sort(h.begin(), h.end());
            printf("A=%s layout=%-6s shift=%d bufs=%d N=%3d : %7.1f / %7.1f / %7.1f   (floor %d)\n", ts ? "tmem" : "smem", layout ? "sw128" : "planes", shift,
                   nbuf, N, (double)h[0] / R, (double)h[grid / 2] / R, (double)h[grid - 1] / R, N / 2);
            fflush(stdout);
          }
  for (int shift = 0; shift < 2; ++shift)
    for (int N : Ns) {
      Args a{N, 0, 0, shift, R, 1, 1};
      mma_rate_kernel<<<grid, 128, 200 * 1024>>>(a, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MN-major N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("A=smem MN-major (both operands) shift=%d N=%3d : %7.1f / %7.1f / %7.1f   (floor %d)\n", shift, N, (double)h[0] / R, (double)h[grid / 2] / R,
             (double)h[grid - 1] / R, N / 2);
      fflush(stdout);
    }
  return 0;
}
