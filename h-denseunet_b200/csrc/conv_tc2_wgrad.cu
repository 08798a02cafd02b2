// Weight gradient, second generation ("tc2"): bf16 operand tensors + TMA (cp.async.bulk.tensor) + tcgen05.
//
//   dW[tap][ci][co] += sum_pos A[pos + off(tap)][ci] * dY[pos][co]        (hybridnet.py:264-298, 11-45, 146-176, 415)
//
// The first-generation kernel (conv_tc_wgrad.cu) fetches the raw fp32 patch with per-thread cp.async and turns it into
// the bf16 MMA operand with 8 SIMT warps inside the GEMM kernel -- once per (tap group x column tile), i.e. six times
// per patch for the 64-wide decoder tail -- and that instruction stream, not the tensor pipe, sets its pace
// (tensor pipe 7 % active).  Here the operand transform happens ONCE, in an HBM-bound pre-pass:
//
//   act_pack_bf16_kernel : A' = bf16( max(a*x+b, 0) [+ second source] ) on the virtual (up-sampled) grid, NDHWC
//                          dY' = bf16(dY)
//
// and the GEMM kernel is a pure copy-engine + tensor-core pipeline: one thread issues TMA tile loads of the patch
// (tile + halo; out-of-bounds rows/columns/slabs/channels arrive as zeros, which IS the convolution's zero padding) and
// of the dY tile into a 4-6 deep shared-memory ring, one thread issues tcgen05.mma, four warps drain the TMEM
// accumulators at the end.  Every filter tap of a group still reads the SAME staged patch through a descriptor whose
// start address is shifted by (th*PW + tw) pixels (the patch-in-smem trick of conv_tc.cu).
//
// Shared-memory operand layouts (both operands MN-major: channels contiguous, the contraction runs over positions):
//   layout 0  "planes"  : 8-channel chunk j of the patch is one TMA box {8, PW, PH} = a dense plane, pixel q at q*16 B,
//                         plane j at j*planeA: the canonical SWIZZLE_NONE MN-major layout (LBO = PW*16 between the two
//                         8-pixel groups of a K step, SBO = planeA between channel chunks) -- identical to the layout
//                         the first-generation kernel builds with SIMT stores.
//   layout 1  "sw128"   : 64-channel block b is one TMA box {64, PW, PH} with CU_TENSOR_MAP_SWIZZLE_128B: pixel q is a
//                         128-byte row at q*128 B (16-byte chunks XOR-swizzled by the row's address bits 7-9), the
//                         canonical SWIZZLE_128B MN-major layout with SBO = PW*128 (next tile row) and LBO = blockA
//                         (next 64 channels).  8x fewer TMA rows per stage.
// The dY tile always uses planes (box {8, 8, 16}: 128 pixels x 16 B per chunk).
#include <cuda.h>
#include <stdlib.h>
#include "hdn_common.cuh"
#include "tc_common.cuh"

int hdn_tc_stem(const hdn_conv* c);   // conv_tc.cu

namespace {

constexpr int W2_THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer (+ TMEM owner), warps 2-5: epilogue
constexpr int W2_MAX_STAGES = 6;

struct Wg2Params {
  CUtensorMap tmA;                    // bf16 A' : (C, W, H, D, N)   or flat (C, M)
  CUtensorMap tmB;                    // bf16 dY': (C, W, H, D, N)   or flat (C, M)
  int D, H, W;
  int kd, kh, kw;
  int Cin, Cout;
  int BN, G, groups, ci_tiles, co_tiles;
  int flat, layout, order;
  int debug;                          // timing experiments (HDN_TC2_DEBUG, results invalid): 1 = no TMA traffic, 2 = K-major operand bits, 4 = N forced to 64
  int PH, PW;
  int unitA, boxA, nunitA_max;        // stride / TMA box bytes of one A unit (plane or 64-channel block); units per stage
  int cwA;                            // channels per A unit
  int stage_bytes, offB, NS;
  int pd_lo, ph_lo, pw_lo;
  int tiles_w, tiles_h;
  long long n_pos_tiles;
  float* dw;
  int tmem_cols;
};

// ---- TMA tile loads ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// SWIZZLE_128B shared-memory matrix descriptor (layout_type 2 at bits 61-63); see tc::make_smem_desc for the fields
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return tc::make_smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(W2_THREADS, 1) conv_wgrad_tc2_kernel(const __grid_constant__ Wg2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the swizzle pattern is a function of the shared-memory address bits: keep every stage 1024-byte aligned
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.NS * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + W2_MAX_STAGES;
  uint64_t* acc_full = bars + 2 * W2_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // ---- CTA work item: (column tile, input-channel tile, tap group) x position split
  int wi = blockIdx.x;
  const int co_t = wi % p.co_tiles; wi /= p.co_tiles;
  const int ci_t = wi % p.ci_tiles; wi /= p.ci_tiles;
  const int grp = wi;
  const int split = blockIdx.y, nsplit = gridDim.y;
  const int ci0 = ci_t * 128, co0 = co_t * p.BN;
  const int taps_hw = p.kh * p.kw;
  const int tap0 = grp * p.G;
  const int gdz = tap0 / taps_hw;                       // depth tap of the group (G <= kh*kw: one depth tap per group)
  const int gth0 = (tap0 % taps_hw) / p.kw;             // first kernel row of the group
  const int gtw0 = (p.G == 1) ? (tap0 % p.kw) : 0;      // single-tap groups stage only that tap's columns

  if (tid == 0) {
    for (int i = 0; i < p.NS; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // position tile t -> (n, d, h0, w0) or the flat row m0; false when the group's depth slab lies outside the volume
  auto tile_valid = [&](long long t, int& n_img, int& d0, int& h0, int& w0, long long& m0) -> bool {
    if (p.flat) { m0 = t * 128; n_img = d0 = h0 = w0 = 0; return true; }
    unsigned r = (unsigned)t;
    const unsigned tw_ = r % (unsigned)p.tiles_w; r /= (unsigned)p.tiles_w;
    const unsigned th_ = r % (unsigned)p.tiles_h; r /= (unsigned)p.tiles_h;
    d0 = (int)(r % (unsigned)p.D); n_img = (int)(r / (unsigned)p.D);
    h0 = (int)th_ * 16; w0 = (int)tw_ * 8; m0 = 0;
    const int vd = d0 - p.pd_lo + gdz;
    return vd >= 0 && vd < p.D;
  };

  const int nuA = min(p.nunitA_max, (min(128, p.Cin - ci0) + p.cwA - 1) / p.cwA);   // A units that hold real channels
  const int nuB = (min(p.BN, p.Cout - co0) + 7) >> 3;                                // dY chunk planes that hold real columns
  const uint32_t tx_bytes = (uint32_t)nuA * (uint32_t)p.boxA + (uint32_t)nuB * 2048u;        // bytes the copy engine reports per stage

  if (warp == 0) {
    // =================================================================== TMA producer (all lanes converged, one elected lane issues)
    {
      if (tc::elect_one_sync()) { tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmB); }
      int st = 0;
      uint32_t ph = 0;
      for (long long t = split; t < p.n_pos_tiles; t += nsplit) {
        int n_img, d0, h0, w0;
        long long m0;
        if (!tile_valid(t, n_img, d0, h0, w0, m0)) continue;
        if (p.debug & 1) break;                           // timing experiment: the MMA stream alone
        tc::mbar_wait(&empty[st], ph ^ 1);
        if (tc::elect_one_sync()) {
          tc::mbar_arrive_expect_tx(&full[st], tx_bytes);
          uint8_t* sA = smem + (size_t)st * p.stage_bytes;
          uint8_t* sB = sA + p.offB;
          if (p.flat) {
            for (int u = 0; u < nuA; ++u) tma_load_2d(sA + (size_t)u * p.unitA, &p.tmA, &full[st], ci0 + u * p.cwA, (int)m0);
            for (int u = 0; u < nuB; ++u) tma_load_2d(sB + (size_t)u * 2048, &p.tmB, &full[st], co0 + u * 8, (int)m0);
          } else {
            const int aw = w0 - p.pw_lo + gtw0, ah = h0 - p.ph_lo + gth0, ad = d0 - p.pd_lo + gdz;
            for (int u = 0; u < nuA; ++u) tma_load_5d(sA + (size_t)u * p.unitA, &p.tmA, &full[st], ci0 + u * p.cwA, aw, ah, ad, n_img);
            for (int u = 0; u < nuB; ++u) tma_load_5d(sB + (size_t)u * 2048, &p.tmB, &full[st], co0 + u * 8, w0, h0, d0, n_img);
          }
        }
        __syncwarp();
        if (++st == p.NS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer (all lanes converged, one elected lane issues)
    {
      const uint32_t idesc = (p.debug & 2) ? tc::make_idesc_bf16(128, (p.debug & 4) ? 64 : p.BN, 0, 0)
                                           : tc::make_idesc_bf16(128, (p.debug & 4) ? 64 : p.BN, 1, 1);
      // A: K group (8 pixels) stride = one patch row; MN group stride = one unit (plane / 64-channel block)
      const uint32_t pix = p.layout ? 128u : 16u;                      // bytes per patch pixel inside a unit
      const uint64_t adesc_hi = p.layout ? make_smem_desc_sw128(0, (uint32_t)p.unitA, (uint32_t)p.PW * pix)
                                         : tc::make_smem_desc(0, (uint32_t)p.PW * pix, (uint32_t)p.unitA);
      const uint64_t bdesc_hi = tc::make_smem_desc(0, 128u, 2048u);
      const uint32_t pixu = pix >> 4;                                  // 16-byte units per pixel
      const uint32_t row2 = 2u * (uint32_t)p.PW * pixu;                // one K step = 16 positions = two patch rows
      const int t20 = tap0 % taps_hw;
      const int twc0 = (p.G == 1) ? 0 : (t20 % p.kw);
      const uint32_t tap_units0 = (uint32_t)((t20 / p.kw - gth0) * p.PW + (t20 % p.kw - gtw0)) * pixu;
      int st = 0;
      uint32_t ph = 0, acc = 0;
      for (long long t = split; t < p.n_pos_tiles; t += nsplit) {
        int n_img, d0, h0, w0;
        long long m0;
        if (!tile_valid(t, n_img, d0, h0, w0, m0)) continue;
        if (!(p.debug & 1)) tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        if (tc::elect_one_sync()) {
        const uint8_t* sA = smem + (size_t)st * p.stage_bytes;
        const uint64_t ad0 = adesc_hi | (uint64_t)((tc::smem_u32(sA) >> 4) & 0x3FFF);
        const uint64_t bd0 = bdesc_hi | (uint64_t)((tc::smem_u32(sA + p.offB) >> 4) & 0x3FFF);
        if (p.order == 0) {
          // tap-major: the 8 K steps of a tap back to back (a dependent chain on one accumulator)
          uint32_t tap_units = tap_units0;
          int twc = twc0;
          uint32_t tm = tmem_base;
          for (int g = 0; g < p.G; ++g, tm += (uint32_t)p.BN) {
            const uint64_t ad = ad0 + tap_units;
            tc::umma_bf16(tm, ad, bd0, idesc, acc);
#pragma unroll
            for (int r = 1; r < 8; ++r) tc::umma_bf16(tm, ad + (uint64_t)(r * row2), bd0 + (uint64_t)(r * 16), idesc, 1u);
            if (++twc == p.kw) { twc = 0; tap_units += (uint32_t)(p.PW - p.kw + 1) * pixu; } else tap_units += pixu;
          }
        } else {
          // K-step-major: consecutive MMAs go to DIFFERENT accumulators (one per tap), so an MMA never waits for the
          // accumulation of the one issued just before it; the dY descriptor is shared by the G MMAs of a K step
          for (int r = 0; r < 8; ++r) {
            uint32_t tap_units = tap_units0;
            int twc = twc0;
            uint32_t tm = tmem_base;
            const uint64_t adr = ad0 + (uint64_t)(r * row2), bdr = bd0 + (uint64_t)(r * 16);
            const uint32_t a2 = r ? 1u : acc;
            for (int g = 0; g < p.G; ++g, tm += (uint32_t)p.BN) {
              tc::umma_bf16(tm, adr + tap_units, bdr, idesc, a2);
              if (++twc == p.kw) { twc = 0; tap_units += (uint32_t)(p.PW - p.kw + 1) * pixu; } else tap_units += pixu;
            }
          }
        }
        tc::umma_commit(&empty[st]);
        }
        __syncwarp();
        acc = 1;
        if (++st == p.NS) { st = 0; ph ^= 1; }
      }
      if (tc::elect_one_sync()) tc::umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // =================================================================== epilogue: TMEM -> dW (vector reductions)
    // did this CTA accumulate anything?  (a tap group whose depth slab is outside for every tile of the split did not)
    int produced = 0;
    if (lane == 0) {
      for (long long t = split; t < p.n_pos_tiles && !produced; t += nsplit) {
        int n_img, d0, h0, w0;
        long long m0;
        if (tile_valid(t, n_img, d0, h0, w0, m0)) produced = 1;
      }
    }
    produced = __shfl_sync(0xffffffffu, produced, 0);
    tc::mbar_wait_sleep(acc_full, 0);
    tc::tc_fence_after();
    const int q4 = warp & 3;                              // TMEM lane quarter this warp may read
    const int ci = ci0 + q4 * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    float v[16];
    for (int g = 0; produced && g < p.G; ++g) {
      const int tap = tap0 + g;
      for (int ce = 0; ce < p.BN; ce += 16) {
        tc::tmem_ld16(taddr + (uint32_t)(g * p.BN + ce), v);
        if (ci < p.Cin) {
          float* q = p.dw + ((long long)tap * p.Cin + ci) * p.Cout + co0 + ce;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            if (co0 + ce + i < p.Cout)                      // Cout % 8 == 0: a quad is all-in or all-out
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + i), "f"(v[i]), "f"(v[i + 1]),
                           "f"(v[i + 2]), "f"(v[i + 3])
                           : "memory");
          }
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---- operand pre-pass: fp32 sources (+ BN -> Scale -> ReLU, Add, UpSampling) -> bf16 NDHWC on the virtual grid ------
struct PackParams {
  int D, H, W, C, nsrc;
  hdn_src src[2];
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;              // optional: bf16 tail x - bf16(x) (the second operand term of bf16x3)
  int interleave;                     // 1: ONE output tensor of 2*ceil32(C) channels per pixel, [head 32 | tail 32] per 32-channel group
                                      //    (a 128-byte row = one K stage of the SWIZZLE_128B form); channels >= C are written as zeros
  long long total;                    // pixels * (C / 8)   (interleave: pixels * (ceil32(C) / 8))
};

__global__ void __launch_bounds__(256) act_pack_bf16_kernel(const __grid_constant__ PackParams p) {
  const int nch = p.interleave ? ((p.C + 31) / 32 * 4) : (p.C >> 3);
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < p.total; idx += (long long)gridDim.x * 256) {
    const long long m = idx / nch;
    const int c = (int)(idx - m * nch) * 8;
    int n, d, h, w;
    hdn_decode(m, p.D, p.H, p.W, n, d, h, w);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s >= p.nsrc || c >= p.C) break;
      const hdn_src& S = p.src[s];
      const float* x = S.t.p + hdn_src_off(S, n, d, h, w) + c;
      float4 x0 = __ldg(reinterpret_cast<const float4*>(x)), x1 = __ldg(reinterpret_cast<const float4*>(x + 4));
      // same arithmetic as every other kernel of the library: one fused multiply-add, then the ReLU (hdn_prologue)
      float4 a0 = make_float4(1.f, 1.f, 1.f, 1.f), a1 = a0, b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (S.pa) { a0 = __ldg(reinterpret_cast<const float4*>(S.pa + c)); a1 = __ldg(reinterpret_cast<const float4*>(S.pa + c + 4)); }
      if (S.pb) { b0 = __ldg(reinterpret_cast<const float4*>(S.pb + c)); b1 = __ldg(reinterpret_cast<const float4*>(S.pb + c + 4)); }
      if (S.pa || S.pb) {
        x0.x = fmaf(a0.x, x0.x, b0.x); x0.y = fmaf(a0.y, x0.y, b0.y); x0.z = fmaf(a0.z, x0.z, b0.z); x0.w = fmaf(a0.w, x0.w, b0.w);
        x1.x = fmaf(a1.x, x1.x, b1.x); x1.y = fmaf(a1.y, x1.y, b1.y); x1.z = fmaf(a1.z, x1.z, b1.z); x1.w = fmaf(a1.w, x1.w, b1.w);
      }
      if (S.relu) {
        x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
        x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
      }
      v0.x += x0.x; v0.y += x0.y; v0.z += x0.z; v0.w += x0.w; v1.x += x1.x; v1.y += x1.y; v1.z += x1.z; v1.w += x1.w;
    }
    uint4 o;
    if (p.interleave) {
      uint4 t;
      tc::pack_split_bf16x2(v0.x, v0.y, o.x, t.x); tc::pack_split_bf16x2(v0.z, v0.w, o.y, t.y);
      tc::pack_split_bf16x2(v1.x, v1.y, o.z, t.z); tc::pack_split_bf16x2(v1.z, v1.w, o.w, t.w);
      const int g = c >> 5, j = (c & 31) >> 3;                        // 32-channel group, 8-channel chunk inside it
      uint4* row = reinterpret_cast<uint4*>(p.out) + m * (long long)(nch * 2) + g * 8;
      row[j] = o;
      row[4 + j] = t;
      continue;
    }
    if (p.out_lo) {
      uint4 t;
      tc::pack_split_bf16x2(v0.x, v0.y, o.x, t.x); tc::pack_split_bf16x2(v0.z, v0.w, o.y, t.y);
      tc::pack_split_bf16x2(v1.x, v1.y, o.z, t.z); tc::pack_split_bf16x2(v1.z, v1.w, o.w, t.w);
      reinterpret_cast<uint4*>(p.out_lo)[idx] = t;
    } else {
      o.x = tc::pack_bf16x2(v0.x, v0.y); o.y = tc::pack_bf16x2(v0.z, v0.w);
      o.z = tc::pack_bf16x2(v1.x, v1.y); o.w = tc::pack_bf16x2(v1.z, v1.w);
    }
    reinterpret_cast<uint4*>(p.out)[idx] = o;
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// bf16 NDHWC tensor (C innermost) as a 5-D tile map with box {bc, bw, bh, 1, 1}, or flat [M][C] with box {bc, 128}
int make_map(CUtensorMap* tm, const void* base, int flat, long long M, int N, int D, int H, int W, int C, int bc, int bw,
             int bh, int swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { hdn_set_error("conv_wgrad tc2: cuTensorMapEncodeTiled is not available from this driver"); return HDN_ERR_CUDA; }
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  int rank;
  if (flat) {
    rank = 2;
    dims[0] = (cuuint64_t)C; dims[1] = (cuuint64_t)M;
    strides[0] = (cuuint64_t)C * 2;
    box[0] = (cuuint32_t)bc; box[1] = 128;
  } else {
    rank = 5;
    dims[0] = (cuuint64_t)C; dims[1] = (cuuint64_t)W; dims[2] = (cuuint64_t)H; dims[3] = (cuuint64_t)D; dims[4] = (cuuint64_t)N;
    strides[0] = (cuuint64_t)C * 2;
    strides[1] = strides[0] * (cuuint64_t)W;
    strides[2] = strides[1] * (cuuint64_t)H;
    strides[3] = strides[2] * (cuuint64_t)D;
    box[0] = (cuuint32_t)bc; box[1] = (cuuint32_t)bw; box[2] = (cuuint32_t)bh; box[3] = 1; box[4] = 1;
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { hdn_set_error("conv_wgrad tc2: cuTensorMapEncodeTiled failed (%d)", (int)r); return HDN_ERR_CUDA; }
  return HDN_OK;
}

struct Wg2Plan {
  int BN, G, groups, ci_tiles, co_tiles, flat, layout, PH, PW, unitA, boxA, nunitA_max, cwA, stage_bytes, offB, NS, tmem_cols;
  int tiles_h, tiles_w, splits;
  long long n_pos_tiles, M;
  size_t smem;
  long long ws_bytes, offA_ws, offB_ws;
};

int g_layout = -1;
int wg2_layout() {
  if (g_layout < 0) {
    const char* e = getenv("HDN_TC2_LAYOUT");
    g_layout = e ? atoi(e) : 0;
    if (g_layout < 0 || g_layout > 1) g_layout = 0;
  }
  return g_layout;
}

bool wg2_plan(const hdn_conv* c, Wg2Plan& best) {
  const int T = c->kd * c->kh * c->kw, taps_hw = c->kh * c->kw;
  const int flat = (T == 1) ? 1 : 0;                     // up-sampling is resolved by the pre-pass
  const int layout = wg2_layout();
  const long long M = (long long)c->N * c->D * c->H * c->W;
  const int ci_tiles = (c->Cin + 127) / 128;
  double best_cost = 1e300;
  bool found = false;
  const int cand[3] = {taps_hw, c->kw, 1};               // taps per group: one depth slab, one kernel row, a single tap
  for (int k = 0; k < 3; ++k) {
    const int G = cand[k];
    if (k > 0 && G == cand[k - 1]) continue;
    int bn_max = (512 / G) / 16 * 16;
    if (bn_max > 256) bn_max = 256;
    if (bn_max < 16) continue;
    Wg2Plan pl;
    memset(&pl, 0, sizeof(pl));
    pl.G = G; pl.groups = T / G; pl.ci_tiles = ci_tiles; pl.flat = flat; pl.layout = layout; pl.M = M;
    pl.co_tiles = (c->Cout + bn_max - 1) / bn_max;
    const int bn = (c->Cout + pl.co_tiles - 1) / pl.co_tiles;
    pl.BN = (bn + 15) / 16 * 16;
    const int rows = (G == taps_hw) ? c->kh : 1, cols = (G == 1) ? 1 : c->kw;
    pl.PH = 16 + rows - 1;
    pl.PW = 8 + cols - 1;
    const int P = pl.PH * pl.PW;
    if (layout == 0) { pl.cwA = 8; pl.nunitA_max = 16; pl.boxA = P * 16; pl.unitA = (P * 16 + 127) / 128 * 128; }
    else             { pl.cwA = 64; pl.nunitA_max = 2; pl.boxA = P * 128; pl.unitA = (P * 128 + 1023) / 1024 * 1024; }
    pl.offB = (pl.nunitA_max * pl.unitA + 1023) / 1024 * 1024;
    pl.stage_bytes = (pl.offB + (pl.BN / 8) * 2048 + 1023) / 1024 * 1024;
    const int budget = 227 * 1024 - 1024 - 256;            // alignment slack + barriers
    pl.NS = budget / pl.stage_bytes;
    if (pl.NS > W2_MAX_STAGES) pl.NS = W2_MAX_STAGES;
    if (pl.NS < 2) continue;
    pl.smem = (size_t)pl.NS * pl.stage_bytes + 1024 + 256;
    pl.tiles_h = (c->H + 15) / 16;
    pl.tiles_w = (c->W + 7) / 8;
    pl.n_pos_tiles = flat ? (M + 127) / 128 : (long long)c->N * c->D * pl.tiles_h * pl.tiles_w;
    int cols_t = 32;
    while (cols_t < G * pl.BN) cols_t *= 2;
    pl.tmem_cols = cols_t;
    // bytes staged per position tile over all work items (what the copy engine moves) -- the pace of narrow layers
    const double cost = (double)pl.ci_tiles * pl.co_tiles * pl.groups * ((double)P * 128.0 + 128.0 * pl.BN);
    if (cost < best_cost) { best_cost = cost; best = pl; found = true; }
  }
  if (!found) return false;
  const long long items = (long long)best.ci_tiles * best.co_tiles * best.groups;
  long long splits = 148 / items;
  if (splits < 1) splits = 1;
  if (splits > best.n_pos_tiles) splits = best.n_pos_tiles;
  best.splits = (int)splits;
  best.offA_ws = 0;
  best.offB_ws = (M * c->Cin * 2 + 255) / 256 * 256;
  best.ws_bytes = best.offB_ws + (M * c->Cout * 2 + 255) / 256 * 256;
  return true;
}

int pack_launch(const hdn_src* srcs, int nsrc, int N, int D, int H, int W, int C, __nv_bfloat16* out, cudaStream_t st,
                __nv_bfloat16* out_lo = nullptr, int interleave = 0) {
  PackParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.D = D; pp.H = H; pp.W = W; pp.C = C; pp.nsrc = nsrc;
  for (int i = 0; i < nsrc; ++i) pp.src[i] = srcs[i];
  pp.out = out;
  pp.out_lo = out_lo;
  pp.interleave = interleave;
  pp.total = (long long)N * D * H * W * (interleave ? (C + 31) / 32 * 4 : C / 8);
  const long long blocks = (pp.total + 255) / 256;
  const unsigned grid = (unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks);
  HDN_LAUNCHED(1), act_pack_bf16_kernel<<<grid, 256, 0, st>>>(pp);
  HDN_CHECK_LAUNCH("act_pack_bf16");
  return HDN_OK;
}

}  // namespace

// shared with the fprop / dgrad kernel's TMA mode (conv_tc.cu)
int hdn_tc2_make_map(CUtensorMap* tm, const void* base, int flat, long long M, int N, int D, int H, int W, int C, int bc, int bw, int bh,
                     int swizzle128) {
  return make_map(tm, base, flat, M, N, D, H, W, C, bc, bw, bh, swizzle128);
}
int hdn_tc2_pack(const hdn_src* srcs, int nsrc, int N, int D, int H, int W, int C, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st,
                 int interleave) {
  return pack_launch(srcs, nsrc, N, D, H, W, C, hi, st, lo, interleave);
}

// HDN_WGRAD_TC2=0 keeps the first-generation weight-gradient kernel for every shape (default 1)
static int g_wgrad_tc2 = -1;
int hdn_wgrad_tc2_enabled() {
  if (g_wgrad_tc2 < 0) {
    const char* e = getenv("HDN_WGRAD_TC2");
    g_wgrad_tc2 = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_wgrad_tc2;
}
void hdn_wgrad_tc2_set(int v) { g_wgrad_tc2 = v ? 1 : 0; }
void hdn_tc2_layout_set(int v) { g_layout = v ? 1 : 0; }

// shapes the tc2 weight gradient takes: plain bf16 operands (precision 1), stride-1 "same" 1x1x1 / 1x3x3 / 3x3x3
int hdn_wgrad_tc2_supported(const hdn_conv* c) {
  if (c->precision != 1 || hdn_tc_stem(c)) return 0;
  if (c->sd != 1 || c->sh != 1 || c->sw != 1) return 0;
  const bool k111 = c->kd == 1 && c->kh == 1 && c->kw == 1;
  const bool k133 = c->kd == 1 && c->kh == 3 && c->kw == 3;
  const bool k333 = c->kd == 3 && c->kh == 3 && c->kw == 3;
  if (!(k111 || k133 || k333)) return 0;
  if (c->pd != c->kd / 2 || c->ph != c->kh / 2 || c->pw != c->kw / 2) return 0;
  const hdn_src& s0 = c->src[0];
  if (c->D != s0.D * s0.ud || c->H != s0.H * s0.uh || c->W != s0.W * s0.uw) return 0;
  for (int i = 0; i < c->nsrc; ++i) {
    const hdn_src& s = c->src[i];
    if (s.t.ldc % 4 || s.t.coff % 4 || (reinterpret_cast<uintptr_t>(s.t.p) & 15) != 0) return 0;
  }
  if (c->y.ldc % 4 || c->y.coff % 4 || (reinterpret_cast<uintptr_t>(c->y.p) & 15) != 0) return 0;
  if (c->Cin % 8 || c->Cout % 8) return 0;
  if ((long long)c->N * c->D * c->H * c->W >= (1ll << 31)) return 0;
  Wg2Plan pl;
  return wg2_plan(c, pl) ? 1 : 0;
}

long long hdn_wgrad_tc2_workspace(const hdn_conv* c) {
  Wg2Plan pl;
  if (!hdn_wgrad_tc2_supported(c) || !wg2_plan(c, pl)) return 0;
  return pl.ws_bytes;
}

int hdn_wgrad_tc2_plan_info(const hdn_conv* c, int* out) {
  Wg2Plan pl;
  if (!wg2_plan(c, pl)) return HDN_ERR_UNSUPPORTED;
  out[0] = pl.BN; out[1] = pl.co_tiles; out[2] = pl.ci_tiles; out[3] = pl.cwA; out[4] = pl.G; out[5] = pl.NS;
  out[6] = pl.tmem_cols; out[7] = (int)pl.smem; out[8] = pl.flat; out[9] = pl.PH * pl.PW; out[10] = 0;
  out[11] = 0; out[12] = pl.ci_tiles * pl.co_tiles * pl.groups * pl.splits; out[13] = 1;
  out[14] = pl.BN; out[15] = 128;
  return HDN_OK;
}

int hdn_conv_wgrad_tc2(const hdn_conv* c, float* dw, cudaStream_t st) {
  Wg2Plan pl;
  HDN_CHECK_ARG(wg2_plan(c, pl), "conv_wgrad tc2: no plan for this shape");
  HDN_CHECK_ARG(c->ws != nullptr && c->ws_bytes >= pl.ws_bytes, "conv_wgrad tc2: workspace too small (%lld < %lld bytes)",
                (long long)c->ws_bytes, (long long)pl.ws_bytes);
  HDN_CHECK_ARG((reinterpret_cast<uintptr_t>(c->ws) & 255) == 0, "conv_wgrad tc2: workspace must be 256-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hdn_set_error("conv_wgrad tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
    attr_set = true;
  }
  __nv_bfloat16* abf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(c->ws) + pl.offA_ws);
  __nv_bfloat16* bbf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(c->ws) + pl.offB_ws);
  // pre-pass: both operands to bf16, once
  int rc = pack_launch(c->src, c->nsrc, c->N, c->D, c->H, c->W, c->Cin, abf, st);
  if (rc) return rc;
  hdn_src dys;
  memset(&dys, 0, sizeof(dys));
  dys.t = c->y; dys.D = c->D; dys.H = c->H; dys.W = c->W; dys.ud = dys.uh = dys.uw = 1;
  rc = pack_launch(&dys, 1, c->N, c->D, c->H, c->W, c->Cout, bbf, st);
  if (rc) return rc;

  Wg2Params p;
  memset(&p, 0, sizeof(p));
  rc = make_map(&p.tmA, abf, pl.flat, pl.M, c->N, c->D, c->H, c->W, c->Cin, pl.cwA, pl.PW, pl.PH, pl.layout);
  if (rc) return rc;
  rc = make_map(&p.tmB, bbf, pl.flat, pl.M, c->N, c->D, c->H, c->W, c->Cout, 8, 8, 16, 0);
  if (rc) return rc;
  p.D = c->D; p.H = c->H; p.W = c->W;
  p.kd = c->kd; p.kh = c->kh; p.kw = c->kw;
  p.Cin = c->Cin; p.Cout = c->Cout;
  p.BN = pl.BN; p.G = pl.G; p.groups = pl.groups; p.ci_tiles = pl.ci_tiles; p.co_tiles = pl.co_tiles;
  p.flat = pl.flat; p.layout = pl.layout; p.PH = pl.PH; p.PW = pl.PW;
  { const char* e = getenv("HDN_TC2_ORDER"); p.order = e ? atoi(e) : 0; }       // 0 = tap-major (measured faster on B200, profiles/r02c_*)
  { const char* e = getenv("HDN_TC2_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.unitA = pl.unitA; p.boxA = pl.boxA; p.nunitA_max = pl.nunitA_max; p.cwA = pl.cwA;
  p.stage_bytes = pl.stage_bytes; p.offB = pl.offB; p.NS = pl.NS;
  p.pd_lo = c->kd / 2; p.ph_lo = c->kh / 2; p.pw_lo = c->kw / 2;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.n_pos_tiles = pl.n_pos_tiles;
  p.dw = dw; p.tmem_cols = pl.tmem_cols;
  HDN_CHECK_ARG(pl.n_pos_tiles < (1ll << 31), "conv_wgrad tc2: too many position tiles");
  dim3 grid((unsigned)(pl.ci_tiles * pl.co_tiles * pl.groups), (unsigned)pl.splits);
  HDN_LAUNCHED(1), conv_wgrad_tc2_kernel<<<grid, W2_THREADS, pl.smem, st>>>(p);
  HDN_CHECK_LAUNCH("conv_wgrad_tc2");
  return HDN_OK;
}
