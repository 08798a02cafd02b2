// sm_100a primitives used by the tcgen05 convolution kernels: mbarrier, bulk async copy (TMA
// engine, UBLKCP), tensor-memory allocation, UMMA descriptors, tcgen05.mma / commit / ld.
// Plain inline PTX; bit layouts follow the PTX ISA "matrix descriptor" / "instruction
// descriptor" tables for tcgen05 (kind::f16, SWIZZLE_NONE canonical layouts).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// waiting role that must not steal issue slots from the warps doing the work on the same scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// One lane of a CONVERGED warp.  Single-thread roles (tcgen05.mma issue, TMA issue) run their loops with all 32 lanes
// converged and put only the issue itself under this predicate: ptxas then emits the uniform-datapath instruction
// (UTCHMMA / UTMALDG) once, straight-line.  Under `if (lane == 0)` the region is divergent for the compiler and every
// such instruction is wrapped in its own ELECT / BRA.U.ANY serialisation loop with R2UR moves in front of it
// (~9 instructions and ~50 issue cycles per MMA: measured, profiles/r02d_tc2_mma_issue.txt).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %2;\n\t"
      "@px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bulk async copy global -> shared (TMA engine, 1-D) ---------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMA tile loads (cp.async.bulk.tensor; the map is a CUtensorMap in kernel-parameter space) ----------------
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- tensor memory ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE.  Canonical layouts (units of 16 bytes):
//   K-major  : ((8, n), 2) : ((1, SBO), LBO)   8 rows x 16 B core matrices; SBO = stride between
//              8-row groups (M/N direction), LBO = stride between the two K-adjacent core matrices.
//   MN-major : ((1, n), (8, k)) : ((X, SBO), (1, LBO))  core matrix = 8 K-rows of 16 B (8 MN
//              elements each); SBO = stride between MN-adjacent core matrices, LBO = stride
//              between K-adjacent ones.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, dense, no negate.
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                         // c_format = F32
  d |= 1u << 7;                         // a_format = BF16
  d |= 1u << 10;                        // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// Same, with the operand formats spelled out: 0 = F16, 1 = BF16 (kind::f16 takes either for A and for B).
__host__ __device__ __forceinline__ uint32_t make_idesc_f16kind(int M, int N, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                         // c_format = F32
  d |= (uint32_t)(a_fmt & 7) << 7;
  d |= (uint32_t)(b_fmt & 7) << 10;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// D[tmem] (+)= A[smem] * B[smem];  one thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  // the registers are only valid after wait::ld: tie them to the wait so no use is hoisted above it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// bf16x3 operand split: (a, b) -> head = bf16_rn(x), tail = bf16_rn(x - head); head + tail carries ~16 significant bits
__device__ __forceinline__ void pack_split_bf16x2(float a, float b, uint32_t& head, uint32_t& tail) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  head = *reinterpret_cast<const uint32_t*>(&h);
  tail = pack_bf16x2(a - hf.x, b - hf.y);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// same split with the tail kept in IEEE half precision: the residual x - head is <= 2^-9 |x|, and a half carries 11
// significant bits of it against bfloat16's 8, so head + tail carries ~19 significant bits instead of ~16 (the MMA pairs a
// bf16 operand with an f16 one: kind::f16 takes the two formats independently for A and B)
__device__ __forceinline__ void pack_split_bf16_f16x2(float a, float b, uint32_t& head, uint32_t& tail) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  head = *reinterpret_cast<const uint32_t*>(&h);
  tail = pack_f16x2(a - hf.x, b - hf.y);
}
template <int SPLIT>
__device__ __forceinline__ void pack_split_any(float a, float b, uint32_t& head, uint32_t& tail) {
  if (SPLIT == 2) pack_split_bf16_f16x2(a, b, head, tail); else pack_split_bf16x2(a, b, head, tail);
}

// Operand transform, warp-per-chunk form: one warp turns ONE 8-channel chunk of a raw fp32 stage (pixel stride RS
// floats) into bf16 for the pixels q0, q0 + qstep, ... (lane = pixel).  The chunk's per-channel (a, b) are read once and
// stay in registers for the whole stage, and the prologue form is a compile-time constant, so a pixel costs two 16-byte
// reads, 8 FFMA + 8 FMNMX (MODE 1), 4 packs and one 16-byte store.
//   MODE 0: y = x            MODE 1: y = max(a*x + b, 0)   (BatchNorm -> Scale -> ReLU folded; hdn_src.pa/pb/relu all set)
//   SPLIT : 1 / 2: also store the tail (x - head) at dtail (bf16x3), as bfloat16 (1) or as IEEE half (2)
// vq[q] = source-patch pixel of virtual pixel q, or -1 => zero padding (applied after the prologue, like ZeroPadding).
template <int MODE, int SPLIT>
__device__ __forceinline__ void xform_px(float4 va, float4 vb, const float4& a0, const float4& a1, const float4& b0,
                                         const float4& b1, bool ok, uint4& o, uint4& t) {
  if (MODE == 1) {
    va.x = fmaxf(fmaf(a0.x, va.x, b0.x), 0.f); va.y = fmaxf(fmaf(a0.y, va.y, b0.y), 0.f);
    va.z = fmaxf(fmaf(a0.z, va.z, b0.z), 0.f); va.w = fmaxf(fmaf(a0.w, va.w, b0.w), 0.f);
    vb.x = fmaxf(fmaf(a1.x, vb.x, b1.x), 0.f); vb.y = fmaxf(fmaf(a1.y, vb.y, b1.y), 0.f);
    vb.z = fmaxf(fmaf(a1.z, vb.z, b1.z), 0.f); vb.w = fmaxf(fmaf(a1.w, vb.w, b1.w), 0.f);
  }
  if (SPLIT) {
    pack_split_any<SPLIT>(va.x, va.y, o.x, t.x); pack_split_any<SPLIT>(va.z, va.w, o.y, t.y);
    pack_split_any<SPLIT>(vb.x, vb.y, o.z, t.z); pack_split_any<SPLIT>(vb.z, vb.w, o.w, t.w);
    if (!ok) t = make_uint4(0u, 0u, 0u, 0u);
  } else {
    o.x = pack_bf16x2(va.x, va.y); o.y = pack_bf16x2(va.z, va.w);
    o.z = pack_bf16x2(vb.x, vb.y); o.w = pack_bf16x2(vb.z, vb.w);
  }
  if (!ok) o = make_uint4(0u, 0u, 0u, 0u);                // zero padding / channels past K: applied after the prologue
}
template <int MODE, int SPLIT>
__device__ __forceinline__ void xform_chunk(const float* rawc, int RS, const int* vq, int P, int q0, int qstep,
                                            const float* a8, const float* b8, bool cvalid, uint8_t* dchunk, uint8_t* dtail) {
  float4 a0 = make_float4(1.f, 1.f, 1.f, 1.f), a1 = a0, b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (MODE == 1) {
    a0 = *reinterpret_cast<const float4*>(a8); a1 = *reinterpret_cast<const float4*>(a8 + 4);
    b0 = *reinterpret_cast<const float4*>(b8); b1 = *reinterpret_cast<const float4*>(b8 + 4);
  }
  // two pixels per trip, branch-free (a padded pixel reads row 0 and is zeroed by a select), so the four 16-byte
  // reads of a trip are in flight together: the producers are few warps per scheduler and otherwise latency-bound
  for (int q = q0; q < P; q += 2 * qstep) {
    const int qb = q + qstep;
    const bool hasb = qb < P;
    const int sqa = vq[q], sqb = hasb ? vq[qb] : -1;
    const float* rowa = rawc + max(sqa, 0) * RS;
    const float* rowb = rawc + max(sqb, 0) * RS;
    const float4 xa0 = *reinterpret_cast<const float4*>(rowa), xa1 = *reinterpret_cast<const float4*>(rowa + 4);
    const float4 xb0 = *reinterpret_cast<const float4*>(rowb), xb1 = *reinterpret_cast<const float4*>(rowb + 4);
    uint4 oa, ta, ob, tb;
    xform_px<MODE, SPLIT>(xa0, xa1, a0, a1, b0, b1, sqa >= 0 && cvalid, oa, ta);
    xform_px<MODE, SPLIT>(xb0, xb1, a0, a1, b0, b1, sqb >= 0 && cvalid, ob, tb);
    *reinterpret_cast<uint4*>(dchunk + (uint32_t)q * 16u) = oa;
    if (SPLIT) *reinterpret_cast<uint4*>(dtail + (uint32_t)q * 16u) = ta;
    if (hasb) {
      *reinterpret_cast<uint4*>(dchunk + (uint32_t)qb * 16u) = ob;
      if (SPLIT) *reinterpret_cast<uint4*>(dtail + (uint32_t)qb * 16u) = tb;
    }
  }
}
// Two-source form (A = f0(src0) + f1(src1): the Add in front of a convolution, merge.py:207-211): same mapping, the
// (a, b) of both sources in registers, four 16-byte reads in flight per pixel.  Zero padding depends on the virtual
// pixel alone, so vq0[q] < 0 exactly when vq1[q] < 0.
template <int MODE>
__device__ __forceinline__ void xform_apply(float4& va, float4& vb, const float4& a0, const float4& a1, const float4& b0,
                                            const float4& b1) {
  if (MODE == 1) {
    va.x = fmaxf(fmaf(a0.x, va.x, b0.x), 0.f); va.y = fmaxf(fmaf(a0.y, va.y, b0.y), 0.f);
    va.z = fmaxf(fmaf(a0.z, va.z, b0.z), 0.f); va.w = fmaxf(fmaf(a0.w, va.w, b0.w), 0.f);
    vb.x = fmaxf(fmaf(a1.x, vb.x, b1.x), 0.f); vb.y = fmaxf(fmaf(a1.y, vb.y, b1.y), 0.f);
    vb.z = fmaxf(fmaf(a1.z, vb.z, b1.z), 0.f); vb.w = fmaxf(fmaf(a1.w, vb.w, b1.w), 0.f);
  }
}
template <int M0, int M1, int SPLIT>
__device__ __forceinline__ void xform_chunk2(const float* raw0c, const float* raw1c, int RS, const int* vq0, const int* vq1,
                                             int P, int q0, int qstep, const float* a80, const float* b80, const float* a81,
                                             const float* b81, bool cvalid, uint8_t* dchunk, uint8_t* dtail) {
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 p0 = one, p1 = one, r0 = zero, r1 = zero, s0 = one, s1 = one, t0 = zero, t1 = zero;
  if (M0 == 1) {
    p0 = *reinterpret_cast<const float4*>(a80); p1 = *reinterpret_cast<const float4*>(a80 + 4);
    r0 = *reinterpret_cast<const float4*>(b80); r1 = *reinterpret_cast<const float4*>(b80 + 4);
  }
  if (M1 == 1) {
    s0 = *reinterpret_cast<const float4*>(a81); s1 = *reinterpret_cast<const float4*>(a81 + 4);
    t0 = *reinterpret_cast<const float4*>(b81); t1 = *reinterpret_cast<const float4*>(b81 + 4);
  }
  for (int q = q0; q < P; q += qstep) {
    const int sq0 = vq0[q], sq1 = vq1[q];
    const float* rowa = raw0c + max(sq0, 0) * RS;
    const float* rowb = raw1c + max(sq1, 0) * RS;
    float4 xa0 = *reinterpret_cast<const float4*>(rowa), xa1 = *reinterpret_cast<const float4*>(rowa + 4);
    float4 xb0 = *reinterpret_cast<const float4*>(rowb), xb1 = *reinterpret_cast<const float4*>(rowb + 4);
    xform_apply<M0>(xa0, xa1, p0, p1, r0, r1);
    xform_apply<M1>(xb0, xb1, s0, s1, t0, t1);
    xa0.x += xb0.x; xa0.y += xb0.y; xa0.z += xb0.z; xa0.w += xb0.w;
    xa1.x += xb1.x; xa1.y += xb1.y; xa1.z += xb1.z; xa1.w += xb1.w;
    uint4 o, t;
    xform_px<0, SPLIT>(xa0, xa1, one, one, zero, zero, sq0 >= 0 && sq1 >= 0 && cvalid, o, t);
    *reinterpret_cast<uint4*>(dchunk + (uint32_t)q * 16u) = o;
    if (SPLIT) *reinterpret_cast<uint4*>(dtail + (uint32_t)q * 16u) = t;
  }
}
// the two-source shapes the networks contain: (BN+ReLU, BN+ReLU) = fianl_conv (hybridnet.py:414-415) and
// (plain, BN+ReLU) = the skip-add decoder (denseunet.py:190-209).  Returns false for any other combination.
__device__ __forceinline__ bool xform_chunk2_any(int m0, int m1, int split, const float* raw0c, const float* raw1c, int RS,
                                                 const int* vq0, const int* vq1, int P, int q0, int qstep, const float* a80,
                                                 const float* b80, const float* a81, const float* b81, bool cvalid,
                                                 uint8_t* dchunk, uint8_t* dtail) {
  if (m0 == 1 && m1 == 1) {
    if (split == 2)  xform_chunk2<1, 1, 2>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    else if (split)  xform_chunk2<1, 1, 1>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    else             xform_chunk2<1, 1, 0>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    return true;
  }
  if (m0 == 0 && m1 == 1) {
    if (split == 2)  xform_chunk2<0, 1, 2>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    else if (split)  xform_chunk2<0, 1, 1>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    else             xform_chunk2<0, 1, 0>(raw0c, raw1c, RS, vq0, vq1, P, q0, qstep, a80, b80, a81, b81, cvalid, dchunk, dtail);
    return true;
  }
  return false;
}

// dispatch on the two runtime-uniform switches
__device__ __forceinline__ void xform_chunk_any(int mode, int split, const float* rawc, int RS, const int* vq, int P, int q0,
                                                int qstep, const float* a8, const float* b8, bool cvalid, uint8_t* dchunk,
                                                uint8_t* dtail) {
  if (mode == 1) {
    if (split == 2)  xform_chunk<1, 2>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
    else if (split)  xform_chunk<1, 1>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
    else             xform_chunk<1, 0>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
  } else {
    if (split == 2)  xform_chunk<0, 2>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
    else if (split)  xform_chunk<0, 1>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
    else             xform_chunk<0, 0>(rawc, RS, vq, P, q0, qstep, a8, b8, cvalid, dchunk, dtail);
  }
}

}  // namespace tc
