// tcgen05 implicit-GEMM convolution path (precision == 1): bf16 operands, fp32 accumulation in
// tensor memory.  precision == 2 ("bf16x3") is the same kernel with both operands split into a bf16
// head and a bf16 tail (x = hi + lo, lo = bf16(x - hi)) and three MMAs per K step,
// hi*hi + lo*hi + hi*lo, i.e. ~16 significant bits per operand: fp32-grade results (the 1e-3 parity
// bound) from the bf16 tensor pipe.  Covers the stride-1 "same" convolutions of the dense blocks, transitions and
// decoders: kernels 1x1x1, 1x3x3, 3x3x3 (hybridnet.py:264-298, 11-45, 235-260, 146-176) and the
// stride-2 7x7(x7) stems in space-to-depth form, in fprop and dgrad; everything the reference
// puts around them (BN -> Scale -> ReLU, ZeroPadding, UpSampling, Add, bias, dropout, batch
// statistics, ReLU/BN backward sums) is fused into the operand producer or the epilogue.
//
// Persistent kernel, one CTA per SM, each CTA loops over 128 x BN output tiles (128 GEMM rows =
// 16 x 8 output pixels of one (n, d) slice, or 128 consecutive positions for 1x1x1).  Every filter tap is the SAME
// shared-memory patch (tile + halo of one depth slab, one channel block) read through a descriptor whose start address is
// shifted by (th*PW + tw) pixels and whose 8-row-group stride is one patch row: no im2col gather, each input element is
// staged once per tile instead of once per tap.  The kernel is instantiated per (pass, folded bf16x3, operand path):
//
// TMA forms (every stride-1 layer; 512 threads).  The A operand was written as bf16 (head, and tail for bf16x3) by the
// pre-pass of conv_tc2_wgrad.cu -- max(a*x+b, 0) (+ second source, up-sampling in the index) on the virtual grid.
//   warp 0      one elected lane issues cp.async.bulk.tensor tile loads of the patch (zero padding, image borders and channel
//               tails = the copy engine's out-of-bounds fill) into a 3-4 stage ring
//   warp 1      weight loader: pre-packed bf16 weight blocks (core-matrix order) with cp.async.bulk, one copy and one
//               full/empty barrier pair per filter ROW (3-4 taps)
//   warp 2      MMA issuer (+ TMEM owner): all lanes run the loop converged, the tcgen05.mma of a whole filter row and its
//               tcgen05.commit are issued inside one elect.sync region (see the comment at the role)
//   warps 4-15  three epilogue sets: tcgen05.ld the finished accumulator while the next tile's MMAs run; a set takes every
//               third 32-column block
// SIMT form (the two stride-2 stems in space-to-depth form; 448 threads).
//   warps 0-7   producers: the raw fp32 patch is fetched with cp.async, transformed once (prologue, bf16 head / tail) and
//               stored in the UMMA no-swizzle K-major layout (chunk j of 8 channels at j*LBO, pixel q at q*16 B)
//   warp 8 weight loader, warp 9 MMA issuer, warps 10-13 epilogue.
#include <cuda.h>
#include <stdlib.h>
#include "hdn_common.cuh"
#include "tc_common.cuh"

// shared with the tc2 weight gradient (conv_tc2_wgrad.cu): bf16 operand pre-pass and tensor maps over its output
int hdn_tc2_make_map(CUtensorMap* tm, const void* base, int flat, long long M, int N, int D, int H, int W, int C, int bc, int bw, int bh,
                     int swizzle128);
int hdn_tc2_pack(const hdn_src* srcs, int nsrc, int N, int D, int H, int W, int C, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st,
                 int interleave);

// -DHDN_TC_TIMING: per-role wait/work cycle counters of CTA 0, printed at kernel end (development aid)
#ifdef HDN_TC_TIMING
#define TT_DECL(n) long long n = 0
#define TT_BEGIN long long tt__0 = clock64()
#define TT_ADD(n) do { long long tt__1 = clock64(); n += tt__1 - tt__0; tt__0 = tt__1; } while (0)
#else
#define TT_DECL(n)
#define TT_BEGIN
#define TT_ADD(n)
#endif

namespace {

// The folded bf16x3 issue scheme (TcParams::fold) is a run-time switch (HDN_TC_X3FOLD / hdn_set_switch).
// -DHDN_NO_FOLD / -DHDN_FPROP_SCALAR build A/B variants of the kernel without the fold code / with the scalar fprop
// epilogue (scripts/build_variants.sh; timings of the variants side by side on one box: profiles/r02n_variants.txt).
#ifdef HDN_NO_FOLD
constexpr bool kFold = false;
#else
constexpr bool kFold = true;
#endif
#ifdef HDN_FPROP_SCALAR
constexpr bool kFpropQuad = false;
#else
constexpr bool kFpropQuad = true;
#endif

constexpr int TC_THREADS = 448;   // SIMT-producer form: warps 0-7 producers, 8 weight loader, 9 MMA issuer, 10-13 epilogue
constexpr int TC_THREADS_TMA = 512;   // TMA forms: warp 0 tile loads, 1 weight loader, 2 MMA issuer, (3 idle,) 4-15 three epilogue sets
constexpr int NPROD = 256;        // producer threads
constexpr int NSA = 2;        // bf16 A-operand stages of the SIMT-producer form
constexpr int MAXNSA = 6;     // ... upper bound (TMA mode: TcParams::nsa stages)
constexpr int NSB_MAX = 16;   // weight-block ring depth upper bound
constexpr int NTAB = 4;       // geometry-table buffers (the copy front runs at most 3 stages ahead of the transform)
constexpr int UB = 2;         // direct (space-to-depth) producer: pixel groups loaded ahead per warp
constexpr int EPI_BYTES = 32 * 33 * 4 + 4 * 32 * 8;   // per epilogue warp: transpose tile + 4 row-offset tables
#ifndef HDN_FB
#define HDN_FB 4
#endif
constexpr int FB = HDN_FB;      // fprop epilogue: rows of the transpose tile a lane reads ahead of its stores
constexpr int MAXC0 = 14;     // copies per producer thread and stage, source 0 (4-tap stems' dgrad: 19x11 px x 16 quads / 256 threads)
constexpr int MAXC1 = 6;      // ... source 1 (only with 32-channel stages: 180 px x 8 quads / 256)

struct TcParams {
  CUtensorMap tmHi, tmLo;     // TMA mode: bf16 head / tail tensors of the A operand on the virtual grid (C, W, H, D, N) or flat (C, M)
  int tma;                    // 1: the A operand was pre-packed to bf16 by hdn_tc2_pack and is staged by TMA tile loads (one box
                              //    {8 channels, PW, PH} per chunk plane; out-of-bounds = the zero padding); no SIMT producers
  int nsa;                    // A-operand stages in the ring (2 in the SIMT-producer form)
  int sw;                     // TMA mode, SWIZZLE_128B form (HDN_TC_SW128): a stage is ONE box {64, PW, PH} of 128-byte pixel rows
                              //   ([head 32 | tail 32] channels in bf16x3), weights packed as swizzled 128-byte rows; K-major SW128
                              //   descriptors with SBO = PW*128 and sub-row K offsets
  int N, D, H, W;             // GEMM row grid = conv output grid = virtual (up-sampled) input grid
  int kd, kh, kw;
  int K, NC;                  // contraction channels, output columns
  int BN, KB, CK, nsb, nraw, tmem_cols;
  int ub;                     // filter taps per weight-ring unit: one bulk copy, one full / empty barrier pair (nsb % ub == 0)
  int flat;                   // 1: rows are 128 consecutive linear positions (1x1x1, no up-sampling)
  int PH, PW, P, Ppad;        // virtual patch rows, cols, pixels, padded (odd) pixel count
  int PHs[2], PWs[2], Ps[2];  // source-resolution patch of each source
  int raw_off[2], ab_off[2], raw_bytes;   // layout of one raw stage (pixel stride CK + 4 floats: bank-conflict-free reads)
  int tab_src[2], tab_vq[2], tab_ints;    // layout of one geometry table
  int tiles_w, tiles_h;
  long long tiles, total_work;
  long long M;
  int nsrc;
  hdn_src src[2];             // A operand sources
  const __nv_bfloat16* wpack;
  int mode;                   // 0 fprop, 1 dgrad
  int fold;                   // bf16x3 with BN <= 128 (HDN_TC_X3FOLD, experiment): weight chunks laid out [head | tail] along N, the
                              //   accumulator is 2*BN columns wide, A_hi x [B_hi | B_lo] is ONE MMA (+ A_lo x B_hi): 2 MMAs per step
  int l2pf;                   // 1: raw patch copies carry the L2::256B prefetch hint (HDN_TC_L2PF, experiment)
  int epi_pf;                 // 1: dgrad epilogue prefetches the next tile's stored values into L2 (HDN_TC_EPIPF, default 1)
  int fastx;                  // 1: warp-per-chunk operand transform (tc::xform_chunk) where the prologue shape allows
  int split;                  // 1: bf16x3 -- stage = 32 channels, A chunks [0,4) head / [4,8) tail, weight block = head | tail
  int tail16;                 // 1: the tails of both operands are IEEE half instead of bfloat16 (HDN_TC_TAIL16; ~19 instead of ~16 significant bits)
  int pd_lo, ph_lo, pw_lo;    // patch origin = tile origin - p*_lo (padding in front of tap 0)
  int s2d;                    // 1: A operand is the space-to-depth view of a stride-2 convolution's input:
                              //    s2d pixel (d,h,w) holds channels (rd,rh,rw,c) = x[2d+rd][2h+rh][2w+rw][c], ldc == 4
  int s2d_quads;              //    4-channel quads per s2d pixel: 8 (3-D) or 4 (2-D)
  int scatter;                // 1: dgrad epilogue writes the (rd,rh,rw,c) columns back to x's positions
  // fprop epilogue
  const float* bias;
  hdn_tensor y;
  double* stat_sum;
  double* stat_sq;
  float drop_keep;
  unsigned long long drop_seed;
  // dgrad epilogue (one per conv source)
  int nepi;
  hdn_src esrc[2];
  hdn_dgrad_epi epi[2];
};

struct TileC { int n_tile, n_img, d0, h0, w0; long long m0; };

__device__ __forceinline__ TileC tile_decode(const TcParams& p, long long w) {
  TileC t;
  const unsigned tiles = (unsigned)p.tiles;               // work items fit 32 bits (checked on the host)
  t.n_tile = (int)((unsigned)w / tiles);
  unsigned r = (unsigned)w - (unsigned)t.n_tile * tiles;
  t.n_img = t.d0 = t.h0 = t.w0 = 0; t.m0 = 0;
  if (p.flat) {
    t.m0 = (long long)r * 128;
  } else {
    const unsigned tw_ = r % (unsigned)p.tiles_w; r /= (unsigned)p.tiles_w;
    const unsigned th_ = r % (unsigned)p.tiles_h; r /= (unsigned)p.tiles_h;
    t.d0 = (int)(r % (unsigned)p.D); t.n_img = (int)(r / (unsigned)p.D);
    t.h0 = (int)th_ * 16; t.w0 = (int)tw_ * 8;
  }
  return t;
}

// Iterator over the (work item, channel block, depth slab) stages of this CTA, in issue order.
struct StageIt {
  long long w;
  int kb, dz, seq;
  bool done;
  TileC t;
};
__device__ __forceinline__ bool slab_ok(const TcParams& p, const StageIt& it) {
  const int vd = it.t.d0 - p.pd_lo + it.dz;
  return vd >= 0 && vd < p.D;
}
__device__ __forceinline__ void it_settle(const TcParams& p, StageIt& it) {   // move to the first valid stage at/after (kb, dz)
  while (!it.done) {
    if (it.dz >= p.kd) { it.dz = 0; ++it.kb; }
    if (it.kb >= p.KB) {
      it.w += gridDim.x; ++it.seq; it.kb = 0; it.dz = 0;
      if (it.w >= p.total_work) { it.done = true; break; }
      it.t = tile_decode(p, it.w);
    }
    if (slab_ok(p, it)) break;
    ++it.dz;
  }
}
__device__ __forceinline__ void it_init(const TcParams& p, StageIt& it) {
  it.w = blockIdx.x; it.kb = 0; it.dz = 0; it.seq = 0;
  it.done = it.w >= p.total_work;
  if (!it.done) { it.t = tile_decode(p, it.w); it_settle(p, it); }
}
__device__ __forceinline__ void it_next(const TcParams& p, StageIt& it) { ++it.dz; it_settle(p, it); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 prologue4(float4 x, float4 a, float4 b, int relu) {
  float4 r;
  r.x = fmaf(a.x, x.x, b.x); r.y = fmaf(a.y, x.y, b.y); r.z = fmaf(a.z, x.z, b.z); r.w = fmaf(a.w, x.w, b.w);
  if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
  return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// same copy with the L2 prefetch hint: the first 16 bytes of a pixel's 256-byte channel sliver pull the whole sliver into L2
__device__ __forceinline__ void cp_async16_l2(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait(int n) {     // n groups may remain in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
template <int NT> __device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory"); }

// Regroup the two channel quads a lane holds ([4*l8, +4) and [32 + 4*l8, +4)) with its neighbour lane into
// 8-channel chunks and store the 16-byte chunk of pixel q.
__device__ __forceinline__ void store_chunks(uint8_t* dst, uint32_t ppad, int q, bool qok, float4 v0, float4 v1, int lane, bool v1_half = false) {
  const int l8 = lane & 7;
  const bool even = (l8 & 1) == 0;
  const int chunk = even ? (l8 >> 1) : (4 + (l8 >> 1));
  const uint32_t p00 = tc::pack_bf16x2(v0.x, v0.y), p01 = tc::pack_bf16x2(v0.z, v0.w);
  const uint32_t p10 = v1_half ? tc::pack_f16x2(v1.x, v1.y) : tc::pack_bf16x2(v1.x, v1.y), p11 = v1_half ? tc::pack_f16x2(v1.z, v1.w) : tc::pack_bf16x2(v1.z, v1.w);
  const uint32_t s0 = even ? p10 : p00, s1 = even ? p11 : p01;
  const uint32_t x0 = __shfl_xor_sync(0xffffffffu, s0, 1), x1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  uint4 o;
  if (even) { o.x = p00; o.y = p01; o.z = x0; o.w = x1; }     // chunk i      = [own quad A | partner quad A]
  else      { o.x = x0; o.y = x1; o.z = p10; o.w = p11; }     // chunk 4 + i  = [partner quad B | own quad B]
  if (qok) *reinterpret_cast<uint4*>(dst + (uint32_t)chunk * ppad * 16u + (uint32_t)q * 16u) = o;
}

// Instantiated per (pass, folded bf16x3, operand path): the warp roles of one instance share a register allocation (128 per
// thread), and code that an instance never runs still costs it -- with the fold and the SIMT-producer code compiled into
// every launch the 3x3(x3) layers ran 7 % slower (profiles/r02n_variants.txt: one box, five builds side by side).
//   MODE 0 fprop, 1 dgrad;  FOLD: TcParams::fold;  OPER 0 SIMT producers (stems), 1 TMA chunk planes, 2 TMA 128-byte swizzled rows
template <int MODE, bool FOLD, int OPER>
__global__ void __launch_bounds__(OPER ? TC_THREADS_TMA : TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr bool TMA = OPER != 0;
  constexpr bool SW = OPER == 2;
  // In the TMA forms one lane of warp 0 feeds the A ring and the eight transform warps of the SIMT form are not needed: the
  // CTA has 16 warps, twelve of them epilogue warps in three sets (same TMEM lane quarters, 32-column blocks taken round
  // robin), which triples the loads / reductions the epilogue keeps in flight -- the data-gradient epilogues and the 1x1
  // forward epilogue bound their kernels (profiles/r02r_role_timing.txt; 4 -> 8 warps: dense2_x1 dgrad 1.78 -> 1.17 ms).
  constexpr int NEW = TMA ? 12 : 4;                         // epilogue warps
  constexpr int NET = NEW * 32;
  constexpr int NSET = NEW / 4;
  constexpr int W_LOAD = TMA ? 1 : 8, W_MMA = TMA ? 2 : 9, W_EPI0 = TMA ? 4 : 10;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the 128-byte swizzle is a function of the shared-memory address bits 4-9: stages start on 1024-byte boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t A_BYTES = SW ? (((uint32_t)p.P * 128u + 1023u) & ~1023u) : 8u * p.Ppad * 16u;
  const uint32_t B_HALF = (uint32_t)p.BN * (uint32_t)p.CK * 2u;          // one bf16 weight block
  const uint32_t B_BYTES = p.split ? 2u * B_HALF : B_HALF;                 // bf16x3: head block | tail block
  const int nsa = p.nsa;
  uint8_t* sA = smem;
  uint8_t* sB = sA + nsa * A_BYTES;
  uint8_t* sRaw = sB + p.nsb * B_BYTES;
  int* tabs = reinterpret_cast<int*>(sRaw + (size_t)p.nraw * p.raw_bytes);
  float* sstat = reinterpret_cast<float*>(tabs + NTAB * p.tab_ints);        // [2 buffers][4][BN]
  double* dstat = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sstat + 8 * p.BN) + 7) & ~uintptr_t(7));   // [4][BN] per-CTA sums
  uint8_t* sEpi = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dstat + 4 * p.BN) + 15) & ~uintptr_t(15));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + NEW * EPI_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + MAXNSA;
  uint64_t* b_full = a_empty + MAXNSA;
  uint64_t* b_empty = b_full + NSB_MAX;
  uint64_t* acc_full = b_empty + NSB_MAX;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (tid == 0) {
    for (int i = 0; i < MAXNSA; ++i) { tc::mbar_init(&a_full[i], TMA ? 1 : NPROD); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NSB_MAX; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], NET); }
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps_hw = p.kh * p.kw;
  const int T = p.kd * taps_hw;

  if (warp < (TMA ? 1 : 8)) {
    // =================================================================== A producers (TMA forms: warp 0 alone)
    const int ptid = tid;                                 // 0..255
    const int l8 = lane & 7, pg = lane >> 3;
    int built_seq = -1;

    // geometry tables of a tile: per source, [Ps] source-patch pixel -> element offset in the slab (or -1),
    // and [P] virtual patch pixel -> source-patch pixel index (or -1 => zero padding)
    auto build_tables = [&](const StageIt& it) {
      int* tab = tabs + (it.seq % NTAB) * p.tab_ints;
      const TileC& t = it.t;
      for (int s = 0; s < p.nsrc; ++s) {
        const hdn_src& S = p.src[s];
        const int vh0 = t.h0 - p.ph_lo, vw0 = t.w0 - p.pw_lo;
        const int sh0 = (S.uh == 2) ? (vh0 >> 1) : vh0, sw0 = (S.uw == 2) ? (vw0 >> 1) : vw0;
        for (int i = ptid; i < p.Ps[s]; i += NPROD) {
          int off = -1;
          if (p.flat) {
            if (t.m0 + i < p.M) off = i * S.t.ldc;
          } else {
            const int sh = sh0 + i / p.PWs[s], sw = sw0 + i % p.PWs[s];
            const int hlim = p.s2d ? p.H : S.H, wlim = p.s2d ? p.W : S.W;      // s2d: the patch lives on the half-resolution grid
            if (sh >= 0 && sh < hlim && sw >= 0 && sw < wlim) off = p.s2d ? (2 * sh * S.W + 2 * sw) * S.t.ldc : (sh * S.W + sw) * S.t.ldc;
          }
          tab[p.tab_src[s] + i] = off;
        }
        for (int q = ptid; q < p.P; q += NPROD) {
          int sq = -1;
          if (p.flat) {
            if (t.m0 + q < p.M) sq = q;
          } else {
            const int vh = vh0 + q / p.PW, vw = vw0 + q % p.PW;
            if (vh >= 0 && vh < p.H && vw >= 0 && vw < p.W) {
              const int sh = (S.uh == 2) ? (vh >> 1) : vh, sw = (S.uw == 2) ? (vw >> 1) : vw;
              sq = (sh - sh0) * p.PWs[s] + (sw - sw0);        // a valid virtual pixel always maps to a valid source pixel
            }
          }
          tab[p.tab_vq[s] + q] = sq;
        }
      }
      bar_producers();
    };
    auto slab_base = [&](const StageIt& it, int s) -> const float* {
      const hdn_src& S = p.src[s];
      if (p.flat) return S.t.p + it.t.m0 * S.t.ldc + S.t.coff;
      const int vd = it.t.d0 - p.pd_lo + it.dz;
      const int sd = (S.ud == 2) ? (vd >> 1) : vd;
      return S.t.p + ((long long)it.t.n_img * S.D + sd) * S.H * S.W * S.t.ldc + S.t.coff;
    };

    StageIt tr;
    it_init(p, tr);
    int sa = 0;
    uint32_t pha = 0;

    if (TMA) {
      // ---- TMA mode: the operand is already bf16 in global memory; warp 0 (converged, one elected lane issuing) stages one
      // chunk plane per box.  Zero padding, image borders and channel tails are the copy engine's out-of-bounds fill.
      if (warp == 0) {
        if (tc::elect_one_sync()) { tc::tma_prefetch_desc(&p.tmHi); if (p.split) tc::tma_prefetch_desc(&p.tmLo); }
        const uint32_t plane = (uint32_t)p.Ppad * 16u;
        const uint32_t box_bytes = (uint32_t)p.P * 16u;
        while (!tr.done) {
          tc::mbar_wait(&a_empty[sa], pha ^ 1);
          if (SW) {
            if (tc::elect_one_sync()) {
              uint8_t* dst = sA + sa * A_BYTES;
              tc::mbar_arrive_expect_tx(&a_full[sa], (uint32_t)p.P * 128u);
              if (p.flat) tc::tma_load_2d(dst, &p.tmHi, &a_full[sa], tr.kb * 64, (int)tr.t.m0);
              else tc::tma_load_5d(dst, &p.tmHi, &a_full[sa], tr.kb * 64, tr.t.w0 - p.pw_lo, tr.t.h0 - p.ph_lo, tr.t.d0 - p.pd_lo + tr.dz, tr.t.n_img);
            }
          } else
          if (tc::elect_one_sync()) {
            uint8_t* dst = sA + sa * A_BYTES;
            const int c0 = tr.kb * p.CK;
            tc::mbar_arrive_expect_tx(&a_full[sa], 8u * box_bytes);
            const int nhi = p.split ? 4 : 8;
            if (p.flat) {
              for (int j = 0; j < nhi; ++j) tc::tma_load_2d(dst + j * plane, &p.tmHi, &a_full[sa], c0 + 8 * j, (int)tr.t.m0);
              if (p.split)
                for (int j = 0; j < 4; ++j) tc::tma_load_2d(dst + (4 + j) * plane, &p.tmLo, &a_full[sa], c0 + 8 * j, (int)tr.t.m0);
            } else {
              const int aw = tr.t.w0 - p.pw_lo, ah = tr.t.h0 - p.ph_lo, ad = tr.t.d0 - p.pd_lo + tr.dz;
              for (int j = 0; j < nhi; ++j) tc::tma_load_5d(dst + j * plane, &p.tmHi, &a_full[sa], c0 + 8 * j, aw, ah, ad, tr.t.n_img);
              if (p.split)
                for (int j = 0; j < 4; ++j) tc::tma_load_5d(dst + (4 + j) * plane, &p.tmLo, &a_full[sa], c0 + 8 * j, aw, ah, ad, tr.t.n_img);
            }
          }
          __syncwarp();
          if (++sa == nsa) { sa = 0; pha ^= 1; }
          it_next(p, tr);
        }
      }
    } else if (!p.s2d) {
      // ---- asynchronous path: raw fp32 patch -> shared memory with cp.async, several stages ahead
      StageIt is;
      it_init(p, is);
      const int NQ = p.CK >> 2;                            // 16-byte quads per pixel in a raw stage (16 or 8)
      const int NQs = (p.CK == 64) ? 4 : 3;
      const int RS = p.CK + 4, RSB = RS * 4;               // raw pixel stride in floats / bytes
      // Copy list of this thread for the tile the copy front is on, kept in registers: item j moves the 16 bytes at
      // element offset goff[j] (from the slab base + channel block) to raw-slot byte (ptid + 256*j)*16.  Rebuilt once
      // per tile, so a stage costs a handful of instructions per copy.
      int goff0[MAXC0], goff1[MAXC1];
      const int part4 = (ptid & (NQ - 1)) * 4;             // this thread's channel quad inside the stage (256 % NQ == 0)
      auto issue = [&](const StageIt& it, int slot) {
        if (it.seq != built_seq) {
          build_tables(it);
          built_seq = it.seq;
          const int* tab = tabs + (it.seq % NTAB) * p.tab_ints;
          const int n0 = p.Ps[0] * NQ, n1 = p.nsrc > 1 ? p.Ps[1] * NQ : 0;
#pragma unroll
          for (int j = 0; j < MAXC0; ++j) {
            const int i = ptid + j * NPROD;
            int g = -1;
            if (i < n0) { const int o = tab[p.tab_src[0] + (i >> NQs)]; if (o >= 0) g = o + part4; }
            goff0[j] = g;
          }
#pragma unroll
          for (int j = 0; j < MAXC1; ++j) {
            const int i = ptid + j * NPROD;
            int g = -1;
            if (i < n1) { const int o = tab[p.tab_src[1] + (i >> NQs)]; if (o >= 0) g = o + part4; }
            goff1[j] = g;
          }
        }
        uint8_t* raw = sRaw + (size_t)slot * p.raw_bytes;
        const int c0 = it.kb * p.CK;
        const bool cok = c0 + part4 < p.K;
        // item j = 16 bytes of source pixel (ptid + 256 j) >> NQs, quad ptid & (NQ-1); raw pixel stride (CK + 4) floats
        const uint32_t d0 = (uint32_t)(ptid >> NQs) * (uint32_t)RSB + (uint32_t)part4 * 4u, dj = (uint32_t)(NPROD >> NQs) * (uint32_t)RSB;
        {
          const float* base = slab_base(it, 0) + c0;
          uint8_t* rs = raw + p.raw_off[0] + d0;
          if (cok && !p.l2pf) {
#pragma unroll
            for (int j = 0; j < MAXC0; ++j)
              if (goff0[j] >= 0) cp_async16(rs + j * dj, base + goff0[j]);
          } else if (cok) {
#pragma unroll
            for (int j = 0; j < MAXC0; ++j)
              if (goff0[j] >= 0) cp_async16_l2(rs + j * dj, base + goff0[j]);
          }
        }
        if (p.nsrc > 1) {
          const float* base = slab_base(it, 1) + c0;
          uint8_t* rs = raw + p.raw_off[1] + d0;
          if (cok) {
#pragma unroll
            for (int j = 0; j < MAXC1; ++j)
              if (goff1[j] >= 0) cp_async16(rs + j * dj, base + goff1[j]);
          }
        }
        if (ptid < NQ && cok) {
          for (int s = 0; s < p.nsrc; ++s) {
            const hdn_src& S = p.src[s];
            if (S.pa) cp_async16(raw + p.ab_off[s] + ptid * 16, S.pa + c0 + ptid * 4);
            if (S.pb) cp_async16(raw + p.ab_off[s] + p.CK * 4 + ptid * 16, S.pb + c0 + ptid * 4);
          }
        }
      };
      int slot_is = 0, slot_tr = 0;
      for (int i = 0; i < p.nraw - 1; ++i) {
        if (!is.done) { issue(is, slot_is); it_next(p, is); }
        cp_async_commit();
        if (++slot_is == p.nraw) slot_is = 0;
      }
      TT_DECL(t_cp); TT_DECL(t_bar); TT_DECL(t_issue); TT_DECL(t_aempty); TT_DECL(t_xform); TT_DECL(n_stage);
      while (!tr.done) {
        TT_BEGIN;
        cp_async_wait(p.nraw - 2);                         // this thread's copies of stage `tr` have landed
        TT_ADD(t_cp);
        bar_producers();                                   // ... and everyone else's; everyone is also done reading the slot refilled next
        TT_ADD(t_bar);
        if (!is.done) { issue(is, slot_is); it_next(p, is); }
        cp_async_commit();
        if (++slot_is == p.nraw) slot_is = 0;
        TT_ADD(t_issue);
        tc::mbar_wait(&a_empty[sa], pha ^ 1);
        TT_ADD(t_aempty);
        {
          const int* tab = tabs + (tr.seq % NTAB) * p.tab_ints;
          const uint8_t* raw = sRaw + (size_t)slot_tr * p.raw_bytes;
          uint8_t* dst = sA + sa * A_BYTES;
          const int c0 = tr.kb * p.CK;
          // lane = patch pixel, inner loop over the 8-channel chunks: two 16-byte reads (conflict-free thanks to the
          // padded pixel stride), BN/Scale/ReLU with warp-uniform (a, b), one 16-byte store per chunk, no shuffles.
          const float* rawf0 = reinterpret_cast<const float*>(raw + p.raw_off[0]);
          const float* rawf1 = reinterpret_cast<const float*>(raw + p.raw_off[1]);
          const float* ab0 = reinterpret_cast<const float*>(raw + p.ab_off[0]);
          const float* ab1 = reinterpret_cast<const float*>(raw + p.ab_off[1]);
          const int* vq0 = tab + p.tab_vq[0];
          const int* vq1 = tab + p.tab_vq[1];
          const bool two = p.nsrc > 1;
          const bool pa0 = p.src[0].pa != nullptr, pb0 = p.src[0].pb != nullptr, relu0 = p.src[0].relu != 0;
          const bool pa1 = two && p.src[1].pa != nullptr, pb1 = two && p.src[1].pb != nullptr, relu1 = two && p.src[1].relu != 0;
          const int nch = p.CK >> 3, nhalf = nch >> 1;           // chunks per stage (8 or 4); a work item = 32 pixels x half of them
          const int nitems = ((p.P + 31) >> 5) * 2;
          // warp-per-chunk form for the two prologue shapes that cover all single-source operands of the networks:
          // BN -> Scale -> ReLU (every dense-block / transition / decoder input) and none (dY in dgrad)
          const int xmode = (pa0 && pb0 && relu0) ? 1 : ((!pa0 && !pb0 && !relu0) ? 0 : -1);
          const int xmode1 = !two ? -1 : ((pa1 && pb1 && relu1) ? 1 : ((!pa1 && !pb1 && !relu1) ? 0 : -1));
          const bool fast2 = p.fastx >= 2 && two && xmode1 == 1 && xmode >= 0;      // shapes xform_chunk2_any takes
          if (fast2) {
            const int j = warp & (nch - 1), c = j * 8;
            tc::xform_chunk2_any(xmode, xmode1, p.split ? (p.tail16 ? 2 : 1) : 0, rawf0 + c, rawf1 + c, RS, vq0, vq1, p.P, (warp / nch) * 32 + lane,
                                 (8 / nch) * 32, ab0 + c, ab0 + p.CK + c, ab1 + c, ab1 + p.CK + c, c0 + c < p.K,
                                 dst + (uint32_t)j * (uint32_t)p.Ppad * 16u, dst + (uint32_t)(nch + j) * (uint32_t)p.Ppad * 16u);
          } else if (p.fastx && !two && xmode >= 0) {
            const int j = warp & (nch - 1), c = j * 8;          // this warp's chunk; 8 / nch warps share a chunk's pixels
            tc::xform_chunk_any(xmode, p.split ? (p.tail16 ? 2 : 1) : 0, rawf0 + c, RS, vq0, p.P, (warp / nch) * 32 + lane, (8 / nch) * 32,
                                ab0 + c, ab0 + p.CK + c, c0 + c < p.K, dst + (uint32_t)j * (uint32_t)p.Ppad * 16u,
                                dst + (uint32_t)(nch + j) * (uint32_t)p.Ppad * 16u);
          } else
          for (int item = warp; item < nitems; item += 8) {
            const int q = (item >> 1) * 32 + lane;
            const int j0 = (item & 1) * nhalf;
            const bool qok = q < p.P;
            const int sq0 = qok ? vq0[q] : -1;
            const int sq1 = (qok && two) ? vq1[q] : -1;
            const float* row0 = rawf0 + sq0 * RS;
            const float* row1 = rawf1 + sq1 * RS;
            uint8_t* drow = dst + (uint32_t)q * 16u;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if (jj >= nhalf) break;
              const int j = j0 + jj, c = j * 8;
              float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
              if (c0 + c < p.K) {                                  // warp-uniform
                if (sq0 >= 0) {                                    // zero padding is applied AFTER BN/Scale/ReLU
                  va = *reinterpret_cast<const float4*>(row0 + c);
                  vb = *reinterpret_cast<const float4*>(row0 + c + 4);
                  if (pa0 | pb0 | relu0) {
                    const float4 aa = pa0 ? *reinterpret_cast<const float4*>(ab0 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ab_ = pa0 ? *reinterpret_cast<const float4*>(ab0 + c + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ba = pb0 ? *reinterpret_cast<const float4*>(ab0 + p.CK + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bb = pb0 ? *reinterpret_cast<const float4*>(ab0 + p.CK + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    va = prologue4(va, aa, ba, relu0);
                    vb = prologue4(vb, ab_, bb, relu0);
                  }
                }
                if (sq1 >= 0) {
                  float4 ua = *reinterpret_cast<const float4*>(row1 + c), ub = *reinterpret_cast<const float4*>(row1 + c + 4);
                  if (pa1 | pb1 | relu1) {
                    const float4 aa = pa1 ? *reinterpret_cast<const float4*>(ab1 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ab_ = pa1 ? *reinterpret_cast<const float4*>(ab1 + c + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ba = pb1 ? *reinterpret_cast<const float4*>(ab1 + p.CK + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bb = pb1 ? *reinterpret_cast<const float4*>(ab1 + p.CK + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ua = prologue4(ua, aa, ba, relu1);
                    ub = prologue4(ub, ab_, bb, relu1);
                  }
                  va.x += ua.x; va.y += ua.y; va.z += ua.z; va.w += ua.w;
                  vb.x += ub.x; vb.y += ub.y; vb.z += ub.z; vb.w += ub.w;
                }
              }
              uint4 o;
              if (!p.split) {
                o.x = tc::pack_bf16x2(va.x, va.y); o.y = tc::pack_bf16x2(va.z, va.w);
                o.z = tc::pack_bf16x2(vb.x, vb.y); o.w = tc::pack_bf16x2(vb.z, vb.w);
              } else {                                             // head -> chunk j, tail -> chunk nch + j
                uint4 t;
                if (p.tail16) {
                  tc::pack_split_bf16_f16x2(va.x, va.y, o.x, t.x); tc::pack_split_bf16_f16x2(va.z, va.w, o.y, t.y);
                  tc::pack_split_bf16_f16x2(vb.x, vb.y, o.z, t.z); tc::pack_split_bf16_f16x2(vb.z, vb.w, o.w, t.w);
                } else {
                  tc::pack_split_bf16x2(va.x, va.y, o.x, t.x); tc::pack_split_bf16x2(va.z, va.w, o.y, t.y);
                  tc::pack_split_bf16x2(vb.x, vb.y, o.z, t.z); tc::pack_split_bf16x2(vb.z, vb.w, o.w, t.w);
                }
                if (qok) *reinterpret_cast<uint4*>(drow + (uint32_t)(nch + j) * (uint32_t)p.Ppad * 16u) = t;
              }
              if (qok) *reinterpret_cast<uint4*>(drow + (uint32_t)j * (uint32_t)p.Ppad * 16u) = o;
            }
          }
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&a_full[sa]);
        if (++sa == nsa) { sa = 0; pha ^= 1; }
        if (++slot_tr == p.nraw) slot_tr = 0;
        it_next(p, tr);
        TT_ADD(t_xform);
#ifdef HDN_TC_TIMING
        ++n_stage;
#endif
      }
#ifdef HDN_TC_TIMING
      if (blockIdx.x == 0 && (tid == 0 || tid == 255))
        printf("[prod t%d] stages %lld  cp_wait %lld  bar %lld  issue %lld  a_empty %lld  transform %lld (cycles/stage)\n", tid, n_stage,
               t_cp / n_stage, t_bar / n_stage, t_issue / n_stage, t_aempty / n_stage, t_xform / n_stage);
#endif
    } else {
      // ---- direct path (space-to-depth stems): register loads, UB pixel groups in flight per warp
      const bool okA = l8 < p.s2d_quads;
      const hdn_src& S = p.src[0];
      const int s2d_rd = (p.s2d_quads == 8) ? (l8 >> 2) : 0;
      const int s2d_add = ((l8 >> 1) & 1) * S.W * S.t.ldc + (l8 & 1) * 4;     // quad l8 = (rd, rh, rw): (rw, c) contiguous in x
      while (!tr.done) {
        if (tr.seq != built_seq) { build_tables(tr); built_seq = tr.seq; }
        const int* tab = tabs + (tr.seq % NTAB) * p.tab_ints;
        const int vd = tr.t.d0 - p.pd_lo + tr.dz;
        const float* base = S.t.p + ((long long)tr.t.n_img * S.D + (p.s2d_quads == 8 ? 2 * vd + s2d_rd : vd)) * S.H * S.W * S.t.ldc +
                            S.t.coff + s2d_add;
        tc::mbar_wait(&a_empty[sa], pha ^ 1);
        uint8_t* dst = sA + sa * A_BYTES;
        for (int q0 = warp * 4; q0 < p.P; q0 += 32 * UB) {
          float4 r0[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int q = q0 + u * 32 + pg;
            r0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < p.P && okA) {
              const int sq = tab[p.tab_vq[0] + q];
              const int off = sq >= 0 ? tab[p.tab_src[0] + sq] : -1;
              if (off >= 0) r0[u] = ldg4(base + off);
            }
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            if (q0 + u * 32 >= p.P) break;                  // warp-uniform
            const int q = q0 + u * 32 + pg;
            float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.split) {                                  // quad B slot (chunks 4-7) carries the tail of quad A's channels
              const float4 x = r0[u];
              r1.x = x.x - __bfloat162float(__float2bfloat16_rn(x.x)); r1.y = x.y - __bfloat162float(__float2bfloat16_rn(x.y));
              r1.z = x.z - __bfloat162float(__float2bfloat16_rn(x.z)); r1.w = x.w - __bfloat162float(__float2bfloat16_rn(x.w));
            }
            store_chunks(dst, (uint32_t)p.Ppad, q, q < p.P, r0[u], r1, lane, p.split && p.tail16);
          }
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&a_full[sa]);
        if (++sa == nsa) { sa = 0; pha ^= 1; }
        it_next(p, tr);
      }
    }
  } else if (warp == W_LOAD) {
    // =================================================================== weight loader (TMA engine)
    // all 32 lanes run the loop converged; one elected lane issues (see tc::elect_one_sync)
    {
      StageIt it;
      it_init(p, it);
      int sb = 0;
      uint32_t phb = 0;
      const size_t blk = (size_t)p.BN * p.CK * (p.split ? 2 : 1);
      const int nsu = p.nsb / p.ub;                         // ring depth in units
      TT_DECL(tl_wait); TT_DECL(tl_n);
      while (!it.done) {
        const __nv_bfloat16* src = p.wpack + (((size_t)it.t.n_tile * p.KB + it.kb) * T + (size_t)it.dz * taps_hw) * blk;
        for (int t2 = 0; t2 < taps_hw; t2 += p.ub, src += (size_t)p.ub * blk) {       // the taps of a unit are contiguous in wpack
          TT_BEGIN;
          tc::mbar_wait(&b_empty[sb], phb ^ 1);
          TT_ADD(tl_wait);
#ifdef HDN_TC_TIMING
          ++tl_n;
#endif
          if (tc::elect_one_sync()) {
            tc::mbar_arrive_expect_tx(&b_full[sb], (uint32_t)p.ub * B_BYTES);
            tc::bulk_g2s(sB + (size_t)sb * p.ub * B_BYTES, src, (uint32_t)p.ub * B_BYTES, &b_full[sb]);
          }
          __syncwarp();
          if (++sb == nsu) { sb = 0; phb ^= 1; }
        }
        it_next(p, it);
      }
#ifdef HDN_TC_TIMING
      if (blockIdx.x == 0 && lane == 0) printf("[wld] units %lld  wait-for-empty %lld (cycles/unit)  ub %d nsu %d\n", tl_n, tl_wait / max(tl_n, 1ll), p.ub, nsu);
#endif
    }
  } else if (warp == W_MMA) {
    // =================================================================== MMA issuer
    // all 32 lanes run the loop converged (waits, iterators); the tcgen05.mma / commit issue sits under elect_one_sync
    {
      const uint32_t idesc = tc::make_idesc_bf16(128, p.BN, 0, 0);
      // bf16x3 cross terms: the tail operand is bfloat16 (format 1) or IEEE half (format 0, HDN_TC_TAIL16)
      const uint32_t idesc_lh = tc::make_idesc_f16kind(128, p.BN, p.tail16 ? 0 : 1, 1, 0, 0);    // A tail x B head
      const uint32_t idesc_hl = tc::make_idesc_f16kind(128, p.BN, 1, p.tail16 ? 0 : 1, 0, 0);    // A head x B tail
      const uint32_t lbo_a = (uint32_t)p.Ppad * 16u, sbo_a = (uint32_t)p.PW * 16u;
      const uint32_t lbo_b = (uint32_t)p.BN * 16u, sbo_b = 128u;
      StageIt it;
      it_init(p, it);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int cur_seq = -1;
      uint32_t acc = 0, tmem_d = tmem_base;
      // SWIZZLE_128B form: rows (pixels / output channels) are 128 bytes, 8-row groups SBO apart, K advances inside the row
      const uint64_t sw_bit = (uint64_t)2 << 61;
      const uint64_t adesc_hi = SW ? (tc::make_smem_desc(0, 16u, (uint32_t)p.PW * 128u) | sw_bit) : tc::make_smem_desc(0, lbo_a, sbo_a);
      const uint64_t bdesc_hi = SW ? (tc::make_smem_desc(0, 16u, 1024u) | sw_bit) : tc::make_smem_desc(0, lbo_b, sbo_b);
      const uint64_t kstep_a = SW ? 2u : (uint64_t)((2u * lbo_a) >> 4), kstep_b = SW ? 2u : (uint64_t)((2u * lbo_b) >> 4);
      const uint64_t tail_a = SW ? 4u : (uint64_t)(((uint32_t)(p.CK >> 3) * lbo_a) >> 4);
      const uint64_t tail_b = SW ? 4u : (uint64_t)(((uint32_t)(p.CK >> 3) * lbo_b) >> 4);
      const uint32_t pixu = SW ? 8u : 1u;                  // 16-byte units per patch pixel
      // folded bf16x3: chunk stride of the [head | tail] weight block is 2*BN rows; N = 2*BN for A_hi, N = BN for A_lo
      const uint32_t idesc_f2 = tc::make_idesc_bf16(128, 2 * p.BN, 0, 0);
      const uint64_t bdesc_hi_f = tc::make_smem_desc(0, 2u * lbo_b, sbo_b), kstep_b_f = (uint64_t)((4u * lbo_b) >> 4);
      TT_DECL(t_acc); TT_DECL(t_afull); TT_DECL(t_bfull); TT_DECL(t_mma); TT_DECL(n_st);
      // All 32 lanes run the loop converged (waits, iterators, descriptor arithmetic: warp-uniform values that live in uniform
      // registers); the MMAs of a whole weight unit (3-4 taps) and its commit are issued inside ONE elect_one_sync region.
      // Nothing the issuing thread executes overlaps with the MMAs it has issued -- an mbarrier wait or a tcgen05.commit costs
      // the tensor pipe ~60 idle cycles each, an elect + re-convergence ~60 (scripts/micro/mma_rate2.cu) -- so those are paid
      // per unit, not per tap.  A unit is one filter row (ub == kw): consecutive taps are one pixel apart in the patch.
      const int nsu = p.nsb / p.ub;
      const uint32_t bstep = B_BYTES >> 4;                   // weight blocks of a unit are B_BYTES apart
      // the MMAs of one tap (nk K steps; bf16, folded bf16x3 or bf16x3)
      auto issue_tap = [&](uint64_t ad, uint64_t bd, uint64_t bdf, int nk, uint32_t acc0) {
        if (!p.split) {
          tc::umma_bf16(tmem_d, ad, bd, idesc, acc0);
          if (nk > 1) tc::umma_bf16(tmem_d, ad + kstep_a, bd + kstep_b, idesc, 1u);
          if (nk > 2) tc::umma_bf16(tmem_d, ad + 2 * kstep_a, bd + 2 * kstep_b, idesc, 1u);
          if (nk > 3) tc::umma_bf16(tmem_d, ad + 3 * kstep_a, bd + 3 * kstep_b, idesc, 1u);
        } else if (FOLD) {
          uint32_t a2 = acc0;
          for (int s = 0; s < nk; ++s) {
            const uint64_t ah = ad + (uint64_t)s * kstep_a, bf = bdf + (uint64_t)s * kstep_b_f;
            tc::umma_bf16(tmem_d, ah, bf, idesc_f2, a2);             // A_hi x [B_hi | B_lo] -> columns [0, 2BN)
            tc::umma_bf16(tmem_d, ah + tail_a, bf, idesc, 1u);       // A_lo x B_hi         -> columns [0, BN)
            a2 = 1;
          }
        } else {
          // bf16x3: tails first (small terms), head x head last; tail operands sit CK/8 chunks behind the heads
          uint32_t a2 = acc0;
          for (int s = 0; s < nk; ++s) {
            const uint64_t ah = ad + (uint64_t)s * kstep_a, bh = bd + (uint64_t)s * kstep_b;
            tc::umma_bf16(tmem_d, ah + tail_a, bh, idesc_lh, a2);
            tc::umma_bf16(tmem_d, ah, bh + tail_b, idesc_hl, 1u);
            tc::umma_bf16(tmem_d, ah, bh, idesc, 1u);
            a2 = 1;
          }
        }
      };
      while (!it.done) {
        TT_BEGIN;
        if (it.seq != cur_seq) {                            // first stage of a new tile: claim an accumulator buffer
          cur_seq = it.seq;
          const int ab = cur_seq & 1;
          tc::mbar_wait(&acc_empty[ab], ((uint32_t)(cur_seq >> 1) & 1u) ^ 1u);
          tc::tc_fence_after();
          tmem_d = tmem_base + (uint32_t)(ab * (FOLD ? 2 * p.BN : p.BN));
          acc = 0;
        }
        const int cv = min(p.CK, p.K - it.kb * p.CK);
        const int nk = (cv + 15) >> 4;
        TT_ADD(t_acc);
        tc::mbar_wait(&a_full[sa], pha);
        tc::tc_fence_after();
        TT_ADD(t_afull);
        // Descriptors differ only in the 14-bit start-address field (16-byte units, < 2^14 for 227 KB of shared memory), so
        // they are advanced by integer adds on the low word.
        const uint64_t a_desc0 = adesc_hi | (uint64_t)((tc::smem_u32(sA + sa * A_BYTES) >> 4) & 0x3FFF);
        uint32_t tap_units = 0;                              // (th * PW + tw) in 16-byte units
        int twc = 0;
        for (int t2 = 0; t2 < taps_hw; t2 += p.ub) {
          tc::mbar_wait(&b_full[sb], phb);
          tc::tc_fence_after();
          TT_ADD(t_bfull);
          const uint32_t sbu = (tc::smem_u32(sB + (size_t)sb * p.ub * B_BYTES) >> 4) & 0x3FFF;
          const uint64_t ad = a_desc0 + (uint64_t)(tap_units * pixu);
          const uint64_t bd = bdesc_hi | (uint64_t)sbu, bdf = bdesc_hi_f | (uint64_t)sbu;
          if (tc::elect_one_sync()) {
            issue_tap(ad, bd, bdf, nk, acc);
            if (p.ub > 1) issue_tap(ad + pixu, bd + bstep, bdf + bstep, nk, 1u);
            if (p.ub > 2) issue_tap(ad + 2 * pixu, bd + 2 * bstep, bdf + 2 * bstep, nk, 1u);
            if (p.ub > 3) issue_tap(ad + 3 * pixu, bd + 3 * bstep, bdf + 3 * bstep, nk, 1u);
            tc::umma_commit(&b_empty[sb]);                 // unit consumed: one commit releases its ring slot
          }
          __syncwarp();
          acc = 1;
          if (++sb == nsu) { sb = 0; phb ^= 1; }
          if (p.ub == p.kw) tap_units += (uint32_t)p.PW;       // next filter row
          else if (++twc == p.kw) { twc = 0; tap_units += (uint32_t)(p.PW - p.kw + 1); } else ++tap_units;   // single-tap units
          TT_ADD(t_mma);
        }
        const int seq_before = it.seq;
        const int sa_done = sa;
        if (++sa == nsa) { sa = 0; pha ^= 1; }
        it_next(p, it);
        if (tc::elect_one_sync()) {
          tc::umma_commit(&a_empty[sa_done]);
          if (it.done || it.seq != seq_before) tc::umma_commit(&acc_full[seq_before & 1]);   // tile finished
        }
        __syncwarp();
        TT_ADD(t_mma);
#ifdef HDN_TC_TIMING
        ++n_st;
#endif
      }
#ifdef HDN_TC_TIMING
      if (blockIdx.x == 0 && lane == 0)
        printf("[mma] stages %lld  acc_empty %lld  a_full %lld  b_full %lld  issue %lld (cycles/stage)\n", n_st, t_acc / n_st, t_afull / n_st,
               t_bfull / n_st, t_mma / n_st);
#endif
    }
  } else if (warp >= W_EPI0) {
    // =================================================================== epilogue (warps 10-13, or 4-15 in the TMA forms)
    const int ew = warp - W_EPI0;                           // 0..NEW-1
    const int eset = ew >> 2;                               // set s takes the 32-column blocks s, s + NSET, ...
    const int etid = ew * 32 + lane;                        // 0..NET-1
    uint8_t* escr = sEpi + (size_t)ew * EPI_BYTES;          // this warp's transpose tile + row-offset tables
    float* tT = reinterpret_cast<float*>(escr);             // [32][33]
    long long* ro0 = reinterpret_cast<long long*>(escr + 32 * 33 * 4);   // [32]  output row offsets (fprop)
    long long* rox = ro0;                                    // [2][32] stored-value row offsets (dgrad)
    long long* rod = ro0 + 64;                               // [2][32] gradient row offsets (dgrad)
    const int qtr = warp & 3;                               // TMEM lane quarter this warp may read
    const int row = qtr * 32 + lane;
    const int hr = row >> 3, wr = row & 7;
    int seq = 0;
    float v[16];
#ifdef HDN_TC_TIMING
    long long te_wait = 0, te_work = 0;
#endif
    // Per-channel sums (batch statistics in fprop, S1 / S2 in dgrad) are accumulated per CTA in shared-memory doubles, each
    // column owned by one epilogue thread, and reach global memory ONCE per CTA and column tile: a double atomic per tile
    // and column from 148 CTAs onto the same few addresses serialises in L2 (16 384 tiles x 128 atomics for the 64-wide
    // decoder tail) and throttled the whole pipeline.
    for (int i = etid; i < 4 * p.BN; i += NET) dstat[i] = 0.0;
    int flush_ntile = -1;
    auto flush_stats = [&](int nt) {
      if (MODE == 0) {
        for (int c = etid; c < p.BN; c += NET) {
          const int col = nt * p.BN + c;
          if (col < p.NC) { atomicAdd(p.stat_sum + col, dstat[c]); atomicAdd(p.stat_sq + col, dstat[p.BN + c]); }
          dstat[c] = 0.0; dstat[p.BN + c] = 0.0;
        }
      } else {
        for (int e = 0; e < p.nepi; ++e) {
          const hdn_dgrad_epi& E = p.epi[e];
          if (E.mode == 2 || E.s1 == nullptr) continue;
          for (int c = etid; c < p.BN; c += NET) {
            const int col = nt * p.BN + c;
            const double s1 = dstat[(2 * e) * p.BN + c], s2 = dstat[(2 * e + 1) * p.BN + c];
            if (col < p.NC) {
              if (s1 != 0.0) atomicAdd(E.s1 + col, s1);
              if (s2 != 0.0) atomicAdd(E.s2 + col, s2);
            }
            dstat[(2 * e) * p.BN + c] = 0.0; dstat[(2 * e + 1) * p.BN + c] = 0.0;
          }
        }
      }
    };
    for (long long w = blockIdx.x; w < p.total_work; w += gridDim.x, ++seq) {
      const TileC t = tile_decode(p, w);
      const int ab = seq & 1;
      float* st_ = sstat + ab * 4 * p.BN;
      const int n_tile = t.n_tile;
      if (flush_ntile >= 0 && flush_ntile != n_tile) { flush_stats(flush_ntile); flush_ntile = -1; }
      bool rvalid;
      long long m;
      int oh = 0, ow = 0;
      if (p.flat) {
        m = t.m0 + row;
        rvalid = m < p.M;
      } else {
        oh = t.h0 + hr; ow = t.w0 + wr;
        rvalid = oh < p.H && ow < p.W;
        m = (((long long)t.n_img * p.D + t.d0) * p.H + oh) * p.W + ow;
      }
      bool any_s = false;
      if (MODE == 0) any_s = p.stat_sum != nullptr;
      else if (!p.scatter)
        for (int e = 0; e < p.nepi; ++e) any_s = any_s || (p.epi[e].mode != 2 && p.epi[e].s1 != nullptr);
      if (any_s) {
        for (int i = etid; i < 4 * p.BN; i += NET) st_[i] = 0.f;
        bar_epilogue<NET>();
      }
#ifdef HDN_TC_TIMING
      long long te0 = clock64();
#endif
      tc::mbar_wait_sleep(&acc_full[ab], (uint32_t)(seq >> 1) & 1u);
      tc::tc_fence_after();
#ifdef HDN_TC_TIMING
      te_wait += clock64() - te0; te0 = clock64();
#endif
      const uint32_t taddr = tmem_base + ((uint32_t)(qtr * 32) << 16) + (uint32_t)(ab * (FOLD ? 2 * p.BN : p.BN));
      // accumulator columns [col, col+16) of this lane's row; folded bf16x3 adds the A_hi x B_lo half BN columns further
      auto ld_acc = [&](int col, float* dst) {
        tc::tmem_ld16(taddr + (uint32_t)col, dst);
        if (FOLD) {
          float w2[16];
          tc::tmem_ld16(taddr + (uint32_t)(p.BN + col), w2);
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i] += w2[i];
        }
      };

      if (MODE == 0) {
        // Transposed epilogue: the warp's 32 x 32 accumulator block goes through shared memory so that lane = output
        // channel and the 32 lanes write 128 contiguous bytes of one output pixel; the per-channel batch statistics
        // become plain per-lane running sums (no shuffles).
        const bool do_stats = p.stat_sum != nullptr;
        const bool has_bias = p.bias != nullptr, drop = p.drop_keep < 1.0f;
        float* ybase = const_cast<float*>(p.y.p) + p.y.coff;
        // 16-byte stores need quads that are all-in or all-out and 16-byte aligned rows
        const bool vec_ok = kFpropQuad && ((p.NC | p.BN | p.y.ldc | p.y.coff) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y.p) & 15) == 0;
        ro0[lane] = m * (long long)p.y.ldc;
        rod[lane] = m;                                      // linear position (dropout hash index)
        const unsigned vmask = __ballot_sync(0xffffffffu, rvalid);
        __syncwarp();
        for (int cb = eset * 32; cb < p.BN; cb += 32 * NSET) {
          const int ncols = min(32, p.BN - cb);
          float v2[16];
          ld_acc(cb, v);
          if (ncols > 16) ld_acc(cb + 16, v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) tT[lane * 33 + i] = v[i];
          if (ncols > 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tT[lane * 33 + 16 + i] = v2[i];
          }
          __syncwarp();
          if (vec_ok) {
            // Quad form (as in the data-gradient epilogue): lane = (row sub-index, channel quad).  The four rows of a batch
            // are read from the transpose tile BEFORE the first store is issued -- a store through the generic output
            // pointer orders every later shared-memory load behind it, which made the scalar form pay a full LDS round trip
            // per output row (78 cycles per row, 15k cycles per 128 x 192 tile: profiles/r02j_role_timing.txt) -- and a row
            // leaves as one 16-byte store per lane.
            const int rsub = lane >> 3, cq = lane & 7;
            const int col4 = n_tile * p.BN + cb + 4 * cq;
            const bool qok = 4 * cq < ncols && col4 < p.NC;
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_bias && qok) bias4 = make_float4(__ldg(p.bias + col4), __ldg(p.bias + col4 + 1), __ldg(p.bias + col4 + 2), __ldg(p.bias + col4 + 3));
            float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
            for (int it0 = 0; it0 < 8; it0 += FB) {
              float4 tv[FB];
              long long ro[FB], rp[FB];
#pragma unroll
              for (int k = 0; k < FB; ++k) {
                const int rr = 4 * (it0 + k) + rsub;
                const float* tp = tT + rr * 33 + 4 * cq;
                tv[k] = make_float4(tp[0], tp[1], tp[2], tp[3]);
                ro[k] = ro0[rr];
                rp[k] = rod[rr];
              }
#pragma unroll
              for (int k = 0; k < FB; ++k) {
                const int rr = 4 * (it0 + k) + rsub;
                if (!(qok && ((vmask >> rr) & 1u))) continue;
                float4 t4 = make_float4(tv[k].x + bias4.x, tv[k].y + bias4.y, tv[k].z + bias4.z, tv[k].w + bias4.w);
                if (drop) {
                  const uint64_t di = (uint64_t)rp[k] * p.NC + col4;
                  t4.x *= hdn_drop_scale(p.drop_seed, di, p.drop_keep);
                  t4.y *= hdn_drop_scale(p.drop_seed, di + 1, p.drop_keep);
                  t4.z *= hdn_drop_scale(p.drop_seed, di + 2, p.drop_keep);
                  t4.w *= hdn_drop_scale(p.drop_seed, di + 3, p.drop_keep);
                }
                *reinterpret_cast<float4*>(ybase + ro[k] + col4) = t4;
                s1.x += t4.x; s1.y += t4.y; s1.z += t4.z; s1.w += t4.w;
                s2.x += t4.x * t4.x; s2.y += t4.y * t4.y; s2.z += t4.z * t4.z; s2.w += t4.w * t4.w;
              }
            }
            if (do_stats) {
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
                s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
                s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
                s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
                s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
              }
              if (rsub == 0 && qok) {
                float* p1 = &st_[cb + 4 * cq];
                float* p2 = &st_[p.BN + cb + 4 * cq];
                atomicAdd(p1, s1.x); atomicAdd(p1 + 1, s1.y); atomicAdd(p1 + 2, s1.z); atomicAdd(p1 + 3, s1.w);
                atomicAdd(p2, s2.x); atomicAdd(p2 + 1, s2.y); atomicAdd(p2 + 2, s2.z); atomicAdd(p2 + 3, s2.w);
              }
            }
          } else {
            const int col = n_tile * p.BN + cb + lane;
            const bool cok = lane < ncols && col < p.NC;
            const float bias = (has_bias && cok) ? __ldg(p.bias + col) : 0.f;
            float ssum = 0.f, ssq = 0.f;
            if (cok) {
  #pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                if (!((vmask >> rr) & 1u)) continue;             // warp-uniform
                float tv = tT[rr * 33 + lane] + bias;
                const long long ro = ro0[rr];
                if (drop) tv *= hdn_drop_scale(p.drop_seed, (uint64_t)rod[rr] * p.NC + col, p.drop_keep);
                ybase[ro + col] = tv;
                ssum += tv; ssq += tv * tv;
              }
              if (do_stats) { atomicAdd(&st_[cb + lane], ssum); atomicAdd(&st_[p.BN + cb + lane], ssq); }
            }
          }
          __syncwarp();
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&acc_empty[ab]);                   // accumulator buffer may be overwritten
        if (do_stats) {
          bar_epilogue<NET>();
          for (int c = etid; c < p.BN; c += NET) { dstat[c] += (double)st_[c]; dstat[p.BN + c] += (double)st_[p.BN + c]; }
          flush_ntile = n_tile;
        }
      } else if (p.scatter) {
        // dgrad of a stride-2 stem in space-to-depth form: column (rd,rh,rw,c) of s2d pixel (d0,oh,ow) is the
        // gradient of x[2*d0+rd][2*oh+rh][2*ow+rw][c]; the input has no prologue (hybridnet.py:122-123,208-209)
        const hdn_dgrad_epi& E = p.epi[0];
        const hdn_src& S = p.esrc[0];
        for (int cc = 0; cc < p.BN; cc += 16) {
          ld_acc(cc, v);
          if (!rvalid) continue;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int qd = (n_tile * p.BN + cc + i) >> 2;
            if (qd >= p.s2d_quads) continue;
            const int rd = (p.s2d_quads == 8) ? (qd >> 2) : 0, rh = (qd >> 1) & 1, rw = qd & 1;
            const int sd = (p.s2d_quads == 8) ? (2 * t.d0 + rd) : t.d0;
            const long long ms = (((long long)t.n_img * S.D + sd) * S.H + 2 * oh + rh) * S.W + 2 * ow + rw;
            float* q = const_cast<float*>(E.dx.p) + ms * E.dx.ldc + E.dx.coff;
            float4 g = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (E.accumulate) { float4 o = *reinterpret_cast<float4*>(q); g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
            *reinterpret_cast<float4*>(q) = g;
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&acc_empty[ab]);
      } else {
        // dgrad, transposed like fprop: lane = input channel.  For each source and each writer row: dz (summed over the
        // up-sampling sub-positions of the row), ReLU mask from the stored value, per-lane S1/S2 running sums, then
        // dx (+)= a*du or du (+)= du as coalesced 128-byte row accesses.
        unsigned wmask[2] = {0u, 0u};
        for (int e = 0; e < p.nepi; ++e) {
          if (p.epi[e].mode == 2) continue;
          const hdn_src& S = p.esrc[e];
          const hdn_dgrad_epi& E = p.epi[e];
          long long ms = m;
          bool writer = rvalid;
          if (!p.flat) {
            const int sd = (S.ud == 2) ? (t.d0 >> 1) : t.d0, sh = (S.uh == 2) ? (oh >> 1) : oh, sw = (S.uw == 2) ? (ow >> 1) : ow;
            ms = (((long long)t.n_img * S.D + sd) * S.H + sh) * S.W + sw;
            if (S.uw == 2 && (wr & 1)) writer = false;
            if (S.uh == 2 && (hr & 1)) writer = false;
          }
          rox[e * 32 + lane] = ms * (long long)S.t.ldc + S.t.coff;
          rod[e * 32 + lane] = E.mode == 0 ? ms * (long long)E.dx.ldc + E.dx.coff : ms * (long long)p.NC;
          wmask[e] = __ballot_sync(0xffffffffu, writer);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, rvalid);
        __syncwarp();
        if (p.epi_pf && eset == 0) {
          // The stored values this epilogue reads were written a whole forward pass ago: they come from DRAM, and the few
          // 16-byte loads a lane keeps in flight make the epilogue latency-bound.  Pull the NEXT tile's rows into L2 now
          // (one row per thread, a prefetch per 128-byte line): by the time that tile's accumulator is ready they are L2 hits.
          const long long wn = w + gridDim.x;
          if (wn < p.total_work) {
            const TileC tn = tile_decode(p, wn);
            const int ncb = min(p.BN, p.NC - tn.n_tile * p.BN) * 4;            // bytes of this column tile per row
            for (int e = 0; e < p.nepi; ++e) {
              if (p.epi[e].mode == 2) continue;
              const hdn_src& S = p.esrc[e];
              long long ms;
              bool wr_ = true;
              if (p.flat) { ms = tn.m0 + row; wr_ = ms < p.M; }
              else {
                const int ohn = tn.h0 + hr, own = tn.w0 + wr;
                wr_ = ohn < p.H && own < p.W && !(S.uw == 2 && (wr & 1)) && !(S.uh == 2 && (hr & 1));
                const int sd = (S.ud == 2) ? (tn.d0 >> 1) : tn.d0, sh = (S.uh == 2) ? (ohn >> 1) : ohn, sw = (S.uw == 2) ? (own >> 1) : own;
                ms = (((long long)tn.n_img * S.D + sd) * S.H + sh) * S.W + sw;
              }
              if (wr_) {
                const char* px = reinterpret_cast<const char*>(S.t.p + ms * (long long)S.t.ldc + S.t.coff + tn.n_tile * p.BN);
                for (int o = 0; o < ncb; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
              }
            }
          }
        }
        for (int cb = eset * 32; cb < p.BN; cb += 32 * NSET) {
          const int ncols = min(32, p.BN - cb);
          float v2[16];
          ld_acc(cb, v);
          if (ncols > 16) ld_acc(cb + 16, v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) tT[lane * 33 + i] = rvalid ? v[i] : 0.f;
          if (ncols > 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tT[lane * 33 + 16 + i] = rvalid ? v2[i] : 0.f;
          }
          __syncwarp();
          for (int e = 0; e < p.nepi; ++e) {
            const hdn_dgrad_epi& E = p.epi[e];
            if (E.mode == 2) continue;
            const hdn_src& S = p.esrc[e];
            {
              // Quad form: lane = (row sub-index, channel quad).  A lane owns 4 consecutive channels of rows 4*it + rsub:
              // the stored values arrive as 16-byte loads, the gradient leaves as ONE 16-byte store or vector reduction
              // per row (red.global.add.v4.f32) instead of four 4-byte REDs -- the dense blocks' 1x1 data gradients issue
              // ~9.5 G such elements per step and were bound by the SM's reduction issue rate, not by HBM.
              const int rsub = lane >> 3, cq = lane & 7;
              const int col4 = n_tile * p.BN + cb + 4 * cq;
              const bool qok = 4 * cq < ncols && col4 < p.NC;          // NC % 8 == 0: a quad is all-in or all-out
              float4 a4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = b4;
              if (qok) {
                if (S.pa) a4 = ldg4(S.pa + col4);
                if (S.pb) b4 = ldg4(S.pb + col4);
                if (E.s1 && E.center) c4 = ldg4(E.center + col4);
              }
              const bool up_w = S.uw == 2, up_h = S.uh == 2, atom = S.ud == 2, acc = E.accumulate != 0, relu = S.relu != 0;
              const float* xb = S.t.p;
              float* db = E.mode == 0 ? const_cast<float*>(E.dx.p) : E.du;
              const float4 ga = E.mode == 0 ? a4 : make_float4(1.f, 1.f, 1.f, 1.f);
              const long long* rx = rox + e * 32;
              const long long* rd = rod + e * 32;
              const unsigned wm = wmask[e];
              float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
              // An accumulating write is a fire-and-forget vector reduction (executes in L2, no read latency on the SM side).  The
              // read-modify-write alternative (load old gradient, add, store) was measured slower on every layer -- dense2_x1
              // 1.45 -> 2.32 ms, fianl_conv 2.39 -> 2.80 ms (profiles/r02l_dgrad_red_vs_rmw.txt): the extra 16-byte load per
              // row doubles what a lane has to keep in flight.
              // two batches of 4 rows: the 4 stored-value loads of a batch are in flight together
              for (int it0 = 0; it0 < 8; it0 += 4) {
                float4 xs[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int rr = 4 * (it0 + k) + rsub;
                  xs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (qok && ((wm >> rr) & 1u)) xs[k] = ldg4(xb + rx[rr] + col4);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int rr = 4 * (it0 + k) + rsub;
                  if (!(qok && ((wm >> rr) & 1u))) continue;
                  const float* tp = tT + rr * 33 + 4 * cq;
                  float4 dz = make_float4(tp[0], tp[1], tp[2], tp[3]);
                  if (up_w) { const float* t1 = tT + (rr ^ 1) * 33 + 4 * cq; dz.x += t1[0]; dz.y += t1[1]; dz.z += t1[2]; dz.w += t1[3]; }
                  if (up_h) {
                    const float* t8 = tT + (rr ^ 8) * 33 + 4 * cq; dz.x += t8[0]; dz.y += t8[1]; dz.z += t8[2]; dz.w += t8[3];
                    if (up_w) { const float* t9 = tT + (rr ^ 9) * 33 + 4 * cq; dz.x += t9[0]; dz.y += t9[1]; dz.z += t9[2]; dz.w += t9[3]; }
                  }
                  const float4 x = xs[k];
                  if (relu) {
                    if (!(fmaf(a4.x, x.x, b4.x) > 0.f)) dz.x = 0.f;
                    if (!(fmaf(a4.y, x.y, b4.y) > 0.f)) dz.y = 0.f;
                    if (!(fmaf(a4.z, x.z, b4.z) > 0.f)) dz.z = 0.f;
                    if (!(fmaf(a4.w, x.w, b4.w) > 0.f)) dz.w = 0.f;
                  }
                  s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
                  s2.x += dz.x * (x.x - c4.x); s2.y += dz.y * (x.y - c4.y); s2.z += dz.z * (x.z - c4.z); s2.w += dz.w * (x.w - c4.w);
                  float* q = db + rd[rr] + col4;
                  const float4 g = make_float4(ga.x * dz.x, ga.y * dz.y, ga.z * dz.z, ga.w * dz.w);
                  if (atom | acc)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w) : "memory");
                  else
                    *reinterpret_cast<float4*>(q) = g;
                }
              }
              if (E.s1) {                                                // (warp-uniform) fold the 4 row sub-indices of a quad
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                  s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
                  s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
                  s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
                  s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
                }
                if (rsub == 0 && qok) {
                  float* p1 = &st_[(2 * e) * p.BN + cb + 4 * cq];
                  float* p2 = &st_[(2 * e + 1) * p.BN + cb + 4 * cq];
                  atomicAdd(p1, s1.x); atomicAdd(p1 + 1, s1.y); atomicAdd(p1 + 2, s1.z); atomicAdd(p1 + 3, s1.w);
                  atomicAdd(p2, s2.x); atomicAdd(p2 + 1, s2.y); atomicAdd(p2 + 2, s2.z); atomicAdd(p2 + 3, s2.w);
                }
              }
            }
          }
          __syncwarp();
        }
        (void)vmask;
        tc::tc_fence_before();
        tc::mbar_arrive(&acc_empty[ab]);
        if (any_s) {
          bar_epilogue<NET>();
          for (int e = 0; e < p.nepi; ++e) {
            const hdn_dgrad_epi& E = p.epi[e];
            if (E.mode == 2 || E.s1 == nullptr) continue;
            for (int c = etid; c < p.BN; c += NET) {
              dstat[(2 * e) * p.BN + c] += (double)st_[(2 * e) * p.BN + c];
              dstat[(2 * e + 1) * p.BN + c] += (double)st_[(2 * e + 1) * p.BN + c];
            }
          }
          flush_ntile = n_tile;
        }
      }
#ifdef HDN_TC_TIMING
      te_work += clock64() - te0;
#endif
    }
    if (flush_ntile >= 0) flush_stats(flush_ntile);
#ifdef HDN_TC_TIMING
    if (blockIdx.x == 0 && etid == 0) printf("[epi] tiles %d  wait %lld  work %lld (cycles/tile)\n", seq, te_wait / max(seq, 1), te_work / max(seq, 1));
#endif
  }

  // ---- teardown
  tc::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------- weight packing
// out[n_tile][kb][tap][part < nsplit][chunk j < KC/8][n][8]  (bf16)  =  B[n = column][k = kb*KC + j*8 + e]
//   part 0 = bf16(B), part 1 (bf16x3 only) = bf16(B - part 0)
//   role 0 (fprop): B[n][k] = w[tap][k][n]                       K = Cin,  NC = Cout
//   role 1 (dgrad): B[n][k] = w[flip(tap)][n][k]                 K = Cout, NC = Cin
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                           int Cin, int Cout, int kd, int kh, int kw, int BN, int KB,
                                                           int KC, int role, int nsplit, int fold, int tail16, long long total16) {
  // one thread per 16-byte output unit (n, chunk j): 8 consecutive k.  role 0 reads w[tap][k][col] (threads
  // adjacent in n -> coalesced over col); role 1 reads w[tap'][col][k..k+8) (two float4 per thread).
  const int T = kd * kh * kw;
  const int K = role == 0 ? Cin : Cout, NC = role == 0 ? Cout : Cin;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total16; idx += (long long)gridDim.x * 256) {
    long long t = idx;
    const int n = (int)(t % BN); t /= BN;
    const int nj = KC / 8;
    int j, part;                         // fold: [chunk j][part][n] (head and tail rows side by side along N), else [part][chunk j][n]
    if (fold) { part = (int)(t % nsplit); t /= nsplit; j = (int)(t % nj); t /= nj; }
    else      { j = (int)(t % nj); t /= nj; part = (int)(t % nsplit); t /= nsplit; }
    const int tap = (int)(t % T); t /= T;
    const int kb = (int)(t % KB); t /= KB;
    const int nt = (int)t;
    const int k0 = kb * KC + j * 8, col = nt * BN + n;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (col < NC && k0 < K) {          // K % 8 == 0: a chunk is all-in or all-out
      if (role == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(w + ((long long)tap * Cin + k0 + e) * Cout + col);
      } else {
        const float* src = w + ((long long)(T - 1 - tap) * Cin + col) * Cout + k0;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
      }
    }
    if (part) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] -= __bfloat162float(__float2bfloat16_rn(v[e]));
    }
    uint4 o;
    if (part && tail16) {
      o.x = tc::pack_f16x2(v[0], v[1]); o.y = tc::pack_f16x2(v[2], v[3]);
      o.z = tc::pack_f16x2(v[4], v[5]); o.w = tc::pack_f16x2(v[6], v[7]);
    } else {
      o.x = tc::pack_bf16x2(v[0], v[1]); o.y = tc::pack_bf16x2(v[2], v[3]);
      o.z = tc::pack_bf16x2(v[4], v[5]); o.w = tc::pack_bf16x2(v[6], v[7]);
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
  }
}

// SWIZZLE_128B form of the packed weights: out[n_tile][kb][tap][n < BN][64 k] as 128-byte rows whose 16-byte chunks are
// XOR-swizzled with the row index (chunk c of row n sits at (c ^ (n & 7))), the layout a K-major SWIZZLE_128B descriptor
// reads when the block starts on a 1024-byte boundary.  bf16x3: k = [head of channels kb*32 .. +32 | tail of the same].
__global__ void __launch_bounds__(256) pack_weights_sw_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cout,
                                                              int kd, int kh, int kw, int BN, int KB, int role, int split, long long total16) {
  const int T = kd * kh * kw;
  const int K = role == 0 ? Cin : Cout, NC = role == 0 ? Cout : Cin;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total16; idx += (long long)gridDim.x * 256) {
    long long t = idx;
    const int c = (int)(t & 7); t >>= 3;
    const int n = (int)(t % BN); t /= BN;
    const int tap = (int)(t % T); t /= T;
    const int kb = (int)(t % KB); t /= KB;
    const int nt = (int)t;
    const int part = split ? (c >> 2) : 0;
    const int k0 = split ? kb * 32 + (c & 3) * 8 : kb * 64 + c * 8;
    const int col = nt * BN + n;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (col < NC && k0 < K) {
      if (role == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(w + ((long long)tap * Cin + k0 + e) * Cout + col);
      } else {
        const float* src = w + ((long long)(T - 1 - tap) * Cin + col) * Cout + k0;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
      }
    }
    if (part) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] -= __bfloat162float(__float2bfloat16_rn(v[e]));
    }
    uint4 o;
    o.x = tc::pack_bf16x2(v[0], v[1]); o.y = tc::pack_bf16x2(v[2], v[3]);
    o.z = tc::pack_bf16x2(v[4], v[5]); o.w = tc::pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(out)[(idx & ~7ll) + (long long)(c ^ (n & 7))] = o;
  }
}

// Stride-2 stems in space-to-depth form (7-tap kernel, pad 3, stride 2 == 4-tap kernel over the s2d input):
// s2d tap tq (0..3) and sub-position r (0/1) read original tap t = 2*tq + r - 1 (t = -1 does not exist -> 0).
//   role 0 (fprop): B[n = co][k = (rd,rh,rw,c)] = w[t(tq,r)][c][co]
//   role 1 (dgrad): B[n = (rd,rh,rw,c)][k = co] = w[t(3 - tq', r)][c][co]   (taps flipped)
__global__ void __launch_bounds__(256) pack_weights_s2d_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                               int Cin, int Cout, int three_d, int BN, int KB, int KC,
                                                               int role, int nsplit, int fold, int tail16, long long total16) {
  const int TD = three_d ? 4 : 1, T = TD * 16;
  const int quads = three_d ? 8 : 4;
  const int K = role == 0 ? quads * 4 : Cout, NC = role == 0 ? Cout : quads * 4;
  const int k7d = three_d ? 7 : 1;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total16; idx += (long long)gridDim.x * 256) {
    long long t = idx;
    const int n = (int)(t % BN); t /= BN;
    const int nj = KC / 8;
    int j, part;
    if (fold) { part = (int)(t % nsplit); t /= nsplit; j = (int)(t % nj); t /= nj; }
    else      { j = (int)(t % nj); t /= nj; part = (int)(t % nsplit); t /= nsplit; }
    int tap = (int)(t % T); t /= T;
    const int kb = (int)(t % KB); t /= KB;
    const int nt = (int)t;
    if (role == 1) tap = T - 1 - tap;
    const int tqw = tap & 3, tqh = (tap >> 2) & 3, tqd = tap >> 4;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kb * KC + j * 8 + e, col = nt * BN + n;
      v[e] = 0.f;
      if (k < K && col < NC) {
        const int sc = role == 0 ? k : col, co = role == 0 ? col : k;    // s2d channel (rd,rh,rw,c), output channel
        const int c = sc & 3, rw = (sc >> 2) & 1, rh = (sc >> 3) & 1, rd = (sc >> 4) & 1;
        const int tw = 2 * tqw + rw - 1, th = 2 * tqh + rh - 1, td = three_d ? (2 * tqd + rd - 1) : 0;
        if (c < Cin && tw >= 0 && th >= 0 && td >= 0 && tw < 7 && th < 7 && td < k7d)
          v[e] = __ldg(w + ((((long long)td * 7 + th) * 7 + tw) * Cin + c) * Cout + co);
      }
      if (part) v[e] -= __bfloat162float(__float2bfloat16_rn(v[e]));
    }
    uint4 o;
    if (part && tail16) {
      o.x = tc::pack_f16x2(v[0], v[1]); o.y = tc::pack_f16x2(v[2], v[3]);
      o.z = tc::pack_f16x2(v[4], v[5]); o.w = tc::pack_f16x2(v[6], v[7]);
    } else {
      o.x = tc::pack_bf16x2(v[0], v[1]); o.y = tc::pack_bf16x2(v[2], v[3]);
      o.z = tc::pack_bf16x2(v[4], v[5]); o.w = tc::pack_bf16x2(v[6], v[7]);
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
  }
}

__global__ void __launch_bounds__(256) zero_window_kernel(hdn_tensor t, long long M, int C) {
  const long long total = M * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long m = i / C;
    const int c = (int)(i - m * C);
    const_cast<float*>(t.p)[m * t.ldc + t.coff + c] = 0.f;
  }
}

}  // namespace

// HDN_TC_FASTX=0/1/2 selects the operand-transform form of the tcgen05 kernels (read once per process; default 2,
// validated on B200 in round 1: profiles/r01b_*, r01c_*)
int hdn_tc_fastx() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDN_TC_FASTX");
    v = e ? atoi(e) : 2;                  // 0 generic, 1 single-source warp-per-chunk form, 2 also the two-source form
    if (v < 0 || v > 2) v = 2;
  }
  return v;
}

// HDN_TC_X3FOLD (default 1): folded bf16x3 (A_hi x [B_hi | B_lo] as one MMA of N = 2*BN, + A_lo x B_hi): 2 MMAs and 14 KB of
// shared-memory operand reads per K step instead of 3 MMAs and 18 KB (BN = 64); see TcParams::fold.  Also hdn_set_switch().
static int g_tc_x3fold = -1;
int hdn_tc_x3fold() {
  if (g_tc_x3fold < 0) {
    const char* e = getenv("HDN_TC_X3FOLD");
    g_tc_x3fold = (e && atoi(e) == 0) ? 0 : 1;            // default on: forward launches 3-6 % faster, data gradients equal (profiles/r02u_*)
  }
  return g_tc_x3fold;
}
void hdn_tc_x3fold_set(int v) { g_tc_x3fold = v ? 1 : 0; }

// Tails of the bf16x3 operands as IEEE half instead of bfloat16 (3 more significant bits per operand) would need MMAs
// that pair a bf16 operand with an f16 one.  The instruction descriptor has separate A / B format fields, but sm_100a
// rejects the mix: the kernel dies with "an illegal instruction was encountered" (B200, round 2, gpurun_out/r2d_*).
// The code path is kept behind this constant for the record; it is never enabled.
int hdn_tc_tail16() { return 0; }

// HDN_TC_TMA: which fprop / dgrad launches take the TMA mode (bf16 operand pre-pass + tile loads instead of the SIMT
// producers): 0 none, 1 the 1x3x3 / 3x3x3 layers, 2 also the 1x1x1 layers (default: measured 656 vs 664 ms per headline
// step, profiles/r02g_*).  Stems keep their own producer.
static int g_tc_tma = -1;
int hdn_tc_tma() {
  if (g_tc_tma < 0) {
    const char* e = getenv("HDN_TC_TMA");
    g_tc_tma = e ? atoi(e) : 2;
    if (g_tc_tma < 0 || g_tc_tma > 2) g_tc_tma = 2;
  }
  return g_tc_tma;
}
void hdn_tc_tma_set(int v) { g_tc_tma = v < 0 ? 0 : (v > 2 ? 2 : v); }

// HDN_TC_SW128=0|1: TMA mode with SWIZZLE_128B operand rows (one tile load per stage, swizzled weight rows) instead of the
// 16-byte chunk planes; also hdn_set_switch()
static int g_tc_sw128 = -1;
int hdn_tc_sw128() {
  if (g_tc_sw128 < 0) {
    const char* e = getenv("HDN_TC_SW128");
    g_tc_sw128 = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_tc_sw128;
}
void hdn_tc_sw128_set(int v) { g_tc_sw128 = v ? 1 : 0; }

// HDN_TC_L2PF=1: experiment switch, see TcParams::l2pf (default 0)
int hdn_tc_l2pf() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDN_TC_L2PF");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v;
}

namespace {

struct TcPlan {
  int tma, nsa, sw;                 // TMA mode (pre-packed bf16 operand), A-ring depth, SWIZZLE_128B form
  long long op_elems;               // bf16 elements of one pre-packed operand tensor (head; the tail doubles it)
  int fold;
  int BN, n_tiles, KB, CK, nsb, nraw, tmem_cols, flat, PH, PW, P, Ppad, tiles_w, tiles_h;
  int ub;
  int PHs[2], PWs[2], Ps[2], raw_off[2], ab_off[2], raw_bytes, tab_src[2], tab_vq[2], tab_ints;
  long long ws_elems;
  size_t smem;
};

bool tc_shape_ok(const hdn_conv* c) {
  if (c->sd != 1 || c->sh != 1 || c->sw != 1) return false;
  const bool k111 = c->kd == 1 && c->kh == 1 && c->kw == 1;
  const bool k133 = c->kd == 1 && c->kh == 3 && c->kw == 3;
  const bool k333 = c->kd == 3 && c->kh == 3 && c->kw == 3;
  if (!(k111 || k133 || k333)) return false;
  if (c->pd != c->kd / 2 || c->ph != c->kh / 2 || c->pw != c->kw / 2) return false;
  const hdn_src& s0 = c->src[0];
  if (c->D != s0.D * s0.ud || c->H != s0.H * s0.uh || c->W != s0.W * s0.uw) return false;
  for (int i = 0; i < c->nsrc; ++i) {
    const hdn_src& s = c->src[i];
    if (s.t.ldc % 4 || s.t.coff % 4) return false;
    if ((reinterpret_cast<uintptr_t>(s.t.p) & 15) != 0) return false;
  }
  return true;
}
bool tc_y_aligned(const hdn_conv* c) {
  return c->y.ldc % 4 == 0 && c->y.coff % 4 == 0 && (reinterpret_cast<uintptr_t>(c->y.p) & 15) == 0;
}

// Stride-2 stems (hybridnet.py:122-123 3dconv1 7x7x7/2, :208-209 conv1 7x7/2, after ZeroPadding 3) run on the
// tensor cores in space-to-depth form.  Returns 3 / 2 for the 3-D / 2-D stem, 0 otherwise.
int tc_stem(const hdn_conv* c) {
  const bool k3 = c->kd == 7 && c->kh == 7 && c->kw == 7 && c->sd == 2 && c->sh == 2 && c->sw == 2 && c->pd == 3 &&
                  c->ph == 3 && c->pw == 3;
  const bool k2 = c->kd == 1 && c->kh == 7 && c->kw == 7 && c->sd == 1 && c->sh == 2 && c->sw == 2 && c->pd == 0 &&
                  c->ph == 3 && c->pw == 3;
  if (!k3 && !k2) return 0;
  const hdn_src& s = c->src[0];
  if (c->nsrc != 1 || s.ud != 1 || s.uh != 1 || s.uw != 1 || s.pa || s.pb || s.relu) return 0;
  if (s.t.ldc != 4 || s.t.coff != 0 || c->Cin > 4 || (reinterpret_cast<uintptr_t>(s.t.p) & 15) != 0) return 0;
  if ((s.H & 1) || (s.W & 1) || c->H != s.H / 2 || c->W != s.W / 2) return 0;
  if (k3 && ((s.D & 1) || c->D != s.D / 2)) return 0;
  if (k2 && c->D != s.D) return 0;
  if (c->Cout % 8) return 0;
  return k3 ? 3 : 2;
}

struct TcGeom { int kd, kh, kw, pd_lo, ph_lo, pw_lo, K, NC, flat, s2d, quads, scatter; };

TcGeom tc_geom(const hdn_conv* c, int mode) {
  TcGeom g;
  memset(&g, 0, sizeof(g));
  const int stem = tc_stem(c);
  if (stem) {
    g.kd = stem == 3 ? 4 : 1; g.kh = 4; g.kw = 4;
    g.quads = stem == 3 ? 8 : 4;
    if (mode == 0) { g.pd_lo = stem == 3 ? 2 : 0; g.ph_lo = 2; g.pw_lo = 2; g.K = g.quads * 4; g.NC = c->Cout; g.s2d = 1; }
    else { g.pd_lo = stem == 3 ? 1 : 0; g.ph_lo = 1; g.pw_lo = 1; g.K = c->Cout; g.NC = g.quads * 4; g.scatter = 1; }
    return g;
  }
  g.kd = c->kd; g.kh = c->kh; g.kw = c->kw;
  g.pd_lo = c->kd / 2; g.ph_lo = c->kh / 2; g.pw_lo = c->kw / 2;
  g.K = mode == 0 ? c->Cin : c->Cout;
  g.NC = mode == 0 ? c->Cout : c->Cin;
  bool up = false;
  for (int i = 0; i < c->nsrc; ++i) up = up || c->src[i].ud != 1 || c->src[i].uh != 1 || c->src[i].uw != 1;
  g.flat = (c->kd == 1 && c->kh == 1 && c->kw == 1 && !up) ? 1 : 0;
  return g;
}

// Weight ring in units of `ub` filter taps: the issuing thread pays ~60 cycles for every mbarrier wait and every
// tcgen05.commit, and nothing it executes overlaps with the MMAs it has issued (scripts/micro/mma_rate2.cu: 48 cycles per
// isolated M=128 x N=64 MMA, 78 with one wait + one commit per 6 MMAs) -- so one barrier pair covers 3-4 taps, not one.
static void tc_weight_units(TcPlan& pl, int kw) {       // a unit is one filter row (ub == kw) or one tap
  pl.ub = (kw >= 2 && kw <= 4 && pl.nsb >= 2 * kw) ? kw : 1;
  pl.nsb = pl.nsb / pl.ub * pl.ub;
}

TcPlan tc_plan1(const hdn_conv* c, const TcGeom& g, int mode, int extra_tiles) {
  TcPlan pl;
  memset(&pl, 0, sizeof(pl));
  const int tma_level = hdn_tc_tma();
  pl.tma = (!g.s2d && !g.scatter && (long long)c->N * c->D * c->H * c->W < (1ll << 31) &&
            (tma_level >= 2 || (tma_level == 1 && !g.flat))) ? 1 : 0;
  pl.nsa = NSA;
  pl.sw = (pl.tma && hdn_tc_sw128()) ? 1 : 0;
  const int nsplit = c->precision == 2 ? 2 : 1;           // bf16x3: head + tail of every operand
  pl.n_tiles = (g.NC + 255) / 256 + extra_tiles;
  int bn = (g.NC + pl.n_tiles - 1) / pl.n_tiles;
  pl.BN = (bn + 15) / 16 * 16;
  const int nsrc = mode == 0 ? c->nsrc : 1;               // dgrad: the A operand is dY alone
  pl.CK = ((nsrc == 2 && !pl.tma) || nsplit == 2) ? 32 : 64;   // channels per stage (two raw patches, or head + tail chunks, must fit a stage)
  pl.KB = (g.K + pl.CK - 1) / pl.CK;
  pl.flat = g.flat;
  pl.PH = 16 + g.kh - 1;
  pl.PW = 8 + g.kw - 1;
  pl.P = pl.PH * pl.PW;
  pl.Ppad = pl.tma ? (pl.P + 7) / 8 * 8 : (pl.P | 1);       // TMA planes start on 128-byte boundaries
  pl.tiles_h = (c->H + 15) / 16;
  pl.tiles_w = (c->W + 7) / 8;
  int off = 0, toff = 0;
  for (int s = 0; s < 2; ++s) {
    pl.PHs[s] = pl.PH; pl.PWs[s] = pl.PW;
    if (s < nsrc && mode == 0 && !g.s2d) {
      if (c->src[s].uh == 2) pl.PHs[s] = pl.PH / 2 + 1;
      if (c->src[s].uw == 2) pl.PWs[s] = pl.PW / 2 + 1;
    }
    pl.Ps[s] = s < nsrc ? pl.PHs[s] * pl.PWs[s] : 0;
    pl.raw_off[s] = off;
    off += pl.Ps[s] * (pl.CK + 4) * 4;
    pl.tab_src[s] = toff; toff += pl.Ps[s];
    pl.tab_vq[s] = toff; toff += s < nsrc ? pl.P : 0;
  }
  for (int s = 0; s < 2; ++s) { pl.ab_off[s] = off; off += s < nsrc ? 2 * pl.CK * 4 : 0; }
  pl.raw_bytes = (off + 127) / 128 * 128;
  pl.tab_ints = toff;
  const size_t a_bytes = pl.sw ? ((size_t)pl.P * 128 + 1023) / 1024 * 1024 : 8ull * pl.Ppad * 16;
  const size_t b_bytes = (size_t)pl.BN * pl.CK * 2 * nsplit;
  pl.fold = (kFold && nsplit == 2 && hdn_tc_x3fold() && pl.BN <= 128 && pl.tma && !pl.sw) ? 1 : 0;
  int cols = 32;
  while (cols < 2 * pl.BN * (pl.fold ? 2 : 1)) cols *= 2;  // two accumulator buffers
  pl.tmem_cols = cols;
  pl.ws_elems = (long long)pl.n_tiles * pl.KB * (g.kd * g.kh * g.kw) * pl.BN * pl.CK * nsplit;
  const long long budget = 226 * 1024;
  if (pl.tma) {
    // no raw ring, no geometry tables: the A ring gets up to 4 stages next to >= taps_hw (<= 12) weight blocks
    pl.tab_ints = 0; pl.raw_bytes = 0;
    const size_t fix = 8ull * pl.BN * 4 + 32ull * pl.BN + 8 + 16 + 12ull * EPI_BYTES + (2 * MAXNSA + 2 * NSB_MAX + 4) * 8 + 16 + 128 + 1024;
    const int want_b = g.kh * g.kw < 4 ? 4 : (g.kh * g.kw > 12 ? 12 : g.kh * g.kw);
    int nsa = 4;
    while (nsa > 2 && (long long)(fix + nsa * a_bytes + (size_t)want_b * b_bytes) > budget) --nsa;
    pl.nsa = nsa;
    long long room = budget - (long long)fix - (long long)nsa * (long long)a_bytes;
    int nsb = (int)(room / (long long)b_bytes);
    pl.nsb = nsb < 2 ? 2 : (nsb > NSB_MAX ? NSB_MAX : nsb);
    tc_weight_units(pl, g.kw);
    pl.nraw = (room >= 2 * (long long)b_bytes) ? nsa : 0;      // reported in the plan's ring-depth slot; 0 = does not fit (narrower tile)
    pl.smem = fix + (size_t)nsa * a_bytes + (size_t)pl.nsb * b_bytes;
    // one operand tensor: K channels per pixel, or -- SWIZZLE_128B form of bf16x3 -- [head 32 | tail 32] per 32-channel group
    pl.op_elems = (long long)c->N * c->D * c->H * c->W * ((pl.sw && nsplit == 2) ? (g.K + 31) / 32 * 32 : g.K);
    return pl;
  }
  const size_t base = NSA * a_bytes + (size_t)NTAB * pl.tab_ints * 4 + 8ull * pl.BN * 4 + 32ull * pl.BN + 8 + 16 + 4ull * EPI_BYTES +
                      (2 * MAXNSA + 2 * NSB_MAX + 4) * 8 + 16 + 1024;
  // raw fp32 ring: 3 stages when they fit next to two weight blocks, else 2; the weight ring takes what is left:
  // every tile streams ALL its weight blocks from L2, so the bytes in flight there set the pace of narrow layers
  pl.nraw = 0;
  if (!g.s2d) {
    pl.nraw = ((long long)(base + 2 * b_bytes + 3ull * pl.raw_bytes) <= budget) ? 3 : 2;
    if ((long long)(base + 2 * b_bytes + (size_t)pl.nraw * pl.raw_bytes) > budget) pl.nraw = 0;   // does not fit: refused by tc_launch
  }
  long long room = budget - (long long)base - (long long)pl.nraw * pl.raw_bytes;
  int nsb = (int)(room / (long long)b_bytes);
  pl.nsb = nsb < 2 ? 2 : (nsb > NSB_MAX ? NSB_MAX : nsb);
  tc_weight_units(pl, g.kw);
  const size_t fixed = base + pl.nsb * b_bytes;
  pl.smem = fixed + (size_t)pl.nraw * pl.raw_bytes;
  return pl;
}

// narrower column tiles until two raw stages and two weight blocks fit next to the rest
TcPlan tc_plan(const hdn_conv* c, const TcGeom& g, int mode) {
  TcPlan pl = tc_plan1(c, g, mode, 0);
  for (int extra = 1; extra < 8 && !g.s2d && pl.nraw < 2; ++extra) pl = tc_plan1(c, g, mode, extra);
  return pl;
}

template <int MODE, bool FOLD, int OPER>
int tc_run(unsigned grid, size_t smem, cudaStream_t st, const TcParams& p) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE, FOLD, OPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hdn_set_error("conv tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
    attr_set = true;
  }
  HDN_LAUNCHED(1), conv_tc_kernel<MODE, FOLD, OPER><<<grid, OPER ? TC_THREADS_TMA : TC_THREADS, smem, st>>>(p);
  return HDN_OK;
}
int tc_dispatch(int mode, int fold, int oper, unsigned grid, size_t smem, cudaStream_t st, const TcParams& p) {
  const int key = (mode ? 1 : 0) * 100 + (fold ? 1 : 0) * 10 + oper;
  switch (key) {
    case 0: return tc_run<0, false, 0>(grid, smem, st, p);
    case 1: return tc_run<0, false, 1>(grid, smem, st, p);
    case 2: return tc_run<0, false, 2>(grid, smem, st, p);
    case 11: return tc_run<0, true, 1>(grid, smem, st, p);
    case 100: return tc_run<1, false, 0>(grid, smem, st, p);
    case 101: return tc_run<1, false, 1>(grid, smem, st, p);
    case 102: return tc_run<1, false, 2>(grid, smem, st, p);
    case 111: return tc_run<1, true, 1>(grid, smem, st, p);
  }
  hdn_set_error("conv tc: no kernel instance for pass %d fold %d operand path %d", mode, fold, oper);
  return HDN_ERR_UNSUPPORTED;
}

int tc_launch(const hdn_conv* c, const hdn_dgrad_epi* epi, int mode, cudaStream_t st) {
  const TcGeom g = tc_geom(c, mode);
  const TcPlan pl = tc_plan(c, g, mode);
  const int nsplit = c->precision == 2 ? 2 : 1;
  HDN_CHECK_ARG(g.s2d || pl.nraw >= 2, "conv tc: shared memory cannot hold two raw stages (BN=%d)", pl.BN);
  HDN_CHECK_ARG(g.s2d || pl.tma || (pl.Ps[0] * (pl.CK / 4) <= MAXC0 * NPROD && pl.Ps[1] * (pl.CK / 4) <= MAXC1 * NPROD),
                "conv tc: patch of %d / %d pixels exceeds the per-thread copy list", pl.Ps[0], pl.Ps[1]);
  HDN_CHECK_ARG(c->ws != nullptr && c->ws_bytes >= pl.ws_elems * 2, "conv tc: workspace too small (%lld < %lld bytes)",
                (long long)c->ws_bytes, (long long)pl.ws_elems * 2);
  const long long w_bytes = (pl.ws_elems * 2 + 255) / 256 * 256;
  const long long op_bytes = pl.tma ? (pl.op_elems * 2 + 255) / 256 * 256 : 0;
  HDN_CHECK_ARG(!pl.tma || (c->ws_bytes >= w_bytes + op_bytes * nsplit && (reinterpret_cast<uintptr_t>(c->ws) & 255) == 0),
                "conv tc: workspace too small or misaligned for the pre-packed operand (%lld < %lld bytes)", (long long)c->ws_bytes,
                (long long)(w_bytes + op_bytes * nsplit));
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(c->ws);
  const int tail16 = (nsplit == 2 && !pl.fold && hdn_tc_tail16()) ? 1 : 0;
  {
    long long total = pl.ws_elems / 8;
    unsigned gr = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (pl.sw)
      HDN_LAUNCHED(1), pack_weights_sw_kernel<<<gr, 256, 0, st>>>(c->w, wp, c->Cin, c->Cout, c->kd, c->kh, c->kw, pl.BN, pl.KB, mode, nsplit == 2 ? 1 : 0, total);
    else if (g.s2d || g.scatter)
      HDN_LAUNCHED(1), pack_weights_s2d_kernel<<<gr, 256, 0, st>>>(c->w, wp, c->Cin, c->Cout, g.quads == 8 ? 1 : 0, pl.BN, pl.KB, pl.CK, mode, nsplit, pl.fold, tail16, total);
    else
      HDN_LAUNCHED(1), pack_weights_kernel<<<gr, 256, 0, st>>>(c->w, wp, c->Cin, c->Cout, c->kd, c->kh, c->kw, pl.BN, pl.KB, pl.CK, mode, nsplit, pl.fold, tail16, total);
    HDN_CHECK_LAUNCH("pack_weights");
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.tma = pl.tma; p.nsa = pl.nsa; p.sw = pl.sw;
  if (pl.tma) {
    // operand pre-pass (once per launch): fprop -- max(a*x+b, 0) (+ second source) on the virtual grid; dgrad -- dY
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(c->ws) + w_bytes);
    __nv_bfloat16* lo = nsplit == 2 ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(c->ws) + w_bytes + op_bytes) : nullptr;
    int rc;
    const int il = (pl.sw && nsplit == 2) ? 1 : 0;            // interleaved [head | tail] rows in ONE tensor
    if (il) lo = nullptr;
    if (mode == 0) rc = hdn_tc2_pack(c->src, c->nsrc, c->N, c->D, c->H, c->W, c->Cin, hi, lo, st, il);
    else {
      hdn_src dys;
      memset(&dys, 0, sizeof(dys));
      dys.t = c->y; dys.D = c->D; dys.H = c->H; dys.W = c->W; dys.ud = dys.uh = dys.uw = 1;
      rc = hdn_tc2_pack(&dys, 1, c->N, c->D, c->H, c->W, c->Cout, hi, lo, st, il);
    }
    if (rc) return rc;
    const long long Mv = (long long)c->N * c->D * c->H * c->W;
    if (pl.sw) {
      const int C2 = il ? (g.K + 31) / 32 * 64 : g.K;
      rc = hdn_tc2_make_map(&p.tmHi, hi, pl.flat, Mv, c->N, c->D, c->H, c->W, C2, 64, pl.PW, pl.PH, 1);
      if (rc) return rc;
    } else {
      rc = hdn_tc2_make_map(&p.tmHi, hi, pl.flat, Mv, c->N, c->D, c->H, c->W, g.K, 8, pl.PW, pl.PH, 0);
      if (rc) return rc;
      if (lo) { rc = hdn_tc2_make_map(&p.tmLo, lo, pl.flat, Mv, c->N, c->D, c->H, c->W, g.K, 8, pl.PW, pl.PH, 0); if (rc) return rc; }
    }
  }
  p.N = c->N; p.D = c->D; p.H = c->H; p.W = c->W;
  p.kd = g.kd; p.kh = g.kh; p.kw = g.kw;
  p.pd_lo = g.pd_lo; p.ph_lo = g.ph_lo; p.pw_lo = g.pw_lo;
  p.s2d = g.s2d; p.s2d_quads = g.quads; p.scatter = g.scatter;
  p.K = g.K; p.NC = g.NC;
  p.BN = pl.BN; p.KB = pl.KB; p.CK = pl.CK; p.nsb = pl.nsb; p.ub = pl.ub; p.nraw = pl.nraw; p.tmem_cols = pl.tmem_cols;
  p.flat = pl.flat; p.PH = pl.PH; p.PW = pl.PW; p.P = pl.P; p.Ppad = pl.Ppad;
  for (int s = 0; s < 2; ++s) {
    p.PHs[s] = pl.PHs[s]; p.PWs[s] = pl.PWs[s]; p.Ps[s] = pl.Ps[s]; p.raw_off[s] = pl.raw_off[s]; p.ab_off[s] = pl.ab_off[s];
    p.tab_src[s] = pl.tab_src[s]; p.tab_vq[s] = pl.tab_vq[s];
  }
  p.raw_bytes = pl.raw_bytes; p.tab_ints = pl.tab_ints;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.M = (long long)c->N * c->D * c->H * c->W;
  p.wpack = wp;
  p.mode = mode;
  p.split = nsplit == 2 ? 1 : 0;
  p.tail16 = tail16;
  p.fastx = hdn_tc_fastx();
  { static int v = -1; if (v < 0) { const char* e = getenv("HDN_TC_EPIPF"); v = (e && atoi(e) == 0) ? 0 : 1; } p.epi_pf = v; }
  p.l2pf = hdn_tc_l2pf();
  p.fold = pl.fold;
  if (mode == 0) {
    p.nsrc = c->nsrc;
    p.src[0] = c->src[0];
    p.src[1] = c->src[1];
    p.bias = c->bias; p.y = c->y; p.stat_sum = c->stat_sum; p.stat_sq = c->stat_sq;
    p.drop_keep = c->drop_keep; p.drop_seed = c->drop_seed;
  } else {
    // the A operand is dY on the output grid, taken as is
    p.nsrc = 1;
    hdn_src dy;
    memset(&dy, 0, sizeof(dy));
    dy.t = c->y; dy.D = c->D; dy.H = c->H; dy.W = c->W; dy.ud = dy.uh = dy.uw = 1;
    p.src[0] = dy;
    p.nepi = c->nsrc;
    for (int i = 0; i < c->nsrc; ++i) {
      p.esrc[i] = c->src[i];
      p.epi[i] = epi[i];
      if (epi[i].mode == 2) continue;
      const hdn_src& s = c->src[i];
      HDN_CHECK_ARG(epi[i].mode != 0 || (epi[i].dx.ldc % 4 == 0 && epi[i].dx.coff % 4 == 0), "conv_dgrad tc: dx window not 16-byte aligned");
      if (g.scatter)
        HDN_CHECK_ARG(epi[i].mode == 0 && epi[i].s1 == nullptr && epi[i].dx.ldc == 4 && epi[i].dx.coff == 0,
                      "conv_dgrad tc: the stem input gradient must be a plain 4-channel tensor");
      if (s.ud == 2 && !epi[i].accumulate) {
        // depth sub-positions live in different CTAs and are combined with atomics: start from zero
        const long long Ms = (long long)c->N * s.D * s.H * s.W;
        if (epi[i].mode == 0) {
          unsigned gr = (unsigned)((Ms * c->Cin + 255) / 256 > 148 * 16 ? 148 * 16 : (Ms * c->Cin + 255) / 256);
          HDN_LAUNCHED(1), zero_window_kernel<<<gr, 256, 0, st>>>(epi[i].dx, Ms, c->Cin);
          HDN_CHECK_LAUNCH("zero_window");
        } else {
          cudaError_t e = cudaMemsetAsync(epi[i].du, 0, (size_t)Ms * c->Cin * sizeof(float), st);
          if (e != cudaSuccess) { hdn_set_error("conv_dgrad tc: memset: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
        }
      }
    }
  }
  const long long tiles = pl.flat ? (p.M + 127) / 128 : (long long)c->N * c->D * pl.tiles_h * pl.tiles_w;
  p.tiles = tiles;
  p.total_work = tiles * pl.n_tiles;
  HDN_CHECK_ARG(p.total_work < (1ll << 31), "conv tc: too many work items");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  const unsigned grid = (unsigned)(p.total_work < num_sms ? p.total_work : num_sms);   // persistent: one CTA per SM
  const int oper = pl.tma ? (pl.sw ? 2 : 1) : 0;
  HDN_CHECK_ARG(!(pl.fold && oper != 1), "conv tc: the folded bf16x3 form exists for the chunk-plane TMA mode only");
  const int rc = tc_dispatch(mode, pl.fold, oper, grid, pl.smem, st, p);
  if (rc != HDN_OK) return rc;
  HDN_CHECK_LAUNCH(mode == 0 ? "conv_fprop_tc" : "conv_dgrad_tc");
  return HDN_OK;
}

}  // namespace

int hdn_wgrad_tc_supported(const hdn_conv* c);
int hdn_tc_stem(const hdn_conv* c) { return tc_stem(c); }

int hdn_tc_supported(const hdn_conv* c, int pass) {
  if (tc_stem(c)) {
    if (pass == 0) return 1;
    if (pass == 1) return tc_y_aligned(c) ? 1 : 0;
    return hdn_wgrad_tc_supported(c);
  }
  if (!tc_shape_ok(c)) return 0;
  if (pass == 0) return (c->Cin % 8 == 0) ? 1 : 0;
  if (pass == 1) return (c->Cout % 8 == 0 && c->Cin % 8 == 0 && tc_y_aligned(c)) ? 1 : 0;   // dY is the A operand
  return hdn_wgrad_tc_supported(c);
}

long long hdn_tc_workspace_bytes(const hdn_conv* c, int pass) {
  if (!hdn_tc_supported(c, pass) || pass == 2) return 0;
  const TcPlan pl = tc_plan(c, tc_geom(c, pass), pass);
  const long long w_bytes = (pl.ws_elems * 2 + 255) / 256 * 256;
  if (!pl.tma) return pl.ws_elems * 2;
  return w_bytes + (pl.op_elems * 2 + 255) / 256 * 256 * (c->precision == 2 ? 2 : 1);      // packed weights + bf16 head (+ tail) operand
                                                                                        // (SWIZZLE_128B form: one interleaved tensor of the same size)
}

// launch plan of the fprop / dgrad kernel for this descriptor (host arithmetic only; see hdn_conv_tc_plan in hdn.h)
int hdn_tc_plan_info(const hdn_conv* c, int pass, int* out) {
  const TcGeom g = tc_geom(c, pass);
  const TcPlan pl = tc_plan(c, g, pass);
  const long long M = (long long)c->N * c->D * c->H * c->W;
  const long long tiles = pl.flat ? (M + 127) / 128 : (long long)c->N * c->D * pl.tiles_h * pl.tiles_w;
  out[0] = pl.BN; out[1] = pl.n_tiles; out[2] = pl.KB; out[3] = pl.CK; out[4] = pl.nsb; out[5] = pl.nraw;
  out[6] = pl.tmem_cols; out[7] = (int)pl.smem; out[8] = pl.flat; out[9] = pl.P; out[10] = c->precision == 2 ? 1 : 0;
  out[11] = g.s2d; out[12] = (tiles * pl.n_tiles < (1ll << 31)) ? (int)(tiles * pl.n_tiles) : -1;
  out[13] = pl.tma || (pl.Ps[0] * (pl.CK / 4) <= MAXC0 * NPROD && pl.Ps[1] * (pl.CK / 4) <= MAXC1 * NPROD);   // per-thread copy list holds the patch (SIMT-producer form)
  out[14] = g.K; out[15] = g.NC;
  return HDN_OK;
}

int hdn_conv_fprop_tc(const hdn_conv* c, cudaStream_t st) { return tc_launch(c, nullptr, 0, st); }
int hdn_conv_dgrad_tc(const hdn_conv* c, const hdn_dgrad_epi* e, cudaStream_t st) { return tc_launch(c, e, 1, st); }
