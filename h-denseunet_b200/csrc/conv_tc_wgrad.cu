// tcgen05 weight-gradient kernel (precision == 1):  dW[tap][ci][co] += sum_pos A[pos + off(tap)][ci] * dY[pos][co]
// for the stride-1 "same" convolutions (1x1x1, 1x3x3, 3x3x3), A being the convolution's (virtual) input
// max(a*x+b,0) (+ second source, up-sampling folded into the index) exactly as in fprop.
//
// GEMM view per tap: D[ci (M=128)][co (N=BN)] += At[ci][pos] * dY[pos][co], contraction over positions.
// Both operands are "MN-major" (channels contiguous), which is how NDHWC activations already lie in
// memory, so the producers use the same chunk layout as fprop: 8-channel chunk j at j*SBO, pixel q at
// q*16 bytes.  One position tile = 16 x 8 pixels; its input PATCH (tile + halo) is staged once and every
// tap of the tap group reads it through a shifted descriptor; each tap owns BN TMEM columns.
// A CTA owns (128 input channels) x (BN output channels) x (tap group) and loops over its share of
// the position tiles (split-K over CTAs); the epilogue adds the fp32 accumulators into dW with
// vector reductions (REDG.ADD.F32x4).
//
//   warps 0-7  producers (patch + dY tile -> bf16 chunks in shared memory), then epilogue
//   warp  8    MMA issuer (one elected thread)
//
// precision == 2 ("bf16x3"): the same MMA stream on a folded tile.  A CTA owns 64 input channels whose bf16
// heads fill GEMM rows 0-63 and whose tails (x - head) fill rows 64-127; the dY tile carries BN head columns
// followed by BN tail columns (N = 2*BN).  One M=128 x N=2BN MMA then forms all four head/tail products and
// the epilogue reduces the four quadrants into the same dW entries with the vector reductions it already uses.
#include "hdn_common.cuh"
#include "tc_common.cuh"

// -DHDN_TC_TIMING: per-role wait/work cycle counters of CTA 0, printed at kernel end (development aid, as in conv_tc.cu)
#ifdef HDN_TC_TIMING
#define WT_DECL(n) long long n = 0
#define WT_BEGIN long long wt__0 = clock64()
#define WT_ADD(n) do { long long wt__1 = clock64(); n += wt__1 - wt__0; wt__0 = wt__1; } while (0)
#else
#define WT_DECL(n)
#define WT_BEGIN
#define WT_ADD(n)
#endif

int hdn_tc_stem(const hdn_conv* c);   // conv_tc.cu
int hdn_tc_fastx();                   // conv_tc.cu

namespace {

constexpr int WG_THREADS = 288;
constexpr int WG_PROD = 256;
constexpr int NS = 2;          // stage ring depth
constexpr int WUB = 3;         // pixel groups loaded ahead per producer warp

struct WgParams {
  int N, D, H, W;              // output grid (= virtual input grid)
  int kd, kh, kw;
  int Cin, Cout;
  int BN, G, groups;           // column tile, taps per group, number of tap groups
  int fastx;                   // 1: warp-per-chunk operand transform (tc::xform_chunk)
  int split, BNe, CI;          // bf16x3 folding: BNe = GEMM N (2*BN when split), CI = input channels per CTA (64 when split, else 128)
  int ci_tiles, co_tiles;
  int flat;
  int PH, PW, P, Ppad;         // patch of ONE tap group (rows PH, cols PW)
  int row0_mode;               // 0: group covers all kh rows; 1: group g covers kernel row (g % kh) only
  int pd_lo, ph_lo, pw_lo;     // padding in front of tap 0
  int s2d, s2d_quads;          // stride-2 stem in space-to-depth form (see conv_tc.cu); Cin here = 4 * quads
  int cin_real;                //   real input channels of the stem (<= 4)
  // asynchronous producer (everything but the stems): raw fp32 ring and per-tile geometry tables
  int nraw, raw_bytes, CW;     // CW: input channels per raw sub-stage (64, or 32 with two sources)
  int PHs[2], PWs[2], Ps[2], raw_off[2], ab_off[2];
  int tab_src[2], tab_vq[2], tab_dy, tab_ints;
  int tiles_w, tiles_h;
  long long n_pos_tiles;       // position tiles per slice set
  long long M;
  int nsrc;
  hdn_src src[2];
  hdn_tensor dy;
  float* dw;
  int tmem_cols;
};

__device__ __forceinline__ float4 ldg4w(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 pro4(float4 x, float4 a, float4 b, int relu) {
  float4 r;
  r.x = fmaf(a.x, x.x, b.x); r.y = fmaf(a.y, x.y, b.y); r.z = fmaf(a.z, x.z, b.z); r.w = fmaf(a.w, x.w, b.w);
  if (relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
  return r;
}

// Stage one 64-channel block [cbase, cbase+64) of `npx` pixels into chunk layout at dst (8 chunks, stride
// ppad*16 bytes).  offs[s*npx + q] is the element offset of pixel q in source s (or -1 => zeros).
__device__ __forceinline__ void stage_block(const hdn_src* src, int nsrc, const float* const* base, const int* offs,
                                            int npx, int ppad, int cbase, int climit, uint8_t* dst, int nchunks,
                                            int pwarp, int lane, int s2d_quads = 0, uint8_t* dst_lo = nullptr) {
  const int l8 = lane & 7, pg = lane >> 3;
  const bool even = (l8 & 1) == 0;
  const int chunk = even ? (l8 >> 1) : (4 + (l8 >> 1));
  const int cA = cbase + l8 * 4, cB = cbase + 32 + l8 * 4;
  bool okA = cA < climit, okB = cB < climit;
  if (s2d_quads) { okA = cbase == 0 && l8 < s2d_quads; okB = false; }   // base[] already points at this lane's 8 floats - cA
  float4 a0[2], b0[2], a1[2], b1[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    a0[s] = a1[s] = make_float4(1.f, 1.f, 1.f, 1.f);
    b0[s] = b1[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < nsrc) {
      if (src[s].pa) { if (okA) a0[s] = ldg4w(src[s].pa + cA); if (okB) a1[s] = ldg4w(src[s].pa + cB); }
      if (src[s].pb) { if (okA) b0[s] = ldg4w(src[s].pb + cA); if (okB) b1[s] = ldg4w(src[s].pb + cB); }
    }
  }
  for (int q0 = pwarp * 4; q0 < npx; q0 += 32 * WUB) {
    float4 r0[WUB][2], r1[WUB][2];
    bool inb[WUB][2];
#pragma unroll
    for (int u = 0; u < WUB; ++u) {
      const int q = q0 + u * 32 + pg;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        inb[u][s] = false;
        r0[u][s] = r1[u][s] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < nsrc && q < npx) {
          const int off = offs[s * npx + q];
          if (off >= 0) {
            inb[u][s] = true;
            const float* g = base[s] + off;
            if (okA) r0[u][s] = ldg4w(g + cA);
            if (okB) r1[u][s] = ldg4w(g + cB);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < WUB; ++u) {
      if (q0 + u * 32 >= npx) break;
      const int q = q0 + u * 32 + pg;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s < nsrc && inb[u][s]) {
          const int relu = src[s].relu;
          if (okA) { float4 t = pro4(r0[u][s], a0[s], b0[s], relu); v0.x += t.x; v0.y += t.y; v0.z += t.z; v0.w += t.w; }
          if (okB) { float4 t = pro4(r1[u][s], a1[s], b1[s], relu); v1.x += t.x; v1.y += t.y; v1.z += t.z; v1.w += t.w; }
        }
      }
      uint32_t p00, p01, p10, p11, t00 = 0, t01 = 0, t10 = 0, t11 = 0;
      if (dst_lo) {
        tc::pack_split_bf16x2(v0.x, v0.y, p00, t00); tc::pack_split_bf16x2(v0.z, v0.w, p01, t01);
        tc::pack_split_bf16x2(v1.x, v1.y, p10, t10); tc::pack_split_bf16x2(v1.z, v1.w, p11, t11);
      } else {
        p00 = tc::pack_bf16x2(v0.x, v0.y); p01 = tc::pack_bf16x2(v0.z, v0.w);
        p10 = tc::pack_bf16x2(v1.x, v1.y); p11 = tc::pack_bf16x2(v1.z, v1.w);
      }
      const uint32_t s0 = even ? p10 : p00, s1 = even ? p11 : p01;
      const uint32_t x0 = __shfl_xor_sync(0xffffffffu, s0, 1), x1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      uint4 o;
      if (even) { o.x = p00; o.y = p01; o.z = x0; o.w = x1; }
      else      { o.x = x0; o.y = x1; o.z = p10; o.w = p11; }
      if (q < npx && chunk < nchunks) *reinterpret_cast<uint4*>(dst + (uint32_t)chunk * ppad * 16u + (uint32_t)q * 16u) = o;
      if (dst_lo) {                                          // warp-uniform: the tails, regrouped the same way
        const uint32_t u0 = even ? t10 : t00, u1 = even ? t11 : t01;
        const uint32_t y0 = __shfl_xor_sync(0xffffffffu, u0, 1), y1 = __shfl_xor_sync(0xffffffffu, u1, 1);
        uint4 ot;
        if (even) { ot.x = t00; ot.y = t01; ot.z = y0; ot.w = y1; }
        else      { ot.x = y0; ot.y = y1; ot.z = t10; ot.w = t11; }
        if (q < npx && chunk < nchunks) *reinterpret_cast<uint4*>(dst_lo + (uint32_t)chunk * ppad * 16u + (uint32_t)q * 16u) = ot;
      }
    }
  }
}

__device__ __forceinline__ void wg_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
constexpr int NTW = 3;         // geometry-table buffers

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t A_BYTES = 16u * p.Ppad * 16u;                       // 128 channels of the patch
  const uint32_t B_BYTES = (uint32_t)((p.BNe + 7) / 8) * 129u * 16u; // dY tile, BNe columns x 128 pixels
  uint8_t* sA = smem;
  uint8_t* sB = sA + NS * A_BYTES;
  int* offs = reinterpret_cast<int*>(sB + NS * B_BYTES);             // [NS][2*P + 128]   (direct path)
  const int OFFS_PER = 2 * p.P + 128;
  int* tabs = offs + NS * OFFS_PER;                                   // [NTW][tab_ints]   (asynchronous path)
  uint8_t* sRaw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tabs + NTW * p.tab_ints) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + (size_t)p.nraw * p.raw_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* acc_full = bars + 2 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // ---- CTA work item
  int wi = blockIdx.x;
  const int co_t = wi % p.co_tiles; wi /= p.co_tiles;
  const int ci_t = wi % p.ci_tiles; wi /= p.ci_tiles;
  const int grp = wi;                                   // tap group
  const int split = blockIdx.y, nsplit = gridDim.y;
  const int ci0 = ci_t * p.CI, co0 = co_t * p.BN;
  const int taps_hw = p.kh * p.kw;
  // taps of this group: tap = tap0 + g, g in [0, G)
  const int tap0 = grp * p.G;
  const int gdz = tap0 / taps_hw;                       // depth tap of the group (G <= kh*kw => one dz per group)
  const int gth0 = (tap0 % taps_hw) / p.kw;             // first kernel row of the group
  const int gtw0 = (p.G == 1) ? (tap0 % p.kw) : 0;      // single-tap groups stage only that tap's columns
  const int hd = p.pd_lo, hh = p.ph_lo, hw = p.pw_lo;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { tc::mbar_init(&full[i], WG_PROD); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // position-tile iteration (identical in every role): tile t -> (n, d, th, tw) or flat index
  auto tile_valid = [&](long long t, int& n_img, int& d0, int& h0, int& w0, long long& m0) -> bool {
    if (p.flat) { m0 = t * 128; n_img = d0 = h0 = w0 = 0; return true; }
    unsigned r = (unsigned)t;                            // position tiles fit 32 bits (checked on the host)
    const unsigned tw_ = r % (unsigned)p.tiles_w; r /= (unsigned)p.tiles_w;
    const unsigned th_ = r % (unsigned)p.tiles_h; r /= (unsigned)p.tiles_h;
    d0 = (int)(r % (unsigned)p.D); n_img = (int)(r / (unsigned)p.D);
    h0 = (int)th_ * 16; w0 = (int)tw_ * 8; m0 = 0;
    const int vd = d0 - hd + gdz;
    return vd >= 0 && vd < p.D;                          // the group's depth slab lies outside: no contribution
  };

  if (warp < 8) {
    // =================================================================== producers
    int st = 0;
    uint32_t ph = 0;
    bool produced = false;
    if (p.nraw >= 2) {
      // ---- asynchronous path: raw fp32 sub-stages (64 channels of the patch, or 64 channels of the dY tile) are fetched
      // with cp.async one sub-stage ahead, then transformed into the bf16 operand stage of the tile
      const int CW = p.CW, NQ = CW >> 2, NQs = (CW == 64) ? 4 : 3, RS = CW + 4;   // RS: raw pixel stride (floats), padded
      const int nA = (min(p.CI, p.Cin - ci0) + CW - 1) / CW, nB = (p.BN + 63) >> 6, nsub = nA + nB;
      const uint32_t a_tail = 8u * (uint32_t)p.Ppad * 16u;           // bf16x3: tails of the CTA's 64 channels = GEMM rows 64-127
      const uint32_t b_tail = (uint32_t)(p.BN >> 3) * 129u * 16u;     //         tails of the BN columns = columns BN..2BN-1
      const int climit = min(p.Cout, co0 + p.BN);
      struct SubIt { long long t; int u, seq; bool done; int n_img, d0, h0, w0; long long m0; };
      auto settle = [&](SubIt& it) {
        while (it.t < p.n_pos_tiles && !tile_valid(it.t, it.n_img, it.d0, it.h0, it.w0, it.m0)) it.t += nsplit;
        it.done = it.t >= p.n_pos_tiles;
      };
      auto advance = [&](SubIt& it) {
        if (++it.u == nsub) { it.u = 0; it.t += nsplit; ++it.seq; settle(it); }
      };
      SubIt is, tr;
      is.t = split; is.u = 0; is.seq = 0; settle(is);
      tr = is;
      int built_seq = -1;
      auto build_tables = [&](const SubIt& it) {
        int* tab = tabs + (it.seq % NTW) * p.tab_ints;
        for (int s = 0; s < p.nsrc; ++s) {
          const hdn_src& S = p.src[s];
          const int vh0 = it.h0 - hh + gth0, vw0 = it.w0 - hw + gtw0;
          const int sh0 = (S.uh == 2) ? (vh0 >> 1) : vh0, sw0 = (S.uw == 2) ? (vw0 >> 1) : vw0;
          for (int i = tid; i < p.Ps[s]; i += WG_PROD) {
            int off = -1;
            if (p.flat) {
              if (it.m0 + i < p.M) off = i * S.t.ldc;
            } else {
              const int sh = sh0 + i / p.PWs[s], sw = sw0 + i % p.PWs[s];
              if (sh >= 0 && sh < S.H && sw >= 0 && sw < S.W) off = (sh * S.W + sw) * S.t.ldc;
            }
            tab[p.tab_src[s] + i] = off;
          }
          for (int q = tid; q < p.P; q += WG_PROD) {
            int sq = -1;
            if (p.flat) {
              if (it.m0 + q < p.M) sq = q;
            } else {
              const int vh = vh0 + q / p.PW, vw = vw0 + q % p.PW;
              if (vh >= 0 && vh < p.H && vw >= 0 && vw < p.W) {
                const int sh = (S.uh == 2) ? (vh >> 1) : vh, sw = (S.uw == 2) ? (vw >> 1) : vw;
                sq = (sh - sh0) * p.PWs[s] + (sw - sw0);
              }
            }
            tab[p.tab_vq[s] + q] = sq;
          }
        }
        for (int r = tid; r < 128; r += WG_PROD) {
          int off = -1;
          if (p.flat) {
            if (it.m0 + r < p.M) off = r * p.dy.ldc;
          } else {
            const int oh = it.h0 + (r >> 3), ow = it.w0 + (r & 7);
            if (oh < p.H && ow < p.W) off = (oh * p.W + ow) * p.dy.ldc;
          }
          tab[p.tab_dy + r] = off;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      };
      auto issue = [&](const SubIt& it, int slot) {
        if (it.seq != built_seq) { build_tables(it); built_seq = it.seq; }
        const int* tab = tabs + (it.seq % NTW) * p.tab_ints;
        uint8_t* raw = sRaw + (size_t)slot * p.raw_bytes;
        if (it.u < nA) {
          const int c0 = ci0 + it.u * CW;
          for (int s = 0; s < p.nsrc; ++s) {
            const hdn_src& S = p.src[s];
            const float* base;
            if (p.flat) base = S.t.p + it.m0 * S.t.ldc + S.t.coff;
            else {
              const int vd = it.d0 - hd + gdz;
              const int sd = (S.ud == 2) ? (vd >> 1) : vd;
              base = S.t.p + ((long long)it.n_img * S.D + sd) * S.H * S.W * S.t.ldc + S.t.coff;
            }
            const int* spix = tab + p.tab_src[s];
            uint8_t* rs = raw + p.raw_off[s];
            const int n16 = p.Ps[s] * NQ;
            for (int i = tid; i < n16; i += WG_PROD) {
              const int sq = i >> NQs, part = i & (NQ - 1);
              const int off = spix[sq], c = c0 + part * 4;
              if (off >= 0 && c < p.Cin) wg_cp_async16(rs + (size_t)sq * (RS * 4) + part * 16, base + off + c);
            }
            if (tid < NQ && c0 + tid * 4 < p.Cin) {
              if (S.pa) wg_cp_async16(raw + p.ab_off[s] + tid * 16, S.pa + c0 + tid * 4);
              if (S.pb) wg_cp_async16(raw + p.ab_off[s] + CW * 4 + tid * 16, S.pb + c0 + tid * 4);
            }
          }
        } else {
          const int c0 = co0 + (it.u - nA) * 64;
          const float* base = p.flat ? (p.dy.p + it.m0 * p.dy.ldc + p.dy.coff)
                                     : (p.dy.p + ((long long)it.n_img * p.D + it.d0) * p.H * p.W * p.dy.ldc + p.dy.coff);
          const int* dyo = tab + p.tab_dy;
          for (int i = tid; i < 128 * 16; i += WG_PROD) {
            const int q = i >> 4, part = i & 15;
            const int off = dyo[q], c = c0 + part * 4;
            if (off >= 0 && c < climit) wg_cp_async16(raw + (size_t)q * (68 * 4) + part * 16, base + off + c);
          }
        }
      };
      int slot_is = 0, slot_tr = 0;
      for (int i = 0; i < p.nraw - 1; ++i) {
        if (!is.done) { issue(is, slot_is); advance(is); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++slot_is == p.nraw) slot_is = 0;
      }
      WT_DECL(t_cp); WT_DECL(t_bar); WT_DECL(t_issue); WT_DECL(t_empty); WT_DECL(t_xa); WT_DECL(t_xb); WT_DECL(n_sub);
      while (!tr.done) {
        WT_BEGIN;
        if (p.nraw == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
        WT_ADD(t_cp);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        WT_ADD(t_bar);
        if (!is.done) { issue(is, slot_is); advance(is); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++slot_is == p.nraw) slot_is = 0;
        WT_ADD(t_issue);
        if (tr.u == 0) { produced = true; tc::mbar_wait(&empty[st], ph ^ 1); }
        WT_ADD(t_empty);
#ifdef HDN_TC_TIMING
        const bool wt_is_a = tr.u < nA;
#endif
        const int* tab = tabs + (tr.seq % NTW) * p.tab_ints;
        const uint8_t* raw = sRaw + (size_t)slot_tr * p.raw_bytes;
        if (tr.u < nA) {
          // lane = patch pixel, inner loop over the 8-channel chunks of this sub-stage (see conv_tc.cu)
          const int c0 = ci0 + tr.u * CW;
          uint8_t* dst = sA + st * A_BYTES + (uint32_t)tr.u * (uint32_t)(CW / 8) * p.Ppad * 16u;
          const float* rawf0 = reinterpret_cast<const float*>(raw + p.raw_off[0]);
          const float* rawf1 = reinterpret_cast<const float*>(raw + p.raw_off[1]);
          const float* ab0 = reinterpret_cast<const float*>(raw + p.ab_off[0]);
          const float* ab1 = reinterpret_cast<const float*>(raw + p.ab_off[1]);
          const int* vq0 = tab + p.tab_vq[0];
          const int* vq1 = tab + p.tab_vq[1];
          const bool two = p.nsrc > 1;
          const bool pa0 = p.src[0].pa != nullptr, pb0 = p.src[0].pb != nullptr;
          const int relu0 = p.src[0].relu;
          const bool pa1 = two && p.src[1].pa != nullptr, pb1 = two && p.src[1].pb != nullptr;
          const int relu1 = two ? p.src[1].relu : 0;
          const int nhalf = CW >> 4;                             // chunks per work item (half of the sub-stage's chunks)
          const int nitems = ((p.P + 31) >> 5) * 2;
          const int xmode = (pa0 && pb0 && relu0 != 0) ? 1 : ((!pa0 && !pb0 && relu0 == 0) ? 0 : -1);
          const int xmode1 = !two ? -1 : ((pa1 && pb1 && relu1 != 0) ? 1 : ((!pa1 && !pb1 && relu1 == 0) ? 0 : -1));
          if (p.fastx >= 2 && two && xmode1 == 1 && xmode >= 0) {
            const int nch = CW >> 3, j = warp & (nch - 1), c = j * 8;
            tc::xform_chunk2_any(xmode, xmode1, p.split != 0 ? 1 : 0, rawf0 + c, rawf1 + c, RS, vq0, vq1, p.P, (warp / nch) * 32 + lane,
                                 (8 / nch) * 32, ab0 + c, ab0 + CW + c, ab1 + c, ab1 + CW + c, c0 + c < p.Cin,
                                 dst + (uint32_t)j * (uint32_t)p.Ppad * 16u, dst + a_tail + (uint32_t)j * (uint32_t)p.Ppad * 16u);
          } else if (p.fastx && !two && xmode >= 0) {
            // warp-per-chunk form (tc::xform_chunk): the chunk's (a, b) stay in registers for the whole sub-stage
            const int nch = CW >> 3, j = warp & (nch - 1), c = j * 8;
            tc::xform_chunk_any(xmode, p.split != 0 ? 1 : 0, rawf0 + c, RS, vq0, p.P, (warp / nch) * 32 + lane, (8 / nch) * 32,
                                ab0 + c, ab0 + CW + c, c0 + c < p.Cin, dst + (uint32_t)j * (uint32_t)p.Ppad * 16u,
                                dst + a_tail + (uint32_t)j * (uint32_t)p.Ppad * 16u);
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
              if (c0 + c < p.Cin) {
                if (sq0 >= 0) {
                  va = *reinterpret_cast<const float4*>(row0 + c);
                  vb = *reinterpret_cast<const float4*>(row0 + c + 4);
                  if (pa0 | pb0 | (relu0 != 0)) {
                    const float4 aa = pa0 ? *reinterpret_cast<const float4*>(ab0 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ab_ = pa0 ? *reinterpret_cast<const float4*>(ab0 + c + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ba = pb0 ? *reinterpret_cast<const float4*>(ab0 + CW + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bb = pb0 ? *reinterpret_cast<const float4*>(ab0 + CW + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    va = pro4(va, aa, ba, relu0);
                    vb = pro4(vb, ab_, bb, relu0);
                  }
                }
                if (sq1 >= 0) {
                  float4 ua = *reinterpret_cast<const float4*>(row1 + c), ub = *reinterpret_cast<const float4*>(row1 + c + 4);
                  if (pa1 | pb1 | (relu1 != 0)) {
                    const float4 aa = pa1 ? *reinterpret_cast<const float4*>(ab1 + c) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ab_ = pa1 ? *reinterpret_cast<const float4*>(ab1 + c + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 ba = pb1 ? *reinterpret_cast<const float4*>(ab1 + CW + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bb = pb1 ? *reinterpret_cast<const float4*>(ab1 + CW + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ua = pro4(ua, aa, ba, relu1);
                    ub = pro4(ub, ab_, bb, relu1);
                  }
                  va.x += ua.x; va.y += ua.y; va.z += ua.z; va.w += ua.w;
                  vb.x += ub.x; vb.y += ub.y; vb.z += ub.z; vb.w += ub.w;
                }
              }
              uint4 o;
              if (!p.split) {
                o.x = tc::pack_bf16x2(va.x, va.y); o.y = tc::pack_bf16x2(va.z, va.w);
                o.z = tc::pack_bf16x2(vb.x, vb.y); o.w = tc::pack_bf16x2(vb.z, vb.w);
              } else {
                uint4 t;
                tc::pack_split_bf16x2(va.x, va.y, o.x, t.x); tc::pack_split_bf16x2(va.z, va.w, o.y, t.y);
                tc::pack_split_bf16x2(vb.x, vb.y, o.z, t.z); tc::pack_split_bf16x2(vb.z, vb.w, o.w, t.w);
                if (qok) *reinterpret_cast<uint4*>(drow + a_tail + (uint32_t)j * (uint32_t)p.Ppad * 16u) = t;
              }
              if (qok) *reinterpret_cast<uint4*>(drow + (uint32_t)j * (uint32_t)p.Ppad * 16u) = o;
            }
          }
        } else {
          // dY sub-stage: 128 pixels x up to 64 channels, no prologue; warp -> (32 pixels, 4 chunks)
          const int cb = (tr.u - nA) * 64;
          const int c0 = co0 + cb;
          uint8_t* dst = sB + st * B_BYTES + (uint32_t)(cb / 8) * 129u * 16u;
          const int nch = (p.BN - cb + 7) / 8;                   // chunks that exist in the B stage for this block
          const int* dyo = tab + p.tab_dy;
          const float* rawf = reinterpret_cast<const float*>(raw);
          const int q = (warp >> 1) * 32 + lane;
          const int j0 = (warp & 1) * 4;
          const bool has = dyo[q] >= 0;
          const float* row = rawf + q * 68;                       // dY raw pixel stride: 64 + 4 floats
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj, c = j * 8;
            if (j >= nch) break;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (has && c0 + c < climit) { va = *reinterpret_cast<const float4*>(row + c); vb = *reinterpret_cast<const float4*>(row + c + 4); }
            uint4 o;
            if (!p.split) {
              o.x = tc::pack_bf16x2(va.x, va.y); o.y = tc::pack_bf16x2(va.z, va.w);
              o.z = tc::pack_bf16x2(vb.x, vb.y); o.w = tc::pack_bf16x2(vb.z, vb.w);
            } else {
              uint4 t;
              tc::pack_split_bf16x2(va.x, va.y, o.x, t.x); tc::pack_split_bf16x2(va.z, va.w, o.y, t.y);
              tc::pack_split_bf16x2(vb.x, vb.y, o.z, t.z); tc::pack_split_bf16x2(vb.z, vb.w, o.w, t.w);
              *reinterpret_cast<uint4*>(dst + b_tail + (uint32_t)j * 129u * 16u + (uint32_t)q * 16u) = t;
            }
            *reinterpret_cast<uint4*>(dst + (uint32_t)j * 129u * 16u + (uint32_t)q * 16u) = o;
          }
        }
        if (tr.u == nsub - 1) {
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&full[st]);
          if (++st == NS) { st = 0; ph ^= 1; }
        }
        if (++slot_tr == p.nraw) slot_tr = 0;
        advance(tr);
#ifdef HDN_TC_TIMING
        if (wt_is_a) WT_ADD(t_xa); else WT_ADD(t_xb);
        ++n_sub;
#endif
      }
#ifdef HDN_TC_TIMING
      if (blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 255) && n_sub > 0)
        printf("[wgrad prod t%d] sub-stages %lld (nA %d nB %d)  cp_wait %lld  bar %lld  issue %lld  empty %lld  xform_A %lld  xform_dY %lld (cycles/sub-stage)\n",
               tid, n_sub, nA, nB, t_cp / n_sub, t_bar / n_sub, t_issue / n_sub, t_empty / n_sub, t_xa / n_sub, t_xb / n_sub);
#endif
    } else {
    // ---- direct path (space-to-depth stems)
    for (long long t = split; t < p.n_pos_tiles; t += nsplit) {
      int n_img, d0, h0, w0;
      long long m0;
      if (!tile_valid(t, n_img, d0, h0, w0, m0)) continue;
      produced = true;
      tc::mbar_wait(&empty[st], ph ^ 1);
      int* of = offs + st * OFFS_PER;
      // pixel offsets of the patch (per source) and of the dY tile
      for (int i = tid; i < p.nsrc * p.P + 128; i += WG_PROD) {
        int off = -1;
        if (i < p.nsrc * p.P) {
          const int s = i / p.P, q = i - s * p.P;
          const hdn_src& S = p.src[s];
          if (p.flat) {
            if (m0 + q < p.M) off = q * S.t.ldc;
          } else {
            const int vh = h0 - hh + gth0 + q / p.PW, vw = w0 - hw + gtw0 + q % p.PW;
            if (vh >= 0 && vh < p.H && vw >= 0 && vw < p.W) {
              if (p.s2d) off = (2 * vh * S.W + 2 * vw) * S.t.ldc;
              else {
                const int sh = (S.uh == 2) ? (vh >> 1) : vh, sw = (S.uw == 2) ? (vw >> 1) : vw;
                off = (sh * S.W + sw) * S.t.ldc;
              }
            }
          }
          of[i] = off;
        } else {
          const int r = i - p.nsrc * p.P;
          if (p.flat) {
            if (m0 + r < p.M) off = r * p.dy.ldc;
          } else {
            const int oh = h0 + (r >> 3), ow = w0 + (r & 7);
            if (oh < p.H && ow < p.W) off = (oh * p.W + ow) * p.dy.ldc;
          }
          of[2 * p.P + r] = off;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* base[2] = {nullptr, nullptr};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s < p.nsrc) {
          const hdn_src& S = p.src[s];
          if (p.flat) base[s] = S.t.p + m0 * S.t.ldc + S.t.coff;
          else {
            const int vd = d0 - hd + gdz;
            if (p.s2d) {
              const int l8 = lane & 7;
              const int rd = (p.s2d_quads == 8) ? (l8 >> 2) : 0;
              base[s] = S.t.p + ((long long)n_img * S.D + (p.s2d_quads == 8 ? 2 * vd + rd : vd)) * S.H * S.W * S.t.ldc + S.t.coff +
                        ((l8 >> 1) & 1) * S.W * S.t.ldc + (l8 & 1) * 4 - (ci0 + l8 * 4);
            } else {
              const int sd = (S.ud == 2) ? (vd >> 1) : vd;
              base[s] = S.t.p + ((long long)n_img * S.D + sd) * S.H * S.W * S.t.ldc + S.t.coff;
            }
          }
        }
      }
      uint8_t* dA = sA + st * A_BYTES;
      if (p.split) {
        stage_block(p.src, p.nsrc, base, of, p.P, p.Ppad, ci0, p.Cin, dA, 8, warp, lane, p.s2d ? p.s2d_quads : 0, dA + 8u * p.Ppad * 16u);
      } else {
        stage_block(p.src, p.nsrc, base, of, p.P, p.Ppad, ci0, p.Cin, dA, 8, warp, lane, p.s2d ? p.s2d_quads : 0);
        stage_block(p.src, p.nsrc, base, of, p.P, p.Ppad, ci0 + 64, p.Cin, dA + 8u * p.Ppad * 16u, 8, warp, lane, p.s2d ? p.s2d_quads : 0);
      }
      // dY tile: plain tensor, no prologue
      hdn_src dys;
      dys.t = p.dy; dys.pa = nullptr; dys.pb = nullptr; dys.relu = 0;
      const float* dbase[2];
      dbase[0] = p.flat ? (p.dy.p + m0 * p.dy.ldc + p.dy.coff)
                        : (p.dy.p + ((long long)n_img * p.D + d0) * p.H * p.W * p.dy.ldc + p.dy.coff);
      dbase[1] = nullptr;
      uint8_t* dB = sB + st * B_BYTES;
      for (int cb = 0; cb < p.BN; cb += 64)
        stage_block(&dys, 1, dbase, of + 2 * p.P, 128, 129, co0 + cb, min(p.Cout, co0 + p.BN), dB + (uint32_t)(cb / 8) * 129u * 16u,
                    (p.BN - cb + 7) / 8, warp, lane, 0,
                    p.split ? dB + (uint32_t)((p.BN + cb) / 8) * 129u * 16u : nullptr);
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[st]);
      if (++st == NS) { st = 0; ph ^= 1; }
    }

    }

    // =================================================================== epilogue: TMEM -> dW (vector reductions)
    tc::mbar_wait(acc_full, 0);
    tc::tc_fence_after();
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may read
    const int ci = ci0 + (p.split ? ((q4 * 32 + lane) & 63) : (q4 * 32 + lane));   // bf16x3: rows 64-127 are the tails of rows 0-63
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool vec = (p.Cout % 4 == 0);
    float v[16];
    // a CTA whose tap group never met a valid depth slab issued no MMA: its accumulators are undefined
    for (int g = (warp >> 2); produced && g < p.G; g += 2) {       // warps 0-3 take even taps, warps 4-7 odd
      const int tap = tap0 + g;
      for (int ce = 0; ce < p.BNe; ce += 16) {
        tc::tmem_ld16(taddr + (uint32_t)(g * p.BNe + ce), v);
        const int cc = ce >= p.BN ? ce - p.BN : ce;            // bf16x3: columns BN.. are the tail columns of 0..BN-1
        long long wrow = (long long)tap * p.Cin + ci;        // row of dW [tap][ci]
        bool rok = ci < p.Cin;
        if (p.s2d) {
          // s2d tap (tqd,tqh,tqw) x s2d channel (rd,rh,rw,c)  ->  original tap t = 2*tq + r - 1 (per axis), channel c
          const int c = ci & 3, rw = (ci >> 2) & 1, rh = (ci >> 3) & 1, rd = (ci >> 4) & 1;
          const int tqw = tap & 3, tqh = (tap >> 2) & 3, tqd = tap >> 4;
          const int tw = 2 * tqw + rw - 1, th = 2 * tqh + rh - 1, td = (p.s2d_quads == 8) ? (2 * tqd + rd - 1) : 0;
          const int k7d = (p.s2d_quads == 8) ? 7 : 1;
          rok = rok && c < p.cin_real && tw >= 0 && th >= 0 && td >= 0 && tw < 7 && th < 7 && td < k7d;
          wrow = (((long long)td * 7 + th) * 7 + tw) * p.cin_real + c;
        }
        if (rok) {
          float* q = p.dw + wrow * p.Cout + co0 + cc;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int col = co0 + cc + i;
            if (vec) {
              if (col < p.Cout)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + i), "f"(v[i]), "f"(v[i + 1]),
                             "f"(v[i + 2]), "f"(v[i + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col + j < p.Cout) atomicAdd(q + i + j, v[i + j]);
            }
          }
        }
      }
    }
  } else {
    // =================================================================== MMA issuer
    // all 32 lanes run the loop converged; the tcgen05.mma / commit issue sits under tc::elect_one_sync
    {
      const uint32_t idesc = tc::make_idesc_bf16(128, p.BNe, 1, 1);
      const uint32_t a_lbo = (uint32_t)p.PW * 16u, a_sbo = (uint32_t)p.Ppad * 16u;   // K-group / MN-group strides
      const uint32_t b_lbo = 128u, b_sbo = 129u * 16u;
      int st = 0;
      uint32_t ph = 0, acc = 0;
      const uint64_t adesc_hi = tc::make_smem_desc(0, a_lbo, a_sbo), bdesc_hi = tc::make_smem_desc(0, b_lbo, b_sbo);
      const int row2 = 2 * p.PW;                             // two patch rows (= 16 pixels of the tile) per K step, 16-byte units
      // first tap of the group relative to the group's patch
      const int t20 = tap0 % taps_hw;
      const int twc0 = (p.G == 1) ? 0 : (t20 % p.kw);
      const uint32_t tap_units0 = (uint32_t)((t20 / p.kw - gth0) * p.PW + (t20 % p.kw - gtw0));
      WT_DECL(t_full); WT_DECL(t_mma); WT_DECL(n_tile);
      for (long long t = split; t < p.n_pos_tiles; t += nsplit) {
        int n_img, d0, h0, w0;
        long long m0;
        if (!tile_valid(t, n_img, d0, h0, w0, m0)) continue;
        WT_BEGIN;
        tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        WT_ADD(t_full);
        if (tc::elect_one_sync()) {
        // lean issue loop (one elected lane): descriptors are advanced by integer adds on the 14-bit start-address field
        const uint64_t ad0 = adesc_hi | (uint64_t)((tc::smem_u32(sA + st * A_BYTES) >> 4) & 0x3FFF);
        const uint64_t bd0 = bdesc_hi | (uint64_t)((tc::smem_u32(sB + st * B_BYTES) >> 4) & 0x3FFF);
        uint32_t tap_units = tap_units0;                     // ((th - gth0) * PW + (tw - gtw0)) in 16-byte units
        int twc = twc0;
        uint32_t tm = tmem_base;
        for (int g = 0; g < p.G; ++g, tm += (uint32_t)p.BNe) {
          const uint64_t ad = ad0 + tap_units;
          tc::umma_bf16(tm, ad, bd0, idesc, acc);            // overwrite only on the very first K step of each tap
#pragma unroll
          for (int r = 1; r < 8; ++r) tc::umma_bf16(tm, ad + (uint64_t)(r * row2), bd0 + (uint64_t)(r * 16), idesc, 1u);
          if (++twc == p.kw) { twc = 0; tap_units += (uint32_t)(p.PW - p.kw + 1); } else ++tap_units;
        }
        tc::umma_commit(&empty[st]);
        }
        __syncwarp();
        acc = 1;
        if (++st == NS) { st = 0; ph ^= 1; }
        WT_ADD(t_mma);
#ifdef HDN_TC_TIMING
        ++n_tile;
#endif
      }
      if (tc::elect_one_sync()) tc::umma_commit(acc_full);
      __syncwarp();
#ifdef HDN_TC_TIMING
      if (blockIdx.x == 0 && blockIdx.y == 0 && n_tile > 0 && lane == 0)
        printf("[wgrad mma] tiles %lld  full_wait %lld  issue %lld (cycles/tile)  G %d BNe %d\n", n_tile, t_full / n_tile, t_mma / n_tile, p.G, p.BNe);
#endif
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

struct WgPlan {
  int split, BNe, CI;
  int BN, G, groups, ci_tiles, co_tiles, flat, PH, PW, P, Ppad, tiles_h, tiles_w, tmem_cols, row_mode;
  int nraw, raw_bytes, CW, PHs[2], PWs[2], Ps[2], raw_off[2], ab_off[2], tab_src[2], tab_vq[2], tab_dy, tab_ints;
  long long n_pos_tiles;
  size_t smem;
  int splits;
};

struct WgGeom { int kd, kh, kw, pd_lo, ph_lo, pw_lo, Cin, s2d, quads; };

WgGeom wg_geom(const hdn_conv* c) {
  WgGeom g;
  const int stem = hdn_tc_stem(c);
  if (stem) {
    g.kd = stem == 3 ? 4 : 1; g.kh = 4; g.kw = 4;
    g.pd_lo = stem == 3 ? 2 : 0; g.ph_lo = 2; g.pw_lo = 2;
    g.quads = stem == 3 ? 8 : 4; g.Cin = 4 * g.quads; g.s2d = 1;
  } else {
    g.kd = c->kd; g.kh = c->kh; g.kw = c->kw;
    g.pd_lo = c->kd / 2; g.ph_lo = c->kh / 2; g.pw_lo = c->kw / 2;
    g.Cin = c->Cin; g.s2d = 0; g.quads = 0;
  }
  return g;
}

bool wg_plan(const hdn_conv* c, const WgGeom& gm, WgPlan& best) {
  const int T = gm.kd * gm.kh * gm.kw, taps_hw = gm.kh * gm.kw;
  bool up = false;
  for (int i = 0; i < c->nsrc; ++i) up = up || c->src[i].ud != 1 || c->src[i].uh != 1 || c->src[i].uw != 1;
  const int flat = (T == 1 && !up) ? 1 : 0;
  const int split = c->precision == 2 ? 1 : 0;            // bf16x3: 64 channels (head + tail rows) x 2*BN (head + tail columns)
  const int CI = split ? 64 : 128;
  const int ci_tiles = (gm.Cin + CI - 1) / CI;
  double best_cost = 1e300;
  bool found = false;
  const int cand[3] = {taps_hw, gm.kw, 1};       // taps per group: one depth slab, one kernel row, a single tap
  for (int k = 0; k < 3; ++k) {
    const int G = cand[k];
    if (k > 0 && G == cand[k - 1]) continue;
    int bn_top = (512 / G) / (split ? 2 : 1) / 16 * 16;
    if (bn_top > (split ? 128 : 256)) bn_top = split ? 128 : 256;
    if (bn_top < 16) continue;
    // bf16x3 doubles the dY stage: narrow the column tile until the asynchronous producer's raw ring fits as well
    for (int bn_max = bn_top; bn_max >= 16; bn_max = split ? (bn_max / 2) / 16 * 16 : 0) {
    WgPlan pl;
    pl.split = split; pl.CI = CI;
    pl.G = G; pl.groups = T / G; pl.ci_tiles = ci_tiles; pl.flat = flat;
    pl.co_tiles = (c->Cout + bn_max - 1) / bn_max;
    int bn = (c->Cout + pl.co_tiles - 1) / pl.co_tiles;
    pl.BN = (bn + 15) / 16 * 16;
    pl.BNe = split ? 2 * pl.BN : pl.BN;
    const int rows = (G == taps_hw) ? gm.kh : 1;  // kernel rows covered by a group
    const int cols = (G == 1) ? 1 : gm.kw;
    pl.row_mode = (G == taps_hw) ? 0 : 1;
    pl.PH = 16 + rows - 1;
    pl.PW = 8 + cols - 1;
    if (G == 1) pl.PW = 8;
    pl.P = pl.PH * pl.PW;
    pl.Ppad = pl.P | 1;
    pl.tiles_h = (c->H + 15) / 16;
    pl.tiles_w = (c->W + 7) / 8;
    const long long M = (long long)c->N * c->D * c->H * c->W;
    pl.n_pos_tiles = flat ? (M + 127) / 128 : (long long)c->N * c->D * pl.tiles_h * pl.tiles_w;
    int cols_t = 32;
    while (cols_t < G * pl.BNe) cols_t *= 2;
    pl.tmem_cols = cols_t;
    const size_t a_bytes = 16ull * pl.Ppad * 16, b_bytes = (size_t)((pl.BNe + 7) / 8) * 129 * 16;
    // asynchronous producer geometry (not for the space-to-depth stems)
    pl.CW = (c->nsrc == 2) ? 32 : 64;
    int roff = 0, toff = 0;
    for (int s = 0; s < 2; ++s) {
      pl.PHs[s] = pl.PH; pl.PWs[s] = pl.PW;
      if (s < c->nsrc && !gm.s2d && !flat) {
        if (c->src[s].uh == 2) pl.PHs[s] = pl.PH / 2 + 1;
        if (c->src[s].uw == 2) pl.PWs[s] = pl.PW / 2 + 1;
      }
      pl.Ps[s] = s < c->nsrc ? pl.PHs[s] * pl.PWs[s] : 0;
      pl.raw_off[s] = roff; roff += pl.Ps[s] * (pl.CW + 4) * 4;
      pl.tab_src[s] = toff; toff += pl.Ps[s];
      pl.tab_vq[s] = toff; toff += s < c->nsrc ? pl.P : 0;
    }
    for (int s = 0; s < 2; ++s) { pl.ab_off[s] = roff; roff += s < c->nsrc ? 2 * pl.CW * 4 : 0; }
    if (roff < 128 * 68 * 4) roff = 128 * 68 * 4;        // a dY sub-stage: 128 pixels x (64 + 4 pad) floats
    pl.raw_bytes = (roff + 127) / 128 * 128;
    pl.tab_dy = toff; toff += 128;
    pl.tab_ints = toff;
    const size_t fixed = NS * (a_bytes + b_bytes) + NS * (2ull * pl.P + 128) * 4 + (size_t)NTW * pl.tab_ints * 4 + 128 + 16 +
                         (2 * NS + 1) * 8 + 16;
    pl.nraw = 0;
    if (!gm.s2d) {
      if (fixed + 3ull * pl.raw_bytes <= 226 * 1024) pl.nraw = 3;
      else if (fixed + 2ull * pl.raw_bytes <= 226 * 1024) pl.nraw = 2;
    }
    pl.smem = fixed + (size_t)pl.nraw * pl.raw_bytes;
    if (pl.smem > 226 * 1024) continue;
    double cost = (double)pl.ci_tiles * pl.co_tiles * pl.groups * ((double)pl.P * CI * c->nsrc + 128.0 * pl.BN);
    if (!gm.s2d && pl.nraw == 0) cost *= 4.0;            // falls back to the synchronous producer: much slower
    if (cost < best_cost) { best_cost = cost; best = pl; found = true; }
    }
  }
  if (!found) return false;
  const long long items = (long long)best.ci_tiles * best.co_tiles * best.groups;
  long long splits = 148 / items;
  if (splits < 1) splits = 1;
  if (splits > best.n_pos_tiles) splits = best.n_pos_tiles;
  best.splits = (int)splits;
  return true;
}

}  // namespace

int hdn_wgrad_tc_supported(const hdn_conv* c) {
  if (hdn_tc_stem(c)) {
    if (c->y.ldc % 4 || c->y.coff % 4 || (reinterpret_cast<uintptr_t>(c->y.p) & 15) != 0) return 0;
    WgPlan pl;
    return wg_plan(c, wg_geom(c), pl) ? 1 : 0;
  }
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
  WgPlan pl;
  return wg_plan(c, wg_geom(c), pl) ? 1 : 0;
}

// launch plan of the wgrad kernel (host arithmetic only; see hdn_conv_tc_plan in hdn.h)
int hdn_wgrad_plan_info(const hdn_conv* c, int* out) {
  WgPlan pl;
  if (!wg_plan(c, wg_geom(c), pl)) return HDN_ERR_UNSUPPORTED;
  out[0] = pl.BN; out[1] = pl.co_tiles; out[2] = pl.ci_tiles; out[3] = pl.CW; out[4] = pl.G; out[5] = pl.nraw;
  out[6] = pl.tmem_cols; out[7] = (int)pl.smem; out[8] = pl.flat; out[9] = pl.P; out[10] = pl.split;
  out[11] = hdn_tc_stem(c) ? 1 : 0; out[12] = pl.ci_tiles * pl.co_tiles * pl.groups * pl.splits; out[13] = 1;
  out[14] = pl.BNe; out[15] = pl.CI;
  return HDN_OK;
}

int hdn_conv_wgrad_tc(const hdn_conv* c, float* dw, cudaStream_t st) {
  WgPlan pl;
  const WgGeom gm = wg_geom(c);
  HDN_CHECK_ARG(wg_plan(c, gm, pl), "conv_wgrad tc: no plan for this shape");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { hdn_set_error("conv_wgrad tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
    attr_set = true;
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.N = c->N; p.D = c->D; p.H = c->H; p.W = c->W;
  p.kd = gm.kd; p.kh = gm.kh; p.kw = gm.kw;
  p.pd_lo = gm.pd_lo; p.ph_lo = gm.ph_lo; p.pw_lo = gm.pw_lo;
  p.s2d = gm.s2d; p.s2d_quads = gm.quads; p.cin_real = c->Cin;
  p.Cin = gm.Cin; p.Cout = c->Cout;
  p.split = pl.split; p.BNe = pl.BNe; p.CI = pl.CI;
  p.fastx = hdn_tc_fastx();
  p.BN = pl.BN; p.G = pl.G; p.groups = pl.groups; p.ci_tiles = pl.ci_tiles; p.co_tiles = pl.co_tiles;
  p.flat = pl.flat; p.PH = pl.PH; p.PW = pl.PW; p.P = pl.P; p.Ppad = pl.Ppad; p.row0_mode = pl.row_mode;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.n_pos_tiles = pl.n_pos_tiles;
  p.M = (long long)c->N * c->D * c->H * c->W;
  p.nsrc = c->nsrc; p.src[0] = c->src[0]; p.src[1] = c->src[1];
  p.dy = c->y; p.dw = dw; p.tmem_cols = pl.tmem_cols;
  p.nraw = pl.nraw; p.raw_bytes = pl.raw_bytes; p.CW = pl.CW;
  for (int s = 0; s < 2; ++s) {
    p.PHs[s] = pl.PHs[s]; p.PWs[s] = pl.PWs[s]; p.Ps[s] = pl.Ps[s]; p.raw_off[s] = pl.raw_off[s]; p.ab_off[s] = pl.ab_off[s];
    p.tab_src[s] = pl.tab_src[s]; p.tab_vq[s] = pl.tab_vq[s];
  }
  p.tab_dy = pl.tab_dy; p.tab_ints = pl.tab_ints;
  HDN_CHECK_ARG(pl.n_pos_tiles < (1ll << 31), "conv_wgrad tc: too many position tiles");
  dim3 grid((unsigned)(pl.ci_tiles * pl.co_tiles * pl.groups), (unsigned)pl.splits);
  HDN_LAUNCHED(1), conv_wgrad_tc_kernel<<<grid, WG_THREADS, pl.smem, st>>>(p);
  HDN_CHECK_LAUNCH("conv_wgrad_tc");
  return HDN_OK;
}
