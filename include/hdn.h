/* hdn.h -- C-ABI of libhdn.so, the sm_100a kernel library behind the H-DenseUNet engine.
 *
 * The reference (xmengli/H-DenseUNet) has no FFI of its own: its hot path is Keras-2.0.8
 * Python calling TensorFlow ops.  Each entry point below replaces the TF op(s) named in
 * its comment (KB = Keras-2.0.8/keras/backend/tensorflow_backend.py in the reference).
 *
 * Conventions
 *   - plain C: POD descriptors, raw device pointers, sizes; no C++/torch types.
 *   - every call returns 0 on success, a negative hdn_status on error; the message is
 *     available from hdn_last_error() (thread-local).  Nothing throws.
 *   - all buffers are caller-owned device memory; launches go to the given cudaStream_t
 *     (passed as void*).  The library keeps no global mutable state except the error string.
 *   - activations are fp32, channels-last 5-D  (N, D, H, W, C); a 2-D tensor has D == 1.
 *     Reference 3-D tensors (N,H,W,S,C) map to D = S (slice axis outermost), which makes the
 *     2-D network's (S,H,W,C) output *be* the 3-D network's input without a copy
 *     (hybridnet.py:355-411).  A tensor argument is a channel window of a wider buffer:
 *     element (n,d,h,w,c) lives at  base[(((n*D+d)*H+h)*W+w)*ldc + coff + c].
 *   - conv weights are [kd][kh][kw][Cin][Cout] fp32 (Keras HWIO / DHWIO with the slice axis
 *     moved first).
 *   - process-wide switches read once from the environment (defaults are the validated forms):
 *       HDN_TC_FASTX=0|1|2  operand-transform form of the tcgen05 kernels (default 2)
 *       HDN_POOL_FAST=0|1   vector max-pool backward (default 1)
 *       HDN_TC_L2PF=0|1     L2::256B prefetch hint on the raw patch copies (default 0, experiment)
 *       HDN_TC_X3FOLD=0|1   folded bf16x3 issue scheme: A_hi x [B_hi | B_lo] as one MMA of N = 2*BN plus A_lo x B_hi
 *                           (2 MMAs per K step instead of 3; layers with BN <= 128; default 1); also hdn_set_switch()
 *       HDN_TC_TMA=0|1|2    fprop / dgrad operand path: bf16 pre-pass + TMA tile loads for no layer / the 3x3(x3) layers /
 *                           all stride-1 layers (default 2); also hdn_set_switch()
 *       HDN_TC_SW128=0|1    TMA mode with 128-byte swizzled K-major operand rows instead of 16-byte chunk planes
 *                           (default 0: measured equal, profiles/r02l_*); also hdn_set_switch()
 *       HDN_WGRAD_TC2=0|1   weight gradients of the 1x3x3 / 3x3x3 convolutions (precision 1) through the bf16 pre-pass +
 *                           TMA tile-load kernel (default 1; 0 = first-generation kernel); also hdn_set_switch()
 *       HDN_TC2_LAYOUT=0|1  shared-memory operand layout of that kernel: 0 = 16-byte chunk planes (SWIZZLE_NONE),
 *                           1 = 128-byte swizzled 64-channel rows
 */
#ifndef HDN_H_
#define HDN_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HDN_OK = 0,
  HDN_ERR_ARG = -1,      /* bad descriptor (shape / pointer / alignment) */
  HDN_ERR_CUDA = -2,     /* CUDA runtime error (message carries cudaGetErrorString) */
  HDN_ERR_UNSUPPORTED = -3
} hdn_status;

const char* hdn_last_error(void);
int hdn_version(void);
/* Set a process-wide switch by its environment-variable name (tests flip kernels inside one process). */
int hdn_set_switch(const char* name, int value);
/* Kernels this thread has launched through the library since it was loaded (every <<<>>> is counted). */
long long hdn_launch_count(void);

typedef struct {
  const float* p;   /* base pointer (device) */
  int ldc;          /* channel stride of the underlying buffer (elements) */
  int coff;         /* first channel of the window */
} hdn_tensor;

/* One A-operand source of a convolution: a stored tensor, an optional per-channel
 * affine + ReLU applied on load (BatchNorm -> Scale -> ReLU folded to y = max(a*x+b,0),
 * KB:1684 + lib/custom_layers.py:68 + KB:2671), and nearest-neighbour up-sampling factors
 * (UpSampling2D/3D, KB:1764-1771,1797-1827) folded into the load index. */
typedef struct {
  hdn_tensor t;
  int D, H, W;            /* stored (pre-upsample) spatial dims */
  int ud, uh, uw;         /* upsample factors, 1 or 2 */
  const float* pa;        /* [Cin] scale, NULL => 1 */
  const float* pb;        /* [Cin] shift, NULL => 0 */
  int relu;
} hdn_src;

/* Convolution geometry.  Output position (n,od,oh,ow) reads the (virtually up-sampled,
 * zero-padded) input at (od*sd - pd + kd_i, ...).  Replaces ZeroPadding + tf.nn.convolution
 * (KB:2020, KB:3158, KB:3307) and, with nsrc == 2, the Add merge in front of it
 * (merge.py:207-211): A = f1(src[0]) + f2(src[1]). */
typedef struct {
  int N, D, H, W;         /* output grid */
  int Cin, Cout;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int nsrc;
  hdn_src src[2];
  const float* w;         /* [kd][kh][kw][Cin][Cout] */
  const float* bias;      /* [Cout] or NULL (KB:3480) */
  hdn_tensor y;           /* output window (fprop) / dY window (dgrad, wgrad) */
  double* stat_sum;       /* [Cout] += sum_m y, or NULL  (feeds tf.nn.moments, KB:1635) */
  double* stat_sq;        /* [Cout] += sum_m y^2 */
  float drop_keep;        /* 1.0 => no dropout; else y = y*mask/keep (KB:2888) */
  uint64_t drop_seed;
  int precision;          /* 0: fp32 FMA (parity path)  1: tcgen05 bf16 operands, fp32 accumulate
                             2: tcgen05 "bf16x3": operands split into bf16 head + tail, 3 MMAs per step
                                (hi*hi + lo*hi + hi*lo), fp32 accumulate -- fp32-grade results */
  void* ws;               /* precision 1/2: device scratch for the packed bf16 weight blocks of this call */
  int64_t ws_bytes;       /*   >= hdn_conv_tc_workspace(c, pass); may be shared by all calls of one stream */
} hdn_conv;

/* Epilogue of a data-gradient: given dz = dL/d(prologue output) it forms
 * du = dz * [a*x+b > 0] (ReLU backward), accumulates S1[c] += sum du, S2[c] += sum du*(x-center[c])
 * (everything the BN / Scale parameter gradients need; center = the BN mean keeps the sums
 * free of cancellation), and then either
 *   mode 0: dx (+)= a[c] * du            (inference-mode BN: the affine is a constant)
 *   mode 1: du_out (+)= du               (training-mode BN: hdn_bn_bwd_apply finishes it)
 *   mode 2: this source needs no gradient; skipped */
typedef struct {
  hdn_tensor dx;          /* gradient window of the stored source (mode 0) */
  float* du;              /* dense [M_src][Cin] (mode 1) */
  int mode;
  int accumulate;         /* 0: overwrite, 1: += */
  double* s1;             /* [Cin] or NULL */
  double* s2;
  const float* center;    /* [Cin] or NULL (=> 0) */
} hdn_dgrad_epi;

int hdn_conv_fprop(const hdn_conv* c, void* stream);
/* dgrad: c->y is dY.  One epilogue per source. */
int hdn_conv_dgrad(const hdn_conv* c, const hdn_dgrad_epi* epi, void* stream);
/* wgrad: dw [kd][kh][kw][Cin][Cout] += A^T dY ; dbias [Cout] += sum_m dY (NULL to skip). */
int hdn_conv_wgrad(const hdn_conv* c, float* dw, float* dbias, void* stream);
/* 1 if the tcgen05 path (precision == 1) takes this descriptor for the given pass, else 0. */
int hdn_conv_tc_supported(const hdn_conv* c, int pass /*0 fprop 1 dgrad 2 wgrad*/);
/* Bytes of hdn_conv.ws the tcgen05 path needs for this descriptor and pass (0 if unsupported). */
int64_t hdn_conv_tc_workspace(const hdn_conv* c, int pass);

/* Launch plan the tcgen05 path would use for this descriptor and pass (host arithmetic only, no device
 * access): out16 = { BN (column tile), column tiles, K blocks (fprop/dgrad) or input-channel tiles (wgrad),
 * channels per stage, weight-ring depth (fprop/dgrad) or taps per group (wgrad), raw fp32 ring depth,
 * TMEM columns, dynamic shared memory bytes, flat (1x1x1) flag, patch pixels, bf16x3 flag, space-to-depth
 * flag, work items, copy-list-fits flag, GEMM K (fprop/dgrad) or GEMM N (wgrad), GEMM columns (fprop/dgrad)
 * or input channels per CTA (wgrad) }.  Used by the host-side tests to check every convolution of a model
 * against the SM's limits (227 KB shared memory, 512 TMEM columns) without a GPU. */
int hdn_conv_tc_plan(const hdn_conv* c, int pass, int32_t* out16);

/* Pooling (KB:3354-3432).  kind 0: max 3x3(x3)/2 after zero-pad 1 (the ZeroPadding + VALID
 * max-pool of hybridnet.py:128-129,215-216; input passes through the src prologue first);
 * kind 1: average 2x2 stride 2 over H,W (AveragePooling2D (2,2) / AveragePooling3D (2,2,1)). */
typedef struct {
  int kind;
  int N, D, H, W, C;      /* output grid */
  int pool_d;             /* kind 0: 1 => also pool depth (3-D), 0 => 2-D */
  hdn_src src;
  hdn_tensor y;           /* output (fwd) / dY (bwd) */
  unsigned char* argidx;  /* kind 0, optional: dense [N*D*H*W][C] arg-max tap (scan order d,h,w) written by
                             hdn_pool_fwd and read by hdn_pool_bwd; NULL => backward recomputes the windows */
} hdn_pool;
int hdn_pool_fwd(const hdn_pool* p, void* stream);
int hdn_pool_bwd(const hdn_pool* p, const hdn_dgrad_epi* epi, void* stream);

/* BatchNorm (+ optional Scale) folding.  mode 1 = training statistics from (sum, sumsq, count),
 * biased variance, and moving-average update  mov -= (mov-batch)*(1-momentum)  (KB:1635,
 * KNORM:179-185, KB:915-927);  mode 0 = moving statistics (KB:1684).
 * Produces a = gs*g*rstd, b = gs*(beta - mean*g*rstd) + bs, and saves mean / rstd. */
typedef struct {
  int C, mode;
  double count;
  const double* sum; const double* sumsq;
  float* mov_mean; float* mov_var;
  const float* gamma; const float* beta;       /* BN */
  const float* sgamma; const float* sbeta;     /* Scale or NULL */
  float eps, momentum;
  float* a; float* b; float* mean; float* rstd;
} hdn_bn_fold_t;
int hdn_bn_fold(const hdn_bn_fold_t* f, void* stream);

/* Parameter gradients of BN/Scale from S1 = sum du, S2 = sum du*(x-mean); in training mode also
 * the per-channel coefficients (k0,k1,k2) with dx = k0*du + k1*(x-mean) + k2 used by
 * hdn_bn_bwd_apply. */
typedef struct {
  int C, mode;
  double count;
  const double* s1; const double* s2;
  const float* mean; const float* rstd;
  const float* gamma; const float* beta; const float* sgamma;
  float* dgamma; float* dbeta; float* dsgamma; float* dsbeta;   /* += ; any may be NULL */
  float* k0; float* k1; float* k2;                               /* training mode only */
} hdn_bn_grad_t;
int hdn_bn_param_grad(const hdn_bn_grad_t* g, void* stream);
/* dx (+)= k0[c]*du[m][c] + k1[c]*(x[m][c]-mean[c]) + k2[c] over M rows. */
int hdn_bn_bwd_apply(const float* du, hdn_tensor x, hdn_tensor dx, int64_t M, int C,
                     const float* k0, const float* k1, const float* k2, const float* mean,
                     int accumulate, void* stream);

/* Per-channel sum and sum of squares of a window (batch statistics of pooled features that
 * no convolution epilogue produced): sum[c] += sum_m y, sq[c] += sum_m y^2. */
int hdn_col_stats(hdn_tensor y, int64_t M, int C, double* sum, double* sq, void* stream);

/* In-place dropout backward on a gradient window: g *= mask/keep (same hash as fprop). */
int hdn_dropout_bwd(hdn_tensor g, int64_t M, int C, float keep, uint64_t seed, void* stream);

/* Weighted soft-max cross entropy, forward + gradient in one pass (loss.py:5-46).
 * logits/dlogits (N,D,H,W,3); labels float (N,D,H,W); only depth d in [d0,d1) counts
 * (loss.py:6-7 crop).  Pass 1 accumulates acc[0] = sum w*log p, acc[1] = #counted voxels;
 * pass 2 writes dlogits = w[y]*(p - onehot)*[p_y >= 1e-10] / count * gscale. */
int hdn_wce_accum(const float* logits, const float* labels, int64_t N, int D, int64_t HW,
                  int d0, int d1, double* acc, void* stream);
int hdn_wce_grad(const float* logits, const float* labels, float* dlogits, int64_t N, int D,
                 int64_t HW, int d0, int d1, const double* acc, float gscale, void* stream);

/* Hybrid glue (hybridnet.py:385-396, 409-411).
 * triplets: vol (B,S,H,W) -> out (B*S,1,H,W,ldc), channel k < 3 of slice s = vol[clamp(s-1+k)], channels
 *           3..ldc-1 zero (ldc = 4 keeps pixels 16-byte aligned for the tensor-core stem).
 * cat4:     out (B,S,H,W,4) = [vol, 250*logits2d(3)];  cat4_bwd: dlogits += 250*dout[...,1:4]. */
int hdn_triplets(const float* vol, float* out, int B, int S, int64_t HW, int ldc, void* stream);
int hdn_cat4(const float* vol, const float* logits, float* out, int64_t M, float k, void* stream);
int hdn_cat4_bwd(const float* dout, float* dlogits, int64_t M, float k, int accumulate, void* stream);

/* Nesterov SGD over a flat arena (optimizers.py:172-181): v = mu*m - lr*g*gs; m = v;
 * p += mu*v - lr*g*gs. */
int hdn_sgd_nesterov(float* p, const float* g, float* m, int64_t n, float lr, float mu,
                     float gscale, void* stream);
/* Data-parallel step over NVLink peer memory: rank r owns elements [lo,hi) of the flat
 * arena, sums the `world` peer gradient arenas (peer_g[i] are peer-mapped device pointers),
 * applies the Nesterov update to its shard and pushes the updated parameters into every
 * peer's parameter arena (replaces multi_gpu.py:7-69 + the implicit TF gradient sum). */
int hdn_dp_reduce_sgd(float* const* peer_p, const float* const* peer_g, float* m_local,
                      int world, int rank, int64_t lo, int64_t hi, float lr, float mu,
                      float gscale, void* stream);

/* Step flags of the data-parallel exchange: every process owns a row of `world` uint32 slots inside its peer-mapped
 * arena.  hdn_dp_signal stores `value` into slot [rank] of every peer's row after a system-scope fence (stream-ordered
 * after this GPU's gradient / parameter writes); hdn_dp_wait blocks the stream until every slot of the local row has
 * reached `value`.  They replace the host barriers around hdn_dp_reduce_sgd (the reference has no counterpart: its
 * towers live in one TF session, multi_gpu.py:35-53). */
int hdn_dp_signal(unsigned int* const* peer_flags, int world, int rank, unsigned int value, void* stream);
int hdn_dp_wait(const unsigned int* flags, int world, unsigned int value, void* stream);

/* Host layout -> engine layout of a staged volume: in (N,H,W,S) float32 or int16 (the reference's
 * (b, size, size, cols, 1) arrays, train_hybrid.py:66-133) -> out (N,S,H,W) float32. */
int hdn_layout_nhws_to_nshw(const void* in, float* out, int N, int H, int W, int S, int is_int16, void* stream);

/* Sliding-window accumulation (lib/funcs.py:28-47): softmax over 3 logits of a window
 * (1,S,H,W,3), slices 1..S-2 added into score (Z,H,W,2) [classes 1,2] and count (Z) at z0+1. */
int hdn_window_accumulate(const float* logits, float* score, int* count, int S, int64_t HW,
                          int z0, void* stream);
int hdn_window_finalize(float* score, const int* count, int Z, int64_t HW, void* stream);

/* Post-processing of the probability volumes (test.py:71-115, lib/funcs.py:138-153), volumes in the reference's C-order
 * (X, Y, Z) layout, one uint8 per voxel; ws >= 8 * X*Y*Z + 8 bytes of device scratch (labels + sizes / flags).
 *   threshold : liver = (s_liver >= t_liver) | tumor, tumor = s_tumor >= t_tumor            (test.py:73-77)
 *   dilate    : ndimage.binary_dilation, 6-neighbour structure, one iteration               (test.py:62,95)
 *   largest_component : measure.label (26-connectivity) + regionprops area + box.index(max) (test.py:83-91,96-103)
 *   fill_holes: ndimage.binary_fill_holes (6-connected background not reaching the border)  (test.py:104,109,112)
 *   and / compose: Segmask * liver_labels; liver_res[Segmask == 1] = 2                      (test.py:108,113) */
int hdn_post_threshold(const float* score_liver, const float* score_tumor, unsigned char* liver, unsigned char* tumor,
                       int64_t n, float thres_liver, float thres_tumor, void* stream);
int hdn_post_dilate(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* stream);
int hdn_post_largest_component(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* ws, int64_t ws_bytes,
                               void* stream);
int hdn_post_fill_holes(const unsigned char* in, unsigned char* out, int X, int Y, int Z, void* ws, int64_t ws_bytes, void* stream);
int hdn_post_and(const unsigned char* a, const unsigned char* b, unsigned char* out, int64_t n, void* stream);
int hdn_post_compose(const unsigned char* liver, const unsigned char* tumor, unsigned char* out, int64_t n, void* stream);

/* One training sample cut out of a volume that is resident on the device (train_hybrid.py:40-98, train_2ddense.py:40-69:
 * crop -> mean subtraction -> one of 8 flips / rotations -> skimage `resize` of the label map (order 0, mode 'edge') and
 * of the image (order 3, mode 'constant', cval 0, clip=True, preserve_range=True) to the network size).
 * The volume lies slice-major (S, H, W) on the device (the reference holds (H, W, S) host arrays; transposed once at
 * upload); `vol` is float32 (vol_i16 == 0) or int16, `seg` uint8.
 * Crop = rows [a0, a0+ch) x columns [b0, b0+cw) x slices [c0, c0+cs).  The flipped crop F is F[i][j] =
 * crop[m00*i + m01*j + o0][m10*i + m11*j + o1] (a signed permutation; identity = {1,0,0,1,0,0}).
 * Outputs: x[s*xs_s + y*xs_h + x*xs_w] (fp32, s < cs, y < out_h, x < out_w) and, for slices [ys0, ys0+yns),
 * y[(s-ys0)*out_h*out_w + y*out_w + x] (fp32 class index); counts[3] (device, int32, may be null) += histogram of the
 * label values written (train_hybrid.py:126-131 discards a batch that misses a class).
 * scratch: >= 16 bytes of device memory (crop minimum / maximum for the clip). */
typedef struct {
  const void* vol;
  const unsigned char* seg;
  int32_t vol_i16;
  int32_t VS, VH, VW;
  int32_t a0, b0, c0;
  int32_t ch, cw, cs;
  int32_t m00, m01, m10, m11, o0, o1;
  float mean;
  int32_t out_h, out_w;
  int64_t xs_s, xs_h, xs_w;
  int32_t ys0, yns;
} hdn_aug;
int hdn_aug_sample(const hdn_aug* a, float* x, float* y, int32_t* counts, void* scratch, void* stream);

/* Raw device memory + CUDA IPC for the parameter / gradient arenas that the per-GPU
 * processes map into each other (handle = 64 bytes, cudaIpcMemHandle_t). */
int hdn_dev_malloc(void** out, int64_t bytes);
int hdn_dev_free(void* p);
int hdn_ipc_get_handle(void* p, unsigned char* handle64);
int hdn_ipc_open(const unsigned char* handle64, void** out);
int hdn_ipc_close(void* p);

#ifdef __cplusplus
}
#endif
#endif
