// extern "C" entry points of libhdn.so: error string, convolution dispatch (fp32 FMA parity
// path vs tcgen05 path), raw device-memory + CUDA-IPC helpers for the data-parallel arenas.
#include <stdarg.h>
#include <string.h>
#include "hdn_common.cuh"

static thread_local char g_err[512] = "";

void hdn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int hdn_validate_conv(const hdn_conv* c);
int hdn_conv_fprop_simt(const hdn_conv* c, cudaStream_t st);
int hdn_conv_dgrad_simt(const hdn_conv* c, const hdn_dgrad_epi* epi, cudaStream_t st);
int hdn_conv_wgrad_simt(const hdn_conv* c, float* dw, cudaStream_t st);
int hdn_colsum(hdn_tensor y, int64_t M, int C, float* out, cudaStream_t st);
// tcgen05 path (conv_tc.cu)
int hdn_tc_supported(const hdn_conv* c, int pass);
long long hdn_tc_workspace_bytes(const hdn_conv* c, int pass);
int hdn_conv_fprop_tc(const hdn_conv* c, cudaStream_t st);
int hdn_conv_dgrad_tc(const hdn_conv* c, const hdn_dgrad_epi* epi, cudaStream_t st);
int hdn_conv_wgrad_tc(const hdn_conv* c, float* dw, cudaStream_t st);
// second-generation weight gradient: bf16 operand pre-pass + TMA tile loads + tcgen05 (conv_tc2_wgrad.cu)
int hdn_wgrad_tc2_enabled();
void hdn_wgrad_tc2_set(int v);
void hdn_tc2_layout_set(int v);
void hdn_tc_tma_set(int v);
void hdn_tc_sw128_set(int v);
void hdn_tc_x3fold_set(int v);
int hdn_wgrad_tc2_supported(const hdn_conv* c);
long long hdn_wgrad_tc2_workspace(const hdn_conv* c);
int hdn_wgrad_tc2_plan_info(const hdn_conv* c, int* out);
int hdn_conv_wgrad_tc2(const hdn_conv* c, float* dw, cudaStream_t st);
static bool use_wgrad_tc2(const hdn_conv* c) { return hdn_wgrad_tc2_enabled() && hdn_wgrad_tc2_supported(c); }

extern "C" const char* hdn_last_error(void) { return g_err; }
extern "C" int hdn_version(void) { return 200; }   // 200: tc2 weight gradient (TMA tile loads), hdn_set_switch, post-processing, layout kernels

extern "C" int hdn_conv_tc_supported(const hdn_conv* c, int pass) {
  if (!c || hdn_validate_conv(c) != HDN_OK) return 0;
  return hdn_tc_supported(c, pass);
}

extern "C" int64_t hdn_conv_tc_workspace(const hdn_conv* c, int pass) {
  if (!c || hdn_validate_conv(c) != HDN_OK) return 0;
  if (pass == 2 && hdn_tc_supported(c, 2) && use_wgrad_tc2(c)) return (int64_t)hdn_wgrad_tc2_workspace(c);
  return (int64_t)hdn_tc_workspace_bytes(c, pass);
}

thread_local long long hdn_tl_launches = 0;
extern "C" long long hdn_launch_count(void) { return hdn_tl_launches; }

// Process-wide switches (the environment variables of the same name set the defaults, see the top of hdn.h).
extern "C" int hdn_set_switch(const char* name, int value) {
  HDN_CHECK_ARG(name != nullptr, "set_switch: null name");
  if (!strcmp(name, "HDN_WGRAD_TC2")) { hdn_wgrad_tc2_set(value); return HDN_OK; }
  if (!strcmp(name, "HDN_TC2_LAYOUT")) { hdn_tc2_layout_set(value); return HDN_OK; }
  if (!strcmp(name, "HDN_TC_TMA")) { hdn_tc_tma_set(value); return HDN_OK; }
  if (!strcmp(name, "HDN_TC_SW128")) { hdn_tc_sw128_set(value); return HDN_OK; }
  if (!strcmp(name, "HDN_TC_X3FOLD")) { hdn_tc_x3fold_set(value); return HDN_OK; }
  hdn_set_error("set_switch: unknown switch %s", name);
  return HDN_ERR_ARG;
}

int hdn_tc_plan_info(const hdn_conv* c, int pass, int* out);
int hdn_wgrad_plan_info(const hdn_conv* c, int* out);
extern "C" int hdn_conv_tc_plan(const hdn_conv* c, int pass, int32_t* out16) {
  int rc = hdn_validate_conv(c);
  if (rc) return rc;
  HDN_CHECK_ARG(out16 != nullptr && pass >= 0 && pass <= 2, "conv_tc_plan: bad arguments");
  if (!hdn_tc_supported(c, pass)) { hdn_set_error("conv_tc_plan: tcgen05 path does not take this shape"); return HDN_ERR_UNSUPPORTED; }
  if (pass == 2 && use_wgrad_tc2(c)) return hdn_wgrad_tc2_plan_info(c, out16);
  return pass == 2 ? hdn_wgrad_plan_info(c, out16) : hdn_tc_plan_info(c, pass, out16);
}

extern "C" int hdn_conv_fprop(const hdn_conv* c, void* stream) {
  int rc = hdn_validate_conv(c);
  if (rc) return rc;
  if (c->precision == 1 || c->precision == 2) {
    if (!hdn_tc_supported(c, 0)) { hdn_set_error("conv_fprop: tcgen05 path does not take this shape"); return HDN_ERR_UNSUPPORTED; }
    return hdn_conv_fprop_tc(c, (cudaStream_t)stream);
  }
  return hdn_conv_fprop_simt(c, (cudaStream_t)stream);
}

extern "C" int hdn_conv_dgrad(const hdn_conv* c, const hdn_dgrad_epi* epi, void* stream) {
  int rc = hdn_validate_conv(c);
  if (rc) return rc;
  HDN_CHECK_ARG(epi != nullptr, "conv_dgrad: null epilogue");
  for (int i = 0; i < c->nsrc; ++i) {
    if (epi[i].mode == 2) continue;
    HDN_CHECK_ARG(epi[i].mode == 0 || epi[i].mode == 1, "conv_dgrad: epilogue %d has bad mode %d", i, epi[i].mode);
    HDN_CHECK_ARG(epi[i].mode == 0 ? epi[i].dx.p != nullptr : epi[i].du != nullptr,
                  "conv_dgrad: epilogue %d has no destination", i);
    HDN_CHECK_ARG((epi[i].s1 == nullptr) == (epi[i].s2 == nullptr), "conv_dgrad: s1/s2 must both be set or both NULL");
  }
  if (c->precision == 1 || c->precision == 2) {
    if (!hdn_tc_supported(c, 1)) { hdn_set_error("conv_dgrad: tcgen05 path does not take this shape"); return HDN_ERR_UNSUPPORTED; }
    return hdn_conv_dgrad_tc(c, epi, (cudaStream_t)stream);
  }
  return hdn_conv_dgrad_simt(c, epi, (cudaStream_t)stream);
}

extern "C" int hdn_conv_wgrad(const hdn_conv* c, float* dw, float* dbias, void* stream) {
  int rc = hdn_validate_conv(c);
  if (rc) return rc;
  HDN_CHECK_ARG(dw != nullptr, "conv_wgrad: null dw");
  if (dbias) {
    rc = hdn_colsum(c->y, (int64_t)c->N * c->D * c->H * c->W, c->Cout, dbias, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (c->precision == 1 || c->precision == 2) {
    if (!hdn_tc_supported(c, 2)) { hdn_set_error("conv_wgrad: tcgen05 path does not take this shape"); return HDN_ERR_UNSUPPORTED; }
    if (use_wgrad_tc2(c)) return hdn_conv_wgrad_tc2(c, dw, (cudaStream_t)stream);
    return hdn_conv_wgrad_tc(c, dw, (cudaStream_t)stream);
  }
  return hdn_conv_wgrad_simt(c, dw, (cudaStream_t)stream);
}

// ---- raw device memory + IPC (arenas shared between the per-GPU processes) ----------------
extern "C" int hdn_dev_malloc(void** out, int64_t bytes) {
  HDN_CHECK_ARG(out && bytes > 0, "dev_malloc: bad arguments");
  cudaError_t e = cudaMalloc(out, (size_t)bytes);
  if (e != cudaSuccess) { hdn_set_error("cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  return HDN_OK;
}
extern "C" int hdn_dev_free(void* p) {
  cudaError_t e = cudaFree(p);
  if (e != cudaSuccess) { hdn_set_error("cudaFree: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  return HDN_OK;
}
extern "C" int hdn_ipc_get_handle(void* p, unsigned char* handle64) {
  HDN_CHECK_ARG(p && handle64, "ipc_get_handle: null");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { hdn_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  memcpy(handle64, &h, sizeof(h));
  return HDN_OK;
}
extern "C" int hdn_ipc_open(const unsigned char* handle64, void** out) {
  HDN_CHECK_ARG(handle64 && out, "ipc_open: null");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { hdn_set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  return HDN_OK;
}
extern "C" int hdn_ipc_close(void* p) {
  cudaError_t e = cudaIpcCloseMemHandle(p);
  if (e != cudaSuccess) { hdn_set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e)); return HDN_ERR_CUDA; }
  return HDN_OK;
}
