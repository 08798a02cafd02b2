// tcgen05 implicit-GEMM convolution path (precision == 1).  Placeholder until the kernels land.
#include "hdn_common.cuh"
int hdn_tc_supported(const hdn_conv* c, int pass) { (void)c; (void)pass; return 0; }
int hdn_conv_fprop_tc(const hdn_conv* c, cudaStream_t st) { (void)c; (void)st; hdn_set_error("tc path not built"); return HDN_ERR_UNSUPPORTED; }
int hdn_conv_dgrad_tc(const hdn_conv* c, const hdn_dgrad_epi* e, cudaStream_t st) { (void)c; (void)e; (void)st; hdn_set_error("tc path not built"); return HDN_ERR_UNSUPPORTED; }
int hdn_conv_wgrad_tc(const hdn_conv* c, float* dw, cudaStream_t st) { (void)c; (void)dw; (void)st; hdn_set_error("tc path not built"); return HDN_ERR_UNSUPPORTED; }
