// Tensor-core (tcgen05 / TMEM) implicit-GEMM convolution — placeholder entry points until the
// kernels land; they fail loudly and the Python side keeps use_tc off.
#include "common.cuh"

extern "C" size_t ttg_pack_weight_tc_bytes(int Cout, int Cin, int ksize) { return (size_t)Cout * Cin * ksize * ksize * 2; }
extern "C" int ttg_pack_weight_tc(const float*, void*, int, int, int, int, void*) {
  return ttg_set_error(TTG_ERR_UNSUPPORTED, "conv2d_tc: not built in this revision");
}
extern "C" int ttg_conv2d_tc(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, void*) {
  return ttg_set_error(TTG_ERR_UNSUPPORTED, "conv2d_tc: not built in this revision");
}
extern "C" int ttg_conv2d_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, int, void*, void*) {
  return ttg_set_error(TTG_ERR_UNSUPPORTED, "conv2d_wgrad_tc: not built in this revision");
}
extern "C" size_t ttg_conv2d_wgrad_tc_workspace_bytes(int, int, int) { return 16; }
