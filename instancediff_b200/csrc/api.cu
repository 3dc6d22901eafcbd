// Error plumbing + watchdog word shared by every kernel file of libidiff_sm100.so.
#include "host_common.h"

namespace idiff {
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return IDIFF_OK;
}
}  // namespace idiff

extern "C" {

int idiff_abi_version(void) { return 1; }

int idiff_sizeof_gemm_params(void) { return (int)sizeof(idiff_gemm_params); }
int idiff_sizeof_gn_fuse(void) { return (int)sizeof(idiff_gn_fuse); }

const char* idiff_last_error(void) { return idiff::g_err; }

int idiff_watchdog_status(int clear) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return idiff::fail(IDIFF_ERR_CUDA, "watchdog sync: %s", cudaGetErrorString(e));
  const int a = idiff::watchdog_conv(clear), b = idiff::watchdog_attn(clear), c = idiff::watchdog_stem(clear),
            d = idiff::watchdog_lattn(clear), r = idiff::watchdog_rowpair(clear);
  if (a < 0 || b < 0 || c < 0 || d < 0 || r < 0) return idiff::fail(IDIFF_ERR_CUDA, "watchdog read failed");
  const int v = a != 0 ? a : (b != 0 ? b : (c != 0 ? c : (d != 0 ? d : r)));
  if (v != 0) idiff::fail(IDIFF_ERR_WATCHDOG, "device pipeline wait timed out at site %d", v);
  return v;
}

}  // extern "C"
