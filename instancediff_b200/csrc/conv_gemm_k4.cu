// 4x4 stride-2 (pad 1) downsampling instantiations of the implicit-GEMM engine.
#include "conv_gemm_kernel.cuh"

namespace idiff {

cudaError_t launch_conv_k4(const KArgs& a, int amode, int grid, int smem, cudaStream_t st) {
  if (a.p.epi != IDIFF_EPI_PLAIN || amode != AMODE_NONE) return cudaErrorNotSupported;
  switch (a.p.NT) {
    case 64: return launch_one<64, 4, IDIFF_EPI_PLAIN, AMODE_NONE>(a, grid, smem, st);
    case 128: return launch_one<128, 4, IDIFF_EPI_PLAIN, AMODE_NONE>(a, grid, smem, st);
    case 256: return launch_one<256, 4, IDIFF_EPI_PLAIN, AMODE_NONE>(a, grid, smem, st);
    default: return cudaErrorNotSupported;
  }
}
int watchdog_conv_k4(int clear) { return watchdog_read_tu(clear); }

}  // namespace idiff
