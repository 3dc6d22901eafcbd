// 3x3 (pad 1) instantiations of the implicit-GEMM engine: plain epilogue, raw or affine+SiLU A operand.
#include "conv_gemm_kernel.cuh"

namespace idiff {

template <int NT>
static cudaError_t launch_nt(const KArgs& a, int amode, int grid, int smem, cudaStream_t st) {
  switch (amode) {
    case AMODE_NONE: return launch_one<NT, 3, IDIFF_EPI_PLAIN, AMODE_NONE>(a, grid, smem, st);
    case AMODE_AFFINE_SILU: return launch_one<NT, 3, IDIFF_EPI_PLAIN, AMODE_AFFINE_SILU>(a, grid, smem, st);
    default: return cudaErrorNotSupported;
  }
}

cudaError_t launch_conv_k3(const KArgs& a, int amode, int grid, int smem, cudaStream_t st) {
  if (a.p.epi != IDIFF_EPI_PLAIN) return cudaErrorNotSupported;
  switch (a.p.NT) {
    case 64: return launch_nt<64>(a, amode, grid, smem, st);
    case 128: return launch_nt<128>(a, amode, grid, smem, st);
    case 256: return launch_nt<256>(a, amode, grid, smem, st);
    default: return cudaErrorNotSupported;
  }
}
int watchdog_conv_k3(int clear) { return watchdog_read_tu(clear); }

}  // namespace idiff
