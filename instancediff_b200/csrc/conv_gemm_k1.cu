// 1x1 / linear-layer instantiations of the implicit-GEMM engine: plain (raw / affine / affine+SiLU A operand),
// q-softmax (linear-attention qkv), GEGLU (feed-forward) and output-LayerNorm epilogues.
#include "conv_gemm_kernel.cuh"

namespace idiff {

template <int NT>
static cudaError_t launch_plain(const KArgs& a, int amode, int grid, int smem, cudaStream_t st) {
  switch (amode) {
    case AMODE_NONE: return launch_one<NT, 1, IDIFF_EPI_PLAIN, AMODE_NONE>(a, grid, smem, st);
    case AMODE_AFFINE: return launch_one<NT, 1, IDIFF_EPI_PLAIN, AMODE_AFFINE>(a, grid, smem, st);
    case AMODE_AFFINE_SILU: return launch_one<NT, 1, IDIFF_EPI_PLAIN, AMODE_AFFINE_SILU>(a, grid, smem, st);
    default: return cudaErrorNotSupported;
  }
}

cudaError_t launch_conv_k1(const KArgs& a, int amode, int grid, int smem, cudaStream_t st) {
  const int NT = a.p.NT;
  switch (a.p.epi) {
    case IDIFF_EPI_PLAIN:
      return NT == 64 ? launch_plain<64>(a, amode, grid, smem, st)
             : NT == 128 ? launch_plain<128>(a, amode, grid, smem, st)
                         : launch_plain<256>(a, amode, grid, smem, st);
    case IDIFF_EPI_QSOFTMAX:
      if (NT != 128 || amode != AMODE_NONE) return cudaErrorNotSupported;
      return launch_one<128, 1, IDIFF_EPI_QSOFTMAX, AMODE_NONE>(a, grid, smem, st);
    case IDIFF_EPI_GEGLU:
      if (NT != 256 || amode != AMODE_NONE) return cudaErrorNotSupported;
      return launch_one<256, 1, IDIFF_EPI_GEGLU, AMODE_NONE>(a, grid, smem, st);
    case IDIFF_EPI_LN_OUT:
      if (amode != AMODE_NONE) return cudaErrorNotSupported;
      return NT == 64 ? launch_one<64, 1, IDIFF_EPI_LN_OUT, AMODE_NONE>(a, grid, smem, st)
             : NT == 128 ? launch_one<128, 1, IDIFF_EPI_LN_OUT, AMODE_NONE>(a, grid, smem, st)
                         : launch_one<256, 1, IDIFF_EPI_LN_OUT, AMODE_NONE>(a, grid, smem, st);
    default: return cudaErrorNotSupported;
  }
}
int watchdog_conv_k1(int clear) { return watchdog_read_tu(clear); }

}  // namespace idiff
