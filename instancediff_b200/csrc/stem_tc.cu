// Stem of the drift/noise UNet on tcgen05:  cat([x - mu, mu]) -> 7x7 conv (pad 3), 2 -> 64 channels, bf16 NHWC out.
//
// The CUDA-core version of this layer was FMA-bound (26 GFLOP of fp32 FMAs, 1.08 ms at B = 32, 256x256: more than
// any tensor-core layer of the network).  Here one CTA computes a 16x8-pixel tile as ONE small GEMM
//   D[128 pixels][64] = A[128][224] * Wk[64][224]^T
// whose A rows are built in shared memory from an fp32 patch (in-kernel im2col, no extra HBM traffic):
//   k' = tap*2 + ci for the 49 taps (98 values), k' = 98 / 99 carry the constant 1 (the bias rides in the weights
//   as a bf16 hi/lo pair), padded to 112; the block is stored TWICE: bf16(a) and bf16(a - bf16(a)), so the
//   activations keep ~16 mantissa bits (x_t feeds the SDE recursion directly) while the weights are bf16 like every
//   other layer.  Weights are pre-packed on the host into the shared-memory image (packing.py::pack_stem_weight).
// Persistent CTAs (2 per SM: gather of one overlaps MMA / epilogue of the other), 256 threads: two threads per
// accumulator row split the im2col groups and the output columns.
// Spec: SURVEY.md App. A (init_conv); serves `self.model(x, self.mu, t*scale)`, utils/sde_utils.py:198.
#include "common.cuh"
#include "host_common.h"

namespace idiff {

constexpr int STM_TH = 16, STM_TW = 8;
constexpr int STM_PH = STM_TH + 6, STM_PW = STM_TW + 6;     // 22 x 14 patch
constexpr int STM_PP = 24;                                  // patch row pitch in shared memory (float2 units):
                                                            // == 8 mod 16, so the 8-byte gathers of a half warp
                                                            // (2 tile rows x 8 pixels) hit 32 distinct banks
constexpr int STM_GH = 14;                                  // 8-wide K groups per half (hi / lo): 112 values
constexpr int STM_G = 2 * STM_GH;                           // 28 groups, K = 224 = 14 MMAs of K = 16
constexpr int STM_LBO_A = 128 * 16;                         // one K group of A: 128 rows x 16 B
constexpr int STM_LBO_B = 64 * 16;                          // one K group of B: 64 rows x 16 B
constexpr int STM_OFF_B = STM_G * STM_LBO_A;                // 57344
constexpr int STM_OFF_P = STM_OFF_B + STM_G * STM_LBO_B;    // 86016
constexpr int STM_OFF_BAR = STM_OFF_P + STM_PH * STM_PP * 8;
constexpr int STM_SMEM = STM_OFF_BAR + 16;

int watchdog_stem(int clear) { return watchdog_read_tu(clear); }

// im2col groups G0 .. G0+NG-1 of one pixel row: 4 taps x 2 channels each, stored as bf16 hi and lo parts.  All tap
// offsets are compile-time constants.
template <int G0, int NG>
IDIFF_DEVINL void stem_gather(const float2* __restrict__ pp, uint8_t* sA, int row) {
#pragma unroll
  for (int g = G0; g < G0 + NG; ++g) {
    float v[8];
#pragma unroll
    for (int t4 = 0; t4 < 4; ++t4) {
      const int tap = 4 * g + t4;
      if (tap < 49) {
        const float2 pv = pp[(tap / 7) * STM_PP + (tap % 7)];
        v[2 * t4] = pv.x;
        v[2 * t4 + 1] = pv.y;
      } else {
        v[2 * t4] = v[2 * t4 + 1] = tap == 49 ? 1.f : 0.f;     // k' = 98, 99: the bias rows of the weights
      }
    }
    const uint4 hi = pack_bf16x8(v);
    float h[8];
    unpack_bf16x8(hi, h);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] -= h[e];
    *reinterpret_cast<uint4*>(sA + g * STM_LBO_A + row * 16) = hi;
    *reinterpret_cast<uint4*>(sA + (STM_GH + g) * STM_LBO_A + row * 16) = pack_bf16x8(v);
  }
}

__global__ void __launch_bounds__(256)
stem_tc_kernel(const float* __restrict__ x, const float* __restrict__ mu, const uint4* __restrict__ wpk,
               __nv_bfloat16* __restrict__ out, int H, int W, int tiles_x, int tiles_y, int total) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* sA = sm;
  uint8_t* sB = sm + STM_OFF_B;
  float2* patch = reinterpret_cast<float2*>(sm + STM_OFF_P);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + STM_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7;                  // accumulator row (tile pixel) / which half of the work
  const int ti = row >> 3, tj = row & 7;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  for (int i = tid; i < STM_G * STM_LBO_B / 16; i += 256) reinterpret_cast<uint4*>(sB)[i] = __ldg(wpk + i);
  // the all-zero tail group of each half is written once
  *reinterpret_cast<uint4*>(sA + (part ? STM_G - 1 : STM_GH - 1) * STM_LBO_A + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, 64, 0);
  const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
  uint32_t phase = 0;

  // patch loader split into a register fetch and a shared-memory store: the NEXT tile's 22x14 patch is in flight
  // while the current tile is gathered, multiplied and stored (the serialised version exposed one global-load
  // latency per tile)
  constexpr int NPF = (STM_PH * STM_PW + 255) / 256;            // 2 patch pixels per thread
  float2 pf[NPF];
  auto fetch_patch = [&](int tile) {
    const int tx = tile % tiles_x, r = tile / tiles_x, ty = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty * STM_TH, ox0 = tx * STM_TW;
    const float* xb = x + (size_t)b * H * W;
    const float* mb = mu + (size_t)b * H * W;
#pragma unroll
    for (int k = 0; k < NPF; ++k) {
      const int i = tid + 256 * k;
      const int py = i / STM_PW, px = i - py * STM_PW, iy = oy0 + py - 3, ix = ox0 + px - 3;
      float2 v = make_float2(0.f, 0.f);
      if (i < STM_PH * STM_PW && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        v.y = __ldg(mb + (size_t)iy * W + ix);
        v.x = __ldg(xb + (size_t)iy * W + ix) - v.y;             // channel 0 = x - mu, channel 1 = mu
      }
      pf[k] = v;
    }
  };
  if ((int)blockIdx.x < total) fetch_patch(blockIdx.x);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = tile % tiles_x, r = tile / tiles_x, ty = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty * STM_TH, ox0 = tx * STM_TW;
#pragma unroll
    for (int k = 0; k < NPF; ++k)
      if (tid + 256 * k < STM_PH * STM_PW) {
        const int i = tid + 256 * k, py = i / STM_PW;
        patch[py * STM_PP + (i - py * STM_PW)] = pf[k];
      }
    __syncthreads();
    // in-kernel im2col: this thread's pixel row, its 7 (part 0) or 6 (part 1) of the 13 groups of 4 taps x 2
    // channels, as bf16 hi and lo parts
    const float2* pp = patch + ti * STM_PP + tj;
    if (part == 0) stem_gather<0, 7>(pp, sA, row);
    else stem_gather<7, 6>(pp, sA, row);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < STM_G / 2; ++kk)
          umma_bf16(tmem, umma_desc(a0 + kk * 2 * STM_LBO_A, STM_LBO_A, 128),
                    umma_desc(b0 + kk * 2 * STM_LBO_B, STM_LBO_B, 128), idesc, kk > 0 ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    if (tile + (int)gridDim.x < total) fetch_patch(tile + gridDim.x);
    mbar_wait(bar, phase, 301);
    phase ^= 1u;
    tc_fence_after();
    const int oy = oy0 + ti, ox = ox0 + tj;
    const bool valid = oy < H && ox < W;
    uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)b * H + (valid ? oy : 0)) * W + (valid ? ox : 0)) * 64);
    {
      float v[32];                                             // warps 0-3: columns 0..31, warps 4-7: columns 32..63
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * 32), v);
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[part * 4 + q] = pack_bf16x8(v + 8 * q);
      }
    }
    tc_fence_before();       // the next tile's barriers order these TMEM reads before its first MMA
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace idiff

extern "C" {
using namespace idiff;

int idiff_stem_packed_bytes(void) { return STM_G * STM_LBO_B; }

int idiff_stem_conv7_tc(const float* x, const float* mu, const void* w_packed, void* out, int B, int H, int W,
                        void* stream) {
  IDIFF_REQUIRE(x && mu && w_packed && out && B > 0 && H > 0 && W > 0, "stem_conv7_tc: bad arguments");
  IDIFF_REQUIRE(aligned16(w_packed) && aligned16(out), "stem_conv7_tc: 16 B alignment");
  static DeviceOnce once;
  int num_sms = 0;
  {
    cudaError_t e = per_device_setup(once, &num_sms, [] {
      return cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STM_SMEM);
    });
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "stem_conv7_tc setup: %s", cudaGetErrorString(e));
  }
  const int tiles_x = (W + STM_TW - 1) / STM_TW, tiles_y = (H + STM_TH - 1) / STM_TH;
  const int total = tiles_x * tiles_y * B;
  const int grid = total < 2 * num_sms ? total : 2 * num_sms;
  stem_tc_kernel<<<grid, 256, STM_SMEM, as_stream(stream)>>>(x, mu, (const uint4*)w_packed, (__nv_bfloat16*)out, H, W,
                                                             tiles_x, tiles_y, total);
  return check_launch("stem_conv7_tc");
}

}  // extern "C"
