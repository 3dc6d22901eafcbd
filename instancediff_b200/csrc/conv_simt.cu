// CUDA-core convolutions for the two shapes that are not tensor-core shaped, plus a plain reference
// convolution used only by the tests to validate the tcgen05 engine.
//   stem : cat([x - mu, mu]) -> 7x7, 2 -> nf channels   (K = 98, 0.4 % of the FLOPs, HBM/LSU bound)
//   head : 3x3, nf -> 1 channel                          (9 FLOP/B, HBM bound)
// Spec: SURVEY.md App. A; serves `self.model(x, self.mu, t*scale)`, utils/sde_utils.py:198.
#include "common.cuh"
#include "host_common.h"

namespace idiff {

// ---- stem: 16x16 pixel tile per CTA, one pixel x 64 output channels per thread ---------------------
constexpr int ST = 16, SK = 7, SP = ST + SK - 1;  // 22
template <int NOUT>
__global__ void __launch_bounds__(256)
stem_conv7_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ w,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W) {
  __shared__ float patch[2][SP][SP + 1];
  __shared__ __align__(16) float ws[SK * SK * 2][NOUT];      // [tap*2+ci][co]
  const int b = blockIdx.z, oy0 = blockIdx.y * ST, ox0 = blockIdx.x * ST;
  const float* xb = x + (size_t)b * H * W;
  const float* mb = mu + (size_t)b * H * W;
  for (int i = threadIdx.x; i < SP * SP; i += 256) {
    const int py = i / SP, px = i - py * SP, iy = oy0 + py - 3, ix = ox0 + px - 3;
    float xv = 0.f, mv = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      mv = mb[(size_t)iy * W + ix];
      xv = xb[(size_t)iy * W + ix] - mv;                     // channel 0 = x - mu, channel 1 = mu
    }
    patch[0][py][px] = xv;
    patch[1][py][px] = mv;
  }
  // w: [NOUT][7][7][2] -> ws[(ky*7+kx)*2+ci][co]
  for (int i = threadIdx.x; i < SK * SK * 2 * NOUT; i += 256) {
    const int co = i / (SK * SK * 2), r = i - co * (SK * SK * 2);
    ws[r][co] = w[i];
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[NOUT];
#pragma unroll
  for (int c = 0; c < NOUT; ++c) acc[c] = bias[c];
  for (int ky = 0; ky < SK; ++ky)
    for (int kx = 0; kx < SK; ++kx)
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const float v = patch[ci][ty + ky][tx + kx];
        const float4* wr = reinterpret_cast<const float4*>(ws[(ky * SK + kx) * 2 + ci]);
#pragma unroll
        for (int c4 = 0; c4 < NOUT / 4; ++c4) {
          const float4 q = wr[c4];
          acc[4 * c4 + 0] = fmaf(v, q.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(v, q.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(v, q.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(v, q.w, acc[4 * c4 + 3]);
        }
      }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < H && ox < W) {
    uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)b * H + oy) * W + ox) * NOUT);
#pragma unroll
    for (int q = 0; q < NOUT / 8; ++q) dst[q] = pack_bf16x8(acc + 8 * q);
  }
}

// ---- head: 3x3, C -> 1, fp32 out ---------------------------------------------------------------------
constexpr int HT = 16, HP = HT + 2;
template <int C>
__global__ void __launch_bounds__(256)
head_conv3_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ w, float bias,
                  float* __restrict__ out, int H, int W) {
  constexpr int PITCH = C * 2 + 16;                              // bytes per patch pixel (+16: bank spread)
  extern __shared__ __align__(16) uint8_t hsm[];
  float* ws = reinterpret_cast<float*>(hsm);                     // [9][C]
  uint8_t* patch = hsm + 9 * C * 4;
  const int b = blockIdx.z, oy0 = blockIdx.y * HT, ox0 = blockIdx.x * HT;
  for (int i = threadIdx.x; i < 9 * C; i += 256) ws[i] = w[i];
  constexpr int VPP = C / 8;
  for (int i = threadIdx.x; i < HP * HP * VPP; i += 256) {
    const int pix = i / VPP, v = i - pix * VPP, py = pix / HP, px = pix - py * HP;
    const int iy = oy0 + py - 1, ix = ox0 + px - 1;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      q = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)b * H + iy) * W + ix) * C + v * 8));
    *reinterpret_cast<uint4*>(patch + pix * PITCH + v * 16) = q;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc = bias;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint8_t* pp = patch + ((ty + ky) * HP + tx + kx) * PITCH;
      const float* wr = ws + (ky * 3 + kx) * C;
#pragma unroll
      for (int v = 0; v < VPP; ++v) {
        float f[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(pp + v * 16), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(f[e], wr[v * 8 + e], acc);
      }
    }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < H && ox < W) out[((size_t)b * H + oy) * W + ox] = acc;
}

// ---- reference conv (tests only): one thread per (pixel, output channel) ----------------------------------
__global__ void conv_ref_kernel(idiff_gemm_params p, const float* __restrict__ w, float* __restrict__ out) {
  const size_t total = (size_t)p.B * p.H * p.W * p.N;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % p.N);
  const size_t m = idx / p.N;
  const int ox = (int)(m % p.W), oy = (int)((m / p.W) % p.H), b = (int)(m / ((size_t)p.W * p.H));
  const int k = p.ksize, s = p.stride, pad = k == 1 ? 0 : 1;
  const int Hin = p.H * s, Win = p.W * s, Hs = p.up0 ? Hin / 2 : Hin, Ws = p.up0 ? Win / 2 : Win;
  const int cin = p.cin0 + p.cin1;
  const int ld0 = p.src0_ld ? p.src0_ld : p.cin0, ld1 = p.src1_ld ? p.src1_ld : p.cin1;
  const __nv_bfloat16* s0 = (const __nv_bfloat16*)p.src0;
  const __nv_bfloat16* s1 = (const __nv_bfloat16*)p.src1;
  float acc = p.bias ? p.bias[n] : 0.f;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const int iy = oy * s - pad + ky, ix = ox * s - pad + kx;
      if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) continue;
      const int sy = p.up0 ? iy / 2 : iy, sx = p.up0 ? ix / 2 : ix;
      for (int c = 0; c < cin; ++c) {
        float v = c < p.cin0 ? __bfloat162float(s0[(((size_t)b * Hs + sy) * Ws + sx) * ld0 + c])
                             : __bfloat162float(s1[(((size_t)b * Hs + sy) * Ws + sx) * ld1 + (c - p.cin0)]);
        if (p.a_scale) {
          v = fmaf(v, p.a_scale[(size_t)b * cin + c], p.a_shift[(size_t)b * cin + c]);
          if (p.a_silu) v = v / (1.f + expf(-v));
          v = __bfloat162float(__float2bfloat16_rn(v));          // the engine feeds bf16 to the tensor core
        }
        acc = fmaf(v, w[(((size_t)n * k + ky) * k + kx) * cin + c], acc);
      }
    }
  out[idx] = acc;
}

}  // namespace idiff

extern "C" {
using namespace idiff;

int idiff_stem_conv7(const float* x, const float* mu, const float* w, const float* bias, void* out, int B, int H,
                     int W, int N, void* stream) {
  IDIFF_REQUIRE(x && mu && w && bias && out && B > 0 && H > 0 && W > 0, "stem_conv7: bad arguments");
  IDIFF_REQUIRE(N == 64, "stem_conv7: nf must be 64 (got %d)", N);
  dim3 grid((W + ST - 1) / ST, (H + ST - 1) / ST, B);
  stem_conv7_kernel<64><<<grid, 256, 0, as_stream(stream)>>>(x, mu, w, bias, (__nv_bfloat16*)out, H, W);
  return check_launch("stem_conv7");
}

int idiff_head_conv3(const void* src, const float* w, float bias, float* out, int B, int H, int W, int C,
                     void* stream) {
  IDIFF_REQUIRE(src && w && out && B > 0 && H > 0 && W > 0, "head_conv3: bad arguments");
  IDIFF_REQUIRE(C == 64, "head_conv3: C must be 64 (got %d)", C);
  dim3 grid((W + HT - 1) / HT, (H + HT - 1) / HT, B);
  const int smem = 9 * 64 * 4 + HP * HP * (64 * 2 + 16);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(head_conv3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  head_conv3_kernel<64><<<grid, 256, smem, as_stream(stream)>>>((const __nv_bfloat16*)src, w, bias, out, H, W);
  return check_launch("head_conv3");
}

int idiff_conv_ref(const idiff_gemm_params* p, const float* w_f32, float* out_f32, void* stream) {
  IDIFF_REQUIRE(p && w_f32 && out_f32 && p->src0, "conv_ref: null pointer");
  const size_t total = (size_t)p->B * p->H * p->W * p->N;
  conv_ref_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(*p, w_f32, out_f32);
  return check_launch("conv_ref");
}

}  // extern "C"
