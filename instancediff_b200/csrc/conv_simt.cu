// The two convolutions that are not shaped for the tcgen05 engine, plus a plain reference convolution used only by
// the tests to validate the engine.
//   stem : cat([x - mu, mu]) -> 7x7, 2 -> nf channels, CUDA cores, fp32: the VALIDATION reference of stem_tc.cu
//   head : 3x3, nf -> 1 channel, warp-level tensor-core MMAs (mma.sync), HBM bound
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

// ---- head tile geometry ------------------------------------------------------------------------------------
constexpr int HT = 16, HP = HT + 2;

// ---- head on tensor cores: 3x3, 64 -> 1, fp32 out ---------------------------------------------------------
// The CUDA-core version of this layer issued 1700 instructions per pixel and ran at 1.4 TB/s; the first mma.sync version
// multiplied every output pixel's 9 shifted patch views (36 ldmatrix.x4 + 36 MMAs per 16 pixels, 1152 B of shared-
// memory reads per pixel, L1 85 % busy: profiles/r01s2 head_conv3 capture).  A 1-channel convolution is 9 dot products
// per INPUT pixel followed by a 9-point gather, so the taps go on the N side instead:
//   phase A  partial[p][tap] = w_tap . x[p] for every pixel p of the 18 x 18 patch: the patch is a linear array of
//            pixels, a warp takes 16 consecutive ones (ldmatrix.x4, ONE A fragment per 16 channels) and multiplies them
//            with three n8 tiles whose columns are (tap, hi) / (tap, lo) pairs -- bf16(w) and bf16(w - bf16(w)): the
//            weights keep ~16 mantissa bits because eps feeds the SDE update directly.  4 ldmatrix + 12 MMAs per 16
//            pixels, 128 B of shared-memory reads per pixel.
//   phase B  out[y][x] = bias + sum_tap partial[(y+dy, x+dx)][tap]: 9 fp32 reads per output pixel, one thread each.
// tcgen05 would spend a 64-cycle M128 slot per K=16 step on an N=24 problem; the warp-level MMA is the right size
// for this 0.04 %-of-FLOPs layer.
constexpr int HM_PITCH = 64 * 2 + 16;                              // bytes per patch pixel (ldmatrix rows: conflict-free)
constexpr int HM_NPIX = HP * HP;                                   // 324
constexpr int HM_MT = (HM_NPIX + 15) / 16;                         // 21 m16 tiles (the last one runs into 12 pad pixels)
constexpr int HM_PIXPAD = HM_MT * 16;
constexpr int HM_WFRAG = 3 * 4 * 32 * 8;                           // [n-tile][k-chunk][lane] -> (b0, b1)
constexpr int HM_PP = 9;                                           // floats per pixel of the partial array (odd: banks)
constexpr int HM_SMEM = HM_WFRAG + HM_PIXPAD * HM_PITCH + HM_PIXPAD * HM_PP * 4;

__global__ void __launch_bounds__(256, 2)
head_conv3_mma_kernel(const __nv_bfloat16* __restrict__ src, const uint4* __restrict__ wpk, float bias,
                      float* __restrict__ out, int H, int W, int tiles_x, int tiles_y, int total) {
  extern __shared__ __align__(16) uint8_t hsm[];
  uint2* wfrag = reinterpret_cast<uint2*>(hsm);
  uint8_t* patch = hsm + HM_WFRAG;
  float* part = reinterpret_cast<float*>(hsm + HM_WFRAG + HM_PIXPAD * HM_PITCH);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B fragments, pre-packed on the host (packing.py::pack_head_weight): lane holds k = 2*(lane%4) + {0,1} (b0) and
  // k + 8 (b1) of column n = lane/4 of n-tile j; column n is tap 4j + n/2, bf16(w) for even n, bf16(w - bf16(w)) for odd
  for (int i = tid; i < HM_WFRAG / 16; i += 256) reinterpret_cast<uint4*>(wfrag)[i] = __ldg(wpk + i);
  for (int i = tid; i < (HM_PIXPAD - HM_NPIX) * HM_PITCH / 16; i += 256)      // pad pixels: defined (never used) values
    reinterpret_cast<uint4*>(patch + HM_NPIX * HM_PITCH)[i] = make_uint4(0u, 0u, 0u, 0u);
  // Persistent CTA: the NEXT tile's patch (11 x 16 B per thread) is in flight while the current one is multiplied.
  constexpr int VPP = 8, NL = (HP * HP * VPP + 255) / 256;
  uint4 q[NL];
  auto fetch = [&](int tile) {
    const int tx = tile % tiles_x, r = tile / tiles_x, ty = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty * HT, ox0 = tx * HT;
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int i = tid + 256 * k;
      const int pix = i / VPP, v = i - pix * VPP, py = pix / HP, px = pix - py * HP;
      const int iy = oy0 + py - 1, ix = ox0 + px - 1;
      q[k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < HP * HP * VPP && iy >= 0 && iy < H && ix >= 0 && ix < W)
        q[k] = __ldg(reinterpret_cast<const uint4*>(src + (((size_t)b * H + iy) * W + ix) * 64 + v * 8));
    }
  };
  const uint32_t patch_u = smem_u32(patch);
  const int g = lane >> 2, qd = lane & 3;
  const int oty = tid >> 4, otx = tid & 15;                 // phase B: this thread's output pixel of the 16 x 16 tile
  if ((int)blockIdx.x < total) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = tile % tiles_x, r = tile / tiles_x, ty0 = r % tiles_y, b = r / tiles_y;
    const int oy0 = ty0 * HT, ox0 = tx * HT;
    __syncthreads();                                   // every thread is done with the previous patch and partials
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int i = tid + 256 * k;
      if (i < HP * HP * VPP) *reinterpret_cast<uint4*>(patch + (i / VPP) * HM_PITCH + (i % VPP) * 16) = q[k];
    }
    __syncthreads();
    if (tile + (int)gridDim.x < total) fetch(tile + gridDim.x);
    // ---- phase A: per-pixel tap partials ----
#pragma unroll 1
    for (int mt = warp; mt < HM_MT; mt += 8) {
      float c[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
      // ldmatrix row of this lane: pixel (lane & 15) of the 16, 8-channel block (lane >> 4)
      const uint32_t a_lane = patch_u + (uint32_t)((mt * 16 + (lane & 15)) * HM_PITCH + (lane >> 4) * 16);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(a_lane + kc * 32));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint2 bf = wfrag[(j * 4 + kc) * 32 + lane];
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                       "{%0, %1, %2, %3};"
                       : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf.x), "r"(bf.y));
        }
      }
      // accumulator columns (2*qd, 2*qd + 1) of n-tile j = (hi, lo) of tap 4j + qd; rows g and g + 8
      float* p0 = part + (mt * 16 + g) * HM_PP;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int tap = 4 * j + qd;
        if (tap < 9) {
          p0[tap] = c[j][0] + c[j][1];
          p0[8 * HM_PP + tap] = c[j][2] + c[j][3];
        }
      }
    }
    __syncthreads();
    // ---- phase B: 9-point gather ----
    {
      const float* pc = part + (oty * HP + otx) * HM_PP;
      float acc = bias;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc += pc[((tap / 3) * HP + (tap % 3)) * HM_PP + tap];
      const int oy = oy0 + oty, ox = ox0 + otx;
      if (oy < H && ox < W) out[((size_t)b * H + oy) * W + ox] = acc;
    }
  }
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

int idiff_head_conv3(const void* src, const void* w, float bias, float* out, int B, int H, int W, int C,
                     void* stream) {
  IDIFF_REQUIRE(src && w && out && B > 0 && H > 0 && W > 0 && aligned16(w) && aligned16(src), "head_conv3: bad arguments");
  IDIFF_REQUIRE(C == 64, "head_conv3: C must be 64 (got %d)", C);
  static DeviceOnce once;
  int num_sms = 0;
  {
    cudaError_t e = per_device_setup(once, &num_sms, [] {
      return cudaFuncSetAttribute(head_conv3_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HM_SMEM);
    });
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "head_conv3 setup: %s", cudaGetErrorString(e));
  }
  const int tiles_x = (W + HT - 1) / HT, tiles_y = (H + HT - 1) / HT, total = tiles_x * tiles_y * B;
  const int grid = total < 2 * num_sms ? total : 2 * num_sms;       // 2 resident CTAs per SM
  head_conv3_mma_kernel<<<grid, 256, HM_SMEM, as_stream(stream)>>>((const __nv_bfloat16*)src, (const uint4*)w, bias, out, H, W,
                                                                   tiles_x, tiles_y, total);
  return check_launch("head_conv3");
}

int idiff_conv_ref(const idiff_gemm_params* p, const float* w_f32, float* out_f32, void* stream) {
  IDIFF_REQUIRE(p && w_f32 && out_f32 && p->src0, "conv_ref: null pointer");
  const size_t total = (size_t)p->B * p->H * p->W * p->N;
  conv_ref_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(*p, w_f32, out_f32);
  return check_launch("conv_ref");
}

}  // extern "C"
