// Attention kernels of the UNet:
//  * self-attention of the SpatialTransformer (head dim 32): S = Q K^T and O = P V on tcgen05 with the
//    accumulators in TMEM, fp32 online softmax in registers, P re-staged through shared memory in the
//    UMMA K-major layout, V consumed as an MN-major operand (no transpose pass).
//  * linear attention (levels 0-2): k-softmax over pixels folded into a streaming context reduction
//    ctx[d][e] = sum_n softmax_n(k)[d,n] v[e,n] / HW, then the per-image effective output weight
//    Weff = Wout * blockdiag(ctx^T) written directly in the packed layout the GEMM engine streams.
// Serves `self.model(x, self.mu, t*scale, **kwargs)`, utils/sde_utils.py:198.
#include "common.cuh"
#include "host_common.h"

namespace idiff {

// =====================================================================================================
// Self-attention.  grid = (ceil(L/128), heads, B), block = 128 threads (thread r owns query row r).
// =====================================================================================================
constexpr int AT_PLANE = 128 * 16 + 32;           // bytes between 8-element planes of a 128-row tile
constexpr int AT_Q = 0;                           // [4 planes]  Q tile   (A of S = Q K^T, K-major)
constexpr int AT_K = AT_Q + 4 * AT_PLANE;         // [4 planes]  K block  (B of S, K-major: keys x d)
constexpr int AT_V = AT_K + 4 * AT_PLANE;         // [4 planes]  V block  (B of O = P V, MN-major: d x keys)
constexpr int AT_P = AT_V + 4 * AT_PLANE;         // [16 planes] P        (A of O, K-major: rows x keys)
constexpr int AT_BAR = AT_P + 16 * AT_PLANE;
constexpr int AT_SMEM = AT_BAR + 64;

static int g_debug_flags = 0;
int watchdog_attn(int clear) { return watchdog_read_tu(clear); }

__global__ void __launch_bounds__(128)
self_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int L, int heads,
                      float scale_log2e, int dbg) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(sm + AT_BAR);
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int C = heads * 32, ld = 3 * C;
  const __nv_bfloat16* base = qkv + (size_t)b * L * ld + h * 32;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // tile loader: 128 rows x 32 channels -> [4 planes][128 rows][8]; rows >= L are zero.  Split into a register
  // fetch and a shared-memory store so the NEXT key block's K / V tiles are in flight while this one is processed.
  const int lc8 = tid & 3, lrow = tid >> 2;
  auto fetch_tile = [&](uint4* q, int row0, int col_off) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // branch-free: clamped row, zeroed afterwards
      const int row = row0 + lrow + 32 * i;
      const int rc = row < L ? row : L - 1;
      q[i] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)rc * ld + col_off + lc8 * 8));
      if (row >= L) q[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto store_tile = [&](int smem_off, const uint4* q) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint4*>(sm + smem_off + lc8 * AT_PLANE + (lrow + 32 * i) * 16) = q[i];
  };
  {
    uint4 qq[4];
    fetch_tile(qq, q0, 0);
    store_tile(AT_Q, qq);
  }

  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, 32, 1);          // B (= V) is MN-major
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  // MMA issue: warp 0 runs the issue code warp-uniformly, one elected lane issues (descriptors stay uniform)
  const bool leader = (warp == 0) && elect_one();
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t smem0 = smem_u32(sm);
  const bool swk = (dbg & 4) != 0, swv = (dbg & 2) != 0;
  const uint32_t k_lbo = swk ? 128 : AT_PLANE, k_sbo = swk ? AT_PLANE : 128;     // K-major operands (Q, K, P)
  const uint32_t v_lbo = swv ? AT_PLANE : 128, v_sbo = swv ? 128 : AT_PLANE;     // MN-major V
  const uint32_t q_lo = umma_desc_lo(smem0 + AT_Q, k_lbo), kk_lo = umma_desc_lo(smem0 + AT_K, k_lbo);
  const uint32_t p_lo = umma_desc_lo(smem0 + AT_P, k_lbo), v_lo = umma_desc_lo(smem0 + AT_V, v_lbo);
  const uint32_t k_hi = umma_desc_hi(k_sbo), v_hi = umma_desc_hi(v_sbo);

  float o[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) o[e] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  const int nblk = (L + 127) / 128;
  uint4 kq[4], vq[4];
  fetch_tile(kq, 0, C);
  fetch_tile(vq, 0, 2 * C);
  for (int j = 0; j < nblk; ++j) {
    const int k0 = j * 128;
    store_tile(AT_K, kq);
    store_tile(AT_V, vq);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16_lohi(tmem_u, q_lo + ((kk * 2 * AT_PLANE) >> 4), k_hi, kk_lo + ((kk * 2 * AT_PLANE) >> 4), k_hi, idesc_s,
                         kk);
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    if (j + 1 < nblk) {                                          // next block's tiles: in flight during softmax + PV
      fetch_tile(kq, k0 + 128, C);
      fetch_tile(vq, k0 + 128, 2 * C);
    }
    mbar_wait(bar_s, j & 1, 201);
    tc_fence_after();
    const bool full = k0 + 128 <= L;                             // no key masking needed in this block

    // ---- online softmax over this key block (scores scaled by 1/sqrt(d), in log2 domain) ----
    float mx = -INFINITY;
    for (int cc = 0; cc < 4; ++cc) {
      float v[32];
      tmem_ld32(lane_addr + cc * 32, v);
      if (full) {
        float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int q = 4; q < 32; ++q) m4[q & 3] = fmaxf(m4[q & 3], v[q]);
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q)
          if (k0 + cc * 32 + q < L) mx = fmaxf(mx, v[q]);
      }
    }
    const float m_new = fmaxf(m_run, mx * scale_log2e);
    const float alpha = exp2f(m_run - m_new);                    // exp2f(-inf) = 0 on the first block
    float psum = 0.f;
    for (int cc = 0; cc < 4; ++cc) {
      float v[32];
      tmem_ld32(lane_addr + cc * 32, v);
      if (full) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          float pv;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"(fmaf(v[q], scale_log2e, -m_new)));
          v[q] = pv;
          s4[q & 3] += pv;
        }
        psum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float pv = (k0 + cc * 32 + q < L) ? exp2f(fmaf(v[q], scale_log2e, -m_new)) : 0.f;
          v[q] = pv;
          psum += pv;
        }
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8)
        *reinterpret_cast<uint4*>(sm + AT_P + (cc * 4 + g8) * AT_PLANE + tid * 16) = pack_bf16x8(v + g8 * 8);
    }
    l_run = l_run * alpha + psum;
    m_run = m_new;
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (leader) {
        // A: P planes (2kk, 2kk+1); B: V rows (keys) 16kk.. as MN-major: LBO = 8-key group pitch (128 B),
        // SBO = pitch between 8-channel groups (plane)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_lohi(tmem_u, p_lo + ((kk * 2 * AT_PLANE) >> 4), k_hi, v_lo + ((kk * 256) >> 4), v_hi, idesc_o, kk);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, j & 1, 202);
    tc_fence_after();
    {
      float v[32];
      tmem_ld32(lane_addr, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha, v[e]);
    }
    tc_fence_before();
    __syncthreads();      // K/V/P smem and the TMEM columns are re-used by the next block
  }

  if (q0 + tid < L) {
    const float inv = 1.f / l_run;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] *= inv;
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * L + q0 + tid) * C + h * 32);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) dst[q4] = pack_bf16x8(o + q4 * 8);
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// =====================================================================================================
// Linear attention context on tensor cores.
//   ctx[h][d][e] = sum_n softmax_n(k)[d,n] v[e,n] / HW
// With the exact per-channel maximum m[d] (pre-pass), E = exp(k - m) in (0,1]:
//   G = E^T [v | 1]   (128 k-channels x 144 columns; fp32, accumulated over pixels by tcgen05.mma)
//   ctx[h][d][e] = G[h*32+d][h*32+e] / (G[h*32+d][128] * HW)
// Both operands are MN-major views of the [8-channel plane][pixel][8] tiles the loaders write (the same
// layout/descriptor roles the self-attention V operand uses).  The Gram matrix over all 4 heads is computed at
// once (M = 128) and only the 4 diagonal 32x32 blocks are kept.
// =====================================================================================================
// pixels per CTA: small enough that every level fills the GPU (a CTA walks its chunk serially; with 4096-pixel
// chunks the 64x64 level ran on 32 CTAs and every level cost ~0.2 ms of latency regardless of its size)
__host__ __device__ inline int la_chunk(int HW) { return HW >= 65536 ? 2048 : HW >= 16384 ? 1024 : 512; }
constexpr int LA_SUB = 128;                                  // pixels per MMA batch (8 x K16)
constexpr int LA_PLANE = LA_SUB * 16 + 32;                   // bytes between 8-channel planes
constexpr int LA_OFF_E = 0;                                  // [16 planes] E tile  (A, MN-major: k-channels x pixels)
constexpr int LA_OFF_V = LA_OFF_E + 16 * LA_PLANE;           // [18 planes] [v | 1] (B, MN-major: 144 columns x pixels)
constexpr int LA_OFF_MAX = LA_OFF_V + 18 * LA_PLANE;         // [128] floats
constexpr int LA_OFF_BAR = LA_OFF_MAX + 512;
constexpr int LA_SMEM = LA_OFF_BAR + 64;
constexpr int LA_NCOL = 144;
constexpr int LA_PART = 128 * 33;                            // floats per (image, chunk): diagonal block + row sum

// per-(image, chunk) column maxima of k: grid = (nchunk, B), block = 256
__global__ void __launch_bounds__(256)
linattn_kmax_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ pmax, int HW) {
  __shared__ float red[16][128];
  const int tid = threadIdx.x, chunk = blockIdx.x, b = blockIdx.y, nchunk = gridDim.x;
  const int c8 = tid & 15, rg = tid >> 4;
  const int r0 = chunk * la_chunk(HW), r1 = min(HW, r0 + la_chunk(HW));
  const __nv_bfloat16* kb = qkv + (size_t)b * HW * 384 + 128 + c8 * 8;
  float mx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) mx[e] = -INFINITY;
  for (int r = r0 + rg; r < r1; r += 16) {
    float f[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(kb + (size_t)r * 384)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], f[e]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rg][c8 * 8 + e] = mx[e];
  __syncthreads();
  if (tid < 128) {
    float m = red[0][tid];
#pragma unroll
    for (int g = 1; g < 16; ++g) m = fmaxf(m, red[g][tid]);
    pmax[((size_t)b * nchunk + chunk) * 128 + tid] = m;
  }
}

// grid = (nchunk, B), block = 256
__global__ void __launch_bounds__(256)
linattn_gram_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ pmax, float* __restrict__ part,
                    int HW, int dbg) {
  extern __shared__ __align__(128) uint8_t sm[];
  float* kmax = reinterpret_cast<float*>(sm + LA_OFF_MAX);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + LA_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, chunk = blockIdx.x, b = blockIdx.y, nchunk = gridDim.x;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid < 128) {                                           // exact channel maximum over the whole image
    float m = -INFINITY;
    for (int c = 0; c < nchunk; ++c) m = fmaxf(m, pmax[((size_t)b * nchunk + c) * 128 + tid]);
    kmax[tid] = m;
  }
  // the two constant "ones" planes of B (columns 128..143): written once
  for (int i = tid; i < 2 * LA_SUB; i += 256) {
    const uint32_t one2 = 0x3F803F80u;                        // bf16(1.0) x2
    *reinterpret_cast<uint4*>(sm + LA_OFF_V + (16 + (i >> 7)) * LA_PLANE + (i & 127) * 16) = make_uint4(one2, one2, one2, one2);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool leader = (warp == 0) && elect_one();
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t smem0 = smem_u32(sm);
  const bool sw = (dbg & 8) != 0;
  const uint32_t lbo = sw ? LA_PLANE : 128, sbo = sw ? 128 : LA_PLANE;    // MN-major: LBO = 8-pixel group pitch
  const uint32_t e_lo = umma_desc_lo(smem0 + LA_OFF_E, lbo), v_lo = umma_desc_lo(smem0 + LA_OFF_V, lbo);
  const uint32_t hi = umma_desc_hi(sbo);
  const uint32_t idesc = umma_idesc_bf16_ex(128, LA_NCOL, 1, 1);

  const int c8 = tid & 15, rg = tid >> 4;                     // 16 pixels per sweep, 8 sweeps per sub-block
  float mloc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) mloc[e] = kmax[c8 * 8 + e];
  const __nv_bfloat16* kb = qkv + (size_t)b * HW * 384 + 128 + c8 * 8;
  const __nv_bfloat16* vb = kb + 128;
  const int r_begin = chunk * la_chunk(HW), r_end = min(HW, r_begin + la_chunk(HW));
  int it = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += LA_SUB, ++it) {
    uint4 qk[8], qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                              // branch-free: clamped row
      const int r = r0 + rg + 16 * i;
      const int rc = r < r_end ? r : r_end - 1;
      qk[i] = __ldg(reinterpret_cast<const uint4*>(kb + (size_t)rc * 384));
      qv[i] = __ldg(reinterpret_cast<const uint4*>(vb + (size_t)rc * 384));
    }
    if (it > 0) mbar_wait(bar, (it - 1) & 1, 301);             // previous MMAs finished reading the tiles
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int px = rg + 16 * i;
      const bool ok = r0 + px < r_end;
      float f[8];
      unpack_bf16x8(qk[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = ok ? __expf(f[e] - mloc[e]) : 0.f;
      *reinterpret_cast<uint4*>(sm + LA_OFF_E + c8 * LA_PLANE + px * 16) = pack_bf16x8(f);
      *reinterpret_cast<uint4*>(sm + LA_OFF_V + c8 * LA_PLANE + px * 16) = ok ? qv[i] : make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < LA_SUB / 16; ++kk)              // K = 16 pixels per MMA: advance by two 8-pixel groups
          umma_bf16_lohi(tmem_u, e_lo + ((kk * 256) >> 4), hi, v_lo + ((kk * 256) >> 4), hi, idesc, (it | kk) != 0 ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
  }
  mbar_wait(bar, (it - 1) & 1, 302);
  tc_fence_after();
  // epilogue: warps 0-3 own TMEM lanes (k-channel rows); keep the head-diagonal 32 columns + the row sum
  if (warp < 4) {
    const int d = warp * 32 + (tid & 31);                     // k channel = row
    float* dst = part + ((size_t)b * nchunk + chunk) * LA_PART + (size_t)d * 33;
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(warp * 32), v);   // head h = warp: columns h*32..
#pragma unroll
    for (int e = 0; e < 32; ++e) dst[e] = v[e];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 128u, v);                    // columns 128..159 (128 = row sum)
    dst[32] = v[0];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// grid = (4 heads, B), block = 256: sum the per-chunk Gram blocks, normalise, then
// Weff[c][h*32+d] = sum_e Wout[c][h*32+e] ctx[h][d][e]   (the columns of this head)
__global__ void __launch_bounds__(256)
linattn_merge_kernel(const float* __restrict__ part, int nchunk, const float* __restrict__ w_out,
                     __nv_bfloat16* __restrict__ weff, int HW, int C) {
  __shared__ float ctx[32 * 33];
  __shared__ float ssum[32];
  const int tid = threadIdx.x, h = blockIdx.x, b = blockIdx.y;
  const float* p0 = part + (size_t)b * nchunk * LA_PART + (size_t)(h * 32) * 33;
  for (int i = tid; i < 32 * 33; i += 256) {                  // rows d = 0..31 of this head, 33 values each
    float acc = 0.f;
    for (int c = 0; c < nchunk; ++c) acc += p0[(size_t)c * LA_PART + i];
    ctx[i] = acc;
  }
  __syncthreads();
  if (tid < 32) ssum[tid] = 1.f / (ctx[tid * 33 + 32] * (float)HW);   // softmax normaliser and v / (h*w)
  __syncthreads();
  __nv_bfloat16* wdst = weff + (size_t)b * C * 128;
  for (int i = tid; i < C * 32; i += 256) {
    const int c = i >> 5, d = i & 31, kc = h * 32 + d;
    const float* wr = w_out + (size_t)c * 128 + h * 32;
    const float* cr = ctx + d * 33;
    float acc = 0.f;
#pragma unroll 8
    for (int e = 0; e < 32; ++e) acc = fmaf(__ldg(wr + e), cr[e], acc);
    acc *= ssum[d];
    // packed B-operand layout of conv_gemm (NT = C, two 64-wide K stages)
    const int ks = kc >> 6, kin = kc & 63;
    wdst[(size_t)ks * C * 64 + (kin >> 3) * (C * 8) + c * 8 + (kin & 7)] = __float2bfloat16_rn(acc);
  }
}

}  // namespace idiff

extern "C" {
using namespace idiff;

int idiff_self_attention(const void* qkv, void* out, int B, int L, int heads, float scale, void* stream) {
  IDIFF_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0, "self_attention: bad arguments");
  IDIFF_REQUIRE(aligned16(qkv) && aligned16(out), "self_attention: 16 B alignment");
  static DeviceOnce once;
  {
    cudaError_t e = per_device_setup(once, nullptr, [] {
      return cudaFuncSetAttribute(self_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    });
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "self_attention attr: %s", cudaGetErrorString(e));
  }
  dim3 grid((unsigned)((L + 127) / 128), (unsigned)heads, (unsigned)B);
  self_attention_kernel<<<grid, 128, AT_SMEM, as_stream(stream)>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, L,
                                                                   heads, scale * 1.4426950408889634f, g_debug_flags);
  return check_launch("self_attention");
}

int idiff_set_debug_flags(int flags) {
  g_debug_flags = flags;
  return IDIFF_OK;
}

size_t idiff_linattn_scratch_floats(int B, int HW) {
  const size_t nchunk = (size_t)(HW + la_chunk(HW) - 1) / la_chunk(HW);
  return (size_t)B * nchunk * (LA_PART + 128);               // Gram partials + per-chunk channel maxima
}

int idiff_linattn_context(const void* qkv, const float* w_out, void* weff_packed, float* scratch, int B, int HW,
                          int C, void* stream) {
  IDIFF_REQUIRE(qkv && w_out && weff_packed && scratch && B > 0 && HW > 0, "linattn_context: bad arguments");
  IDIFF_REQUIRE(C == 64 || C == 128 || C == 256, "linattn_context: C must be 64/128/256");
  IDIFF_REQUIRE(aligned16(qkv), "linattn_context: 16 B alignment");
  const int nchunk = (HW + la_chunk(HW) - 1) / la_chunk(HW);
  float* part = scratch;
  float* pmax = scratch + (size_t)B * nchunk * LA_PART;
  static DeviceOnce once;
  {
    cudaError_t e = per_device_setup(once, nullptr, [] {
      return cudaFuncSetAttribute(linattn_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LA_SMEM);
    });
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "linattn attr: %s", cudaGetErrorString(e));
  }
  dim3 grid((unsigned)nchunk, (unsigned)B);
  linattn_kmax_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)qkv, pmax, HW);
  if (int rc = check_launch("linattn_kmax")) return rc;
  linattn_gram_kernel<<<grid, 256, LA_SMEM, as_stream(stream)>>>((const __nv_bfloat16*)qkv, pmax, part, HW, g_debug_flags);
  if (int rc = check_launch("linattn_gram")) return rc;
  linattn_merge_kernel<<<dim3(4, (unsigned)B), 256, 0, as_stream(stream)>>>(part, nchunk, w_out,
                                                                            (__nv_bfloat16*)weff_packed, HW, C);
  return check_launch("linattn_merge");
}

}  // extern "C"
