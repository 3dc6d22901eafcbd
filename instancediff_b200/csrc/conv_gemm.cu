// Implicit-GEMM convolution / linear engine on tcgen05 tensor cores (sm_100a).
//
// One CTA = one 16x8 output-pixel tile (M = 128 rows) x NT output channels.
//   warps 0-3 : A producers.  Per 64-channel chunk they load the input patch (tile + halo) ONCE from
//               global memory (coalesced 16 B vectors, optional concat of two sources, optional
//               nearest x2 upsampling, optional per-(image,channel) affine + SiLU = fused
//               GroupNorm/time-modulation/activation of the previous layer) and store it in shared
//               memory as [8-channel group][patch pixel][8 ch] -- the UMMA SWIZZLE_NONE K-major
//               canonical layout with a 16 B row pitch.  Every filter tap is then just a different
//               descriptor start address (shifted window), so the patch is re-used k*k times from
//               shared memory.  After the main loop the same warps run the epilogue out of TMEM.
//   warp 4    : B producer.  Streams pre-packed weight tiles (one per (chunk, tap)) with 1-D bulk TMA
//               into a ring, completion on mbarriers.
//   warp 5    : allocates TMEM, issues tcgen05.mma (one thread), commits to the ring barriers.
// Accumulators: 128 lanes x NT fp32 columns of TMEM.
//
// Network spec: SURVEY.md App. A (the reference's models/modules/* is not in the snapshot); the call
// this serves is `noise = self.model(x, self.mu, t*scale, **kwargs)`, utils/sde_utils.py:198.
#include "common.cuh"
#include "host_common.h"

namespace idiff {

constexpr int TILE_H = 16, TILE_W = 8, TILE_M = 128;
constexpr int kLoaderThreads = 128;
constexpr int kThreads = 192;
constexpr int kMaxSA = 4, kMaxSB = 8;

struct Geom {
  int k, s, pad;
  int PR, PC;          // patch rows / cols (input space)
  int nslot;           // PR*PC
  int lboA;            // bytes between 8-channel planes (odd multiple of 16 -> conflict-free stores)
  int sboA;            // bytes between consecutive output rows of the tile (8-pixel groups)
  int stageA;          // bytes per A stage
};

__host__ __device__ inline Geom make_geom(int ksize, int stride) {
  Geom g;
  g.k = ksize;
  g.s = stride;
  g.pad = ksize == 1 ? 0 : 1;
  g.PR = (TILE_H - 1) * stride + ksize;
  g.PC = (TILE_W - 1) * stride + ksize;
  g.nslot = g.PR * g.PC;
  g.lboA = (g.nslot | 1) * 16;
  g.sboA = stride * g.PC * 16;
  g.stageA = ((8 * g.lboA + 127) / 128) * 128;
  return g;
}

struct SmemPlan {
  int SA, SB, stageA, stageB, offA, offB, total;
};

__host__ inline SmemPlan plan_smem(const idiff_gemm_params& p, const Geom& g) {
  SmemPlan s;
  const int nchunks = (p.cin0 + p.cin1) / 64;
  const int nk = nchunks * g.k * g.k;
  s.stageA = g.stageA;
  s.stageB = p.NT * 128;
  const int budget = (p.NT == 256 ? 226 : 113) * 1024 - 1024;
  s.SA = nchunks < 2 ? 1 : 2;
  if (s.SA * s.stageA + 2 * s.stageB > budget) s.SA = 1;
  int sb = (budget - s.SA * s.stageA) / s.stageB;
  if (sb > kMaxSB) sb = kMaxSB;
  if (sb > nk) sb = nk;
  if (sb < 1) sb = 1;
  s.SB = sb;
  s.offA = 1024;
  s.offB = s.offA + s.SA * s.stageA;
  s.total = s.offB + s.SB * s.stageB;
  return s;
}

struct KArgs {
  idiff_gemm_params p;
  Geom g;
  int SA, SB, offA, offB, stageB;
  int tiles_x, tiles_y;
};

// --------------------------------------------------------------------------------------------------
IDIFF_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(kThreads, 2) conv_gemm_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const idiff_gemm_params& p = a.p;
  const Geom& g = a.g;

  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);
  uint64_t* emptyA = fullA + kMaxSA;
  uint64_t* fullB = emptyA + kMaxSA;
  uint64_t* emptyB = fullB + kMaxSB;
  uint64_t* accum_bar = emptyB + kMaxSB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* red = reinterpret_cast<float*>(smem + 512);          // [4 warps][8 groups][2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  const int tile_in_img = blockIdx.x % tiles_per_img;
  const int oy0 = (tile_in_img / a.tiles_x) * TILE_H, ox0 = (tile_in_img % a.tiles_x) * TILE_W;
  const int ntile = blockIdx.y, n0 = ntile * p.NT;
  const int cin = p.cin0 + p.cin1, nchunks = cin >> 6, ntaps = g.k * g.k, nk = nchunks * ntaps;

  if (tid == 0) {
    for (int i = 0; i < a.SA; ++i) { mbar_init(&fullA[i], kLoaderThreads); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < a.SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)p.NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ============================== A producers ==============================================
    const int c8 = tid & 7, prow = tid >> 3;                       // 16 patch pixels per sweep
    // virtual (possibly upsampled) input extent
    const int Hin = p.H * g.s, Win = p.W * g.s;
    const int Hs = p.up0 ? (Hin >> 1) : Hin, Ws = p.up0 ? (Win >> 1) : Win;
    const int iy0 = oy0 * g.s - g.pad, ix0 = ox0 * g.s - g.pad;
    const bool affine = p.a_scale != nullptr;

    for (int ch = 0; ch < nchunks; ++ch) {
      const int sa = ch % a.SA;
      if (ch >= a.SA) mbar_wait(&emptyA[sa], (((ch / a.SA) & 1) ^ 1), 101);
      uint8_t* stage = smem + a.offA + sa * g.stageA + c8 * g.lboA;

      const bool from0 = (ch << 6) < p.cin0;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(from0 ? p.src0 : p.src1);
      const int Cs = from0 ? (p.src0_ld ? p.src0_ld : p.cin0) : (p.src1_ld ? p.src1_ld : p.cin1);
      const int coff = (from0 ? (ch << 6) : ((ch << 6) - p.cin0)) + c8 * 8;
      const __nv_bfloat16* src_b = src + (size_t)b * Hs * Ws * Cs + coff;

      float sc[8], sh[8];
      if (affine) {
        const float* ps = p.a_scale + (size_t)b * cin + (ch << 6) + c8 * 8;
        const float* pt = p.a_shift + (size_t)b * cin + (ch << 6) + c8 * 8;
        const float4 s0 = ldg4(ps), s1 = ldg4(ps + 4), t0 = ldg4(pt), t1 = ldg4(pt + 4);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
      }

      constexpr int BATCH = 6;
      for (int px0 = prow; px0 < g.nslot; px0 += 16 * BATCH) {
        uint4 q[BATCH];
        int slot[BATCH];
        bool inb[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int px = px0 + 16 * i;
          slot[i] = -1;
          inb[i] = false;
          q[i] = make_uint4(0u, 0u, 0u, 0u);
          if (px < g.nslot) {
            const int v = px / g.PC, u = px - v * g.PC;
            slot[i] = g.s == 1 ? px : v * g.PC + (u & 1) * (g.PC >> 1) + (u >> 1);
            const int iy = iy0 + v, ix = ix0 + u;
            if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
              inb[i] = true;
              const int sy = p.up0 ? (iy >> 1) : iy, sx = p.up0 ? (ix >> 1) : ix;
              q[i] = __ldg(reinterpret_cast<const uint4*>(src_b + ((size_t)sy * Ws + sx) * Cs));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          if (slot[i] < 0) continue;
          uint4 o = q[i];
          if (affine && inb[i]) {
            float f[8];
            unpack_bf16x8(q[i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float y = fmaf(f[e], sc[e], sh[e]);
              f[e] = p.a_silu ? silu_fast(y) : y;
            }
            o = pack_bf16x8(f);
          }
          *reinterpret_cast<uint4*>(stage + slot[i] * 16) = o;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&fullA[sa]);
    }

    // ============================== epilogue ==================================================
    mbar_wait(accum_bar, 0, 102);
    tc_fence_after();

    const int r = tid, ti = r >> 3, tj = r & 7;
    const int oy = oy0 + ti, ox = ox0 + tj;
    const bool valid = (oy < p.H) && (ox < p.W);
    const size_t m = ((size_t)b * p.H + (valid ? oy : 0)) * p.W + (valid ? ox : 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int nchunk32 = p.NT >> 5;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* res0 = reinterpret_cast<const __nv_bfloat16*>(p.res0);
    const __nv_bfloat16* res1 = reinterpret_cast<const __nv_bfloat16*>(p.res1);

    float mean_in = 0.f, rstd_in = 1.f;
    if (p.row_stats) {
      mean_in = __ldg(p.row_stats + 2 * m);
      rstd_in = __ldg(p.row_stats + 2 * m + 1);
    }
    // value of accumulator column n after LayerNorm fold + biases
    auto base_value = [&](float acc, int n) -> float {
      float v = acc;
      if (p.row_stats) v = (v - mean_in * __ldg(p.wsum + n)) * rstd_in;
      if (p.bias) v += __ldg(p.bias + n);
      if (p.bias_img) v += __ldg(p.bias_img + (size_t)b * p.N + n);
      return v;
    };
    auto add_residuals = [&](float* v, int ncol0) {   // 32 columns starting at ncol0 (global column)
      if (res0) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8];
          unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(res0 + m * p.N + ncol0 + q4 * 8)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float rv = f[e];
            if (p.res0_scale) {
              const int n = ncol0 + q4 * 8 + e;
              rv = silu_fast(fmaf(rv, __ldg(p.res0_scale + (size_t)b * p.N + n), __ldg(p.res0_shift + (size_t)b * p.N + n)));
            }
            v[q4 * 8 + e] += rv;
          }
        }
      }
      if (res1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8];
          unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(res1 + m * p.N + ncol0 + q4 * 8)), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[q4 * 8 + e] += f[e];
        }
      }
    };
    auto store32 = [&](const float* v, int col0) {      // 32 bf16 at out[m][col0..]
      if (!valid) return;
      uint4* dst = reinterpret_cast<uint4*>(outp + m * p.out_ld + col0);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) dst[q4] = pack_bf16x8(v + q4 * 8);
    };

    if (p.epi == IDIFF_EPI_LN_OUT) {
      // pass 1: LayerNorm statistics over the whole row (NT == N)
      float s1 = 0.f, s2 = 0.f;
      for (int cc = 0; cc < nchunk32; ++cc) {
        float v[32];
        tmem_ld32(taddr + cc * 32, v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float x = base_value(v[q], n0 + cc * 32 + q);
          s1 += x;
          s2 += x * x;
        }
      }
      const float mean = s1 / p.N, var = fmaxf(s2 / p.N - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.ln_eps);
      float o1 = 0.f, o2 = 0.f;
      for (int cc = 0; cc < nchunk32; ++cc) {
        float v[32];
        tmem_ld32(taddr + cc * 32, v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int n = n0 + cc * 32 + q;
          v[q] = (base_value(v[q], n) - mean) * rstd * __ldg(p.ln_g + n);
        }
        if (valid) add_residuals(v, n0 + cc * 32);
#pragma unroll
        for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 += v[q] * v[q]; }
        store32(v, n0 + cc * 32);
      }
      if (p.out_row_stats && valid) {
        const float mo = o1 / p.N, vo = fmaxf(o2 / p.N - mo * mo, 0.f);
        p.out_row_stats[2 * m] = mo;
        p.out_row_stats[2 * m + 1] = rsqrtf(vo + p.ln_eps);
      }
    } else {
      float o1 = 0.f, o2 = 0.f;
      const int cpg = p.gn_groups > 0 ? p.N / p.gn_groups : 32;      // channels per GroupNorm group
      for (int cc = 0; cc < nchunk32; ++cc) {
        float v[32];
        tmem_ld32(taddr + cc * 32, v);
        const int ncol0 = n0 + cc * 32;
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = base_value(v[q], ncol0 + q);

        if (p.gn_groups > 0) {
          // per-(tile, group) partial sums of the conv output (bias included), valid rows only.
          // 8-column sub-sums first (static indexing), then merged according to channels-per-group.
          float a1[4], a2[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float x = v[g8 * 8 + q]; s1 += x; s2 = fmaf(x, x, s2); }
            a1[g8] = valid ? s1 : 0.f;
            a2[g8] = valid ? s2 : 0.f;
          }
          if (cpg >= 16) { a1[0] += a1[1]; a2[0] += a2[1]; a1[2] += a1[3]; a2[2] += a2[3]; }
          if (cpg >= 32) { a1[0] += a1[2]; a2[0] += a2[2]; }
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const bool lead = cpg == 8 || (cpg == 16 && (g8 & 1) == 0) || (cpg >= 32 && g8 == 0);
            if (!lead) continue;                                      // warp-uniform
            const float s1 = warp_sum(a1[g8]), s2 = warp_sum(a2[g8]);
            if (lane == 0) {
              const int gl = (cc * 32 + g8 * 8) / cpg;                // group index local to this N tile
              float* dst = red + (warp * 8 + gl) * 2;
              if (cpg > 32 && (cc * 32) % cpg != 0) { dst[0] += s1; dst[1] += s2; }
              else { dst[0] = s1; dst[1] = s2; }
            }
          }
        }

        if (p.epi == IDIFF_EPI_QSOFTMAX && ncol0 < 128) {
          float mx = v[0];
#pragma unroll
          for (int q = 1; q < 32; ++q) mx = fmaxf(mx, v[q]);
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 32; ++q) { v[q] = __expf(v[q] - mx); s += v[q]; }
          const float inv = p.qscale / s;
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] *= inv;
          store32(v, ncol0);
        } else if (p.epi == IDIFF_EPI_GEGLU) {
          float o[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) o[q] = v[2 * q] * gelu_erf(v[2 * q + 1]);
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(outp + m * p.out_ld + (ncol0 >> 1));
            dst[0] = pack_bf16x8(o);
            dst[1] = pack_bf16x8(o + 8);
          }
        } else {
          if (valid) add_residuals(v, ncol0);
#pragma unroll
          for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 += v[q] * v[q]; }
          store32(v, ncol0);
        }
      }
      if (p.out_row_stats && valid) {
        const float mo = o1 / p.N, vo = fmaxf(o2 / p.N - mo * mo, 0.f);
        p.out_row_stats[2 * m] = mo;
        p.out_row_stats[2 * m + 1] = rsqrtf(vo + p.ln_eps);
      }
      if (p.gn_groups > 0) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int groups_here = p.NT / cpg;
        if (tid < groups_here) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) { s1 += red[(w * 8 + tid) * 2]; s2 += red[(w * 8 + tid) * 2 + 1]; }
          const int gglob = n0 / cpg + tid;
          float* dst = p.gn_partial + (((size_t)b * tiles_per_img + tile_in_img) * p.gn_groups + gglob) * 2;
          dst[0] = s1;
          dst[1] = s2;
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ============================== B producer (bulk TMA) ====================================
    if (lane == 0) {
      const __nv_bfloat16* wbase = reinterpret_cast<const __nv_bfloat16*>(p.w) + (size_t)b * p.w_image_stride +
                                   (size_t)ntile * nk * p.NT * 64;
      for (int ks = 0; ks < nk; ++ks) {
        const int sb = ks % a.SB;
        if (ks >= a.SB) mbar_wait(&emptyB[sb], (((ks / a.SB) & 1) ^ 1), 103);
        mbar_arrive_expect_tx(&fullB[sb], (uint32_t)a.stageB);
        bulk_g2s(smem + a.offB + sb * a.stageB, wbase + (size_t)ks * p.NT * 64, (uint32_t)a.stageB, &fullB[sb]);
      }
    }
  } else {
    // ============================== MMA issuer ================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TILE_M, p.NT, 0);
      const uint32_t lboB = (uint32_t)p.NT * 16, sboB = 128;
      const uint32_t a0 = smem_u32(smem + a.offA), b0 = smem_u32(smem + a.offB);
      int ks = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int sa = ch % a.SA;
        mbar_wait(&fullA[sa], (ch / a.SA) & 1, 104);
        tc_fence_after();
        for (int tap = 0; tap < ntaps; ++tap, ++ks) {
          const int sb = ks % a.SB;
          mbar_wait(&fullB[sb], (ks / a.SB) & 1, 105);
          tc_fence_after();
          const int tr = tap / g.k, tc = tap - tr * g.k;
          const int tapslot = g.s == 1 ? tr * g.PC + tc : tr * g.PC + (tc & 1) * (g.PC >> 1) + (tc >> 1);
          const uint32_t abase = a0 + sa * g.stageA + tapslot * 16;
          const uint32_t bbase = b0 + sb * a.stageB;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint64_t ad, bd;
            if (p.dbg_swap_lbo_sbo) {
              ad = umma_desc(abase + kk * 2 * g.lboA, g.sboA, g.lboA);
              bd = umma_desc(bbase + kk * 2 * lboB, sboB, lboB);
            } else {
              ad = umma_desc(abase + kk * 2 * g.lboA, g.lboA, g.sboA);
              bd = umma_desc(bbase + kk * 2 * lboB, lboB, sboB);
            }
            umma_bf16(tmem_base, ad, bd, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&emptyB[sb]);
        }
        umma_commit(&emptyA[sa]);
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.NT);
  }
}

static int validate(const idiff_gemm_params& p) {
  IDIFF_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0, "conv_gemm: empty output");
  IDIFF_REQUIRE((p.ksize == 1 && p.stride == 1) || (p.ksize == 3 && p.stride == 1) || (p.ksize == 4 && p.stride == 2),
                "conv_gemm: unsupported ksize/stride %d/%d", p.ksize, p.stride);
  IDIFF_REQUIRE(p.cin0 > 0 && p.cin0 % 64 == 0 && p.cin1 >= 0 && p.cin1 % 64 == 0, "conv_gemm: cin must be multiples of 64");
  IDIFF_REQUIRE(p.NT == 64 || p.NT == 128 || p.NT == 256, "conv_gemm: NT must be 64/128/256");
  IDIFF_REQUIRE(p.N > 0 && p.N % p.NT == 0, "conv_gemm: N %% NT != 0");
  IDIFF_REQUIRE(p.src0 && p.w && p.out, "conv_gemm: null src/w/out");
  IDIFF_REQUIRE(p.cin1 == 0 || p.src1, "conv_gemm: cin1 > 0 needs src1");
  IDIFF_REQUIRE((p.a_scale == nullptr) == (p.a_shift == nullptr), "conv_gemm: a_scale/a_shift must come together");
  IDIFF_REQUIRE(!p.up0 || (p.ksize == 3 && p.H % 2 == 0 && p.W % 2 == 0), "conv_gemm: upsample needs ksize 3 and even output");
  IDIFF_REQUIRE((p.row_stats == nullptr) == (p.wsum == nullptr), "conv_gemm: row_stats needs wsum");
  IDIFF_REQUIRE(idiff::aligned16(p.src0) && idiff::aligned16(p.w) && idiff::aligned16(p.out), "conv_gemm: 16 B alignment");
  if (p.epi == IDIFF_EPI_LN_OUT) IDIFF_REQUIRE(p.NT == p.N && p.ln_g, "conv_gemm: LN_OUT needs NT == N and ln_g");
  if (p.epi == IDIFF_EPI_QSOFTMAX) IDIFF_REQUIRE(p.N >= 128 && p.NT <= 128, "conv_gemm: QSOFTMAX needs NT <= 128");
  if (p.epi == IDIFF_EPI_GEGLU) IDIFF_REQUIRE(p.out_ld >= p.N / 2 && !p.res0 && !p.res1, "conv_gemm: GEGLU out_ld");
  else IDIFF_REQUIRE(p.out_ld >= p.N, "conv_gemm: out_ld < N");
  IDIFF_REQUIRE(p.out_ld % 8 == 0, "conv_gemm: out_ld %% 8");
  IDIFF_REQUIRE(p.src0_ld % 8 == 0 && p.src1_ld % 8 == 0 && (p.src0_ld == 0 || p.src0_ld >= p.cin0) &&
                    (p.src1_ld == 0 || p.src1_ld >= p.cin1), "conv_gemm: bad source pitch");
  if (p.out_row_stats) IDIFF_REQUIRE(p.NT == p.N, "conv_gemm: out_row_stats needs NT == N");
  if (p.gn_groups > 0) {
    IDIFF_REQUIRE(p.gn_partial && p.N % p.gn_groups == 0, "conv_gemm: gn_partial / groups");
    const int cpg = p.N / p.gn_groups;
    IDIFF_REQUIRE((cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64) && p.NT % cpg == 0 && p.NT / cpg <= 8,
                  "conv_gemm: unsupported channels-per-group %d", cpg);
    IDIFF_REQUIRE(p.epi == IDIFF_EPI_PLAIN, "conv_gemm: GroupNorm partials need the plain epilogue");
  }
  if (p.res0_scale) IDIFF_REQUIRE(p.res0 && p.res0_shift, "conv_gemm: res0 affine needs res0 and shift");
  return IDIFF_OK;
}

}  // namespace idiff

extern "C" {

int idiff_conv_gemm_smem_bytes(const idiff_gemm_params* p) {
  using namespace idiff;
  IDIFF_REQUIRE(p, "conv_gemm: null params");
  int rc = validate(*p);
  if (rc) return rc;
  return plan_smem(*p, make_geom(p->ksize, p->stride)).total;
}

int idiff_conv_gemm(const idiff_gemm_params* pp, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(pp, "conv_gemm: null params");
  int rc = validate(*pp);
  if (rc) return rc;
  KArgs a;
  a.p = *pp;
  a.g = make_geom(pp->ksize, pp->stride);
  const SmemPlan s = plan_smem(*pp, a.g);
  a.SA = s.SA; a.SB = s.SB; a.offA = s.offA; a.offB = s.offB; a.stageB = s.stageB;
  a.tiles_x = (pp->W + TILE_W - 1) / TILE_W;
  a.tiles_y = (pp->H + TILE_H - 1) / TILE_H;
  IDIFF_REQUIRE(s.total <= 227 * 1024, "conv_gemm: shared memory plan %d B too large", s.total);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv_gemm attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)(a.tiles_x * a.tiles_y * pp->B), (unsigned)(pp->N / pp->NT));
  conv_gemm_kernel<<<grid, kThreads, s.total, as_stream(stream)>>>(a);
  return check_launch("conv_gemm");
}

}  // extern "C"
