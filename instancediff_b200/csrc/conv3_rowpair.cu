// 3x3 convolution with 64 output channels on FULL-WIDTH tcgen05 MMAs ("row-pair" formulation).
//
// Why: with the pixels on M and the 64 output channels on N, a M128 x N64 x K16 tcgen05.mma occupies the tensor
// pipe for 64 cycles -- as long as N = 128 -- because the 4 KB A operand is read from shared memory at 64 B/clk
// (profiles/r01s2_*: sm__pipe_tensor_subpipe_hmma_cycles_active / MMA count = 64).  The generic engine
// (conv_gemm_kernel.cuh) therefore runs the Cout = 64 layers at HALF the tensor peak: 3 x 64 cycles per filter
// column, K16 step and 128 output pixels.
//
// Here one work item is 2 output rows x 128 pixels of one image.  A (M = 128) is ONE INPUT ROW of the strip viewed
// through a shifted descriptor (filter column dx), and the N side stacks the two output rows: an input row
// r in {2i-1, 2i, 2i+1, 2i+2} feeds output row 2i through filter row dy = r - 2i + 1 and output row 2i+1 through
// dy - 1.  With the weights of one filter column stored as the N-stacked blocks [W(dy=2) ; W(dy=1) ; W(dy=0)]
// (64 rows each) the four views of an item are windows of that one buffer:
//     input row 2i-1 : B = [W0]       N =  64 -> accumulator columns [0, 64)     (output row 2i)
//     input row 2i   : B = [W1 ; W0]  N = 128 -> columns [0, 128)                (rows 2i and 2i+1)
//     input row 2i+1 : B = [W2 ; W1]  N = 128 -> columns [0, 128)
//     input row 2i+2 : B = [W2]       N =  64 -> columns [64, 128)               (output row 2i+1)
// i.e. 4 x 64 cycles per filter column, K16 step and 256 output pixels: 1.5x fewer tensor cycles than the generic
// engine, 75 % of the pipe does useful work instead of 50 %.
//
// The input rows live in a RING of row-pair stages in shared memory ([8-channel plane][2 rows x 130 pixels][8 ch],
// the UMMA no-swizzle K-major layout, 16 B per pixel and plane): a CTA walks DOWN a 128-pixel column strip, every
// item loads only its two NEW input rows (the other two are the previous item's), so each input pixel is read from
// global memory 1.016 times (the 16x8 tiles of the generic engine: 1.41 times) and the affine + SiLU transform of
// the fused GroupNorm / time modulation is applied 1.39x less often.  The first half of an item's MMAs uses the
// pair that is already resident, so the tensor pipe works while the loaders fetch the second pair.
//
// Roles (20 warps, as in the generic engine): warps 0-3 / 4-7 two epilogue groups (one accumulator buffer each,
// 32 pixels x 128 columns per warp: bias, GroupNorm partial sums, bf16, 128B-swizzled staging, one TMA store per
// output row), warps 8-15 loaders, warp 16 weight producer (73.7 KB, resident for the CTA's life), warp 17 MMA.
//
// Serves `noise = self.model(x, self.mu, t*scale, **kwargs)` (utils/sde_utils.py:198): the 64 -> 64 convolutions
// of the full-resolution and half-resolution ResBlocks (shape table rows 1 and 3).
#include "conv_gemm_kernel.cuh"

namespace idiff {

constexpr int RP_W = 128;                          // strip width = MMA M
constexpr int RP_ROW = RP_W + 2;                   // pixels of one staged input row (halo left / right)
constexpr int RP_SLOTS = 2 * RP_ROW;               // 260 pixel slots per stage (one row pair)
constexpr int RP_LBO = (RP_SLOTS | 1) * 16;        // plane pitch: odd multiple of 16 B -> conflict-free 16 B stores
constexpr int RP_STAGE = 8 * RP_LBO;               // 33408 B
constexpr int RP_NSTAGE = 3;
constexpr int RP_WBLK = 64 * 16;                   // one filter-row block of one plane: 64 n rows x 16 B
constexpr int RP_WLBO = 3 * RP_WBLK;               // weight plane pitch (blocks dy = 2, 1, 0)
constexpr int RP_WSTAGE = 8 * RP_WLBO;             // one filter column: 24576 B
constexpr int RP_WBYTES = 3 * RP_WSTAGE;           // 73728 B
constexpr int RP_OFF_OUT = kHeader;
constexpr int RP_OFF_A = RP_OFF_OUT + kEpiWarps * kStageTile;
constexpr int RP_OFF_W = RP_OFF_A + RP_NSTAGE * RP_STAGE;
constexpr int RP_OFF_P = RP_OFF_W + RP_WBYTES;
constexpr int RP_SMEM = RP_OFF_P + 256;
static_assert(RP_SMEM <= kSmemLimit, "row-pair kernel: shared memory plan too large");
static_assert(RP_STAGE % 128 == 0 && RP_OFF_A % 128 == 0 && RP_OFF_W % 128 == 0, "alignment");
constexpr int RP_SWEEPS = (RP_SLOTS + 31) / 32;    // 9 sweeps of 32 pixel slots (8 loader threads per pixel)

struct RpArgs {
  const uint8_t* src;         // bf16 [B][H][W][64]
  const float* a_scale;       // [B][64] or nullptr
  const float* a_shift;
  const uint8_t* w;           // packed by packing.py::pack_conv3_rowpair
  const float* bias;          // [64] or nullptr
  float* gn_partial;          // [B][H * nstrips * 4][8][2] or nullptr
  int B, H, W, nstrips, rows2, total_items;
  GnFuse gf;                  // fused GroupNorm finalize (gf.sums != nullptr: no partial rows)
  unsigned long long* prof;   // 16 role counters of CTA 0 (-DIDIFF_PROF builds, params.reserved0 = 1) or nullptr:
                              //  0 kernel cycles  1 items  2 ld: wait empty  3 ld: issue loads  4 ld: data + store
                              //  5 mma: wait tmem  6 mma: wait upper pair  7 mma: wait lower pair  8 mma: issue
                              //  9 epi: wait accumulator  10 epi: work (group 0)
  alignas(64) CUtensorMap tm_out;   // [B][H][W][ld] bf16, box {64, 32, 1, 1}, 128B swizzle
};

// item range of one CTA: contiguous, so that consecutive items share an input row pair
IDIFF_DEVINL int rp_first_item(int cta, int ncta, int total) { return (int)(((long long)cta * total) / ncta); }

// Sum 16 per-thread values over the 32 lanes of a warp (transpose-reduce, 16 shuffles): afterwards every lane L
// holds the warp total of value index (L >> 1).
IDIFF_DEVINL float warp_reduce16(float* t, int lane) {
#pragma unroll
  for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < w) {
        const float send = up ? t[i] : t[i + w];
        const float keep = up ? t[i + w] : t[i];
        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
  return t[0] + __shfl_xor_sync(0xffffffffu, t[0], 1);
}

// packed bf16x2 arithmetic of the loader's affine + SiLU transform (AMODE_AFFINE_SILU_PK)
IDIFF_DEVINL uint32_t fma_bf16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
IDIFF_DEVINL uint32_t tanh_bf16x2(uint32_t a) {
  uint32_t d;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
constexpr int AMODE_AFFINE_SILU_PK = 3;    // affine + SiLU evaluated in packed bf16x2 (12 instead of 36 instructions per vector)

template <int AMODE, bool GN>
__global__ void __launch_bounds__(kThreads, 1) conv3_rowpair_kernel(const __grid_constant__ RpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);
  uint64_t* emptyA = fullA + RP_NSTAGE;
  uint64_t* tmem_full = emptyA + RP_NSTAGE;        // [4]
  uint64_t* tmem_empty = tmem_full + 4;            // [4]
  uint64_t* w_bar = tmem_empty + 4;                // [3] one per filter column
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef IDIFF_PROF
  const bool prof = a.prof != nullptr && blockIdx.x == 0;
#else
  constexpr bool prof = false;
#endif
  const long long t_kernel = prof ? clock64() : 0ll;
  constexpr uint32_t tmem_cols = 512;              // 4 accumulator buffers x 128 columns (2 per MMA warp)
  const int g0 = rp_first_item(blockIdx.x, gridDim.x, a.total_items);
  const int g1 = rp_first_item(blockIdx.x + 1, gridDim.x, a.total_items);

  if (tid == 0) {
    // emptyA: every item commits once per pair it used, or twice when it is the pair's only user (strip start /
    // strip end) -> always 2 arrivals per use of a stage
    for (int i = 0; i < RP_NSTAGE; ++i) { mbar_init(&fullA[i], kLoaderWarps); mbar_init(&emptyA[i], 2); }
    for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps / 2); }
    for (int i = 0; i < 3; ++i) mbar_init(&w_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  static_assert(256 * 112 + 256 * 104 + 128 * 40 <= kThreads * 96, "setmaxnreg budget exceeds the CTA's registers");
  if (warp >= kWarpB) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  else if (warp >= kEpiWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

  if (warp < kEpiWarps) {
    // ============================== epilogue ==================================================
    // group g serves MMA warp g (items of parity g); a warp owns 32 pixels x 128 columns = both output rows
    const int quarter = warp & 3, grp = warp >> 2;
    float* pcache = reinterpret_cast<float*>(smem + RP_OFF_P);
    if (tid < 64) pcache[tid] = a.bias ? __ldg(a.bias + tid) : 0.f;
    asm volatile("bar.sync 3, 256;" ::: "memory");
    uint8_t* tile = smem + RP_OFF_OUT + warp * kStageTile;
    const uint32_t row_off = (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
    const int rows_per_img = a.H * a.nstrips * 4;
    int mine = 0;
    long long pacc9 = 0, pacc10 = 0;
    int g = g0 + grp;
    int strip = g / a.rows2, i = g - strip * a.rows2;
    for (; g < g1; g += 2) {
      const int b = strip / a.nstrips, sx = strip - b * a.nstrips;
      const int x0 = sx * RP_W + quarter * 32, y0 = 2 * i;
      const bool valid = x0 + lane < a.W;
      const int buf = grp * 2 + (mine & 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 128);
      long long tp = PROF_T();
      mbar_wait(&tmem_full[buf], (uint32_t)((mine >> 1) & 1), 202);
      tc_fence_after();
      PROF_ADD(9, tp);
      tp = PROF_T();
#pragma unroll 1
      for (int orow = 0; orow < 2; ++orow) {
        float v[64];
        tmem_ld64(taddr + orow * 64, v);
        if (orow == 1) {                                          // last TMEM read of this item
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        for_cols32(pcache, v, [](float x, float c) { return x + c; });
        for_cols32(pcache + 32, v + 32, [](float x, float c) { return x + c; });
        if (GN) {
          float gl[16];
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const float x = v[g8 * 8 + q]; s1 += x; s2 = fmaf(x, x, s2); }
            gl[2 * g8] = valid ? s1 : 0.f;
            gl[2 * g8 + 1] = valid ? s2 : 0.f;
          }
          const float tot = warp_reduce16(gl, lane);              // lane L: value L >> 1 = group*2 + {sum, sumsq}
          if ((lane & 1) == 0) {
            const int prow = ((y0 + orow) * a.nstrips + sx) * 4 + quarter;
            if (a.gf.sums) gn_fuse_add(a.gf, b, prow, lane >> 1, tot);
            else a.gn_partial[((size_t)b * rows_per_img + prow) * 16 + (lane >> 1)] = tot;
          }
        }
        uint4 pk[8];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) pk[q4] = pack_bf16x8(v + q4 * 8);
        if (lane == 0) bulk_wait_read<0>();                       // the previous row's store has read the tile
        __syncwarp();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          *reinterpret_cast<uint4*>(tile + row_off + (((uint32_t)q4 ^ swz) << 4)) = pk[q4];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&a.tm_out, tile, 0, x0, y0 + orow, b);
          bulk_commit();
        }
      }
      ++mine;
      if (grp == 0) PROF_ADD(10, tp);
      i += 2;
      while (i >= a.rows2) { i -= a.rows2; ++strip; }
    }
    if (lane == 0) bulk_wait_all();
    if (GN && a.gf.sums)                                          // last CTA of the grid: sums -> affine of the next layer
      gn_fuse_finish<kEpiThreads>(a.gf, reinterpret_cast<float2*>(smem + RP_OFF_OUT), reinterpret_cast<volatile int*>(smem + kHeader - 16),
                                  tid, 3, kEpiWarps * kStageTile / 8);
    if (prof && tid == 0) { a.prof[9] = pacc9; a.prof[10] = pacc10; a.prof[1] = g1 - g0; }
  } else if (warp < kWarpB) {
    // ============================== loaders ===================================================
    const int ltid = tid - kEpiThreads;
    const int c8 = ltid & 7, prow = ltid >> 3;                    // 32 pixel slots per sweep, 8 threads per pixel
    constexpr bool kAffine = AMODE != AMODE_NONE;
    constexpr bool kPacked = AMODE == AMODE_AFFINE_SILU_PK;
    // slot s = prow + 32 i of a stage -> (row rr of the pair, column s - 130 rr of the staged row): sweeps 0-3 are
    // row 0, sweeps 5-8 row 1, sweep 4 splits at prow = 2; sweep 8 only has 4 slots (260 = 8 * 32 + 4)
    const uint32_t row1_bits = 0x1E0u | (prow >= RP_ROW - 128 ? 0x10u : 0u);
    const uint32_t live_bits = prow < RP_SLOTS - 32 * (RP_SWEEPS - 1) ? 0x1FFu : 0x0FFu;
    const int row1_off = (a.W - RP_ROW) * 128;                    // byte offset of row 1 relative to slot s of row 0
    float sc[8], sh[8];
    uint32_t sc2[4], sh2[4];
    int aff_key = -1;
    auto load_affine = [&](int b) {
      if (!kAffine || b == aff_key) return;
      aff_key = b;
      const float* ps = a.a_scale + (size_t)b * 64 + c8 * 8;
      const float* pt = a.a_shift + (size_t)b * 64 + c8 * 8;
      const float4 s0 = ldg4(ps), s1 = ldg4(ps + 4), t0 = ldg4(pt), t1 = ldg4(pt + 4);
      sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
      sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] *= 0.5f; sh[e] *= 0.5f; }     // silu(y) = h*tanh(h) + h with h = y/2
      if (kPacked) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { sc2[e] = pack_bf16x2(sc[2 * e], sc[2 * e + 1]); sh2[e] = pack_bf16x2(sh[2 * e], sh[2 * e + 1]); }
      }
    };
    auto transform = [&](const uint4& qv) -> uint4 {
      if (!kAffine) return qv;
      if (kPacked) {
        uint4 o;
        uint32_t h;
        h = fma_bf16x2(qv.x, sc2[0], sh2[0]); o.x = fma_bf16x2(h, tanh_bf16x2(h), h);
        h = fma_bf16x2(qv.y, sc2[1], sh2[1]); o.y = fma_bf16x2(h, tanh_bf16x2(h), h);
        h = fma_bf16x2(qv.z, sc2[2], sh2[2]); o.z = fma_bf16x2(h, tanh_bf16x2(h), h);
        h = fma_bf16x2(qv.w, sc2[3], sh2[3]); o.w = fma_bf16x2(h, tanh_bf16x2(h), h);
        return o;
      }
      float f[8];
      unpack_bf16x8(qv, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float h = fmaf(f[e], sc[e], sh[e]);
        f[e] = fmaf(h, tanh_fast(h), h);
      }
      return pack_bf16x8(f);
    };

    // sequence of row pairs this CTA stages: for every item its lower pair, preceded by the upper pair when the
    // item starts a strip (or the CTA's range)
    int eg = g0, estrip = g0 / a.rows2, ei = g0 - estrip * a.rows2;
    bool head = true;
    int x_strip = -1;
    uint32_t xm = 0;                                              // sweeps whose column lies inside the image
    auto next_event = [&](int& strip_o, int& k_o) -> bool {
      if (eg >= g1) return false;
      strip_o = estrip;
      if (head) { head = false; k_o = ei; return true; }
      k_o = ei + 1;
      ++eg;
      if (++ei == a.rows2) { ei = 0; ++estrip; head = true; }
      return true;
    };
    long long pacc2 = 0, pacc3 = 0, pacc4 = 0;
    // puts one pair's loads in flight: q[i] = pixel slot prow + 32 i (zeros outside the image)
    auto fetch = [&](uint4* q, uint32_t& okm, int& img) -> bool {
      int strip, k;
      if (!next_event(strip, k)) return false;
      long long tp = PROF_T();
      const int b = strip / a.nstrips, sx = strip - b * a.nstrips;
      const int ytop = 2 * k - 1, xl = sx * RP_W - 1;
      img = b;
      if (strip != x_strip) {                                     // column mask of this strip (once per strip)
        x_strip = strip;
        xm = 0;
#pragma unroll
        for (int i = 0; i < RP_SWEEPS; ++i) {
          const int s = prow + 32 * i;
          const int x = xl + (((row1_bits >> i) & 1u) ? s - RP_ROW : s);
          xm |= (x >= 0 && x < a.W ? 1u : 0u) << i;
        }
      }
      const uint8_t* org = a.src + (((long long)b * a.H + ytop) * a.W + xl + prow) * 128 + c8 * 16;
      const uint32_t m = live_bits & xm & ((ytop >= 0 ? ~row1_bits : 0u) | (ytop + 1 < a.H ? row1_bits : 0u));
#pragma unroll
      for (int i = 0; i < RP_SWEEPS; ++i) {
        q[i] = make_uint4(0u, 0u, 0u, 0u);
        if ((m >> i) & 1u)
          q[i] = __ldg(reinterpret_cast<const uint4*>(org + i * 4096 + (((row1_bits >> i) & 1u) ? row1_off : 0)));
      }
      okm = m;
      PROF_ADD(3, tp);
      return true;
    };
    Ring ra;
    ra.init(RP_NSTAGE);
    bool first_lap = true;
    auto consume = [&](const uint4* q, uint32_t okm, int img) {
      load_affine(img);
      long long tp = PROF_T();
      if (!first_lap) mbar_wait(&emptyA[ra.slot], ra.phase ^ 1, 201);
      PROF_ADD(2, tp);
      tp = PROF_T();
      uint8_t* stage = smem + RP_OFF_A + ra.slot * RP_STAGE + c8 * RP_LBO + prow * 16;
#pragma unroll
      for (int i = 0; i < RP_SWEEPS; ++i) {
        if (i < RP_SWEEPS - 1 || (live_bits >> i) & 1u) {
          uint4 o = transform(q[i]);                               // the zero padding applies AFTER the activation
          if (kAffine && !((okm >> i) & 1u)) o = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(stage + i * 512) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullA[ra.slot]);
      ra.advance();
      if (ra.slot == 0) first_lap = false;
      PROF_ADD(4, tp);
    };
    // two register buffers: the next pair's loads are in flight while this one is transformed and stored
    uint4 qa[RP_SWEEPS], qb[RP_SWEEPS];
    uint32_t ma = 0, mb = 0;
    int ia = 0, ib = 0;
    bool ha = fetch(qa, ma, ia), hb = false;
    while (ha) {
      hb = fetch(qb, mb, ib);
      consume(qa, ma, ia);
      if (!hb) break;
      ha = fetch(qa, ma, ia);
      consume(qb, mb, ib);
    }
    if (prof && ltid == 0) { a.prof[2] = pacc2; a.prof[3] = pacc3; a.prof[4] = pacc4; }
  } else if (warp == kWarpB) {
    // ============================== weight producer ===========================================
    if (elect_one()) {
      for (int dx = 0; dx < 3; ++dx) {
        mbar_arrive_expect_tx(&w_bar[dx], (uint32_t)RP_WSTAGE);
        bulk_g2s(smem + RP_OFF_W + dx * RP_WSTAGE, a.w + dx * RP_WSTAGE, (uint32_t)RP_WSTAGE, &w_bar[dx]);
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ============================== two MMA issuers ===========================================
    // Issuer j owns the items of parity j and accumulator buffers 2j, 2j+1.  A tcgen05.mma queue is shallow: what
    // one issuing thread does between two MMAs (waits, descriptor arithmetic) is not overlapped with the tensor
    // pipe, so two warps alternate -- while one waits for a row pair or an accumulator, the other one's MMAs run.
    // Both walk the whole item sequence (the stage ring advances identically) and issue only their own items.
    const int me = warp - kWarpMma;
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc128 = umma_idesc_bf16(TILE_M, 128, 0), idesc64 = umma_idesc_bf16(TILE_M, 64, 0);
    const uint32_t ab_hi = umma_desc_hi(128);                     // A and B: 8-row groups 128 B apart
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem + RP_OFF_A), RP_LBO);
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(smem + RP_OFF_W), RP_WLBO);
    constexpr uint32_t kA_KK = (2 * RP_LBO) >> 4, kB_KK = (2 * RP_WLBO) >> 4, kStageU = RP_STAGE >> 4;
    constexpr uint32_t kW1 = RP_WBLK >> 4, kW0 = (2 * RP_WBLK) >> 4, kWdx = RP_WSTAGE >> 4;
    Ring ra;
    ra.init(RP_NSTAGE);
    int m = 0;                                                    // items issued by this warp
    int i = g0 - (g0 / a.rows2) * a.rows2;
    int top_slot = 0, top_phase = 0;
    bool fresh = true;
    long long pacc5 = 0, pacc6 = 0, pacc7 = 0, pacc8 = 0;
    for (int g = g0; g < g1; ++g) {
      if (fresh) { top_slot = ra.slot; top_phase = ra.phase; ra.advance(); }
      const int bot_slot = ra.slot, bot_phase = ra.phase;
      ra.advance();
      const bool strip_end = (i + 1 == a.rows2) || (g + 1 == g1);
      if (((g - g0) & 1) == me) {
        const int buf = me * 2 + (m & 1);
        long long tp = PROF_T();
        if (m >= 2) mbar_wait(&tmem_empty[buf], (uint32_t)(((m >> 1) & 1) ^ 1), 207);
        PROF_ADD(5, tp);
        tp = PROF_T();
        const uint32_t tacc = tmem_u + (uint32_t)(buf * 128);
        mbar_wait(&fullA[top_slot], (uint32_t)top_phase, 204);
        tc_fence_after();
        PROF_ADD(6, tp);
        tp = PROF_T();
        const uint32_t top_lo = a_lo0 + (uint32_t)top_slot * kStageU;
        // ---- first half: the upper pair.  Input row 2i (second row of the pair) first: its N = 128 MMA
        // initialises all 128 accumulator columns; input row 2i-1 then adds into columns [0, 64).
        {
          uint32_t w_lo = w_lo0;
#pragma unroll 1
          for (int dx = 0; dx < 3; ++dx) {
            if (m == 0) { mbar_wait(&w_bar[dx], 0, 206); tc_fence_after(); }
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                umma_bf16_lohi(tacc, top_lo + (uint32_t)(RP_ROW + dx) + kk * kA_KK, ab_hi, w_lo + kW1 + kk * kB_KK, ab_hi,
                               idesc128, (dx | kk) == 0 ? 0u : 1u);
                umma_bf16_lohi(tacc, top_lo + (uint32_t)dx + kk * kA_KK, ab_hi, w_lo + kW0 + kk * kB_KK, ab_hi, idesc64,
                               1u);
              }
            }
            w_lo += kWdx;
          }
        }
        if (leader) {
          umma_commit(&emptyA[top_slot]);
          if (fresh) umma_commit(&emptyA[top_slot]);              // no item above shares this pair
        }
        __syncwarp();
        PROF_ADD(8, tp);
        tp = PROF_T();
        // ---- second half: the lower pair (input rows 2i+1, 2i+2), shared with the next item of the strip
        mbar_wait(&fullA[bot_slot], (uint32_t)bot_phase, 205);
        tc_fence_after();
        PROF_ADD(7, tp);
        tp = PROF_T();
        const uint32_t bot_lo = a_lo0 + (uint32_t)bot_slot * kStageU;
        {
          uint32_t w_lo = w_lo0;
#pragma unroll 1
          for (int dx = 0; dx < 3; ++dx) {
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                umma_bf16_lohi(tacc, bot_lo + (uint32_t)dx + kk * kA_KK, ab_hi, w_lo + kk * kB_KK, ab_hi, idesc128, 1u);
                umma_bf16_lohi(tacc + 64u, bot_lo + (uint32_t)(RP_ROW + dx) + kk * kA_KK, ab_hi, w_lo + kk * kB_KK, ab_hi,
                               idesc64, 1u);
              }
            }
            w_lo += kWdx;
          }
        }
        if (leader) {
          umma_commit(&emptyA[bot_slot]);
          if (strip_end) umma_commit(&emptyA[bot_slot]);          // no item below shares this pair
          umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        PROF_ADD(8, tp);
        ++m;
      }
      fresh = strip_end;
      top_slot = bot_slot;
      top_phase = bot_phase;
      if (++i == a.rows2) i = 0;
    }
    if (prof && leader && me == 0) { a.prof[5] = pacc5; a.prof[6] = pacc6; a.prof[7] = pacc7; a.prof[8] = pacc8; }
  }

  tc_fence_before();
  __syncthreads();
  if (prof && tid == 0) a.prof[0] = clock64() - t_kernel;
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int watchdog_rowpair(int clear) { return watchdog_read_tu(clear); }

template <int AMODE, bool GN>
static cudaError_t launch_rp(const RpArgs& a, int grid, cudaStream_t st) {
  static DeviceOnce once;
  cudaError_t e = per_device_setup(once, nullptr, [] {
    return cudaFuncSetAttribute(conv3_rowpair_kernel<AMODE, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_SMEM);
  });
  if (e != cudaSuccess) return e;
  conv3_rowpair_kernel<AMODE, GN><<<grid, kThreads, RP_SMEM, st>>>(a);
  return cudaSuccess;
}

int make_map_box(CUtensorMap* tm, const void* base, int cols, int ld, int B, int H, int W, int box_w, int box_h);  // conv_gemm.cu
unsigned long long* prof_buffer();                                                                                  // conv_gemm.cu

}  // namespace idiff

extern "C" {

int idiff_conv3_rowpair_gn_rows(int H, int W) { return H * ((W + idiff::RP_W - 1) / idiff::RP_W) * 4; }

int idiff_conv3_rowpair_supported(const idiff_gemm_params* p) {
  if (!p) return 0;
  return p->ksize == 3 && p->stride == 1 && !p->up0 && p->cin0 == 64 && p->cin1 == 0 && p->N == 64 && p->NT == 64 &&
         p->epi == IDIFF_EPI_PLAIN && !p->res0 && !p->res1 && !p->bias_img && !p->row_stats && !p->out_row_stats &&
         p->w_image_stride == 0 && (p->src0_ld == 0 || p->src0_ld == 64) && p->H % 2 == 0 &&
         (p->a_scale == nullptr || p->a_silu);
}

int idiff_conv3_rowpair(const idiff_gemm_params* p, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(p, "conv3_rowpair: null params");
  IDIFF_REQUIRE(idiff_conv3_rowpair_supported(p), "conv3_rowpair: built for 3x3 stride-1 64 -> 64 layers without residuals "
                "(even H, plain epilogue, A transform none or affine+SiLU)");
  IDIFF_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->src0 && p->w && p->out, "conv3_rowpair: null / empty arguments");
  IDIFF_REQUIRE((p->a_scale == nullptr) == (p->a_shift == nullptr), "conv3_rowpair: a_scale/a_shift must come together");
  IDIFF_REQUIRE(aligned16(p->src0) && aligned16(p->w) && aligned16(p->out), "conv3_rowpair: 16 B alignment");
  IDIFF_REQUIRE(p->out_ld >= 64 && p->out_ld % 8 == 0, "conv3_rowpair: bad out_ld");
  IDIFF_REQUIRE(p->gn_groups == 0 || (p->gn_groups == 8 && (p->gn_partial || p->gn_fuse)), "conv3_rowpair: GroupNorm partials need 8 groups");
  RpArgs a;
  a.src = reinterpret_cast<const uint8_t*>(p->src0);
  a.a_scale = p->a_scale;
  a.a_shift = p->a_shift;
  a.w = reinterpret_cast<const uint8_t*>(p->w);
  a.bias = p->bias;
  a.gn_partial = p->gn_groups ? p->gn_partial : nullptr;
  IDIFF_REQUIRE(gn_fuse_make(a.gf, p->gn_groups ? p->gn_fuse : nullptr, p->B, 64, 8), "conv3_rowpair: incomplete gn_fuse description");
  a.B = p->B; a.H = p->H; a.W = p->W;
  a.nstrips = (p->W + RP_W - 1) / RP_W;
  a.rows2 = p->H / 2;
  const long long total = (long long)p->B * a.nstrips * a.rows2;
  IDIFF_REQUIRE(total < (1ll << 30), "conv3_rowpair: too many items");
  a.total_items = (int)total;
  a.prof = nullptr;
#ifdef IDIFF_PROF
  if (p->reserved0) a.prof = prof_buffer();
#endif
  int rc = make_map_box(&a.tm_out, p->out, 64, p->out_ld, p->B, p->H, p->W, 32, 1);
  if (rc) return rc;
  static DeviceOnce once;
  int num_sms = 0;
  cudaError_t e = per_device_setup(once, &num_sms, [] { return cudaSuccess; });
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv3_rowpair setup: %s", cudaGetErrorString(e));
  const int grid = a.total_items < num_sms ? a.total_items : num_sms;
  const bool gn = a.gn_partial != nullptr || a.gf.sums != nullptr;
  if (p->a_scale && (p->a_silu & 2))          // a_silu bit 1: evaluate affine + SiLU in packed bf16x2
    e = gn ? launch_rp<AMODE_AFFINE_SILU_PK, true>(a, grid, as_stream(stream)) : launch_rp<AMODE_AFFINE_SILU_PK, false>(a, grid, as_stream(stream));
  else if (p->a_scale) e = gn ? launch_rp<AMODE_AFFINE_SILU, true>(a, grid, as_stream(stream)) : launch_rp<AMODE_AFFINE_SILU, false>(a, grid, as_stream(stream));
  else e = gn ? launch_rp<AMODE_NONE, true>(a, grid, as_stream(stream)) : launch_rp<AMODE_NONE, false>(a, grid, as_stream(stream));
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv3_rowpair attr: %s", cudaGetErrorString(e));
  return check_launch("conv3_rowpair");
}

}  // extern "C"
