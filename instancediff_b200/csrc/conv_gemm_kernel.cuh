// Implicit-GEMM convolution / linear engine on tcgen05 tensor cores (sm_100a) -- device side.
//
// Persistent kernel, one CTA per SM, work item = one 16x8 output-pixel tile (M = 128 rows) x NT channels.
// 20 warps, specialised and pipelined ACROSS items through mbarrier rings and a double-buffered TMEM
// accumulator (2 x NT fp32 columns).  The kernel is a template over (NT, kernel size, epilogue kind, A-operand
// transform) so every instantiation only carries the code of its own layer family (the monolithic version was
// instruction-fetch bound: 36 % of the epilogue's stall samples were `no_inst`).
//
//   warps 0-3 / 4-7 : two epilogue groups; group g owns accumulator buffer g (even / odd items of this CTA).  Warp w
//                of a group owns TMEM lanes 32*(w%4).. = 4 tile rows x 8 pixels and ALL NT columns of them, so
//                LayerNorm / row statistics are thread-local and GroupNorm partial sums are warp-local (no
//                cross-warp exchange, no block barrier).  Output goes through a 128B-swizzled shared-memory
//                staging tile and ONE bulk-tensor (TMA) store per 64 channels; residual tiles arrive the same
//                way (TMA load, prefetched before the accumulator is ready) when NT == 64.
//   warps 8-15 : A producers.  Per 64-channel chunk they load the input patch (tile + halo) ONCE from global
//                memory (16 B loads at per-thread constant offsets; border tiles take a masked path), apply the
//                optional per-(image,channel) affine (+ SiLU) = fused GroupNorm / time modulation / activation
//                of the previous layer, and store it as [8-channel group][patch pixel][8 ch] -- the UMMA
//                SWIZZLE_NONE K-major canonical layout with a 16 B row pitch.  Every filter tap is then a
//                descriptor with a shifted start address, so the patch is re-used k*k times from shared memory.
//   warp 16    : weight producer.  Weights stay RESIDENT in shared memory for the whole CTA when they fit
//                (one bulk-TMA burst), otherwise pre-packed (chunk, tap) stages stream through a ring.
//   warp 17    : TMEM owner + MMA issuer (warp-uniform; tcgen05.mma / commit predicated on one elected lane).
//
// Network spec: SURVEY.md App. A (the reference's models/modules/* is not in the snapshot); the call this serves
// is `noise = self.model(x, self.mu, t*scale, **kwargs)`, utils/sde_utils.py:198.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gn_fuse.cuh"
#include "host_common.h"

namespace idiff {

constexpr int TILE_H = 16, TILE_W = 8, TILE_M = 128;
constexpr int kEpiWarps = 8;
constexpr int kLoaderWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32, kLoaderThreads = kLoaderWarps * 32;
constexpr int kWarpB = kEpiWarps + kLoaderWarps;   // 16
constexpr int kWarpMma = kWarpB + 1;               // 17
constexpr int kThreads = 20 * 32;                  // 5 warpgroups (warps 18-19 idle): setmaxnreg works per warpgroup
constexpr int kMaxSA = 4, kMaxSB = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHeader = 1024;                      // barriers + TMEM slot
constexpr int kImgBytes = 2 * 3 * 256 * 4;         // per-image column parameters, one set per epilogue group
constexpr int kStageTile = 4096;                   // one staging tile: 32 pixels x 64 channels bf16 (128B swizzle)

// A-operand transform applied by the loaders
enum { AMODE_NONE = 0, AMODE_AFFINE = 1, AMODE_AFFINE_SILU = 2 };

template <int KS>
struct GeomT {
  static constexpr int S = KS == 4 ? 2 : 1;
  static constexpr int PAD = KS == 1 ? 0 : 1;
  static constexpr int PR = (TILE_H - 1) * S + KS, PC = (TILE_W - 1) * S + KS;
  static constexpr int NSLOT = PR * PC;
  static constexpr int LBO = (NSLOT | 1) * 16;      // odd multiple of 16 B: conflict-free 16 B stores
  static constexpr int SBO = S * PC * 16;           // pitch between the 8-pixel rows of the output tile
  static constexpr int STAGE = ((8 * LBO + 127) / 128) * 128;
  static constexpr int NTAPS = KS * KS;
  __host__ __device__ static constexpr int slot(int v, int u) {
    return S == 1 ? v * PC + u : v * PC + (u & 1) * (PC / 2) + (u >> 1);
  }
};

struct KArgs {
  idiff_gemm_params p;
  int SA, SB, resident, offA, offB, offP, offImg, offOut, offRes;
  int nbuf_out;                       // staging tiles per epilogue warp for the output (1 or 2)
  int res_stride;                     // bytes of ONE residual staging set per epilogue warp (4 KB per TMA residual)
  int res_nbuf;                       // residual staging sets per warp: 2 = the next item's tiles are prefetched
  int tiles_x, tiles_y, ntiles_n, total_items;
  unsigned long long* prof;           // 16 counters of CTA 0 (IDIFF_PROF builds), or nullptr
  GnFuse gf;                          // fused GroupNorm finalize (gf.sums == nullptr: partial rows are written instead)
  alignas(64) CUtensorMap tm_out;     // [B][H][W][out cols] bf16, box {64, 8, 4, 1}, 128B swizzle
  alignas(64) CUtensorMap tm_res0;    // residuals (NT == 64 only), same box
  alignas(64) CUtensorMap tm_res1;
};

// Per-role cycle counters of CTA 0 (enabled with params.reserved0 = 1 in -DIDIFF_PROF builds):
//  0 kernel cycles   1 items of CTA 0
//  2 loader: wait emptyA   3 loader: issue loads   4 loader: wait data + transform + store + arrive
//  5 mma: wait tmem_empty  6 mma: wait fullA       7 mma: wait fullB   8 mma: issue
//  9 epi: wait tmem_full  10 epi: work (group 0)
#ifdef IDIFF_PROF
#define PROF_T() (prof ? clock64() : 0ll)
#define PROF_ADD(i, t0) do { if (prof) { pacc##i += clock64() - (t0); } } while (0)
#else                                   // production build: the hooks compile to nothing
#define PROF_T() 0ll
#define PROF_ADD(i, t0) do { (void)(t0); } while (0)
#endif

IDIFF_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- bulk-tensor (TMA) copies between a shared-memory staging tile and a 4-D bf16 tensor ----------
IDIFF_DEVINL void tma_store_4d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
IDIFF_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
IDIFF_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
IDIFF_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
IDIFF_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Work-item iterator: item -> (n tile, tile x, tile y, image), advanced by the grid stride with carries only
// (the three integer divisions happen once per thread, not once per item).
struct ItemIter {
  int item, nt, tx, ty, b;            // current item
  int s_nt, s_tx, s_ty, s_b;          // decomposition of the stride (gridDim.x)
  int ntiles_n, tiles_x, tiles_y, total;
  IDIFF_DEVINL void init(const KArgs& a, int first, int stride) {
    ntiles_n = a.ntiles_n; tiles_x = a.tiles_x; tiles_y = a.tiles_y; total = a.total_items;
    item = first;
    decompose(first, nt, tx, ty, b);
    decompose(stride, s_nt, s_tx, s_ty, s_b);
  }
  IDIFF_DEVINL void decompose(int v, int& n, int& x, int& y, int& bb) const {
    n = v % ntiles_n;                                  // N tiles of one pixel tile are adjacent: A patch hits L2
    int m = v / ntiles_n;
    x = m % tiles_x;
    m /= tiles_x;
    y = m % tiles_y;
    bb = m / tiles_y;
  }
  IDIFF_DEVINL bool valid() const { return item < total; }
  IDIFF_DEVINL void next() {
    item += s_nt + ntiles_n * (s_tx + tiles_x * (s_ty + tiles_y * s_b));   // == gridDim.x (kept as a sum of parts)
    nt += s_nt;
    int c = nt >= ntiles_n;
    nt -= c ? ntiles_n : 0;
    tx += s_tx + c;
    c = tx >= tiles_x;
    tx -= c ? tiles_x : 0;
    ty += s_ty + c;
    c = ty >= tiles_y;
    ty -= c ? tiles_y : 0;
    b += s_b + c;
  }
  IDIFF_DEVINL int tile_in_img() const { return ty * tiles_x + tx; }
  IDIFF_DEVINL int oy0() const { return ty * TILE_H; }
  IDIFF_DEVINL int ox0() const { return tx * TILE_W; }
};

// ring position with explicit wrap (no modulo: stays in the uniform datapath for the single-issuer warps)
struct Ring {
  int slot, phase, n;
  IDIFF_DEVINL void init(int stages) { slot = 0; phase = 0; n = stages; }
  IDIFF_DEVINL void advance() {
    if (++slot == n) { slot = 0; phase ^= 1; }
  }
};

// v[0..31] op= per-column parameters p[n..n+32) staged in shared memory (8 broadcast LDS.128)
template <typename F>
IDIFF_DEVINL void for_cols32(const float* __restrict__ p, float* v, F f) {
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 t = *reinterpret_cast<const float4*>(p + q4 * 4);   // shared memory, warp-broadcast
    v[q4 * 4] = f(v[q4 * 4], t.x);
    v[q4 * 4 + 1] = f(v[q4 * 4 + 1], t.y);
    v[q4 * 4 + 2] = f(v[q4 * 4 + 2], t.z);
    v[q4 * 4 + 3] = f(v[q4 * 4 + 3], t.w);
  }
}

// Sum 8 per-thread values over the 32 lanes of a warp (transpose-reduce, 9 shuffles): afterwards every lane L
// holds the warp total of value index (L >> 2).
IDIFF_DEVINL float warp_reduce8(float* t, int lane) {
#pragma unroll
  for (int w = 4, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < w) {
        const float send = up ? t[i] : t[i + w];
        const float keep = up ? t[i + w] : t[i];
        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
  float r = t[0] + __shfl_xor_sync(0xffffffffu, t[0], 2);
  return r + __shfl_xor_sync(0xffffffffu, r, 1);
}

IDIFF_DEVINL void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

template <int NT, int KS, int EPI, int AMODE>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ KArgs a) {
  using G = GeomT<KS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const idiff_gemm_params& p = a.p;

  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);
  uint64_t* emptyA = fullA + kMaxSA;
  uint64_t* fullB = emptyA + kMaxSA;
  uint64_t* emptyB = fullB + kMaxSB;
  uint64_t* tmem_full = emptyB + kMaxSB;           // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* wres_bar = tmem_empty + 2;             // [16] resident weights landed (per 64-channel chunk)
  uint64_t* res_bar = wres_bar + 16;               // [8][2] residual tiles landed (per epilogue warp and staging set)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef IDIFF_PROF
  const bool prof = a.prof != nullptr && p.reserved0 != 0 && blockIdx.x == 0;
#else
  constexpr bool prof = false;
#endif
  const long long t_kernel = prof ? clock64() : 0ll;
  const int cin = p.cin0 + p.cin1, nchunks = cin >> 6;
  constexpr int ntaps = G::NTAPS;
  const int nk = nchunks * ntaps;
  constexpr uint32_t tmem_cols = 2 * NT;                          // double-buffered accumulator
  constexpr int stageB = NT * 128;

  if (tid == 0) {
    for (int i = 0; i < a.SA; ++i) { mbar_init(&fullA[i], kLoaderWarps); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < a.SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps / 2); }
    for (int i = 0; i < 16; ++i) mbar_init(&wres_bar[i], 1);
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register re-balancing (per warpgroup).  The CTA owns 640 x 96 = 61440 registers (launch bounds); `inc` can
  // only draw on what `dec` released inside the CTA, so the targets must satisfy
  //   256*EPI + 256*LOADER + 128*OTHER <= 61440   ->   112 / 104 / 40  (= 60416).
  static_assert(256 * 112 + 256 * 104 + 128 * 40 <= kThreads * 96, "setmaxnreg budget exceeds the CTA's registers");
  if (warp >= kWarpB) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  else if (warp >= kEpiWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

  if (warp < kEpiWarps) {
    // ============================== epilogue (TMEM -> registers -> staging -> TMA store) ======
    const int quarter = warp & 3, grp = warp >> 2;                // TMEM lane quarter / accumulator buffer
    const int gtid = tid & 127;
    const int r = quarter * 32 + lane, ti = r >> 3, tj = r & 7;   // accumulator row = tile pixel
    constexpr int NC = NT / 32;                                   // 32-column chunks per row
    const int cpg = p.N >> 3;                                     // channels per GroupNorm group (8 groups over ALL N
                                                                  // channels; an N tile holds NT / cpg whole groups)
    constexpr bool kTmaRes = NT == 64;
    constexpr bool kTmaOut = EPI != IDIFF_EPI_GEGLU;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* res0 = reinterpret_cast<const __nv_bfloat16*>(p.res0);
    const __nv_bfloat16* res1 = reinterpret_cast<const __nv_bfloat16*>(p.res1);
    const bool gn = p.gn_groups > 0;                              // NT is a multiple of the group width
    const float invN = 1.f / (float)p.N;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    // per-layer column parameters staged once per CTA: [bias | wsum | ln_g], N floats each
    float* pcache = reinterpret_cast<float*>(smem + a.offP);
    for (int i = tid; i < p.N; i += kEpiThreads) {
      pcache[i] = p.bias ? __ldg(p.bias + i) : 0.f;
      pcache[p.N + i] = p.wsum ? __ldg(p.wsum + i) : 0.f;
      pcache[2 * p.N + i] = p.ln_g ? __ldg(p.ln_g + i) : 1.f;
    }
    const bool img_params = p.bias_img != nullptr || p.res0_scale != nullptr;
    float* pimg = reinterpret_cast<float*>(smem + a.offImg) + grp * (3 * 256);   // [bias_img | res0_scale | res0_shift][256]
    uint8_t* ostage = smem + a.offOut + warp * (a.nbuf_out * kStageTile);
    uint8_t* rstage_w = smem + a.offRes + warp * (a.res_nbuf * a.res_stride);    // [set][res0 tile | res1 tile]
    const int res1_off = p.res0 ? kStageTile : 0;
    uint64_t* rbar_w = &res_bar[2 * warp];
    const bool res_tma = kTmaRes && (res0 || res1);
    const bool res_pref = a.res_nbuf > 1;
    // puts the residual tiles of item `ri` (this group's u-th item) in flight into staging set u & (nbuf-1)
    auto issue_residuals = [&](const ItemIter& ri, int u) {
      if (lane == 0) {
        const int set = res_pref ? (u & 1) : 0;
        uint8_t* dst = rstage_w + set * a.res_stride;
        uint64_t* bar = rbar_w + set;
        mbar_arrive_expect_tx(bar, (uint32_t)kStageTile * ((res0 ? 1u : 0u) + (res1 ? 1u : 0u)));
        if (res0) tma_load_4d(dst, &a.tm_res0, ri.nt * NT, ri.ox0(), ri.oy0() + 4 * quarter, ri.b, bar);
        if (res1) tma_load_4d(dst + res1_off, &a.tm_res1, ri.nt * NT, ri.ox0(), ri.oy0() + 4 * quarter, ri.b, bar);
      }
    };
    const uint32_t row_off = (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
    const bool two_bufs = a.nbuf_out > 1;
    int obuf = 0;
    asm volatile("bar.sync 3, 256;" ::: "memory");                // pcache visible to both groups

    long long pacc9 = 0, pacc10 = 0;
    int mine = 0, img_key = -1;
    ItemIter it;
    it.init(a, blockIdx.x + grp * gridDim.x, 2 * gridDim.x);      // this group's items: every second one of the CTA
    if (res_tma && res_pref && it.valid()) issue_residuals(it, 0);
    // Residual rows read straight from global memory (NT > 64: no room for TMA staging tiles) are pulled into L2 one
    // item ahead -- the epilogue's loads are synchronous (load -> use), so an HBM miss per 32-column chunk was ~1 us of
    // exposed latency per chunk and group.
    const bool res_l2 = !kTmaRes && (res0 || res1);
    auto prefetch_residuals = [&](const ItemIter& ri) {
      const int oy2 = ri.oy0() + ti, ox2 = ri.ox0() + tj;
      if (oy2 < p.H && ox2 < p.W) {
        const size_t e2 = (((size_t)ri.b * p.H + oy2) * p.W + ox2) * p.N + ri.nt * NT;
#pragma unroll
        for (int l = 0; l < NT / 64; ++l) {
          if (res0) prefetch_l2(res0 + e2 + l * 64);
          if (res1) prefetch_l2(res1 + e2 + l * 64);
        }
      }
    };
    if (res_l2 && it.valid()) prefetch_residuals(it);
    for (; it.valid(); it.next()) {
      const int b = it.b, n0 = it.nt * NT;
      const int oy = it.oy0() + ti, ox = it.ox0() + tj;
      const bool valid = (oy < p.H) && (ox < p.W);
      const size_t m = ((size_t)b * p.H + (valid ? oy : 0)) * p.W + (valid ? ox : 0);   // clamped: loads stay legal
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * NT);
      const int by = it.oy0() + 4 * quarter;                      // first tile row of this warp's TMA box

      // residual tiles by TMA, issued before the accumulator is ready
      const uint8_t* rstage = rstage_w + (res_pref ? (mine & 1) * a.res_stride : 0);
      if (res_tma) {
        __syncwarp();                                             // every lane is done with the tile being replaced
        if (res_pref) {                                           // next item's tiles: a whole item of lead time
          ItemIter nx = it;
          nx.next();
          if (nx.valid()) issue_residuals(nx, mine + 1);
        } else {
          issue_residuals(it, 0);
        }
      }
      if (res_l2) {
        ItemIter nx = it;
        nx.next();
        if (nx.valid()) prefetch_residuals(nx);
      }
      float mean_in = 0.f, rstd_in = 1.f;
      if (p.row_stats) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.row_stats) + m);
        mean_in = st.x;
        rstd_in = st.y;
      }
      // per-image column parameters of this item's N tile -> shared memory, only when (image, N tile) changes
      if (img_params) {
        const int key = b * a.ntiles_n + it.nt;
        if (key != img_key) {
          img_key = key;
          group_bar(grp);                                         // everyone is done reading the previous set
          for (int c = gtid; c < NT; c += 128) {
            const size_t gi = (size_t)b * p.N + n0 + c;
            pimg[c] = p.bias_img ? __ldg(p.bias_img + gi) : 0.f;
            pimg[256 + c] = p.res0_scale ? __ldg(p.res0_scale + gi) : 1.f;
            pimg[512 + c] = p.res0_scale ? __ldg(p.res0_shift + gi) : 0.f;
          }
          group_bar(grp);
        }
      }
      // accumulator chunk -> value after LayerNorm fold + biases (32 columns starting at global column n)
      auto apply_base = [&](float* v, int n) {
        if (p.row_stats) for_cols32(pcache + p.N + n, v, [&](float x, float c) { return fmaf(-mean_in, c, x) * rstd_in; });
        if (p.bias) for_cols32(pcache + n, v, [](float x, float c) { return x + c; });
        if (p.bias_img) for_cols32(pimg + (n - n0), v, [](float x, float c) { return x + c; });
      };
      bool res_waited = false;
      auto add_residuals = [&](float* v, int n) {
        if (!res0 && !res1) return;
        if (kTmaRes) {
          if (!res_waited) {
            mbar_wait(rbar_w + (res_pref ? (mine & 1) : 0), (uint32_t)((res_pref ? (mine >> 1) : mine) & 1), 108);
            res_waited = true;
          }
          const uint32_t cb = (uint32_t)((n - n0) >> 3);          // first 16 B chunk of these 32 columns
          if (res0) {
            float rr[32];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              unpack_bf16x8(*reinterpret_cast<const uint4*>(rstage + row_off + (((cb + q4) ^ swz) << 4)), rr + q4 * 8);
            if (p.res0_scale) {                                   // residual enters as silu(GN(y)) (ResBlock tail)
              for_cols32(pimg + 256 + (n - n0), rr, [](float x, float c) { return x * c; });
              for_cols32(pimg + 512 + (n - n0), rr, [](float x, float c) { return silu_fast(x + c); });
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += rr[q];
          }
          if (res1) {
            float rr[32];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              unpack_bf16x8(*reinterpret_cast<const uint4*>(rstage + res1_off + row_off + (((cb + q4) ^ swz) << 4)),
                            rr + q4 * 8);
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += rr[q];
          }
        } else {
          if (res0) {
            float rr[32];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(res0 + m * p.N + n + q4 * 8)), rr + q4 * 8);
            if (p.res0_scale) {
              for_cols32(pimg + 256 + (n - n0), rr, [](float x, float c) { return x * c; });
              for_cols32(pimg + 512 + (n - n0), rr, [](float x, float c) { return silu_fast(x + c); });
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += rr[q];
          }
          if (res1) {
            float rr[32];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(res1 + m * p.N + n + q4 * 8)), rr + q4 * 8);
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += rr[q];
          }
        }
      };
      // 32 bf16 of this row into the staging tile (128B swizzle); every second chunk completes a 64-channel box
      auto store32 = [&](const float* v, int cc) {
        uint8_t* tile = ostage + obuf * kStageTile;
        if ((cc & 1) == 0) {                              // first half of a box: the tile must be free again
          if (lane == 0) {
            if (two_bufs) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
        }
        const uint32_t cb = (uint32_t)((cc & 1) * 4);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(tile + row_off + (((cb + q4) ^ swz) << 4)) = pack_bf16x8(v + q4 * 8);
        if (cc & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&a.tm_out, tile, n0 + (cc >> 1) * 64, it.ox0(), by, b);
            bulk_commit();
          }
          if (two_bufs) obuf ^= 1;
        }
      };
      auto release_tmem = [&]() {                         // this warp is done reading the accumulator buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[grp]);
      };

      long long tp = PROF_T();
      mbar_wait(&tmem_full[grp], (uint32_t)(mine & 1), 102);
      tc_fence_after();
      PROF_ADD(9, tp);
      tp = PROF_T();

      float o1 = 0.f, o2 = 0.f;
      // The chunk loops below stay ROLLED: one copy of the per-chunk code keeps the epilogue's instruction
      // footprint inside the instruction cache it shares with the loader and MMA warps.
      if (EPI == IDIFF_EPI_LN_OUT) {
        // pass 1: LayerNorm statistics over the whole row (NT == N), thread-local
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < NC; ++cc) {
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          apply_base(v, n0 + cc * 32);
#pragma unroll
          for (int q = 0; q < 32; ++q) { s1 += v[q]; s2 = fmaf(v[q], v[q], s2); }
        }
        const float mean = s1 * invN, var = fmaxf(s2 * invN - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll 1
        for (int cc = 0; cc < NC; ++cc) {
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          if (cc == NC - 1) release_tmem();
          apply_base(v, n0 + cc * 32);
          for_cols32(pcache + 2 * p.N + n0 + cc * 32, v, [&](float x, float c) { return (x - mean) * rstd * c; });
          add_residuals(v, n0 + cc * 32);
          if (p.out_row_stats) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 = fmaf(v[q], v[q], o2); }
          }
          store32(v, cc);
        }
      } else {
#pragma unroll 1
        for (int cc = 0; cc < NC; ++cc) {
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          if (cc == NC - 1) release_tmem();               // last TMEM read of this item: release the buffer early
          const int ncol0 = n0 + cc * 32;
          apply_base(v, ncol0);

          if (EPI == IDIFF_EPI_PLAIN && gn) {
            // partial sums of the conv output (bias included) over this warp's 32 pixels: the four 8-column
            // blocks of the chunk are reduced over the warp together (lane L ends up with value L >> 2 =
            // block*2 + {sum, sum of squares}); blocks of the same group (CPG = 16: pairs, CPG = 32: all four)
            // are then folded and one lane per (group, statistic) writes.
            float gl[8];
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) { const float x = v[g8 * 8 + q]; s1 += x; s2 = fmaf(x, x, s2); }
              gl[2 * g8] = valid ? s1 : 0.f;
              gl[2 * g8 + 1] = valid ? s2 : 0.f;
            }
            float tot = warp_reduce8(gl, lane);
            if (cpg >= 16) tot += __shfl_xor_sync(0xffffffffu, tot, 8);      // warp-uniform conditions
            if (cpg >= 32) tot += __shfl_xor_sync(0xffffffffu, tot, 16);
            const int g8 = lane >> 3, st = (lane >> 2) & 1;
            const int bpg = cpg >> 3;                     // 8-column blocks per group: 1, 2 or 4
            if ((lane & 3) == 0 && (g8 & (bpg - 1)) == 0) {
              const int grp_idx = ncol0 / cpg + g8 / bpg;   // group of the GLOBAL column: N tiles fill disjoint entries
              if (a.gf.sums) gn_fuse_add(a.gf, b, it.tile_in_img() * 4 + quarter, grp_idx * 2 + st, tot);
              else p.gn_partial[(((size_t)b * tiles_per_img + it.tile_in_img()) * 4 + quarter) * 16 + grp_idx * 2 + st] = tot;
            }
          }

          if (EPI == IDIFF_EPI_QSOFTMAX && ncol0 < 128) {
            float m4[4] = {v[0], v[1], v[2], v[3]};       // four independent chains instead of one 32-long one
#pragma unroll
            for (int q = 4; q < 32; ++q) m4[q & 3] = fmaxf(m4[q & 3], v[q]);
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 32; ++q) { v[q] = __expf(v[q] - mx); s4[q & 3] += v[q]; }
            const float inv = p.qscale / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] *= inv;
            store32(v, cc);
          } else if (EPI == IDIFF_EPI_GEGLU) {
            float o[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) o[q] = v[2 * q] * gelu_erf_fast(v[2 * q + 1]);
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(outp + m * p.out_ld + (ncol0 >> 1));
              dst[0] = pack_bf16x8(o);
              dst[1] = pack_bf16x8(o + 8);
            }
          } else {
            add_residuals(v, ncol0);
            if (p.out_row_stats) {
#pragma unroll
              for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 = fmaf(v[q], v[q], o2); }
            }
            store32(v, cc);
          }
        }
      }
      if (p.out_row_stats && valid) {
        const float mo = o1 * invN, vo = fmaxf(o2 * invN - mo * mo, 0.f);
        reinterpret_cast<float2*>(p.out_row_stats)[m] = make_float2(mo, rsqrtf(vo + p.ln_eps));
      }
      ++mine;
      if (grp == 0) PROF_ADD(10, tp);
    }
    if (kTmaOut && lane == 0) bulk_wait_all();              // staging tiles must outlive the last TMA store
    if (EPI == IDIFF_EPI_PLAIN && a.gf.sums)                // last CTA of the grid: sums -> affine of the next layer
      gn_fuse_finish<kEpiThreads>(a.gf, reinterpret_cast<float2*>(smem + a.offOut), reinterpret_cast<volatile int*>(smem + kHeader - 16),
                                  tid, 3, kEpiWarps * a.nbuf_out * kStageTile / 8);
    if (prof && tid == 0) { a.prof[9] = pacc9; a.prof[10] = pacc10; a.prof[1] = mine; }
  } else if (warp < kWarpB) {
    // ============================== A producers ==============================================
    const int ltid = tid - kEpiThreads;
    const int c8 = ltid & 7, prow = ltid >> 3;                        // 32 patch pixels per sweep
    constexpr int SWEEP = kLoaderThreads / 8;
    constexpr bool kSmall = KS != 4;                                  // k = 1, 3: one batch covers the patch
    constexpr int BATCH = kSmall ? (G::NSLOT + SWEEP - 1) / SWEEP : 5; // 16 B loads in flight per thread
    constexpr bool kAffine = AMODE != AMODE_NONE, kSilu = AMODE == AMODE_AFFINE_SILU;
    const int Hin = p.H * G::S, Win = p.W * G::S;                     // virtual (possibly upsampled) input extent
    const int Hs = p.up0 ? (Hin >> 1) : Hin, Ws = p.up0 ? (Win >> 1) : Win;
    long long pacc2 = 0, pacc3 = 0, pacc4 = 0;

    // per-stage source description (stage = one 64-channel chunk of one item)
    struct Src {
      const uint8_t* img;            // image b of the source tensor, channel offset applied (bytes)
      int CsB;                       // pixel pitch in bytes
      int iy0, ix0;
      int key;                       // (image, chunk) -> affine parameters
      bool interior;                 // whole patch inside the image and no upsampling: constant offsets apply
    };
    auto describe = [&](const ItemIter& it, int ch) {
      Src sd;
      const bool from0 = (ch << 6) < p.cin0;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(from0 ? p.src0 : p.src1);
      const int Cs = from0 ? (p.src0_ld ? p.src0_ld : p.cin0) : (p.src1_ld ? p.src1_ld : p.cin1);
      const int coff = (from0 ? (ch << 6) : ((ch << 6) - p.cin0)) + c8 * 8;
      sd.CsB = Cs * 2;
      sd.img = src + ((size_t)it.b * Hs * Ws * Cs + coff) * 2;
      sd.iy0 = it.oy0() * G::S - G::PAD;
      sd.ix0 = it.ox0() * G::S - G::PAD;
      sd.interior = !p.up0 && sd.iy0 >= 0 && sd.ix0 >= 0 && sd.iy0 + G::PR <= Hin && sd.ix0 + G::PC <= Win;
      sd.key = it.b * nchunks + ch;
      return sd;
    };
    // affine parameters of one (image, chunk): cached in registers, reloaded only when the key changes.
    // With SiLU the 0.5 of silu(y) = h*tanh(h) + h, h = y/2, is folded in.
    float sc[8], sh[8];
    int aff_key = -1;
    auto load_affine = [&](int key) {
      if (!kAffine || key == aff_key) return;
      aff_key = key;
      const float* ps = p.a_scale + (size_t)key * 64 + c8 * 8;       // [b][cin] with cin = nchunks*64
      const float* pt = p.a_shift + (size_t)key * 64 + c8 * 8;
      const float4 s0 = ldg4(ps), s1 = ldg4(ps + 4), t0 = ldg4(pt), t1 = ldg4(pt + 4);
      sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
      sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
      if (kSilu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[e] *= 0.5f; sh[e] *= 0.5f; }
      }
    };
    auto transform = [&](const uint4& qv) -> uint4 {
      if (!kAffine) return qv;
      float f[8];
      unpack_bf16x8(qv, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float h = fmaf(f[e], sc[e], sh[e]);
        f[e] = kSilu ? fmaf(h, tanh_fast(h), h) : h;
      }
      return pack_bf16x8(f);
    };

    Ring ra;
    ra.init(a.SA);
    bool first_lap = true;
    ItemIter it;
    it.init(a, blockIdx.x, gridDim.x);
    auto publish = [&]() {                                 // stage written: hand it to the MMA warp
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullA[ra.slot]);
      ra.advance();
      if (ra.slot == 0) first_lap = false;
    };

    if (kSmall) {
      // per-thread constants: pixel offset inside the patch (source pixels) and shared-memory slot offset
      int pix[BATCH], soff[BATCH];
      uint32_t act_mask = 0;
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int px = prow + SWEEP * i;
        const bool in_patch = px < G::NSLOT;
        const int pxc = in_patch ? px : 0;
        const int v = pxc / G::PC, u = pxc - v * G::PC;
        pix[i] = v * Ws + u;
        soff[i] = G::slot(v, u) * 16;
        act_mask |= in_patch ? (1u << i) : 0u;
      }
      // branch-light batch: interior tiles use the constant offsets, border tiles test every pixel
      auto issue = [&](const Src& sd, uint4* q) -> uint32_t {
        if (sd.interior) {
          const uint8_t* org = sd.img + ((long long)sd.iy0 * Ws + sd.ix0) * sd.CsB;
#pragma unroll
          for (int i = 0; i < BATCH; ++i) {
            const bool on = (SWEEP * (i + 1) <= G::NSLOT) || ((act_mask >> i) & 1u);
            if (on) q[i] = __ldg(reinterpret_cast<const uint4*>(org + (long long)pix[i] * sd.CsB));
          }
          return act_mask;
        }
        uint32_t mask = 0;
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int px = prow + SWEEP * i;
          const int pxc = px < G::NSLOT ? px : 0;
          const int v = pxc / G::PC, u = pxc - v * G::PC;
          const int iy = sd.iy0 + v, ix = sd.ix0 + u;
          const bool ok = px < G::NSLOT && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
          const int sy = p.up0 ? (iy >> 1) : iy, sx = p.up0 ? (ix >> 1) : ix;
          q[i] = make_uint4(0u, 0u, 0u, 0u);
          if (ok) q[i] = __ldg(reinterpret_cast<const uint4*>(sd.img + ((long long)sy * Ws + sx) * sd.CsB));
          mask |= ok ? (1u << i) : 0u;
        }
        return mask;
      };

      // Software pipeline over NB register buffers used round-robin (the loop body is unrolled NB times so every
      // buffer index is static): while stage s is transformed and stored, the loads of stages s+1 .. s+NB-1 are
      // in flight -- two further stages for the 1x1 layers, whose 16 KB stages would otherwise expose the HBM
      // latency once per stage.
      constexpr int NB = KS == 1 ? 3 : 2;
      int ch = 0;
      bool more = it.valid();
      uint4 q[NB][BATCH];
      uint32_t mk[NB];
      int key[NB];
      bool has[NB];
      auto fetch = [&](uint4* qq, uint32_t& m, int& k) -> bool {      // puts the next stage's loads in flight
        if (!more) return false;
        long long tp = PROF_T();
        const Src sd = describe(it, ch);
        m = issue(sd, qq);
        k = sd.key;
        if (++ch == nchunks) { ch = 0; it.next(); }
        more = it.valid();
        PROF_ADD(3, tp);
        return true;
      };
      auto consume = [&](const uint4* qq, uint32_t mask_c, int key_c) {
        load_affine(key_c);
        long long tp = PROF_T();
        if (!first_lap) mbar_wait(&emptyA[ra.slot], ra.phase ^ 1, 101);
        PROF_ADD(2, tp);
        tp = PROF_T();
        uint8_t* stage = smem + a.offA + ra.slot * G::STAGE + c8 * G::LBO;
        // one copy of the transform: border pixels are transformed too (their registers hold zeros) and then
        // replaced by the zero padding, which applies AFTER the activation
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const bool on = (SWEEP * (i + 1) <= G::NSLOT) || ((act_mask >> i) & 1u);
          if (on) {
            uint4 o = transform(qq[i]);
            if (!((mask_c >> i) & 1u)) o = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(stage + soff[i]) = o;
          }
        }
        publish();
        PROF_ADD(4, tp);
      };
#pragma unroll
      for (int j = 0; j < NB; ++j) has[j] = fetch(q[j], mk[j], key[j]);
      while (has[0]) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          if (!has[j]) break;
          consume(q[j], mk[j], key[j]);
          has[j] = fetch(q[j], mk[j], key[j]);
        }
      }
    } else {
      for (; it.valid(); it.next()) {
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
          const Src sd = describe(it, ch);
          load_affine(sd.key);
          long long tp = PROF_T();
          if (!first_lap) mbar_wait(&emptyA[ra.slot], ra.phase ^ 1, 101);
          PROF_ADD(2, tp);
          tp = PROF_T();
          uint8_t* stage = smem + a.offA + ra.slot * G::STAGE + c8 * G::LBO;
#pragma unroll 1
          for (int px0 = prow; px0 < G::NSLOT; px0 += SWEEP * BATCH) {
            uint4 q[BATCH];
            int slot[BATCH];
            bool inb[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              const int px = px0 + SWEEP * i;
              const bool in_patch = px < G::NSLOT;
              const int pxc = in_patch ? px : 0;
              const int v = pxc / G::PC, u = pxc - v * G::PC;
              slot[i] = in_patch ? G::slot(v, u) : -1;
              const int iy = sd.iy0 + v, ix = sd.ix0 + u;
              const bool ok = in_patch && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
              const int sy = p.up0 ? (iy >> 1) : iy, sx = p.up0 ? (ix >> 1) : ix;
              inb[i] = ok;
              q[i] = make_uint4(0u, 0u, 0u, 0u);
              if (ok) q[i] = __ldg(reinterpret_cast<const uint4*>(sd.img + ((long long)sy * Ws + sx) * sd.CsB));
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              if (slot[i] >= 0)
                *reinterpret_cast<uint4*>(stage + slot[i] * 16) = inb[i] ? transform(q[i]) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          publish();
          PROF_ADD(4, tp);
        }
      }
    }
    if (prof && ltid == 0) { a.prof[2] = pacc2; a.prof[3] = pacc3; a.prof[4] = pacc4; }
  } else if (warp == kWarpB) {
    // ============================== weight producer (bulk TMA, warp-uniform) ==================
    const bool leader = elect_one();
    const uint8_t* w0 = reinterpret_cast<const uint8_t*>(p.w);
    if (a.resident) {
      // every N tile's weights, once per CTA: [ntile][nk] stages of stageB bytes, contiguous in global memory
      // One barrier per 64-channel chunk (single N tile, <= 16 chunks) so the first MMAs start as soon as the
      // first chunk's taps have landed instead of after the whole burst -- matters for the small low-resolution
      // layers, whose run time is a few microseconds; otherwise one barrier for everything.
      const uint32_t total = (uint32_t)a.ntiles_n * (uint32_t)nk * (uint32_t)stageB;
      const bool per_chunk = a.ntiles_n == 1 && nchunks <= 16;
      const uint32_t part = per_chunk ? (uint32_t)ntaps * (uint32_t)stageB : total;
      if (leader) {
        uint32_t off = 0;
        for (int ch = 0; off < total; ++ch) {
          mbar_arrive_expect_tx(&wres_bar[ch], part);
          for (uint32_t o2 = 0; o2 < part; o2 += 32768) {
            const uint32_t n = part - o2 < 32768u ? part - o2 : 32768u;
            bulk_g2s(smem + a.offB + off + o2, w0 + off + o2, n, &wres_bar[ch]);
          }
          off += part;
        }
      }
    } else {
      Ring rb;
      rb.init(a.SB);
      bool first_lap = true;
      ItemIter it;
      for (it.init(a, blockIdx.x, gridDim.x); it.valid(); it.next()) {
        const uint8_t* wbase = w0 + 2 * ((size_t)it.b * p.w_image_stride + (size_t)it.nt * nk * NT * 64);
        for (int ks = 0; ks < nk; ++ks) {
          if (!first_lap) mbar_wait(&emptyB[rb.slot], rb.phase ^ 1, 103);
          if (leader) {
            mbar_arrive_expect_tx(&fullB[rb.slot], (uint32_t)stageB);
            bulk_g2s(smem + a.offB + rb.slot * stageB, wbase + (size_t)ks * stageB, (uint32_t)stageB, &fullB[rb.slot]);
          }
          rb.advance();
          if (rb.slot == 0) first_lap = false;
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ============================== MMA issuer (warp-uniform, one elected lane issues) =========
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = umma_idesc_bf16(TILE_M, NT, 0);
    constexpr uint32_t lboB = NT * 16, sboB = 128;
    const uint32_t a_hi = umma_desc_hi(G::SBO), b_hi = umma_desc_hi(sboB);
    const uint32_t a0 = smem_u32(smem + a.offA), b0 = smem_u32(smem + a.offB);
    long long pacc5 = 0, pacc6 = 0, pacc7 = 0, pacc8 = 0;
    const bool w_per_chunk = a.resident && a.ntiles_n == 1 && nchunks <= 16;
    if (a.resident && !w_per_chunk) {
      mbar_wait(&wres_bar[0], 0, 106);
      tc_fence_after();
    }
    // all indices below advance with compare-and-wrap only (uniform datapath; no modulo, no division)
    Ring ra, rb;
    ra.init(a.SA);
    rb.init(a.SB > 0 ? a.SB : 1);
    int ab = 0, acc_phase = 0, items_done = 0, nt = blockIdx.x % a.ntiles_n;
    const int nt_step = gridDim.x % a.ntiles_n;
    const uint32_t res_item_stride = (uint32_t)((nk * stageB) >> 4);          // descriptor units per N tile
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      long long tp = PROF_T();
      if (items_done >= 2) mbar_wait(&tmem_empty[ab], acc_phase ^ 1, 107);
      tc_fence_after();
      PROF_ADD(5, tp);
      const uint32_t tacc = tmem_u + (uint32_t)(ab * NT);
      uint32_t b_res = umma_desc_lo(b0, lboB) + (uint32_t)nt * res_item_stride;  // resident: this N tile, chunk 0
      uint32_t accum = 0;
#pragma unroll 1
      for (int ch = 0; ch < nchunks; ++ch) {
        tp = PROF_T();
        mbar_wait(&fullA[ra.slot], ra.phase, 104);
        if (w_per_chunk && items_done == 0) mbar_wait(&wres_bar[ch], 0, 106);   // first item: this chunk's weights
        tc_fence_after();
        PROF_ADD(6, tp);
        // Tap loops stay ROLLED (one copy of the 4-MMA body): the tap's patch offset advances by one pixel per
        // dx and by the rest of the patch row per dy -- adds on uniform registers only.
        uint32_t a_lo = umma_desc_lo(a0 + ra.slot * G::STAGE, G::LBO);
        if (a.resident) {
          tp = PROF_T();
          uint32_t b_lo = b_res;
#pragma unroll 1
          for (int dy = 0; dy < KS; ++dy) {
            // one filter row (KS taps = 4*KS MMAs) per iteration under a single leader branch: the row's tap
            // offsets are immediates, the loop overhead is paid once per row
            if (leader) {
#pragma unroll
              for (int dx = 0; dx < KS; ++dx) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_lohi(tacc, a_lo + (uint32_t)(G::slot(0, dx) - G::slot(0, 0)) + (uint32_t)((kk * 2 * G::LBO) >> 4), a_hi,
                                 b_lo + (uint32_t)((dx * stageB) >> 4) + (uint32_t)((kk * 2 * lboB) >> 4), b_hi, idesc,
                                 (dx | kk) == 0 ? accum : 1u);
              }
            }
            accum = 1u;
            a_lo += (uint32_t)(G::slot(1, 0) - G::slot(0, 0));
            b_lo += (uint32_t)((KS * stageB) >> 4);
          }
          b_res = b_lo;
          PROF_ADD(8, tp);
        } else {
#pragma unroll 1
          for (int dy = 0; dy < KS; ++dy) {
#pragma unroll 1
            for (int dx = 0; dx < KS; ++dx) {
              tp = PROF_T();
              mbar_wait(&fullB[rb.slot], rb.phase, 105);
              tc_fence_after();
              PROF_ADD(7, tp);
              tp = PROF_T();
              const uint32_t b_lo = umma_desc_lo(b0 + rb.slot * stageB, lboB);
              if (leader) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_lohi(tacc, a_lo + (uint32_t)((kk * 2 * G::LBO) >> 4), a_hi,
                                 b_lo + (uint32_t)((kk * 2 * lboB) >> 4), b_hi, idesc, kk == 0 ? accum : 1u);
                umma_commit(&emptyB[rb.slot]);
              }
              accum = 1u;
              rb.advance();
              a_lo += (uint32_t)(G::slot(0, dx + 1) - G::slot(0, dx));
              PROF_ADD(8, tp);
            }
            a_lo += (uint32_t)(G::slot(1, 0) - G::slot(0, KS));
          }
        }
        if (leader) umma_commit(&emptyA[ra.slot]);
        ra.advance();
      }
      if (leader) umma_commit(&tmem_full[ab]);
      __syncwarp();
      ++items_done;
      ab ^= 1;
      if (ab == 0) acc_phase ^= 1;
      nt += nt_step;
      if (nt >= a.ntiles_n) nt -= a.ntiles_n;
    }
    if (prof && leader) { a.prof[5] = pacc5; a.prof[6] = pacc6; a.prof[7] = pacc7; a.prof[8] = pacc8; }
  }

  tc_fence_before();
  __syncthreads();
  if (prof && tid == 0) a.prof[0] = clock64() - t_kernel;
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <int NT, int KS, int EPI, int AMODE>
inline cudaError_t launch_one(const KArgs& a, int grid, int smem, cudaStream_t st) {
  static DeviceOnce once;                     // one flag set per instantiation, one bit per device
  cudaError_t e = per_device_setup(once, nullptr, [] {
    return cudaFuncSetAttribute(conv_gemm_kernel<NT, KS, EPI, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemLimit);
  });
  if (e != cudaSuccess) return e;
  conv_gemm_kernel<NT, KS, EPI, AMODE><<<grid, kThreads, smem, st>>>(a);
  return cudaSuccess;
}

// one translation unit per kernel size (parallel compilation); each returns cudaErrorNotSupported for a
// combination it does not instantiate
cudaError_t launch_conv_k1(const KArgs& a, int amode, int grid, int smem, cudaStream_t st);
cudaError_t launch_conv_k3(const KArgs& a, int amode, int grid, int smem, cudaStream_t st);
cudaError_t launch_conv_k4(const KArgs& a, int amode, int grid, int smem, cudaStream_t st);
int watchdog_conv_k1(int clear);
int watchdog_conv_k3(int clear);
int watchdog_conv_k4(int clear);

}  // namespace idiff
