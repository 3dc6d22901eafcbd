// Implicit-GEMM convolution / linear engine on tcgen05 tensor cores (sm_100a).
//
// Persistent kernel, one CTA per SM, work item = one 16x8 output-pixel tile (M = 128 rows) x NT channels.
// 18 warps, specialised and pipelined ACROSS items through mbarrier rings and a double-buffered TMEM
// accumulator (2 x NT fp32 columns):
//   warps 0-7  : epilogue.  TMEM -> registers (tcgen05.ld 32x32b.x32; warp w owns TMEM lanes 32*(w%4).., the
//                two warps of a lane quarter split the 32-column chunks) -> bias / LayerNorm fold / GroupNorm
//                partial sums / q-softmax / GEGLU / output LayerNorm / residual adds -> bf16 stores.
//   warps 8-15 : A producers.  Per 64-channel chunk they load the input patch (tile + halo) ONCE from global
//                memory (branch-free batches of 16 B loads, optional concat of two sources, optional nearest
//                x2 upsampling, optional per-(image,channel) affine + SiLU = fused GroupNorm / time modulation /
//                activation of the previous layer) and store it as [8-channel group][patch pixel][8 ch] -- the
//                UMMA SWIZZLE_NONE K-major canonical layout with a 16 B row pitch.  Every filter tap is then a
//                descriptor with a shifted start address, so the patch is re-used k*k times from shared memory.
//   warp 16    : weight producer.  Weights stay RESIDENT in shared memory for the whole CTA when they fit
//                (one bulk-TMA burst), otherwise pre-packed (chunk, tap) stages stream through a ring.
//   warp 17    : TMEM owner + MMA issuer.  Runs warp-uniformly; only the tcgen05.mma / commit instructions are
//                predicated on one elected lane, so descriptors stay in uniform registers and consecutive MMAs
//                differ by immediates (tap offsets are compile-time: the kernel is templated on NT and KS).
//
// Network spec: SURVEY.md App. A (the reference's models/modules/* is not in the snapshot); the call this serves
// is `noise = self.model(x, self.mu, t*scale, **kwargs)`, utils/sde_utils.py:198.
#include "common.cuh"
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
constexpr int kHeader = 12288;                     // barriers, GroupNorm scratch, row-statistics exchange
constexpr int kOffRed = 256, kOffXch = 1536, kOffImg = 5632;   // red 512 B; xch 4 KB; per-image params 2 x 3 x 256 floats = 6 KB
constexpr int kHeaderEnd = kOffImg + 2 * 3 * 256 * 4;
static_assert(kHeaderEnd <= kHeader, "header overflow");

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

struct SmemPlan {
  int SA, SB, resident, stageA, stageB, offA, offB, offP, total;
};

__host__ inline int stage_a_bytes(int ks) {
  return ks == 1 ? GeomT<1>::STAGE : ks == 3 ? GeomT<3>::STAGE : GeomT<4>::STAGE;
}

// Weights stay RESIDENT for the whole (persistent) CTA when all of them fit next to at least one A stage;
// otherwise they stream through a ring of SB stages.
__host__ inline SmemPlan plan_smem(const idiff_gemm_params& p) {
  SmemPlan s;
  const int nchunks = (p.cin0 + p.cin1) / 64;
  const int nk = nchunks * p.ksize * p.ksize;
  s.stageA = stage_a_bytes(p.ksize);
  s.stageB = p.NT * 128;
  const int pbytes = ((3 * p.N * 4 + 127) / 128) * 128;                // bias | wsum | ln_g cache
  const int budget = kSmemLimit - kHeader - pbytes;
  const long wbytes = (long)nk * s.stageB * (p.N / p.NT);           // every N tile
  s.resident = (p.w_image_stride == 0 && wbytes + s.stageA <= budget) ? 1 : 0;
  s.offA = kHeader;
  if (s.resident) {
    int sa = (int)((budget - wbytes) / s.stageA);
    s.SA = sa > kMaxSA ? kMaxSA : sa;
    s.SB = 0;
    s.offB = s.offA + s.SA * s.stageA;
    s.offP = s.offB + (int)wbytes;
    s.total = s.offP + pbytes;
  } else {
    s.SA = 2;
    if (2 * s.stageA + 2 * s.stageB > budget) s.SA = 1;
    int sb = (budget - s.SA * s.stageA) / s.stageB;
    if (sb > kMaxSB) sb = kMaxSB;
    if (sb > nk) sb = nk;
    if (sb < 1) sb = 1;
    s.SB = sb;
    if (s.SA == 2 && s.SA * s.stageA + s.SB * s.stageB + s.stageA <= budget) s.SA = 3;   // spare room: 3rd A stage
    s.offB = s.offA + s.SA * s.stageA;
    s.offP = s.offB + s.SB * s.stageB;
    s.total = s.offP + pbytes;
  }
  return s;
}

int watchdog_conv(int clear) { return watchdog_read_tu(clear); }

struct KArgs {
  idiff_gemm_params p;
  int SA, SB, resident, offA, offB, offP;
  int tiles_x, tiles_y, ntiles_n, total_items;
};

// Per-role cycle counters of CTA 0 (enabled with params.reserved0 = 1; read by idiff_debug_read_prof).
//  0 kernel cycles   1 items of CTA 0
//  2 loader: wait emptyA   3 loader: issue loads   4 loader: wait data + transform + store + arrive
//  5 mma: wait tmem_empty  6 mma: wait fullA       7 mma: wait fullB   8 mma: issue
//  9 epi: wait tmem_full  10 epi: work
__device__ unsigned long long g_prof[16];
#ifdef IDIFF_PROF
#define PROF_T() (prof ? clock64() : 0ll)
#define PROF_ADD(i, t0) do { if (prof) { pacc##i += clock64() - (t0); } } while (0)
#else                                   // production build: the hooks compile to nothing
#define PROF_T() 0ll
#define PROF_ADD(i, t0) do { (void)(t0); } while (0)
#endif

IDIFF_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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

IDIFF_DEVINL void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int NT, int KS>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ KArgs a) {
  using G = GeomT<KS>;
  extern __shared__ __align__(128) uint8_t smem[];
  const idiff_gemm_params& p = a.p;

  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);
  uint64_t* emptyA = fullA + kMaxSA;
  uint64_t* fullB = emptyA + kMaxSA;
  uint64_t* emptyB = fullB + kMaxSB;
  uint64_t* tmem_full = emptyB + kMaxSB;           // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* wres_bar = tmem_empty + 2;             // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);
  float* red = reinterpret_cast<float*>(smem + kOffRed);     // [2 items][8 warps][4 local groups][2]
  float2* xch = reinterpret_cast<float2*>(smem + kOffXch);   // [2 kinds][2 halves][128 rows]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef IDIFF_PROF
  const bool prof = p.reserved0 != 0 && blockIdx.x == 0;
#else
  constexpr bool prof = false;
#endif
  const long long t_kernel = prof ? clock64() : 0ll;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int cin = p.cin0 + p.cin1, nchunks = cin >> 6;
  constexpr int ntaps = G::NTAPS;
  const int nk = nchunks * ntaps;
  constexpr uint32_t tmem_cols = 2 * NT;                          // double-buffered accumulator
  constexpr int stageB = NT * 128;

  if (tid == 0) {
    for (int i = 0; i < a.SA; ++i) { mbar_init(&fullA[i], kLoaderThreads); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < a.SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    mbar_init(wres_bar, 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register re-balancing (per warpgroup).  The CTA owns 640 x 96 = 61440 registers (launch bounds); `inc` can
  // only draw on what `dec` released inside the CTA, so the targets must satisfy
  //   256*EPI + 256*LOADER + 128*OTHER <= 61440   ->   128 / 88 / 40  (= 60416).
  static_assert(256 * 128 + 256 * 88 + 128 * 40 <= kThreads * 96, "setmaxnreg budget exceeds the CTA's registers");
  if (warp >= kWarpB) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  else if (warp >= kEpiWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");

  if (warp < kEpiWarps) {
    // ============================== epilogue (TMEM -> registers -> global) ====================
    const int quarter = warp & 3, half = warp >> 2;               // TMEM lane quarter / which chunks
    const int r = quarter * 32 + lane, ti = r >> 3, tj = r & 7;   // accumulator row = tile pixel
    constexpr int NC = NT / 32, NCH = NC / 2;                     // chunks per row / per thread
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* res0 = reinterpret_cast<const __nv_bfloat16*>(p.res0);
    const __nv_bfloat16* res1 = reinterpret_cast<const __nv_bfloat16*>(p.res1);
    const bool gn = p.gn_groups > 0;                              // requires NT == N: exactly 8 groups per tile
    const float invN = 1.f / (float)p.N;
    // per-layer column parameters staged once per CTA: [bias | wsum | ln_g], N floats each
    float* pcache = reinterpret_cast<float*>(smem + a.offP);
    for (int i = tid; i < p.N; i += kEpiThreads) {
      pcache[i] = p.bias ? __ldg(p.bias + i) : 0.f;
      pcache[p.N + i] = p.wsum ? __ldg(p.wsum + i) : 0.f;
      pcache[2 * p.N + i] = p.ln_g ? __ldg(p.ln_g + i) : 1.f;
    }
    const bool img_params = p.bias_img != nullptr || p.res0_scale != nullptr;
    float* pimg_all = reinterpret_cast<float*>(smem + kOffImg);   // [2 items][bias_img | res0_scale | res0_shift][256]
    epi_bar();

    long long pacc9 = 0, pacc10 = 0;
    int it_local = 0;
    ItemIter it;
    for (it.init(a, blockIdx.x, gridDim.x); it.valid(); it.next(), ++it_local) {
      const int b = it.b, n0 = it.nt * NT;
      const int ab = it_local & 1;
      const int oy = it.oy0() + ti, ox = it.ox0() + tj;
      const bool valid = (oy < p.H) && (ox < p.W);
      const size_t m = ((size_t)b * p.H + (valid ? oy : 0)) * p.W + (valid ? ox : 0);   // clamped: loads stay legal
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * NT);
      float* redi = red + ab * (kEpiWarps * 8);

      float mean_in = 0.f, rstd_in = 1.f;
      if (p.row_stats) {
        mean_in = __ldg(p.row_stats + 2 * m);
        rstd_in = __ldg(p.row_stats + 2 * m + 1);
      }
      // per-image column parameters of this item's N tile -> shared memory (before waiting for the accumulator)
      float* pimg = pimg_all + ab * (3 * 256);
      if (img_params) {
        if (tid < NT) {
          const size_t gi = (size_t)b * p.N + n0 + tid;
          pimg[tid] = p.bias_img ? __ldg(p.bias_img + gi) : 0.f;
          pimg[256 + tid] = p.res0_scale ? __ldg(p.res0_scale + gi) : 1.f;
          pimg[512 + tid] = p.res0_scale ? __ldg(p.res0_shift + gi) : 0.f;
        }
        epi_bar();
      }
      // accumulator chunk -> value after LayerNorm fold + biases (32 columns starting at global column n)
      auto apply_base = [&](float* v, int n) {
        if (p.row_stats) for_cols32(pcache + p.N + n, v, [&](float x, float c) { return (x - mean_in * c) * rstd_in; });
        if (p.bias) for_cols32(pcache + n, v, [](float x, float c) { return x + c; });
        if (p.bias_img) for_cols32(pimg + (n - n0), v, [](float x, float c) { return x + c; });
      };
      auto add_residuals = [&](float* v, int n) {
        if (res0) {
          float rr[32];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(res0 + m * p.N + n + q4 * 8)), rr + q4 * 8);
          if (p.res0_scale) {                                       // residual enters as silu(GN(y)) (ResBlock tail)
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
      };
      auto store32 = [&](const float* v, int col0) {      // 32 bf16 at out[m][col0..]
        if (!valid) return;
        uint4* dst = reinterpret_cast<uint4*>(outp + m * p.out_ld + col0);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) dst[q4] = pack_bf16x8(v + q4 * 8);
      };
      auto release_tmem = [&]() {                         // this warp is done reading the accumulator buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[ab]);
      };
      // sum (s1, s2) of the two column halves of a row through shared memory
      auto exchange = [&](int kind, float& s1, float& s2) {
        xch[(kind * 2 + half) * 128 + r] = make_float2(s1, s2);
        epi_bar();
        const float2 o = xch[(kind * 2 + (half ^ 1)) * 128 + r];
        s1 += o.x;
        s2 += o.y;
      };

      long long tp = PROF_T();
      mbar_wait(&tmem_full[ab], (it_local >> 1) & 1, 102);
      tc_fence_after();
      PROF_ADD(9, tp);
      tp = PROF_T();

      float o1 = 0.f, o2 = 0.f;
      bool exchanged = false;
      if (p.epi == IDIFF_EPI_LN_OUT) {
        // pass 1: LayerNorm statistics over the whole row (NT == N); each warp half sees half of the columns
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int cc = 2 * j + half;
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          apply_base(v, n0 + cc * 32);
#pragma unroll
          for (int q = 0; q < 32; ++q) { s1 += v[q]; s2 = fmaf(v[q], v[q], s2); }
        }
        exchange(0, s1, s2);
        exchanged = true;
        const float mean = s1 * invN, var = fmaxf(s2 * invN - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int cc = 2 * j + half;
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          if (j == NCH - 1) release_tmem();
          apply_base(v, n0 + cc * 32);
          for_cols32(pcache + 2 * p.N + n0 + cc * 32, v, [&](float x, float c) { return (x - mean) * rstd * c; });
          add_residuals(v, n0 + cc * 32);
          if (p.out_row_stats) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 = fmaf(v[q], v[q], o2); }
          }
          store32(v, n0 + cc * 32);
        }
      } else {
        float gl[8];                                      // GroupNorm: [local group 0..3][sum, sum of squares]
#pragma unroll
        for (int i = 0; i < 8; ++i) gl[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int cc = 2 * j + half;
          float v[32];
          tmem_ld32(taddr + cc * 32, v);
          if (j == NCH - 1) release_tmem();               // last TMEM read of this item: release the buffer early
          const int ncol0 = n0 + cc * 32;
          apply_base(v, ncol0);

          if (gn && valid) {
            // partial sums of the conv output (bias included).  Local group index of the 8-column block g8 of
            // this thread's chunk j: NT=64 -> g8 ; NT=128 -> 2j + g8/2 ; NT=256 -> j   (compile time)
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              constexpr int CPG = NT / 8;
              const int lg = CPG == 8 ? g8 : CPG == 16 ? 2 * j + (g8 >> 1) : j;
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) { const float x = v[g8 * 8 + q]; s1 += x; s2 = fmaf(x, x, s2); }
              gl[2 * lg] += s1;
              gl[2 * lg + 1] += s2;
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
            add_residuals(v, ncol0);
            if (p.out_row_stats) {
#pragma unroll
              for (int q = 0; q < 32; ++q) { o1 += v[q]; o2 = fmaf(v[q], v[q], o2); }
            }
            store32(v, ncol0);
          }
        }
        if (gn) {
          const float tot = warp_reduce8(gl, lane);       // lane L: warp total of local value index L >> 2
          if ((lane & 3) == 0) redi[warp * 8 + (lane >> 2)] = tot;
          epi_bar();
          if (tid < 16) {
            // global group Gi, statistic st -> (column half h, local group l) that accumulated it
            constexpr int CPG = NT / 8;
            const int Gi = tid >> 1, st = tid & 1;
            const int h = CPG == 8 ? (Gi >> 2) : CPG == 16 ? ((Gi >> 1) & 1) : (Gi & 1);
            const int l = CPG == 8 ? (Gi & 3) : CPG == 16 ? (2 * (Gi >> 2) + (Gi & 1)) : (Gi >> 1);
            float sum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) sum += redi[(h * 4 + q) * 8 + l * 2 + st];
            p.gn_partial[(((size_t)b * tiles_per_img + it.tile_in_img()) * 8 + Gi) * 2 + st] = sum;
          }
        }
      }
      if (p.out_row_stats) {
        exchange(1, o1, o2);
        exchanged = true;
        if (valid && half == 0) {
          const float mo = o1 * invN, vo = fmaxf(o2 * invN - mo * mo, 0.f);
          p.out_row_stats[2 * m] = mo;
          p.out_row_stats[2 * m + 1] = rsqrtf(vo + p.ln_eps);
        }
      }
      if (exchanged) epi_bar();                            // exchange buffers are re-used by the next item
      PROF_ADD(10, tp);
    }
    if (prof && tid == 0) { g_prof[9] = pacc9; g_prof[10] = pacc10; g_prof[1] = it_local; }
  } else if (warp < kWarpB) {
    // ============================== A producers ==============================================
    const int ltid = tid - kEpiThreads;
    const int c8 = ltid & 7, prow = ltid >> 3;                        // 32 patch pixels per sweep
    constexpr int SWEEP = kLoaderThreads / 8;
    constexpr int BATCH = 6;                                          // 16 B loads in flight per thread
    constexpr bool kSmall = G::NSLOT <= SWEEP * BATCH;                // k = 1, 3: one batch covers the patch
    const int Hin = p.H * G::S, Win = p.W * G::S;                     // virtual (possibly upsampled) input extent
    const int Hs = p.up0 ? (Hin >> 1) : Hin, Ws = p.up0 ? (Win >> 1) : Win;
    const bool affine = p.a_scale != nullptr;
    long long pacc2 = 0, pacc3 = 0, pacc4 = 0;
    // patch coordinates of this thread's pixels: identical for every tile and chunk
    int pv[BATCH], pu[BATCH], pslot[BATCH];
#pragma unroll
    for (int i = 0; i < BATCH; ++i) {
      const int px = prow + SWEEP * i;
      const bool in_patch = px < G::NSLOT;
      const int pxc = in_patch ? px : 0;
      pv[i] = pxc / G::PC;
      pu[i] = pxc - pv[i] * G::PC;
      pslot[i] = in_patch ? G::slot(pv[i], pu[i]) : -1;
    }

    // per-stage source description (stage = one 64-channel chunk of one item)
    struct Src {
      const __nv_bfloat16* base;     // image b, channel offset applied
      const __nv_bfloat16* origin;   // base + patch origin (iy0, ix0) -- only meaningful for interior tiles
      int Cs, iy0, ix0, b, ch;
      bool interior;                 // whole patch inside the image and no upsampling: pre-computed offsets apply
    };
    auto describe = [&](const ItemIter& it, int ch) {
      Src sd;
      const bool from0 = (ch << 6) < p.cin0;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(from0 ? p.src0 : p.src1);
      sd.Cs = from0 ? (p.src0_ld ? p.src0_ld : p.cin0) : (p.src1_ld ? p.src1_ld : p.cin1);
      const int coff = (from0 ? (ch << 6) : ((ch << 6) - p.cin0)) + c8 * 8;
      sd.base = src + (size_t)it.b * Hs * Ws * sd.Cs + coff;
      sd.iy0 = it.oy0() * G::S - G::PAD;
      sd.ix0 = it.ox0() * G::S - G::PAD;
      sd.interior = !p.up0 && sd.iy0 >= 0 && sd.ix0 >= 0 && sd.iy0 + G::PR <= Hin && sd.ix0 + G::PC <= Win;
      sd.origin = sd.base + ((long long)sd.iy0 * Ws + sd.ix0) * sd.Cs;
      sd.b = it.b;
      sd.ch = ch;
      return sd;
    };
    // branch-free batch: every load is issued from an always-valid address; returns the in-image mask
    const uint32_t patch_mask = [&] {
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < BATCH; ++i) m |= pslot[i] >= 0 ? (1u << i) : 0u;
      return m;
    }();
    auto issue = [&](const Src& sd, uint4* q) -> uint32_t {
      if (sd.interior) {
        // (pv*Ws + pu)*Cs relative to the patch origin: one 32-bit multiply-add per vector, no bounds tests
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int rel = (pslot[i] >= 0) ? (pv[i] * Ws + pu[i]) * sd.Cs : 0;
          q[i] = __ldg(reinterpret_cast<const uint4*>(sd.origin + rel));
        }
        return patch_mask;
      }
      uint32_t mask = 0;
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int iy = sd.iy0 + pv[i], ix = sd.ix0 + pu[i];
        const bool ok = pslot[i] >= 0 && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
        const int sy = p.up0 ? (iy >> 1) : iy, sx = p.up0 ? (ix >> 1) : ix;
        const size_t off = ok ? ((size_t)sy * Ws + sx) * sd.Cs : 0;
        mask |= ok ? (1u << i) : 0u;
        q[i] = __ldg(reinterpret_cast<const uint4*>(sd.base + off));
      }
      return mask;
    };
    // affine with the 0.5 of silu(y) = h*tanh(h) + h, h = y/2, folded in when SiLU follows
    const float half_if_silu = p.a_silu ? 0.5f : 1.0f;
    auto load_affine = [&](const Src& sd, float* sc, float* sh) {
      const float* ps = p.a_scale + (size_t)sd.b * cin + (sd.ch << 6) + c8 * 8;
      const float* pt = p.a_shift + (size_t)sd.b * cin + (sd.ch << 6) + c8 * 8;
      const float4 s0 = ldg4(ps), s1 = ldg4(ps + 4), t0 = ldg4(pt), t1 = ldg4(pt + 4);
      sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
      sh[0] = t0.x; sh[1] = t0.y; sh[2] = t0.z; sh[3] = t0.w; sh[4] = t1.x; sh[5] = t1.y; sh[6] = t1.z; sh[7] = t1.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] *= half_if_silu; sh[e] *= half_if_silu; }
    };
    auto transform_store = [&](uint8_t* stage, const uint4& qv, bool ok, int slot, const float* sc, const float* sh) {
      if (slot < 0) return;
      uint4 o = ok ? qv : make_uint4(0u, 0u, 0u, 0u);
      if (affine && ok) {
        float f[8];
        unpack_bf16x8(qv, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float h = fmaf(f[e], sc[e], sh[e]);
          f[e] = p.a_silu ? fmaf(h, tanh_fast(h), h) : h;
        }
        o = pack_bf16x8(f);
      }
      *reinterpret_cast<uint4*>(stage + slot * 16) = o;
    };

    Ring ra;
    ra.init(a.SA);
    bool first_lap = true;
    ItemIter it;
    it.init(a, blockIdx.x, gridDim.x);
    if (kSmall) {
      // Software pipeline: the loads of stage s+1 are in flight while stage s is transformed and stored.
      int ch = 0;
      bool have = it.valid();
      uint4 qn[BATCH];
      uint32_t mask_n = 0;
      Src cur;
      if (have) {
        cur = describe(it, ch);
        mask_n = issue(cur, qn);
      }
      while (have) {
        uint4 qc[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) qc[i] = qn[i];
        const uint32_t mask_c = mask_n;
        const Src now = cur;
        // advance to the next stage and put its loads in flight
        if (++ch == nchunks) { ch = 0; it.next(); }
        have = it.valid();
        long long tp = PROF_T();
        if (have) {
          cur = describe(it, ch);
          mask_n = issue(cur, qn);
        }
        PROF_ADD(3, tp);
        tp = PROF_T();
        if (!first_lap) mbar_wait(&emptyA[ra.slot], ra.phase ^ 1, 101);
        PROF_ADD(2, tp);
        tp = PROF_T();
        uint8_t* stage = smem + a.offA + ra.slot * G::STAGE + c8 * G::LBO;
        float sc[8], sh[8];
        if (affine) load_affine(now, sc, sh);
        if (mask_c == patch_mask) {                      // whole patch inside the image: no zero-fill selects
#pragma unroll
          for (int i = 0; i < BATCH; ++i) transform_store(stage, qc[i], true, pslot[i], sc, sh);
        } else {
#pragma unroll
          for (int i = 0; i < BATCH; ++i) transform_store(stage, qc[i], (mask_c >> i) & 1u, pslot[i], sc, sh);
        }
        fence_proxy_async_smem();
        mbar_arrive(&fullA[ra.slot]);
        ra.advance();
        if (ra.slot == 0) first_lap = false;
        PROF_ADD(4, tp);
      }
    } else {
      for (; it.valid(); it.next()) {
        for (int ch = 0; ch < nchunks; ++ch) {
          const Src sd = describe(it, ch);
          long long tp = PROF_T();
          if (!first_lap) mbar_wait(&emptyA[ra.slot], ra.phase ^ 1, 101);
          PROF_ADD(2, tp);
          tp = PROF_T();
          uint8_t* stage = smem + a.offA + ra.slot * G::STAGE + c8 * G::LBO;
          float sc[8], sh[8];
          if (affine) load_affine(sd, sc, sh);
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
              const size_t off = ok ? ((size_t)sy * Ws + sx) * sd.Cs : 0;
              inb[i] = ok;
              q[i] = __ldg(reinterpret_cast<const uint4*>(sd.base + off));
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) transform_store(stage, q[i], inb[i], slot[i], sc, sh);
          }
          fence_proxy_async_smem();
          mbar_arrive(&fullA[ra.slot]);
          ra.advance();
          if (ra.slot == 0) first_lap = false;
          PROF_ADD(4, tp);
        }
      }
    }
    if (prof && ltid == 0) { g_prof[2] = pacc2; g_prof[3] = pacc3; g_prof[4] = pacc4; }
  } else if (warp == kWarpB) {
    // ============================== weight producer (bulk TMA, warp-uniform) ==================
    const bool leader = elect_one();
    const uint8_t* w0 = reinterpret_cast<const uint8_t*>(p.w);
    if (a.resident) {
      // every N tile's weights, once per CTA: [ntile][nk] stages of stageB bytes, contiguous in global memory
      const uint32_t total = (uint32_t)a.ntiles_n * (uint32_t)nk * (uint32_t)stageB;
      if (leader) {
        mbar_arrive_expect_tx(wres_bar, total);
        for (uint32_t off = 0; off < total; off += 32768) {
          const uint32_t n = total - off < 32768u ? total - off : 32768u;
          bulk_g2s(smem + a.offB + off, w0 + off, n, wres_bar);
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
    if (a.resident) {
      mbar_wait(wres_bar, 0, 106);
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
      for (int ch = 0; ch < nchunks; ++ch) {
        tp = PROF_T();
        mbar_wait(&fullA[ra.slot], ra.phase, 104);
        tc_fence_after();
        PROF_ADD(6, tp);
        const uint32_t a_lo0 = umma_desc_lo(a0 + ra.slot * G::STAGE, G::LBO);
#pragma unroll
        for (int tap = 0; tap < ntaps; ++tap) {
          uint32_t b_lo0;
          if (a.resident) {
            b_lo0 = b_res + (uint32_t)((tap * stageB) >> 4);
          } else {
            tp = PROF_T();
            mbar_wait(&fullB[rb.slot], rb.phase, 105);
            tc_fence_after();
            PROF_ADD(7, tp);
            b_lo0 = umma_desc_lo(b0 + rb.slot * stageB, lboB);
          }
          tp = PROF_T();
          const int tapslot = G::slot(tap / KS, tap % KS);                       // compile time
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_lo = a_lo0 + (uint32_t)((tapslot * 16 + kk * 2 * G::LBO) >> 4);
              const uint32_t b_lo = b_lo0 + (uint32_t)((kk * 2 * lboB) >> 4);
              umma_bf16_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, (ch | tap | kk) != 0 ? 1u : 0u);
            }
            if (!a.resident) umma_commit(&emptyB[rb.slot]);
          }
          if (!a.resident) rb.advance();
          PROF_ADD(8, tp);
        }
        if (leader) umma_commit(&emptyA[ra.slot]);
        ra.advance();
        b_res += (uint32_t)((ntaps * stageB) >> 4);
      }
      if (leader) umma_commit(&tmem_full[ab]);
      __syncwarp();
      ++items_done;
      ab ^= 1;
      if (ab == 0) acc_phase ^= 1;
      nt += nt_step;
      if (nt >= a.ntiles_n) nt -= a.ntiles_n;
    }
    if (prof && leader) { g_prof[5] = pacc5; g_prof[6] = pacc6; g_prof[7] = pacc7; g_prof[8] = pacc8; }
  }

  tc_fence_before();
  __syncthreads();
  if (prof && tid == 0) g_prof[0] = clock64() - t_kernel;
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
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
    IDIFF_REQUIRE(p.gn_partial && p.gn_groups == 8 && p.NT == p.N,
                  "conv_gemm: fused GroupNorm partials need 8 groups and NT == N");
    IDIFF_REQUIRE(p.epi == IDIFF_EPI_PLAIN, "conv_gemm: GroupNorm partials need the plain epilogue");
  }
  if (p.res0_scale) IDIFF_REQUIRE(p.res0 && p.res0_shift, "conv_gemm: res0 affine needs res0 and shift");
  return IDIFF_OK;
}

}  // namespace idiff

extern "C" {

int idiff_debug_read_prof(unsigned long long* out16_host) {
  using namespace idiff;
  IDIFF_REQUIRE(out16_host, "debug_read_prof: null");
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out16_host, g_prof, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "debug_read_prof: %s", cudaGetErrorString(e));
  return IDIFF_OK;
}

int idiff_conv_gemm_smem_bytes(const idiff_gemm_params* p) {
  using namespace idiff;
  IDIFF_REQUIRE(p, "conv_gemm: null params");
  int rc = validate(*p);
  if (rc) return rc;
  return plan_smem(*p).total;
}

}  // extern "C"

template <int NT, int KS>
static cudaError_t launch_one(const idiff::KArgs& a, int grid, int smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(idiff::conv_gemm_kernel<NT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         idiff::kSmemLimit);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  idiff::conv_gemm_kernel<NT, KS><<<grid, idiff::kThreads, smem, st>>>(a);
  return cudaSuccess;
}
template <int NT>
static cudaError_t launch_nt(const idiff::KArgs& a, int grid, int smem, cudaStream_t st) {
  switch (a.p.ksize) {
    case 1: return launch_one<NT, 1>(a, grid, smem, st);
    case 3: return launch_one<NT, 3>(a, grid, smem, st);
    default: return launch_one<NT, 4>(a, grid, smem, st);
  }
}

extern "C" {

int idiff_conv_gemm(const idiff_gemm_params* pp, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(pp, "conv_gemm: null params");
  int rc = validate(*pp);
  if (rc) return rc;
  KArgs a;
  a.p = *pp;
  const SmemPlan s = plan_smem(*pp);
  a.SA = s.SA; a.SB = s.SB; a.resident = s.resident; a.offA = s.offA; a.offB = s.offB; a.offP = s.offP;
  a.tiles_x = (pp->W + TILE_W - 1) / TILE_W;
  a.tiles_y = (pp->H + TILE_H - 1) / TILE_H;
  a.ntiles_n = pp->N / pp->NT;
  a.total_items = a.tiles_x * a.tiles_y * pp->B * a.ntiles_n;
  IDIFF_REQUIRE(s.total <= kSmemLimit && s.SA >= 1, "conv_gemm: shared memory plan %d B too large", s.total);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { num_sms = 0; return fail(IDIFF_ERR_CUDA, "conv_gemm setup: %s", cudaGetErrorString(e)); }
  }
  // persistent: one CTA per SM (TMEM: 2*NT columns each), items strided by the grid size
  const int grid = a.total_items < num_sms ? a.total_items : num_sms;
  cudaError_t e = pp->NT == 64 ? launch_nt<64>(a, grid, s.total, as_stream(stream))
                  : pp->NT == 128 ? launch_nt<128>(a, grid, s.total, as_stream(stream))
                                  : launch_nt<256>(a, grid, s.total, as_stream(stream));
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv_gemm attr: %s", cudaGetErrorString(e));
  return check_launch("conv_gemm");
}

}  // extern "C"
