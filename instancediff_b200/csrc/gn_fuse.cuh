// GroupNorm finalize folded into the PRODUCING kernel (no separate idiff_gn_finalize launch).
//
// The producer's epilogue warps add their (sum, sum of squares) of a warp tile as 64-bit FIXED-POINT integers into
// sums[image][slot][group][2] with atomic adds.  Integer addition is associative, so the totals do not depend on the
// order of the adds, on the launch grid or on how a batch is sharded over GPUs -- the property the fp32 partial-sum rows
// + fixed-order reduction of idiff_gn_finalize bought with 128 KB of partials per image and a kernel of its own.  The
// slots only spread the atomics of one image over IDIFF_GN_SLOTS addresses per entry.
// The LAST CTA of the grid to finish (arrival counter, the threadFenceReduction pattern) turns the totals into the
// per-(image, channel) affine the next layer applies on load
//     scale = rstd * gamma * (1 + ts),   shift = (beta - mean * rstd * gamma) * (1 + ts) + tb
// and leaves sums / arrivals zeroed for the next launch.  Everything is stream-ordered; a CUDA graph replays it as is.
//
// Fixed point: sums in units of 2^-20, sums of squares in units of 2^-16.  A warp tile's fp32 partial (256 - 1024
// values) carries 24 bits, so the quantisation (<= 5e-7 / 8e-6 per partial) is below its own rounding for all but
// near-zero partials; totals overflow only for a group mean square above ~2e8 (bf16 activations never get there).
// Serves the GroupNorm of `self.model(...)`, utils/sde_utils.py:198 (ResBlock / SpatialTransformer of SURVEY App. A).
#pragma once
#include "common.cuh"
#include "host_common.h"

namespace idiff {

constexpr int kGnSlots = IDIFF_GN_SLOTS;
constexpr float kGnUnit1 = 1048576.f;   // 2^20
constexpr float kGnUnit2 = 65536.f;     // 2^16

struct GnFuse {                          // device view of idiff_gn_fuse (+ sizes)
  unsigned long long* sums;              // [B][kGnSlots][G][2]
  unsigned int* arrivals;
  const float *gamma, *beta, *t_scale, *t_shift;
  float *scale_out, *shift_out;
  int t_ld, count, G, C, B;
  float eps;
};

// host: fill the device view from the public struct (nullptr f -> disabled); returns false on a bad description
inline bool gn_fuse_make(GnFuse& d, const idiff_gn_fuse* f, int B, int C, int G) {
  d = GnFuse{};
  if (!f) return true;
  if ((C != 64 && C != 128 && C != 256) || G <= 0 || C % G) return false;
  if (!f->sums || !f->arrivals || !f->gamma || !f->beta || !f->scale_out || !f->shift_out || f->count_per_group <= 0 ||
      (f->t_scale == nullptr) != (f->t_shift == nullptr) || (reinterpret_cast<uintptr_t>(f->sums) & 15u))
    return false;
  d.sums = reinterpret_cast<unsigned long long*>(f->sums);
  d.arrivals = f->arrivals;
  d.gamma = f->gamma; d.beta = f->beta; d.t_scale = f->t_scale; d.t_shift = f->t_shift;
  d.scale_out = f->scale_out; d.shift_out = f->shift_out;
  d.t_ld = f->t_ld; d.count = f->count_per_group; d.G = G; d.C = C; d.B = B; d.eps = f->eps;
  return true;
}

// one statistic of one warp tile: entry = group * 2 + st (st 0: sum, 1: sum of squares)
IDIFF_DEVINL void gn_fuse_add(const GnFuse& f, int b, int slot, int entry, float v) {
  const long long q = __float2ll_rn(v * ((entry & 1) ? kGnUnit2 : kGnUnit1));
  atomicAdd(f.sums + ((size_t)b * kGnSlots + (slot & (kGnSlots - 1))) * (2 * f.G) + entry, (unsigned long long)q);
}

IDIFF_DEVINL void gn_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Called by the NT threads that issued gn_fuse_add (all of them, converged, tid in [0, NT)) after their last add.
// stat: shared-memory scratch for `stat_cap` (mean, rstd) pairs; flag: one shared-memory word; bar_id: a named barrier
// these NT threads may use.  C must divide NT (64 / 128 / 256 channels, 256 threads): a thread owns ONE channel, so
// gamma / beta (and the time rows when the batch shares them) are loaded once -- before the arrival, where their
// latency hides behind the fence and the atomic -- and the per-image loop is arithmetic and two stores.
template <int NT>
IDIFF_DEVINL void gn_fuse_finish(const GnFuse& f, float2* stat, volatile int* flag, int tid, int bar_id, int stat_cap) {
  const int c = tid % f.C, b0 = tid / f.C, bstep = NT / f.C, cpg = f.C / f.G, g = c / cpg;
  const float gam = __ldg(f.gamma + c), bet = __ldg(f.beta + c);
  const bool timed = f.t_scale != nullptr, shared_t = f.t_ld == 0;
  float ts = 1.f, tb = 0.f;
  if (timed && shared_t) { ts = 1.f + f.t_scale[c]; tb = f.t_shift[c]; }
  gn_named_bar(bar_id, NT);                                    // every thread's adds are ordered before thread 0's fence
  if (tid == 0) {
    __threadfence();                                           // cumulative: publishes the whole CTA's adds before the arrival
    *flag = atomicAdd(f.arrivals, 1u) == gridDim.x * gridDim.y - 1u ? 1 : 0;
  }
  gn_named_bar(bar_id, NT);
  if (!*flag) return;
  __threadfence();
  const double u1 = 1.0 / ((double)kGnUnit1 * (double)f.count), u2 = 1.0 / ((double)kGnUnit2 * (double)f.count);
  const int img_cap = stat_cap / f.G;                          // images per pass (all of them for every plan shape)
  for (int base = 0; base < f.B; base += img_cap) {
    const int nimg = min(img_cap, f.B - base), n = nimg * f.G;
    for (int i = tid; i < n; i += NT) {
      const int bi = i / f.G, gi = i - bi * f.G;
      ulonglong2* p = reinterpret_cast<ulonglong2*>(f.sums + ((size_t)(base + bi) * kGnSlots * f.G + gi) * 2);
      long long s1 = 0, s2 = 0;
#pragma unroll
      for (int s = 0; s < kGnSlots; ++s) {
        const ulonglong2 v = __ldcg(p + (size_t)s * f.G);      // the adds were performed in L2
        s1 += (long long)v.x;
        s2 += (long long)v.y;
      }
#pragma unroll
      for (int s = 0; s < kGnSlots; ++s) p[(size_t)s * f.G] = make_ulonglong2(0ull, 0ull);   // ready for the next launch
      const double mean = (double)s1 * u1;
      const double var = fmax((double)s2 * u2 - mean * mean, 0.0);
      stat[i] = make_float2((float)mean, rsqrtf((float)var + f.eps));
    }
    gn_named_bar(bar_id, NT);
#pragma unroll 4
    for (int bi = b0; bi < nimg; bi += bstep) {
      const int b = base + bi;
      const float2 st = stat[bi * f.G + g];
      float sc = st.y * gam;
      float sh = bet - st.x * sc;
      if (timed) {
        if (!shared_t) { ts = 1.f + f.t_scale[(size_t)b * f.t_ld + c]; tb = f.t_shift[(size_t)b * f.t_ld + c]; }
        sc *= ts;
        sh = fmaf(sh, ts, tb);
      }
      f.scale_out[(size_t)b * f.C + c] = sc;
      f.shift_out[(size_t)b * f.C + c] = sh;
    }
    gn_named_bar(bar_id, NT);
  }
  if (tid == 0) *f.arrivals = 0u;
}

}  // namespace idiff
