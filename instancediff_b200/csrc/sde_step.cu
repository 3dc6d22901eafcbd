// Fused reverse-SDE Euler-Maruyama update (HBM-bound, one pass).
//
// Replaces utils/sde_utils.py:187-188 (score from noise), :178-179 (reverse drift), :184-185
// (dispersion) and :45-46 (x - drift - dispersion): 12 ATen launches / ~108 B per element in the
// reference, 20 B per element here with pre-drawn z (read x, eps, mu, z; write x) and 16 B with
// the in-kernel Philox draw.  Arithmetic keeps the reference's fp32 operation ORDER with explicit
// round-to-nearest intrinsics (no FMA contraction) so the result matches the CPU oracle bit for bit.
#include "common.cuh"
#include "host_common.h"

namespace idiff {

// ---- Philox4x32-10 ----------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  IDIFF_DEVINL static uint4 gen(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
      const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
      ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
      key.x += W0;
      key.y += W1;
    }
    return ctr;
  }
};

// 4 standard normals for the 4-element group `grp` of stream (seed, step).
IDIFF_DEVINL float4 philox_normal4(uint64_t seed, uint64_t grp, uint32_t step) {
  const uint4 r = Philox::gen(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), step, 0u),
                              make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = fmaf((float)r.x, k, 0.5f * k), u1 = fmaf((float)r.y, k, 0.5f * k);
  const float u2 = fmaf((float)r.z, k, 0.5f * k), u3 = fmaf((float)r.w, k, 0.5f * k);
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

struct StepCoef {
  float theta, sigma, sbar, dt, sqrt_dt;
  uint32_t step;
};

// `flags`: bit 0 = `e` already holds the score; bit 1 = probability-flow ODE drift (:181-182) instead of :178-179
IDIFF_DEVINL float sde_update(float x, float e, float mu, float z, const StepCoef& c, int flags, bool has_noise) {
  // utils/sde_utils.py:188   score = -noise / sigma_bar
  const float score = (flags & 1) ? e : __fdiv_rn(-e, c.sbar);
  // :179   (theta*(mu - x) - sigma**2 * score) * dt        :182   ... - 0.5 * sigma**2 * score ...
  const float t1 = __fmul_rn(c.theta, __fsub_rn(mu, x));
  const float s2 = __fmul_rn(c.sigma, c.sigma);
  const float t2 = __fmul_rn((flags & 2) ? __fmul_rn(0.5f, s2) : s2, score);
  const float drift = __fmul_rn(__fsub_rn(t1, t2), c.dt);
  float out = __fsub_rn(x, drift);                       // :42 / :46 left-associated
  if (has_noise) {
    // :185   sigma * (randn * sqrt(dt))
    const float disp = __fmul_rn(c.sigma, __fmul_rn(z, c.sqrt_dt));
    out = __fsub_rn(out, disp);
  }
  return out;
}

template <bool kVec, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
sde_step_kernel(float* __restrict__ xo, const float* __restrict__ x, const float* __restrict__ eps,
                const float* __restrict__ mu, const float* __restrict__ z, const float* __restrict__ coef,
                int is_score, int philox, uint64_t seed, uint64_t elem_offset, const uint64_t* __restrict__ rng_dev,
                size_t n) {
  if (rng_dev) {                       // {seed, elem_offset} in device memory: one captured graph serves every item
    seed = __ldg(rng_dev);
    elem_offset = __ldg(rng_dev + 1);
  }
  StepCoef c;
  c.theta = __ldg(coef + 0);
  c.sigma = __ldg(coef + 1);
  c.sbar = __ldg(coef + 2);
  c.dt = __ldg(coef + 3);
  c.sqrt_dt = __ldg(coef + 4);
  c.step = __float_as_uint(__ldg(coef + 5));
  const bool noise = philox || z != nullptr;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (kVec) {
    // two independent float4 items per thread per sweep: all loads are issued before the arithmetic
    const size_t n4 = n >> 2;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
      const size_t i1 = i0 + stride;
      const bool two = i1 < n4;
      const size_t j1 = two ? i1 : i0;
      const float4 xa = __ldcs(reinterpret_cast<const float4*>(x) + i0), xb = __ldcs(reinterpret_cast<const float4*>(x) + j1);
      const float4 ea = __ldcs(reinterpret_cast<const float4*>(eps) + i0), eb = __ldcs(reinterpret_cast<const float4*>(eps) + j1);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 ma = mu ? __ldg(reinterpret_cast<const float4*>(mu) + i0) : zero;
      const float4 mb = mu ? __ldg(reinterpret_cast<const float4*>(mu) + j1) : zero;
      float4 za = zero, zb = zero;
      if (philox) {
        za = philox_normal4(seed, (elem_offset >> 2) + i0, c.step);
        zb = philox_normal4(seed, (elem_offset >> 2) + j1, c.step);
      } else if (z) {
        za = __ldcs(reinterpret_cast<const float4*>(z) + i0);
        zb = __ldcs(reinterpret_cast<const float4*>(z) + j1);
      }
      float4 o;
      o.x = sde_update(xa.x, ea.x, ma.x, za.x, c, is_score, noise);
      o.y = sde_update(xa.y, ea.y, ma.y, za.y, c, is_score, noise);
      o.z = sde_update(xa.z, ea.z, ma.z, za.z, c, is_score, noise);
      o.w = sde_update(xa.w, ea.w, ma.w, za.w, c, is_score, noise);
      reinterpret_cast<float4*>(xo)[i0] = o;
      if (two) {
        o.x = sde_update(xb.x, eb.x, mb.x, zb.x, c, is_score, noise);
        o.y = sde_update(xb.y, eb.y, mb.y, zb.y, c, is_score, noise);
        o.z = sde_update(xb.z, eb.z, mb.z, zb.z, c, is_score, noise);
        o.w = sde_update(xb.w, eb.w, mb.w, zb.w, c, is_score, noise);
        reinterpret_cast<float4*>(xo)[i1] = o;
      }
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float zv = 0.f;
      if (philox) {
        const uint64_t g = elem_offset + i;
        const float4 q = philox_normal4(seed, g >> 2, c.step);
        const int l = (int)(g & 3);
        zv = l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
      } else if (z) {
        zv = z[i];
      }
      xo[i] = sde_update(x[i], eps[i], mu ? mu[i] : 0.f, zv, c, is_score, noise);
    }
  }
}

// Forward-process training states (utils/sde_utils.py:169-170, 331-336): per sample b with its own timestep,
//   state_mean = mu + (x0 - mu) * decay[b]          decay[b] = exp(-thetas_cumsum[t_b] * dt)   (:170)
//   x_t        = z * sbar[b] + state_mean           sbar[b]  = sigma_bars[t_b]                 (:334-336)
// One pass instead of the reference's 5 elementwise launches; same fp32 operation order, no FMA contraction.
IDIFF_DEVINL float random_state(float x0, float mu, float z, float decay, float sbar) {
  const float mean = __fadd_rn(mu, __fmul_rn(__fsub_rn(x0, mu), decay));
  return __fadd_rn(__fmul_rn(z, sbar), mean);
}

template <bool kVec>
__global__ void __launch_bounds__(256)
random_states_kernel(float* __restrict__ xt, float* __restrict__ z_out, const float* __restrict__ x0,
                     const float* __restrict__ mu, const float* __restrict__ z, const float* __restrict__ decay,
                     const float* __restrict__ sbar, int philox, uint64_t seed, uint64_t elem_offset, uint32_t step,
                     size_t per_sample, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (kVec) {
    const size_t n4 = n >> 2, ps4 = per_sample >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const size_t b = i / ps4;
      const float d = __ldg(decay + b), sb = __ldg(sbar + b);
      const float4 a = __ldcs(reinterpret_cast<const float4*>(x0) + i);
      const float4 m = __ldcs(reinterpret_cast<const float4*>(mu) + i);
      const float4 q = philox ? philox_normal4(seed, (elem_offset >> 2) + i, step)
                              : __ldcs(reinterpret_cast<const float4*>(z) + i);
      float4 o;
      o.x = random_state(a.x, m.x, q.x, d, sb);
      o.y = random_state(a.y, m.y, q.y, d, sb);
      o.z = random_state(a.z, m.z, q.z, d, sb);
      o.w = random_state(a.w, m.w, q.w, d, sb);
      reinterpret_cast<float4*>(xt)[i] = o;
      if (z_out) reinterpret_cast<float4*>(z_out)[i] = q;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const size_t b = i / per_sample;
      float zv;
      if (philox) {
        const uint64_t g = elem_offset + i;
        const float4 q = philox_normal4(seed, g >> 2, step);
        const int l = (int)(g & 3);
        zv = l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
      } else {
        zv = z[i];
      }
      xt[i] = random_state(x0[i], mu[i], zv, __ldg(decay + b), __ldg(sbar + b));
      if (z_out) z_out[i] = zv;
    }
  }
}

__global__ void __launch_bounds__(256)
noise_state_kernel(float* __restrict__ xo, const float* __restrict__ mu, const float* __restrict__ z,
                   float max_sigma, int philox, uint64_t seed, uint64_t elem_offset, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float zv;
    if (philox) {
      const uint64_t g = elem_offset + i;
      const float4 q = philox_normal4(seed, g >> 2, 0xFFFFFFFFu);   // step word reserved for x_T
      const int l = (int)(g & 3);
      zv = l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
    } else {
      zv = z[i];
    }
    // utils/sde_utils.py:341   tensor + randn * max_sigma
    xo[i] = __fadd_rn(mu[i], __fmul_rn(zv, max_sigma));
  }
}

__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ out, uint64_t seed, uint64_t elem_offset, uint32_t step, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t g = elem_offset + i;
    const float4 q = philox_normal4(seed, g >> 2, step);
    const int l = (int)(g & 3);
    out[i] = l == 0 ? q.x : l == 1 ? q.y : l == 2 ? q.z : q.w;
  }
}

// Select the coefficient row for the step held in *t_counter, publish the model time, decrement.
// Lets one captured CUDA graph serve every step of utils/sde_utils.py:248.
__global__ void step_select_kernel(const float* __restrict__ table, int* __restrict__ t_counter,
                                   float* __restrict__ cur_row, float* __restrict__ cur_time, float sample_scale) {
  const int t = *t_counter;
  if (threadIdx.x < 8) cur_row[threadIdx.x] = table[(size_t)t * 8 + threadIdx.x];
  if (threadIdx.x == 8) *cur_time = (float)t * sample_scale;       // :198   t * scale
  __syncthreads();
  if (threadIdx.x == 0) *t_counter = t - 1;
}

// Same, plus the time conditioning of the step: row t of a [T+1][S] table (every ResBlock's scale / shift projections
// of the time embedding, which depend on t only and are built once per sampler) is copied to `ss_out`, so the two
// time-embedding launches leave the captured step.
__global__ void step_select_ss_kernel(const float* __restrict__ table, int* __restrict__ t_counter,
                                      float* __restrict__ cur_row, float* __restrict__ cur_time, float sample_scale,
                                      const float* __restrict__ ss_table, float* __restrict__ ss_out, int S) {
  const int t = *t_counter;
  if (threadIdx.x < 8) cur_row[threadIdx.x] = table[(size_t)t * 8 + threadIdx.x];
  if (threadIdx.x == 8) *cur_time = (float)t * sample_scale;
  const float* src = ss_table + (size_t)t * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) ss_out[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) *t_counter = t - 1;
}

static int grid_for(size_t work_items, int block) {
  size_t g = (work_items + block - 1) / block;
  const size_t cap = 148 * 8;                 // 8 resident 256-thread CTAs per SM, grid-stride beyond
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace idiff

extern "C" {

static int sde_step_launch(float* x_out, const float* x, const float* eps, const float* mu, const float* z,
                           const float* coef, int input_is_score, int use_philox, uint64_t seed, uint64_t elem_offset,
                           const uint64_t* rng_dev, bool offset_mult4, size_t n, void* stream) {
  using namespace idiff;
  if (n == 0) return IDIFF_OK;                       // empty batch: nothing to do (pointers may be null)
  IDIFF_REQUIRE(x_out && x && eps && coef, "sde_step: null pointer");
  const bool vec = (n % 4 == 0) && aligned16(x_out) && aligned16(x) && aligned16(eps) && (!mu || aligned16(mu)) &&
                   (!z || aligned16(z)) && offset_mult4;
  if (vec) {
    // ONE resident wave: 4 CTAs per SM (the launch bound caps the kernel at 64 registers), every thread walks the
    // tensor with the grid stride, two float4 items in flight per trip.  Measured inside a captured graph with inputs
    // coming from HBM (tools/sweep_sde.py rotates buffer sets past the L2 size; B200, 16 B per element):
    //   elements        one CTA per 512 items, <= 8 per SM (76 reg: 3 resident)    one wave of 4 CTAs per SM
    //   2.1M  (B = 32)        7.69 us  4.36 TB/s                                    7.19 us  4.66 TB/s
    //   4.2M  (B = 64)       13.93 us  4.82 TB/s                                   12.88 us  5.21 TB/s
    //   16.8M (B = 256)      45.67 us  5.88 TB/s                                   42.85 us  6.26 TB/s
    // 5 / 6 CTAs per SM (48 / 40 registers, spills) and a one-item software-pipelined loop were slower at every size.
    const size_t n4 = n / 4;
    size_t blocks = (n4 + 511) / 512;
    if (blocks > (size_t)148 * 4) blocks = (size_t)148 * 4;
    if (blocks < 1) blocks = 1;
    sde_step_kernel<true, 4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        x_out, x, eps, mu, z, coef, input_is_score, use_philox, seed, elem_offset, rng_dev, n);
  } else {
    sde_step_kernel<false, 1><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        x_out, x, eps, mu, z, coef, input_is_score, use_philox, seed, elem_offset, rng_dev, n);
  }
  return check_launch("sde_step");
}

int idiff_sde_step(float* x_out, const float* x, const float* eps, const float* mu, const float* z,
                   const float* coef, int input_is_score, int use_philox, uint64_t seed, uint64_t elem_offset,
                   size_t n, void* stream) {
  return sde_step_launch(x_out, x, eps, mu, z, coef, input_is_score, use_philox, seed, elem_offset, nullptr,
                         elem_offset % 4 == 0, n, stream);
}

int idiff_sde_step_rng(float* x_out, const float* x, const float* eps, const float* mu, const float* coef,
                       int input_is_score, const uint64_t* rng_dev, int offset_is_multiple_of_4, size_t n,
                       void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(n == 0 || rng_dev, "sde_step_rng: null rng pointer");
  return sde_step_launch(x_out, x, eps, mu, nullptr, coef, input_is_score, 1, 0, 0, rng_dev,
                         offset_is_multiple_of_4 != 0, n, stream);
}

int idiff_sde_pack_table(const float* theta, const float* sigma, const float* sigma_bar, int T1, float dt,
                         double sqrt_dt, float* out) {
  IDIFF_REQUIRE(theta && sigma && sigma_bar && out && T1 > 0, "sde_pack_table: bad arguments");
  for (int t = 0; t < T1; ++t) {
    float* r = out + (size_t)t * 8;
    r[0] = theta[t];
    r[1] = sigma[t];
    r[2] = sigma_bar[t];
    r[3] = dt;
    r[4] = (float)sqrt_dt;      // Python double -> fp32 scalar, as torch does for `randn * math.sqrt(dt)`
    uint32_t step = (uint32_t)t;
    float as_float;
    static_assert(sizeof(as_float) == sizeof(step), "");
    __builtin_memcpy(&as_float, &step, 4);
    r[5] = as_float;
    r[6] = 0.f;
    r[7] = 0.f;
  }
  return IDIFF_OK;
}

int idiff_step_select(const float* table, int* t_counter, float* cur_row, float* cur_time, float sample_scale,
                      void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(table && t_counter && cur_row && cur_time, "step_select: null pointer");
  step_select_kernel<<<1, 32, 0, as_stream(stream)>>>(table, t_counter, cur_row, cur_time, sample_scale);
  return check_launch("step_select");
}

int idiff_step_select_ss(const float* table, int* t_counter, float* cur_row, float* cur_time, float sample_scale,
                         const float* ss_table, float* ss_out, int S, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(table && t_counter && cur_row && cur_time && ss_table && ss_out && S > 0, "step_select_ss: bad arguments");
  step_select_ss_kernel<<<1, 512, 0, as_stream(stream)>>>(table, t_counter, cur_row, cur_time, sample_scale, ss_table, ss_out, S);
  return check_launch("step_select_ss");
}

int idiff_noise_state(float* x_out, const float* mu, const float* z, float max_sigma, int use_philox,
                      uint64_t seed, uint64_t elem_offset, size_t n, void* stream) {
  using namespace idiff;
  if (n == 0) return IDIFF_OK;
  IDIFF_REQUIRE(x_out && mu && (z || use_philox), "noise_state: null pointer");
  noise_state_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x_out, mu, z, max_sigma, use_philox, seed,
                                                                      elem_offset, n);
  return check_launch("noise_state");
}

int idiff_random_states(float* xt_out, float* z_out, const float* x0, const float* mu, const float* z,
                        const float* decay, const float* sigma_bar, int use_philox, uint64_t seed,
                        uint64_t elem_offset, uint32_t step_word, size_t per_sample, size_t n, void* stream) {
  using namespace idiff;
  if (n == 0) return IDIFF_OK;
  IDIFF_REQUIRE(xt_out && x0 && mu && decay && sigma_bar && (z || use_philox), "random_states: null pointer");
  IDIFF_REQUIRE(per_sample > 0 && n % per_sample == 0, "random_states: n must be a multiple of per_sample");
  const bool vec = per_sample % 4 == 0 && elem_offset % 4 == 0 && aligned16(xt_out) && aligned16(x0) &&
                   aligned16(mu) && (!z || aligned16(z)) && (!z_out || aligned16(z_out));
  if (vec)
    random_states_kernel<true><<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        xt_out, z_out, x0, mu, z, decay, sigma_bar, use_philox, seed, elem_offset, step_word, per_sample, n);
  else
    random_states_kernel<false><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
        xt_out, z_out, x0, mu, z, decay, sigma_bar, use_philox, seed, elem_offset, step_word, per_sample, n);
  return check_launch("random_states");
}

int idiff_philox_normal(float* out, uint64_t seed, uint64_t elem_offset, uint32_t step, size_t n, void* stream) {
  using namespace idiff;
  if (n == 0) return IDIFF_OK;
  IDIFF_REQUIRE(out, "philox_normal: null pointer");
  philox_normal_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(out, seed, elem_offset, step, n);
  return check_launch("philox_normal");
}

}  // extern "C"
