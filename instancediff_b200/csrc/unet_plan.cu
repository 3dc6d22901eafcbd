// C-level network object: weight packing, launch plan and the reverse-SDE loop behind plain C entry points
// (idiff_unet_create / load_weight / finalize / set_context / forward / reverse_sde / destroy, include/idiff.h).
//
// This is the boundary SURVEY.md section 8b asks for: a host in any language hands over the fp32 parameters by their
// state_dict names and device pointers to x / mu / image embedding, and gets eps (one forward) or the restored state
// (the whole loop of utils/sde_utils.py:244-261) back -- without re-implementing the ~135-launch plan.  Host-only code:
// it packs the weights exactly like instancediff_b200/packing.py and replays the same C-ABI kernel entry points the
// Python plan (unet.py::_Plan, kept as the instrumented plan of the tests and of bench.py) calls, in the same order
// with the same parameters.  Architecture: SURVEY.md App. A; model call: utils/sde_utils.py:198.
#include <cuda_bf16.h>

#include <functional>
#include <stdlib.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "host_common.h"

namespace idiff {
namespace {

struct HostT {
  std::vector<float> v;
  std::vector<int> shape;
  size_t numel() const { return v.size(); }
};

inline uint16_t bf16_bits(float f) {
  const __nv_bfloat16 h = __float2bfloat16_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
inline float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

// ---- packing (instancediff_b200/packing.py) ---------------------------------------------------------------------
// w [N][Cin][k][k] (or [N][Cin]) -> [n_tile][chunk][tap][c8][n_local][e] bf16
std::vector<uint16_t> pack_conv_weight(const float* w, int N, int Cin, int k, int NT) {
  const int taps = k * k, nch = Cin / 64;
  std::vector<uint16_t> out((size_t)N * Cin * taps);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) {
        const int nt = n / NT, nl = n % NT, ch = c / 64, c8 = (c % 64) / 8, e = c % 8;
        const size_t idx = ((((((size_t)nt * nch + ch) * taps + t) * 8 + c8) * NT + nl) * 8) + e;
        out[idx] = bf16_bits(w[((size_t)n * Cin + c) * taps + t]);
      }
  return out;
}
// w [64][64][3][3] -> [dx][c8][blk = 2 - dy][n][e] bf16 (idiff_conv3_rowpair)
std::vector<uint16_t> pack_conv3_rowpair(const float* w) {
  std::vector<uint16_t> out(64 * 64 * 9);
  for (int n = 0; n < 64; ++n)
    for (int c = 0; c < 64; ++c)
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const int c8 = c / 8, e = c % 8, blk = 2 - dy;
          out[((((size_t)dx * 8 + c8) * 3 + blk) * 64 + n) * 8 + e] = bf16_bits(w[(((size_t)n * 64 + c) * 3 + dy) * 3 + dx]);
        }
  return out;
}
// LayerNorm folded into the following linear layer: W' = W diag(g), wsum from the bf16-rounded W', extra = W beta
void fold_layernorm(const std::vector<float>& w, int N, int K, const float* gain, const float* beta, std::vector<float>& wf,
                    std::vector<float>& wsum, std::vector<float>& extra) {
  wf.resize((size_t)N * K);
  wsum.assign(N, 0.f);
  extra.assign(beta ? N : 0, 0.f);
  for (int n = 0; n < N; ++n) {
    double s = 0.0, ex = 0.0;
    for (int k = 0; k < K; ++k) {
      const float x = w[(size_t)n * K + k] * gain[k];
      wf[(size_t)n * K + k] = x;
      s += (double)bf16_round(x);
      if (beta) ex += (double)w[(size_t)n * K + k] * (double)beta[k];
    }
    wsum[n] = (float)s;
    if (beta) extra[n] = (float)ex;
  }
}
std::vector<uint16_t> pack_stem_weight(const float* w /*[64][2][7][7]*/, const float* bias) {
  std::vector<float> full(64 * 224, 0.f);
  for (int n = 0; n < 64; ++n) {
    for (int ci = 0; ci < 2; ++ci)
      for (int t = 0; t < 49; ++t) {
        const float x = w[((size_t)n * 2 + ci) * 49 + t];
        full[(size_t)n * 224 + t * 2 + ci] = x;
        full[(size_t)n * 224 + 112 + t * 2 + ci] = x;
      }
    const float bh = bf16_round(bias[n]);
    full[(size_t)n * 224 + 98] = bh;
    full[(size_t)n * 224 + 99] = bias[n] - bh;
  }
  std::vector<uint16_t> out(28 * 64 * 8);
  for (int n = 0; n < 64; ++n)
    for (int g = 0; g < 28; ++g)
      for (int e = 0; e < 8; ++e) out[((size_t)g * 64 + n) * 8 + e] = bf16_bits(full[(size_t)n * 224 + g * 8 + e]);
  return out;
}
std::vector<uint16_t> pack_head_weight(const float* w /*[1][64][3][3]*/) {
  // [tap][k-chunk][k] from [c][ky][kx]; then the mma.sync B fragments (packing.py::pack_head_weight)
  std::vector<uint16_t> out(3 * 4 * 32 * 4, 0);
  for (int j = 0; j < 3; ++j)
    for (int lane = 0; lane < 32; ++lane) {
      const int n = lane / 4, k0 = (lane % 4) * 2, tap = 4 * j + n / 2;
      if (tap >= 9) continue;
      for (int kc = 0; kc < 4; ++kc) {
        const int ks[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
        for (int q = 0; q < 4; ++q) {
          const int c = kc * 16 + ks[q];
          const float x = w[(size_t)c * 9 + tap];
          const float hi = bf16_round(x);
          out[(((size_t)j * 4 + kc) * 32 + lane) * 4 + q] = bf16_bits(n % 2 == 0 ? hi : x - hi);
        }
      }
    }
  return out;
}

// ---- packed layer entries (device pointers into the arena) --------------------------------------------------------
struct Entry {
  std::string name;              // set for layers that may be re-packed for NT = 128 (N >= 256)
  const void* w = nullptr;       // packed bf16
  const void* w_rp = nullptr;    // row-pair packing (3x3 64 -> 64 only)
  const float* bias = nullptr;
  const float* wsum = nullptr;
  int N = 0, NT = 0, cin = 0, k = 1;
};
struct Norm { const float* g = nullptr; const float* b = nullptr; };
struct LinAttn {
  const void *wq = nullptr, *wk = nullptr;
  const float *wv = nullptr, *w_f32 = nullptr, *bias = nullptr, *g = nullptr;
  bool fused = false;
};
struct Cross { const float *wv = nullptr, *wo = nullptr, *bo = nullptr; int C = 0; };

struct Act { void* t = nullptr; int H = 0, W = 0, C = 0; float* stats = nullptr; };

}  // namespace

struct Plan;

struct Net {
  int nf = 64, td = 256, context_dim = 512, S = 0;
  std::vector<int> dims;                                  // [nf, nf*m...]
  std::map<std::string, HostT> raw;                        // fp32 parameters by state_dict name
  bool finalized = false;
  uint8_t* arena = nullptr;
  std::map<std::string, Entry> conv;
  std::map<std::string, Norm> norm;
  std::map<std::string, LinAttn> lattn;
  std::map<std::string, Cross> cross;
  std::map<std::string, const float*> vec;                 // loose fp32 vectors (prenorm gains ...)
  std::map<std::string, int> ss_off;
  bool fuse_gn = false;                                      // IDIFF_GN_FUSE=1: finalize folded into the producing conv (see unet.py)
  bool fuse_gn_st = true;                                    // IDIFF_NO_GN_FUSE_ST=1: chan_ln + gn_stats + gn_finalize launches
  const void *stem_w = nullptr, *head_w = nullptr;
  float head_bias = 0.f;
  const float *w1t = nullptr, *b1 = nullptr, *w2t = nullptr, *b2 = nullptr, *wss = nullptr, *bss = nullptr;
  std::map<std::vector<int>, std::unique_ptr<Plan>> plans;  // key {B, H, W, shared_time}
  std::map<int, std::vector<float*>> crossvec;              // per batch size: [3 layers][B][C]
  std::map<std::string, void*> variants;                    // "<layer>:<NT>" -> weights packed for another N tile
  std::map<std::string, std::vector<float>> variant_src;    // fp32 weights (after LayerNorm folding) of those layers
  const void* packed_variant(const std::string& name, int NT, int N, int cin, int k);
  int n_levels() const { return (int)dims.size() - 1; }
  std::vector<std::string> spatial_layers() const {
    return {"downs." + std::to_string(n_levels() - 1) + ".2", "mid_attn", "ups.0.2"};
  }
  ~Net();
};

struct Plan {
  Net* net;
  int B, H, W;
  bool shared_time;
  std::vector<std::function<int(void*)>> ops;
  std::vector<void*> owned;                                // cudaMalloc'd buffers
  std::map<std::string, void*> scratch;                    // shared temporaries by (name, bytes)
  std::vector<std::unique_ptr<idiff_gemm_params>> params;
  std::vector<std::unique_ptr<idiff_gn_fuse>> fuses;        // GroupNorm finalize folded into the producing launch
  void* gn_sums = nullptr;                                  // zeroed once; every fused launch leaves it zeroed again
  unsigned int* gn_arrivals = nullptr;
  std::map<std::string, idiff_gemm_params*> ctx_slots;
  float *eps = nullptr, *temb = nullptr, *ss = nullptr;
  void* x_first = nullptr;
  // loop state of idiff_unet_reverse_sde
  int* counter = nullptr;
  float *row = nullptr, *timev = nullptr;
  unsigned long long* rng = nullptr;
  int n_launch = 0;
  int err = 0;

  Plan(Net* n, int b, int h, int w, bool st) : net(n), B(b), H(h), W(w), shared_time(st) {}
  ~Plan() {
    for (void* p : owned) cudaFree(p);
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { err = 1; return nullptr; }
    owned.push_back(p);
    return p;
  }
  void* tmp(const std::string& name, size_t bytes) {
    const std::string key = name + ":" + std::to_string(bytes);
    auto it = scratch.find(key);
    if (it != scratch.end()) return it->second;
    void* p = alloc(bytes);
    scratch[key] = p;
    return p;
  }
  Act act(int h, int w, int c, bool stats = false, const char* tmp_name = nullptr) {
    Act a;
    a.H = h; a.W = w; a.C = c;
    const size_t bytes = (size_t)B * h * w * c * 2;
    a.t = tmp_name ? tmp(tmp_name, bytes) : alloc(bytes);
    if (stats) a.stats = (float*)alloc((size_t)B * h * w * 2 * sizeof(float));
    return a;
  }

  struct GemmOpt {
    int k = 1, stride = 1, up = 0, a_silu = 0, epi = IDIFF_EPI_PLAIN, cin0 = -1, src0_ld = 0, NT = -1, out_ld = -1;
    const float *a_scale = nullptr, *a_shift = nullptr, *row_stats = nullptr, *res0_scale = nullptr, *res0_shift = nullptr, *ln_g = nullptr;
    float *gn_partial = nullptr, *out_stats = nullptr;
    const idiff_gn_fuse* gn_fuse = nullptr;
    const void *res0 = nullptr, *res1 = nullptr, *w_override = nullptr;
    bool bias = true;
    const char* bias_img_slot = nullptr;
    long long w_image_stride = 0;
    float qscale = 1.f, ln_eps = 1e-5f;
  };
  bool rowpair_ok(const Entry& e, const Act* src1, const Act& out, int k, int stride, int up) const {
    if (!e.w_rp || src1 || k != 3 || stride != 1 || up) return false;
    const int strips = (out.W + 127) / 128;
    return out.H % 2 == 0 && (double)out.W / (strips * 128) >= 0.74;
  }
  int gn_rows(const Entry& e, const Act* src1, const Act& out) const {
    return rowpair_ok(e, src1, out, 3, 1, 0) ? idiff_conv3_rowpair_gn_rows(out.H, out.W) : idiff_conv_gemm_gn_rows(out.H, out.W);
  }
  void gemm(const Act& src0, const Act* src1, const Entry& e, const Act& out, const GemmOpt& o) {
    params.emplace_back(new idiff_gemm_params());
    idiff_gemm_params* p = params.back().get();
    memset(p, 0, sizeof(*p));
    p->B = B; p->H = out.H; p->W = out.W;
    p->ksize = o.k; p->stride = o.stride; p->up0 = o.up;
    p->cin0 = o.cin0 < 0 ? src0.C : o.cin0;
    p->cin1 = src1 ? src1->C : 0;
    p->src0_ld = o.src0_ld;
    p->N = e.N;
    p->NT = o.NT < 0 ? e.NT : o.NT;
    p->a_silu = o.a_silu; p->epi = o.epi;
    p->gn_groups = (o.gn_partial || o.gn_fuse) ? 8 : 0;
    p->gn_fuse = o.gn_fuse;
    p->out_ld = o.out_ld < 0 ? out.C : o.out_ld;
    p->src0 = src0.t; p->src1 = src1 ? src1->t : nullptr;
    p->a_scale = o.a_scale; p->a_shift = o.a_shift;
    p->w = o.w_override ? o.w_override : e.w;
    p->w_image_stride = o.w_image_stride;
    p->bias = o.bias ? e.bias : nullptr;
    p->row_stats = o.row_stats;
    p->wsum = o.row_stats ? e.wsum : nullptr;
    p->res0 = o.res0; p->res1 = o.res1;
    p->res0_scale = o.res0_scale; p->res0_shift = o.res0_shift;
    p->ln_g = o.ln_g;
    p->out = out.t;
    p->gn_partial = o.gn_partial;
    p->out_row_stats = o.out_stats;
    p->qscale = o.qscale; p->ln_eps = o.ln_eps;
    if (o.bias_img_slot) ctx_slots[o.bias_img_slot] = p;
    // small grids: split a 256-wide N tile in two (same rule and same bit-identical results as unet.py::_Plan.gemm)
    if (o.NT < 0 && p->NT == 256 && !e.name.empty() && !o.w_override && o.epi == IDIFF_EPI_PLAIN && !o.res0 && !o.res1 && !o.out_stats) {
      const int tiles = B * ((out.H + 15) / 16) * ((out.W + 7) / 8);
      if (tiles * (p->N / 256) <= 111) {
        const void* v = net->packed_variant(e.name, 128, e.N, e.cin, e.k);
        if (v) { p->NT = 128; p->w = v; }
      }
    }
    const bool rp = !o.w_override && rowpair_ok(e, src1, out, o.k, o.stride, o.up) && idiff_conv3_rowpair_supported(p);
    if (rp) {
      p->w = e.w_rp;
      if (o.a_silu) p->a_silu = 3;                          // packed bf16x2 affine + SiLU (the Python plan's default)
      ops.push_back([p](void* s) { return idiff_conv3_rowpair(p, s); });
    } else {
      ops.push_back([p](void* s) { return idiff_conv_gemm(p, s); });
    }
    ++n_launch;
  }
  // idiff_gn_fuse for a producer launch (same role as unet.py::_Plan.gn_fuse_desc)
  const idiff_gn_fuse* gn_fuse_desc(const Norm& nm, int C, int count, float eps_, int t_off, const char* tag, float** sc_out,
                                    float** sh_out) {
    if (!gn_sums) {
      const size_t bytes = (size_t)B * IDIFF_GN_SLOTS * 32 * 2 * 8;
      gn_sums = alloc(bytes);
      gn_arrivals = (unsigned int*)alloc(16);
      // plan-build time, once: the zeroing runs on the legacy stream, the plan's launches on the caller's (possibly
      // non-blocking) stream -- make it complete before the first of them can start
      if (gn_sums && gn_arrivals && (cudaMemset(gn_sums, 0, bytes) != cudaSuccess || cudaMemset(gn_arrivals, 0, 16) != cudaSuccess ||
                                     cudaDeviceSynchronize() != cudaSuccess)) err = 1;
    }
    float* sc = (float*)tmp(std::string(tag) + "_sc", (size_t)B * C * 4);
    float* sh = (float*)tmp(std::string(tag) + "_sh", (size_t)B * C * 4);
    fuses.emplace_back(new idiff_gn_fuse());
    idiff_gn_fuse* f = fuses.back().get();
    memset(f, 0, sizeof(*f));
    f->sums = gn_sums; f->arrivals = gn_arrivals;
    f->gamma = nm.g; f->beta = nm.b;
    if (t_off >= 0) {
      f->t_scale = ss + t_off;
      f->t_shift = ss + t_off + C;
      f->t_ld = shared_time ? 0 : net->S;
    }
    f->scale_out = sc; f->shift_out = sh;
    f->count_per_group = count; f->eps = eps_;
    *sc_out = sc;
    *sh_out = sh;
    return f;
  }
  void gn_finalize(const float* partial, int ntile, const Norm& nm, int C, int G, int count, float eps_, int t_off,
                   const char* tag, float** sc_out, float** sh_out) {
    float* sc = (float*)tmp(std::string(tag) + "_sc", (size_t)B * C * 4);
    float* sh = (float*)tmp(std::string(tag) + "_sh", (size_t)B * C * 4);
    const float *ts = nullptr, *tb = nullptr;
    int t_ld = 0;
    if (t_off >= 0) {
      ts = ss + t_off;
      tb = ss + t_off + C;
      t_ld = shared_time ? 0 : net->S;
    }
    const int Bv = B;
    ops.push_back([=](void* s) { return idiff_gn_finalize(partial, ntile, nm.g, nm.b, ts, tb, t_ld, sc, sh, Bv, C, G, count, eps_, s); });
    ++n_launch;
    *sc_out = sc;
    *sh_out = sh;
  }
  Act resblock(const std::string& prefix, const Act& src0, const Act* src1, int cout, bool want_stats = false) {
    const int h = src0.H, w = src0.W, cin = src0.C + (src1 ? src1->C : 0), count = h * w * (cout / 8);
    const Entry &c1 = net->conv.at(prefix + ".conv1"), &c2 = net->conv.at(prefix + ".conv2");
    Act y1 = act(h, w, cout, false, "y1"), y2 = act(h, w, cout, false, "y2");
    float *sc1, *sh1, *sc2, *sh2;
    GemmOpt o1, o2;
    o1.k = 3;
    o2.k = 3; o2.a_silu = 1;
    if (net->fuse_gn) {
      o1.gn_fuse = gn_fuse_desc(net->norm.at(prefix + ".norm1"), cout, count, 1e-5f, net->ss_off.at(prefix), "gn1", &sc1, &sh1);
      gemm(src0, src1, c1, y1, o1);
      o2.a_scale = sc1; o2.a_shift = sh1;
      o2.gn_fuse = gn_fuse_desc(net->norm.at(prefix + ".norm2"), cout, count, 1e-5f, -1, "gn2", &sc2, &sh2);
      gemm(y1, nullptr, c2, y2, o2);
    } else {
      const int nt1 = gn_rows(c1, src1, y1), nt2 = gn_rows(c2, nullptr, y2);
      float* part1 = (float*)tmp("gnp1", (size_t)B * nt1 * 16 * 4);
      float* part2 = (float*)tmp("gnp2", (size_t)B * nt2 * 16 * 4);
      o1.gn_partial = part1;
      gemm(src0, src1, c1, y1, o1);
      gn_finalize(part1, nt1, net->norm.at(prefix + ".norm1"), cout, 8, count, 1e-5f, net->ss_off.at(prefix), "gn1", &sc1, &sh1);
      o2.a_scale = sc1; o2.a_shift = sh1; o2.gn_partial = part2;
      gemm(y1, nullptr, c2, y2, o2);
      gn_finalize(part2, nt2, net->norm.at(prefix + ".norm2"), cout, 8, count, 1e-5f, -1, "gn2", &sc2, &sh2);
    }
    Act out = act(h, w, cout, want_stats);
    if (cin == cout) {
      const void *y = y2.t, *res = src0.t;
      void* dst = out.t;
      float* st = out.stats;
      const int Bv = B, hw = h * w;
      ops.push_back([=](void* s) { return idiff_block_tail(y, sc2, sh2, res, dst, st, 1e-5f, Bv, hw, cout, s); });
      ++n_launch;
    } else {
      GemmOpt o3;
      o3.k = 1; o3.res0 = y2.t; o3.res0_scale = sc2; o3.res0_shift = sh2; o3.out_stats = out.stats; o3.NT = cout;
      gemm(src0, src1, net->conv.at(prefix + ".res_conv"), out, o3);
    }
    return out;
  }
  Act linear_attn(const std::string& prefix, const Act& x) {
    const int h = x.H, w = x.W, C = x.C, HW = h * w;
    const std::string f = prefix + ".fn";
    const LinAttn& la = net->lattn.at(f);
    const int Bv = B;
    if (la.fused && HW % 128 == 0) {
      void* weff = tmp("la_weff", (size_t)B * C * 128 * 2);
      float* sc = (float*)tmp("laf_scratch", idiff_linattn_fused_scratch_floats(B, HW, C) * 4);
      Act out = act(h, w, C);
      const void* xin = x.t;
      const float* st = x.stats;
      void* dst = out.t;
      ops.push_back([=](void* s) {
        return idiff_linattn_fused(xin, st, la.wq, la.wk, la.wv, la.w_f32, la.bias, la.g, weff, dst, sc, Bv, HW, C, 0.17677669529663687f,
                                   1e-5f, s);
      });
      n_launch += 3;
      return out;
    }
    Act qkv = act(h, w, 384, false, "la_qkv");
    GemmOpt o;
    o.k = 1; o.bias = false; o.row_stats = x.stats; o.epi = IDIFF_EPI_QSOFTMAX; o.qscale = 0.17677669529663687f;
    gemm(x, nullptr, net->conv.at(f + ".to_qkv"), qkv, o);
    void* weff = tmp("la_weff", (size_t)B * C * 128 * 2);
    float* sc = (float*)tmp("la_scratch", idiff_linattn_scratch_floats(B, HW) * 4);
    const void* q = qkv.t;
    ops.push_back([=](void* s) { return idiff_linattn_context(q, la.w_f32, weff, sc, Bv, HW, C, s); });
    n_launch += 3;
    Act out = act(h, w, C);
    Entry e;
    e.N = C; e.NT = C; e.w = weff; e.bias = la.bias;
    GemmOpt o2;
    o2.k = 1; o2.cin0 = 128; o2.src0_ld = 384; o2.w_image_stride = (long long)C * 128; o2.epi = IDIFF_EPI_LN_OUT; o2.ln_g = la.g;
    o2.res0 = x.t; o2.w_override = weff;
    gemm(qkv, nullptr, e, out, o2);
    return out;
  }
  Act spatial_attn(const std::string& prefix, const Act& x) {
    const int h = x.H, w = x.W, C = x.C, HW = h * w;
    const size_t rows = (size_t)B * HW;
    const std::string f = prefix + ".fn";
    const int Bv = B;
    Act y = act(h, w, C, false, "st_y");
    float *sc, *sh;
    if (net->fuse_gn_st && HW % 32 == 0) {              // channel LayerNorm + GroupNorm(32) statistics + finalize: one launch
      const idiff_gn_fuse* fz = gn_fuse_desc(net->norm.at(f + ".norm"), C, HW * (C / 32), 1e-6f, -1, "gn32", &sc, &sh);
      const void* xin = x.t;
      const float* g = net->vec.at(prefix + ".prenorm");
      void* dst = y.t;
      ops.push_back([=](void* s) { return idiff_chan_ln_gn(xin, g, dst, 1e-5f, Bv, HW, C, 32, fz, s); });
      ++n_launch;
    } else {
      {
        const void* xin = x.t;
        const float* g = net->vec.at(prefix + ".prenorm");
        void* dst = y.t;
        ops.push_back([=](void* s) { return idiff_chan_ln(xin, g, dst, 1e-5f, rows, C, s); });
      }
      const int ntile = idiff_gn_stats_ntile(HW);
      float* part = (float*)tmp("st_gnp", (size_t)B * ntile * 32 * 2 * 4);
      {
        const void* src = y.t;
        ops.push_back([=](void* s) { return idiff_gn_stats(src, part, Bv, HW, C, 32, s); });
      }
      n_launch += 2;
      gn_finalize(part, ntile, net->norm.at(f + ".norm"), C, 32, HW * (C / 32), 1e-6f, -1, "gn32", &sc, &sh);
    }
    Act h0 = act(h, w, C, true);
    GemmOpt o;
    o.k = 1; o.a_scale = sc; o.a_shift = sh; o.out_stats = h0.stats;
    gemm(y, nullptr, net->conv.at(f + ".proj_in"), h0, o);
    Act qkv = act(h, w, 3 * C, false, "st_qkv");
    GemmOpt oq;
    oq.k = 1; oq.row_stats = h0.stats;
    gemm(h0, nullptr, net->conv.at(f + ".attn1.qkv"), qkv, oq);
    Act att = act(h, w, C, false, "st_att");
    {
      const void* q = qkv.t;
      void* dst = att.t;
      const int heads = C / 32;
      ops.push_back([=](void* s) { return idiff_self_attention(q, dst, Bv, HW, heads, 0.17677669529663687f, s); });
      ++n_launch;
    }
    Act h2 = act(h, w, C, true);
    GemmOpt oo;
    oo.k = 1; oo.res0 = h0.t; oo.out_stats = h2.stats;
    slot_names.push_back(prefix);
    oo.bias_img_slot = slot_names.back().c_str();
    gemm(att, nullptr, net->conv.at(f + ".attn1.to_out"), h2, oo);
    Act ff = act(h, w, 4 * C, false, "st_ff");
    GemmOpt of;
    of.k = 1; of.row_stats = h2.stats; of.epi = IDIFF_EPI_GEGLU; of.out_ld = 4 * C;
    gemm(h2, nullptr, net->conv.at(f + ".ff.proj"), ff, of);
    Act h3 = act(h, w, C, false, "st_h3");
    GemmOpt o3;
    o3.k = 1; o3.res0 = h2.t;
    gemm(ff, nullptr, net->conv.at(f + ".ff.out"), h3, o3);
    Act out = act(h, w, C);
    GemmOpt o4;
    o4.k = 1; o4.res0 = y.t; o4.res1 = x.t;
    gemm(h3, nullptr, net->conv.at(f + ".proj_out"), out, o4);
    return out;
  }
  std::vector<std::string> slot_names;                       // keeps the c_str() of the context slots alive
  Act convl(const std::string& name, const Act& x, int cout, int k, int stride, int up, const void* res0) {
    const int h = stride == 2 ? x.H / 2 : (up ? x.H * 2 : x.H), w = stride == 2 ? x.W / 2 : (up ? x.W * 2 : x.W);
    Act out = act(h, w, cout);
    GemmOpt o;
    o.k = k; o.stride = stride; o.up = up; o.res0 = res0;
    gemm(x, nullptr, net->conv.at(name), out, o);
    return out;
  }
  void build() {
    slot_names.reserve(8);
    const int nf = net->nf, n = net->n_levels();
    eps = (float*)alloc((size_t)B * H * W * 4);
    const int Bt = shared_time ? 1 : B;
    temb = (float*)alloc((size_t)Bt * net->td * 4);
    ss = (float*)alloc((size_t)Bt * net->S * 4);
    counter = (int*)alloc(16);
    row = (float*)alloc(32);
    timev = (float*)alloc(16);
    rng = (unsigned long long*)alloc(16);
    n_launch += 3;                                           // time embedding (2 kernels) + stem
    Act xf = act(H, W, nf);
    x_first = xf.t;
    Act h = xf;
    std::vector<Act> skips;
    for (int i = 0; i < n; ++i) {
      const bool last = i == n - 1;
      const int di = net->dims[i], dn = net->dims[i + 1];
      const std::string d = "downs." + std::to_string(i);
      h = resblock(d + ".0", h, nullptr, di);
      skips.push_back(h);
      h = resblock(d + ".1", h, nullptr, di, !last);
      h = last ? spatial_attn(d + ".2", h) : linear_attn(d + ".2", h);
      skips.push_back(h);
      h = last ? convl(d + ".3", h, dn, 3, 1, 0, nullptr) : convl(d + ".3", h, dn, 4, 2, 0, nullptr);
    }
    h = resblock("mid_block1", h, nullptr, net->dims.back());
    h = spatial_attn("mid_attn", h);
    h = resblock("mid_block2", h, nullptr, net->dims.back());
    for (int i = 0; i < n; ++i) {
      const int lvl = n - 1 - i, di = net->dims[lvl], dn = net->dims[lvl + 1];
      const bool last = lvl == n - 1;
      const std::string u = "ups." + std::to_string(i);
      Act s1 = skips.back();
      skips.pop_back();
      h = resblock(u + ".0", h, &s1, dn);
      Act s2 = skips.back();
      skips.pop_back();
      h = resblock(u + ".1", h, &s2, dn, !last);
      h = last ? spatial_attn(u + ".2", h) : linear_attn(u + ".2", h);
      if (lvl > 0) h = convl(u + ".3.conv", h, di, 3, 1, 1, nullptr);
      else h = convl(u + ".3", h, di, 3, 1, 0, xf.t);         // epilogue also adds the stem output (x + x_)
    }
    h = resblock("final_res", h, &xf, nf);
    {
      const void* src = h.t;
      const void* hw = net->head_w;
      const float hb = net->head_bias;
      float* dst = eps;
      const int Bv = B, Hv = H, Wv = W;
      ops.push_back([=](void* s) { return idiff_head_conv3(src, hw, hb, dst, Bv, Hv, Wv, nf, s); });
      ++n_launch;
    }
  }
  int run(const float* x, const float* mu, const float* t_dev, float t_scalar, void* stream) {
    const int Bt = shared_time ? 1 : B;
    int rc = idiff_time_embed(t_dev, t_scalar, net->w1t, net->b1, net->w2t, net->b2, net->wss, net->bss, temb, ss, Bt, net->nf, net->S, stream);
    if (rc) return rc;
    rc = idiff_stem_conv7_tc(x, mu, net->stem_w, x_first, B, H, W, stream);
    if (rc) return rc;
    for (auto& op : ops)
      if ((rc = op(stream)) != 0) return rc;
    return 0;
  }
};

const void* Net::packed_variant(const std::string& name, int NT, int N, int cin, int k) {
  const std::string key = name + ":" + std::to_string(NT);
  auto it = variants.find(key);
  if (it != variants.end()) return it->second;
  auto src = variant_src.find(name);
  if (src == variant_src.end()) return nullptr;
  const std::vector<uint16_t> pk = pack_conv_weight(src->second.data(), N, cin, k, NT);
  void* dev = nullptr;
  if (cudaMalloc(&dev, pk.size() * 2) != cudaSuccess || cudaMemcpy(dev, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
    return nullptr;
  variants[key] = dev;
  return dev;
}

Net::~Net() {
  plans.clear();
  for (auto& kv : variants) cudaFree(kv.second);
  for (auto& kv : crossvec)
    for (float* p : kv.second) cudaFree(p);
  if (arena) cudaFree(arena);
}

namespace {

// arena builder: host image + fix-ups of the pointers that refer into it
struct ArenaBuilder {
  std::vector<uint8_t> host;
  std::vector<std::pair<const void**, size_t>> fix;
  template <typename T>
  void add(const std::vector<T>& v, const void** dst) {
    const size_t off = (host.size() + 255) / 256 * 256;
    host.resize(off + v.size() * sizeof(T));
    memcpy(host.data() + off, v.data(), v.size() * sizeof(T));
    fix.emplace_back(dst, off);
  }
  template <typename T>
  void addf(const std::vector<T>& v, const float** dst) { add(v, reinterpret_cast<const void**>(dst)); }
};

int finalize(Net* net) {
  auto& R = net->raw;
  auto need = [&](const std::string& k) -> const HostT* {
    auto it = R.find(k);
    return it == R.end() ? nullptr : &it->second;
  };
  ArenaBuilder ab;
  std::string missing;
  for (auto& kv : net->variants) cudaFree(kv.second);        // re-finalize: packings of the previous weights
  net->variants.clear();
  net->variant_src.clear();
  auto get = [&](const std::string& k) -> const HostT& {
    static const HostT empty;
    const HostT* t = need(k);
    if (!t) { if (missing.empty()) missing = k; return empty; }
    return *t;
  };
  auto conv_entry = [&](const std::string& name, int NT, const std::vector<float>* w_over, const std::vector<float>* b_over, int N, int cin, int k) {
    Entry& e = net->conv[name];
    e.N = N; e.NT = NT; e.cin = cin; e.k = k;
    const std::vector<float>& w = w_over ? *w_over : get(name + ".weight").v;
    if (w.size() != (size_t)N * cin * k * k) return;
    ab.add(pack_conv_weight(w.data(), N, cin, k, NT), &e.w);
    if (N >= 256) { e.name = name; net->variant_src[name] = w; }
    if (N == 64 && cin == 64 && k == 3) ab.add(pack_conv3_rowpair(w.data()), &e.w_rp);
    if (b_over) ab.addf(*b_over, &e.bias);
    else if (need(name + ".bias")) ab.addf(get(name + ".bias").v, &e.bias);
  };
  auto plain_conv = [&](const std::string& name) {
    const HostT& w = get(name + ".weight");
    if (w.shape.size() < 2) return;
    const int N = w.shape[0], cin = w.shape[1], k = w.shape.size() == 4 ? w.shape[2] : 1;
    conv_entry(name, N < 256 ? N : 256, nullptr, nullptr, N, cin, k);
  };
  auto resblock = [&](const std::string& p) {
    plain_conv(p + ".conv1");
    plain_conv(p + ".conv2");
    if (need(p + ".res_conv.weight")) plain_conv(p + ".res_conv");
    for (const char* nn : {"norm1", "norm2"}) {
      Norm& nm = net->norm[p + "." + nn];
      ab.addf(get(p + "." + nn + ".weight").v, &nm.g);
      ab.addf(get(p + "." + nn + ".bias").v, &nm.b);
    }
  };
  std::vector<std::string> res_names;
  auto attn = [&](const std::string& p, bool spatial, int dim) {
    const std::string f = p + ".fn";
    const std::vector<float>& g_pre = get(p + ".prenorm.g").v;
    if ((int)g_pre.size() != dim) return;
    if (!spatial) {
      const std::vector<float>& wq = get(f + ".to_qkv.weight").v;                    // [384][dim]
      if (wq.size() != (size_t)384 * dim) return;
      std::vector<float> wf, wsum, extra;
      fold_layernorm(wq, 384, dim, g_pre.data(), nullptr, wf, wsum, extra);
      conv_entry(f + ".to_qkv", 128, &wf, nullptr, 384, dim, 1);
      net->conv[f + ".to_qkv"].bias = nullptr;
      ab.addf(wsum, &net->conv[f + ".to_qkv"].wsum);
      LinAttn& la = net->lattn[f];
      ab.addf(get(f + ".to_out.weight").v, &la.w_f32);
      ab.addf(get(f + ".to_out.bias").v, &la.bias);
      ab.addf(get(f + ".out_norm.g").v, &la.g);
      if (dim == 64 || dim == 128) {
        la.fused = true;
        std::vector<float> qv(wf.begin(), wf.begin() + (size_t)128 * dim), kv(wf.begin() + (size_t)128 * dim, wf.begin() + (size_t)256 * dim),
            vv(wf.begin() + (size_t)256 * dim, wf.end());
        ab.add(pack_conv_weight(qv.data(), 128, dim, 1, 128), &la.wq);
        ab.add(pack_conv_weight(kv.data(), 128, dim, 1, 128), &la.wk);
        ab.addf(vv, &la.wv);
      }
    } else {
      ab.addf(g_pre, &net->vec[p + ".prenorm"]);
      Norm& nm = net->norm[f + ".norm"];
      ab.addf(get(f + ".norm.weight").v, &nm.g);
      ab.addf(get(f + ".norm.bias").v, &nm.b);
      plain_conv(f + ".proj_in");
      std::vector<float> wqkv;
      for (const char* nn : {"to_q", "to_k", "to_v"}) {
        const std::vector<float>& w = get(f + ".attn1." + nn + ".weight").v;
        wqkv.insert(wqkv.end(), w.begin(), w.end());
      }
      if (wqkv.size() != (size_t)3 * dim * dim) return;
      std::vector<float> wf, wsum, extra;
      fold_layernorm(wqkv, 3 * dim, dim, get(f + ".norm1.weight").v.data(), get(f + ".norm1.bias").v.data(), wf, wsum, extra);
      conv_entry(f + ".attn1.qkv", 256, &wf, &extra, 3 * dim, dim, 1);
      ab.addf(wsum, &net->conv[f + ".attn1.qkv"].wsum);
      plain_conv(f + ".attn1.to_out");
      // GEGLU: value / gate rows interleaved, then LayerNorm(norm3) folded
      const std::vector<float>&pw = get(f + ".ff.proj.weight").v, &pb = get(f + ".ff.proj.bias").v;
      const int F = 4 * dim;
      if (pw.size() != (size_t)2 * F * dim) return;
      std::vector<float> wi((size_t)2 * F * dim), bi(2 * F);
      for (int j = 0; j < F; ++j) {
        memcpy(&wi[(size_t)(2 * j) * dim], &pw[(size_t)j * dim], dim * 4);
        memcpy(&wi[(size_t)(2 * j + 1) * dim], &pw[(size_t)(F + j) * dim], dim * 4);
        bi[2 * j] = pb[j];
        bi[2 * j + 1] = pb[F + j];
      }
      fold_layernorm(wi, 2 * F, dim, get(f + ".norm3.weight").v.data(), get(f + ".norm3.bias").v.data(), wf, wsum, extra);
      for (int j = 0; j < 2 * F; ++j) bi[j] += extra[j];
      conv_entry(f + ".ff.proj", 256, &wf, &bi, 2 * F, dim, 1);
      ab.addf(wsum, &net->conv[f + ".ff.proj"].wsum);
      plain_conv(f + ".ff.out");
      plain_conv(f + ".proj_out");
      Cross& cr = net->cross[p];
      cr.C = dim;
      ab.addf(get(f + ".attn2.to_v.weight").v, &cr.wv);
      ab.addf(get(f + ".attn2.to_out.weight").v, &cr.wo);
      ab.addf(get(f + ".attn2.to_out.bias").v, &cr.bo);
    }
  };
  // stem / head
  {
    const HostT &w = get("init_conv.weight"), &b = get("init_conv.bias");
    if (w.numel() == 64 * 2 * 49 && b.numel() == 64) ab.add(pack_stem_weight(w.v.data(), b.v.data()), &net->stem_w);
    const HostT &hw = get("final_conv.weight"), &hb = get("final_conv.bias");
    if (hw.numel() == 64 * 9 && hb.numel() >= 1) {
      ab.add(pack_head_weight(hw.v.data()), &net->head_w);
      net->head_bias = hb.v[0];
    }
  }
  const int n = net->n_levels();
  for (int i = 0; i < n; ++i) {
    const std::string d = "downs." + std::to_string(i);
    for (int j = 0; j < 2; ++j) { resblock(d + "." + std::to_string(j)); res_names.push_back(d + "." + std::to_string(j)); }
    attn(d + ".2", i == n - 1, net->dims[i]);
    plain_conv(d + ".3");
  }
  resblock("mid_block1"); res_names.push_back("mid_block1");
  attn("mid_attn", true, net->dims.back());
  resblock("mid_block2"); res_names.push_back("mid_block2");
  for (int i = 0; i < n; ++i) {
    const int lvl = n - 1 - i;
    const std::string u = "ups." + std::to_string(i);
    for (int j = 0; j < 2; ++j) { resblock(u + "." + std::to_string(j)); res_names.push_back(u + "." + std::to_string(j)); }
    attn(u + ".2", lvl == n - 1, net->dims[lvl + 1]);
    plain_conv(lvl > 0 ? u + ".3.conv" : u + ".3");
  }
  resblock("final_res"); res_names.push_back("final_res");
  // time embedding: transposed MLP weights + every ResBlock's Linear concatenated
  {
    const HostT &w1 = get("time_lin1.weight"), &w2 = get("time_lin2.weight");
    const int nf = net->nf, td = net->td;
    if (w1.numel() == (size_t)td * nf && w2.numel() == (size_t)td * td) {
      std::vector<float> w1t((size_t)nf * td), w2t((size_t)td * td);
      for (int o = 0; o < td; ++o)
        for (int i = 0; i < nf; ++i) w1t[(size_t)i * td + o] = w1.v[(size_t)o * nf + i];
      for (int o = 0; o < td; ++o)
        for (int i = 0; i < td; ++i) w2t[(size_t)i * td + o] = w2.v[(size_t)o * td + i];
      ab.addf(w1t, &net->w1t);
      ab.addf(w2t, &net->w2t);
    }
    ab.addf(get("time_lin1.bias").v, &net->b1);
    ab.addf(get("time_lin2.bias").v, &net->b2);
    std::vector<float> wss, bss;
    int off = 0;
    for (const std::string& nm : res_names) {
      const HostT &w = get(nm + ".mlp.weight"), &b = get(nm + ".mlp.bias");
      net->ss_off[nm] = off;
      off += w.shape.empty() ? 0 : w.shape[0];
      wss.insert(wss.end(), w.v.begin(), w.v.end());
      bss.insert(bss.end(), b.v.begin(), b.v.end());
    }
    net->S = off;
    ab.addf(wss, &net->wss);
    ab.addf(bss, &net->bss);
  }
  if (!missing.empty()) return fail(IDIFF_ERR_ARG, "unet_finalize: parameter '%s' was not loaded", missing.c_str());
  if (net->arena) { cudaFree(net->arena); net->arena = nullptr; }
  if (cudaMalloc(&net->arena, ab.host.size() + 256) != cudaSuccess) return fail(IDIFF_ERR_CUDA, "unet_finalize: arena allocation failed");
  if (cudaMemcpy(net->arena, ab.host.data(), ab.host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(IDIFF_ERR_CUDA, "unet_finalize: arena upload failed");
  for (auto& fx : ab.fix) *fx.first = net->arena + fx.second;
  net->plans.clear();
  net->finalized = true;
  return IDIFF_OK;
}

Plan* get_plan(Net* net, int B, int H, int W, bool shared) {
  const std::vector<int> key = {B, H, W, shared ? 1 : 0};
  auto it = net->plans.find(key);
  if (it != net->plans.end()) return it->second.get();
  std::unique_ptr<Plan> p(new Plan(net, B, H, W, shared));
  p->build();
  if (p->err) { fail(IDIFF_ERR_CUDA, "unet plan: workspace allocation failed"); return nullptr; }
  auto cv = net->crossvec.find(B);
  if (cv != net->crossvec.end()) {
    const std::vector<std::string> names = net->spatial_layers();
    for (size_t i = 0; i < names.size(); ++i) {
      auto s = p->ctx_slots.find(names[i]);
      if (s != p->ctx_slots.end()) s->second->bias_img = cv->second[i];
    }
  }
  Plan* raw = p.get();
  net->plans[key] = std::move(p);
  return raw;
}

}  // namespace
}  // namespace idiff

extern "C" {
using namespace idiff;

struct idiff_unet { Net net; };

int idiff_unet_create(const idiff_unet_cfg* cfg, idiff_unet** out) {
  IDIFF_REQUIRE(cfg && out, "unet_create: null argument");
  IDIFF_REQUIRE(cfg->in_nc == 2 && cfg->out_nc == 1 && cfg->nf == 64 && cfg->n_levels == 4 && cfg->ch_mult[0] == 1 && cfg->ch_mult[1] == 2 &&
                    cfg->ch_mult[2] == 4 && cfg->ch_mult[3] == 4 && cfg->down_kernel == 4,
                "unet_create: kernels are built for in_nc=2, out_nc=1, nf=64, ch_mult=[1,2,4,4], down_kernel=4 (Configurations/config.yml:109-113)");
  IDIFF_REQUIRE(cfg->context_dim > 0, "unet_create: context_dim");
  idiff_unet* u = new idiff_unet();
  u->net.nf = cfg->nf;
  u->net.td = cfg->nf * 4;
  u->net.context_dim = cfg->context_dim;
  {
    const char* e = getenv("IDIFF_GN_FUSE");                  // A/B switches, same as unet.py
    u->net.fuse_gn = e && e[0] == '1';
    e = getenv("IDIFF_NO_GN_FUSE_ST");
    u->net.fuse_gn_st = !(e && e[0] == '1');
  }
  u->net.dims.push_back(cfg->nf);
  for (int i = 0; i < cfg->n_levels; ++i) u->net.dims.push_back(cfg->nf * cfg->ch_mult[i]);
  *out = u;
  return IDIFF_OK;
}

void idiff_unet_destroy(idiff_unet* u) { delete u; }

int idiff_unet_load_weight(idiff_unet* u, const char* name, const float* data_host, int ndim, const int64_t* shape) {
  IDIFF_REQUIRE(u && name && data_host && ndim >= 0 && ndim <= 4 && (ndim == 0 || shape), "unet_load_weight: bad arguments");
  HostT t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back((int)shape[i]); n *= (size_t)shape[i]; }
  t.v.assign(data_host, data_host + n);
  u->net.raw[name] = std::move(t);
  u->net.finalized = false;
  return IDIFF_OK;
}

int idiff_unet_finalize(idiff_unet* u) {
  IDIFF_REQUIRE(u, "unet_finalize: null");
  return finalize(&u->net);
}

int idiff_unet_set_context(idiff_unet* u, const float* ctx, int B, void* stream) {
  IDIFF_REQUIRE(u && ctx && B > 0, "unet_set_context: bad arguments");
  IDIFF_REQUIRE(u->net.finalized, "unet_set_context: call idiff_unet_finalize first");
  Net& net = u->net;
  const std::vector<std::string> names = net.spatial_layers();
  auto& bufs = net.crossvec[B];
  if (bufs.empty()) {
    for (const std::string& nm : names) {
      float* p = nullptr;
      if (cudaMalloc(&p, (size_t)B * net.cross.at(nm).C * 4) != cudaSuccess) return fail(IDIFF_ERR_CUDA, "unet_set_context: allocation failed");
      bufs.push_back(p);
    }
    for (auto& kv : net.plans)
      if (kv.first[0] == B)
        for (size_t i = 0; i < names.size(); ++i) {
          auto s = kv.second->ctx_slots.find(names[i]);
          if (s != kv.second->ctx_slots.end()) s->second->bias_img = bufs[i];
        }
  }
  for (size_t i = 0; i < names.size(); ++i) {
    const Cross& c = net.cross.at(names[i]);
    const int rc = idiff_cross_vec(ctx, c.wv, c.wo, c.bo, bufs[i], B, net.context_dim, c.C, stream);
    if (rc) return rc;
  }
  return IDIFF_OK;
}

int idiff_unet_forward(idiff_unet* u, const float* x, const float* mu, const float* t_dev, float t_scalar, float* eps_out, int B,
                       int H, int W, void* stream) {
  IDIFF_REQUIRE(u && x && mu && eps_out && B > 0, "unet_forward: bad arguments");
  IDIFF_REQUIRE(u->net.finalized, "unet_forward: call idiff_unet_finalize first");
  IDIFF_REQUIRE(H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, "unet_forward: H and W must be multiples of 16 (pad on the host side)");
  IDIFF_REQUIRE(u->net.crossvec.count(B), "unet_forward: call idiff_unet_set_context for this batch size first");
  Plan* p = get_plan(&u->net, B, H, W, t_dev == nullptr);
  if (!p) return IDIFF_ERR_CUDA;
  int rc = p->run(x, mu, t_dev, t_scalar, stream);
  if (rc) return rc;
  if (cudaMemcpyAsync(eps_out, p->eps, (size_t)B * H * W * 4, cudaMemcpyDeviceToDevice, as_stream(stream)) != cudaSuccess)
    return fail(IDIFF_ERR_CUDA, "unet_forward: copy of the result failed");
  return IDIFF_OK;
}

int idiff_unet_reverse_sde(idiff_unet* u, float* x, const float* mu, const float* table, int T, float sample_scale, uint64_t seed,
                           uint64_t elem_offset, int B, int H, int W, void* stream) {
  IDIFF_REQUIRE(u && x && mu && table && B > 0 && T >= 0, "unet_reverse_sde: bad arguments");
  IDIFF_REQUIRE(u->net.finalized, "unet_reverse_sde: call idiff_unet_finalize first");
  IDIFF_REQUIRE(H % 16 == 0 && W % 16 == 0, "unet_reverse_sde: H and W must be multiples of 16");
  IDIFF_REQUIRE(u->net.crossvec.count(B), "unet_reverse_sde: call idiff_unet_set_context for this batch size first");
  if (T == 0) return IDIFF_OK;
  Plan* p = get_plan(&u->net, B, H, W, true);
  if (!p) return IDIFF_ERR_CUDA;
  cudaStream_t st = as_stream(stream);
  const unsigned long long rng_h[2] = {seed, elem_offset};
  if (cudaMemcpyAsync(p->counter, &T, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(p->rng, rng_h, 16, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return fail(IDIFF_ERR_CUDA, "unet_reverse_sde: loop state upload failed");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(IDIFF_ERR_CUDA, "unet_reverse_sde: sync failed");   // rng_h / T live on this stack
  const size_t n = (size_t)B * H * W;
  for (int step = 0; step < T; ++step) {                     // for t in reversed(range(1, T + 1))  (utils/sde_utils.py:248)
    int rc = idiff_step_select(table, p->counter, p->row, p->timev, sample_scale, stream);
    if (!rc) rc = p->run(x, mu, p->timev, 0.f, stream);
    if (!rc) rc = idiff_sde_step_rng(x, x, p->eps, mu, p->row, 0, reinterpret_cast<const uint64_t*>(p->rng), elem_offset % 4 == 0, n, stream);
    if (rc) return rc;
  }
  return IDIFF_OK;
}

int idiff_unet_num_launches(idiff_unet* u, int B, int H, int W) {
  IDIFF_REQUIRE(u && u->net.finalized && B > 0 && H % 16 == 0 && W % 16 == 0, "unet_num_launches: bad arguments");
  Plan* p = get_plan(&u->net, B, H, W, true);
  return p ? p->n_launch : IDIFF_ERR_CUDA;
}

}  // extern "C"
