// Host side of the implicit-GEMM convolution / linear engine: argument validation, shared-memory plan,
// TMA tensor maps of the output / residual tensors, dispatch to the per-kernel-size translation units
// (conv_gemm_k1.cu / _k3.cu / _k4.cu instantiate conv_gemm_kernel.cuh).
#include "conv_gemm_kernel.cuh"

namespace idiff {

struct SmemPlan {
  int SA, SB, resident, stageA, stageB, offA, offB, offP, offImg, offOut, offRes, nbuf_out, res_stride, res_nbuf, total;
};

static int stage_a_bytes(int ks) {
  return ks == 1 ? GeomT<1>::STAGE : ks == 3 ? GeomT<3>::STAGE : GeomT<4>::STAGE;
}
static bool tma_residuals(const idiff_gemm_params& p) { return p.NT == 64 && (p.res0 || p.res1); }

// Layout: [header][output staging][residual staging][A stages][weights][column params][per-image params].
// Weights stay RESIDENT for the whole (persistent) CTA when all of them fit next to at least two A stages;
// otherwise they stream through a ring of SB stages.
static SmemPlan plan_smem_n(const idiff_gemm_params& p, int nbuf_out, int res_nbuf) {
  SmemPlan s;
  const int nchunks = (p.cin0 + p.cin1) / 64;
  const int nk = nchunks * p.ksize * p.ksize;
  s.stageA = stage_a_bytes(p.ksize);
  s.stageB = p.NT * 128;
  s.nbuf_out = p.epi == IDIFF_EPI_GEGLU ? 0 : nbuf_out;
  s.res_stride = tma_residuals(p) ? kStageTile * ((p.res0 ? 1 : 0) + (p.res1 ? 1 : 0)) : 0;
  const int pbytes = ((3 * p.N * 4 + 127) / 128) * 128;                // bias | wsum | ln_g cache
  const int ibytes = (p.bias_img || p.res0_scale) ? kImgBytes : 0;
  s.offOut = kHeader;
  s.offRes = s.offOut + kEpiWarps * s.nbuf_out * kStageTile;
  s.res_nbuf = s.res_stride ? res_nbuf : 1;
  s.offA = s.offRes + kEpiWarps * s.res_nbuf * s.res_stride;
  const int budget = kSmemLimit - s.offA - pbytes - ibytes;
  const long wbytes = (long)nk * s.stageB * (p.N / p.NT);           // every N tile
  s.resident = (p.w_image_stride == 0 && wbytes + 2 * s.stageA <= budget) ? 1 : 0;
  if (s.resident) {
    int sa = (int)((budget - wbytes) / s.stageA);
    s.SA = sa > kMaxSA ? kMaxSA : sa;
    s.SB = 0;
    s.offB = s.offA + s.SA * s.stageA;
    s.offP = s.offB + (int)wbytes;
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
  }
  s.offImg = s.offP + pbytes;
  s.total = s.offImg + ibytes;
  return s;
}

// A second output staging tile per warp only pays when a row has several 64-channel boxes (NT > 64) and it
// does not cost pipeline depth.
static bool no_worse(const SmemPlan& s2, const SmemPlan& s1) {
  return s2.total <= kSmemLimit && s2.resident == s1.resident && s2.SA >= (s1.SA < 3 ? s1.SA : 3) &&
         (s2.resident || s2.SB >= (s1.SB < 4 ? s1.SB : 4));
}
// The same rule decides on a second set of residual staging tiles (TMA residuals of the NEXT item prefetched).
static SmemPlan plan_smem(const idiff_gemm_params& p) {
  const SmemPlan s1 = plan_smem_n(p, 1, 1);
  if (tma_residuals(p)) {
    const SmemPlan s2 = plan_smem_n(p, 1, 2);
    return no_worse(s2, s1) ? s2 : s1;
  }
  if (p.NT == 64 || p.epi == IDIFF_EPI_GEGLU) return s1;
  const SmemPlan s2 = plan_smem_n(p, 2, 1);
  return no_worse(s2, s1) ? s2 : s1;
}

int watchdog_conv(int clear) {
  const int a = watchdog_conv_k1(clear), b = watchdog_conv_k3(clear), c = watchdog_conv_k4(clear);
  if (a < 0 || b < 0 || c < 0) return -1;
  return a ? a : (b ? b : c);
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
  IDIFF_REQUIRE(p.epi >= IDIFF_EPI_PLAIN && p.epi <= IDIFF_EPI_LN_OUT, "conv_gemm: unknown epilogue %d", p.epi);
  if (p.epi != IDIFF_EPI_PLAIN) IDIFF_REQUIRE(p.ksize == 1, "conv_gemm: fused epilogues are built for 1x1 layers only");
  if (p.epi == IDIFF_EPI_LN_OUT) IDIFF_REQUIRE(p.NT == p.N && p.ln_g, "conv_gemm: LN_OUT needs NT == N and ln_g");
  if (p.epi == IDIFF_EPI_QSOFTMAX) IDIFF_REQUIRE(p.N >= 128 && p.NT == 128, "conv_gemm: QSOFTMAX needs NT == 128");
  if (p.epi == IDIFF_EPI_GEGLU) IDIFF_REQUIRE(p.NT == 256 && p.out_ld >= p.N / 2 && !p.res0 && !p.res1, "conv_gemm: GEGLU needs NT == 256, out_ld >= N/2, no residuals");
  else IDIFF_REQUIRE(p.out_ld >= p.N, "conv_gemm: out_ld < N");
  IDIFF_REQUIRE(p.out_ld % 8 == 0, "conv_gemm: out_ld %% 8");
  IDIFF_REQUIRE(p.src0_ld % 8 == 0 && p.src1_ld % 8 == 0 && (p.src0_ld == 0 || p.src0_ld >= p.cin0) &&
                    (p.src1_ld == 0 || p.src1_ld >= p.cin1), "conv_gemm: bad source pitch");
  if (p.out_row_stats) IDIFF_REQUIRE(p.NT == p.N && p.epi != IDIFF_EPI_GEGLU && p.epi != IDIFF_EPI_QSOFTMAX, "conv_gemm: out_row_stats needs NT == N");
  if (p.gn_groups > 0) {
    IDIFF_REQUIRE((p.gn_partial || p.gn_fuse) && p.gn_groups == 8 && p.N % 64 == 0 && p.N <= 256 && p.NT % (p.N / 8) == 0,
                  "conv_gemm: fused GroupNorm partials need 8 groups of 8, 16 or 32 channels and NT a multiple of the group width");
    IDIFF_REQUIRE(p.epi == IDIFF_EPI_PLAIN, "conv_gemm: GroupNorm partials need the plain epilogue");
  }
  if (p.res0_scale) IDIFF_REQUIRE(p.res0 && p.res0_shift, "conv_gemm: res0 affine needs res0 and shift");
  if (p.a_scale) {
    IDIFF_REQUIRE(p.ksize != 4, "conv_gemm: the A affine is not built for the 4x4 stride-2 layers");
    IDIFF_REQUIRE(p.ksize != 3 || p.a_silu, "conv_gemm: 3x3 layers are built with affine+SiLU or no transform");
  } else {
    IDIFF_REQUIRE(!p.a_silu, "conv_gemm: a_silu needs a_scale/a_shift");
  }
  if (tma_residuals(p)) {
    IDIFF_REQUIRE((!p.res0 || idiff::aligned16(p.res0)) && (!p.res1 || idiff::aligned16(p.res1)), "conv_gemm: residual alignment");
  }
  return IDIFF_OK;
}

// ---- TMA tensor maps ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// bf16 [B][H][W][ld] tensor of which `cols` channels are addressed; box = 64 channels x 8 x 4 pixels (one epilogue
// warp's share of a tile), 128B swizzle.  Out-of-range box parts are clipped (stores) / zero-filled (loads).
int make_map_box(CUtensorMap* tm, const void* base, int cols, int ld, int B, int H, int W, int box_w, int box_h);
static int make_map(CUtensorMap* tm, const void* base, int cols, int ld, int B, int H, int W) {
  return make_map_box(tm, base, cols, ld, B, H, W, TILE_W, 4);
}
// general form: box = 64 channels x box_w x box_h pixels (conv3_rowpair.cu stores 32 x 1 pixel rows)
int make_map_box(CUtensorMap* tm, const void* base, int cols, int ld, int B, int H, int W, int box_w, int box_h) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(IDIFF_ERR_CUDA, "conv_gemm: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IDIFF_ERR_CUDA, "conv_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return IDIFF_OK;
}

// 2-D view [rows][cols] of a bf16 matrix with row pitch ld (elements), box = box_c columns x box_r rows, no swizzle:
// a {8, 128} box lands as 128 rows x 16 B = one 8-channel plane of the UMMA no-swizzle K-major layout
// (linattn_fused.cu loads its raw activation tiles this way).
int make_map_2d(CUtensorMap* tm, const void* base, int cols, long long rows, int ld, int box_c, int box_r) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(IDIFF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IDIFF_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed (%d)", (int)r);
  return IDIFF_OK;
}

static unsigned long long* g_prof_dev = nullptr;
unsigned long long* prof_buffer() {                 // shared with conv3_rowpair.cu (profiling builds)
  if (!g_prof_dev && cudaMalloc(&g_prof_dev, 16 * sizeof(unsigned long long)) != cudaSuccess) g_prof_dev = nullptr;
  return g_prof_dev;
}

}  // namespace idiff

extern "C" {

int idiff_debug_read_prof(unsigned long long* out16_host) {
  using namespace idiff;
  IDIFF_REQUIRE(out16_host, "debug_read_prof: null");
  IDIFF_REQUIRE(g_prof_dev, "debug_read_prof: no profiled launch yet (build with -DIDIFF_PROF, params.reserved0 = 1)");
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out16_host, g_prof_dev, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost);
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

int idiff_conv_gemm_gn_rows(int H, int W) {
  return 4 * ((H + idiff::TILE_H - 1) / idiff::TILE_H) * ((W + idiff::TILE_W - 1) / idiff::TILE_W);
}

int idiff_conv_gemm(const idiff_gemm_params* pp, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(pp, "conv_gemm: null params");
  int rc = validate(*pp);
  if (rc) return rc;
  KArgs a;
  a.p = *pp;
  const SmemPlan s = plan_smem(*pp);
  a.SA = s.SA; a.SB = s.SB; a.resident = s.resident; a.offA = s.offA; a.offB = s.offB; a.offP = s.offP;
  a.offImg = s.offImg; a.offOut = s.offOut; a.offRes = s.offRes; a.nbuf_out = s.nbuf_out; a.res_stride = s.res_stride; a.res_nbuf = s.res_nbuf;
  a.tiles_x = (pp->W + TILE_W - 1) / TILE_W;
  a.tiles_y = (pp->H + TILE_H - 1) / TILE_H;
  a.ntiles_n = pp->N / pp->NT;
  a.total_items = a.tiles_x * a.tiles_y * pp->B * a.ntiles_n;
  a.prof = nullptr;
  IDIFF_REQUIRE(gn_fuse_make(a.gf, pp->gn_groups > 0 ? pp->gn_fuse : nullptr, pp->B, pp->N, 8),
                "conv_gemm: incomplete gn_fuse description");
  IDIFF_REQUIRE(s.total <= kSmemLimit && s.SA >= 1, "conv_gemm: shared memory plan %d B too large", s.total);
#ifdef IDIFF_PROF
  if (pp->reserved0) {
    if (!g_prof_dev && cudaMalloc(&g_prof_dev, 16 * sizeof(unsigned long long)) != cudaSuccess)
      return fail(IDIFF_ERR_CUDA, "conv_gemm: profile buffer allocation failed");
    a.prof = g_prof_dev;
  }
#endif
  if (pp->epi != IDIFF_EPI_GEGLU) {
    rc = make_map(&a.tm_out, pp->out, pp->N, pp->out_ld, pp->B, pp->H, pp->W);
    if (rc) return rc;
  }
  if (tma_residuals(*pp)) {
    if (pp->res0) rc = make_map(&a.tm_res0, pp->res0, pp->N, pp->N, pp->B, pp->H, pp->W);
    if (!rc && pp->res1) rc = make_map(&a.tm_res1, pp->res1, pp->N, pp->N, pp->B, pp->H, pp->W);
    if (rc) return rc;
  }
  static DeviceOnce once;
  int num_sms = 0;
  {
    cudaError_t e = per_device_setup(once, &num_sms, [] { return cudaSuccess; });
    if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv_gemm setup: %s", cudaGetErrorString(e));
  }
  // persistent: one CTA per SM (TMEM: 2*NT columns each), items strided by the grid size
  const int grid = a.total_items < num_sms ? a.total_items : num_sms;
  const int amode = pp->a_scale ? (pp->a_silu ? AMODE_AFFINE_SILU : AMODE_AFFINE) : AMODE_NONE;
  cudaError_t e = pp->ksize == 1   ? launch_conv_k1(a, amode, grid, s.total, as_stream(stream))
                  : pp->ksize == 3 ? launch_conv_k3(a, amode, grid, s.total, as_stream(stream))
                                   : launch_conv_k4(a, amode, grid, s.total, as_stream(stream));
  if (e == cudaErrorNotSupported) return fail(IDIFF_ERR_UNSUPPORTED, "conv_gemm: combination not built (k%d NT%d epi%d amode%d)", pp->ksize, pp->NT, pp->epi, amode);
  if (e != cudaSuccess) return fail(IDIFF_ERR_CUDA, "conv_gemm attr: %s", cudaGetErrorString(e));
  return check_launch("conv_gemm");
}

}  // extern "C"
