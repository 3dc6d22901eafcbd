"""tcgen05 implicit-GEMM engine (idiff_conv_gemm) vs an fp32 torch reference of the same op.

Tolerance: inputs/weights are bf16 on both sides, accumulation fp32, output stored as bf16, so the
bound is the bf16 rounding of the output: max |err| <= 1e-2 * max|ref| (north-star tolerance).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16r, conv_reference, describe, no_tf32, rand_act, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    no_tf32()
    yield
    from instancediff_b200 import _lib
    _lib.watchdog()


def _run(B, H, W, cin0, cin1, N, k, stride=1, up=0, affine=False, silu=False, NT=None, seed=0, swap=0, **extra):
    from instancediff_b200 import ops
    from instancediff_b200.packing import pack_conv_weight
    g = torch.Generator().manual_seed(seed)
    Hs, Ws = (H * stride, W * stride) if not up else (H // 2, W // 2)
    src0 = rand_act(B, Hs, Ws, cin0, g)
    src1 = rand_act(B, Hs, Ws, cin1, g) if cin1 else None
    cin = cin0 + cin1
    w = (torch.rand(N, cin, k, k, generator=g) * 2 - 1).cuda() / math.sqrt(cin * k * k)
    bias = (torch.rand(N, generator=g) * 2 - 1).cuda() * 0.1
    sc = sh = None
    if affine:
        sc = (torch.rand(B, cin, generator=g) + 0.5).cuda()
        sh = (torch.rand(B, cin, generator=g) - 0.5).cuda()
    NT = min(N, 256) if NT is None else NT
    out = torch.zeros(B, H, W, N, dtype=torch.bfloat16, device="cuda")
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=k, stride=stride, cin0=cin0, cin1=cin1, up0=up, N=N, NT=NT,
                             a_silu=int(silu), epi=0, out_ld=N, src0=src0, src1=src1, a_scale=sc, a_shift=sh,
                             w=pack_conv_weight(w, NT).cuda(), bias=bias, out=out, dbg_swap_lbo_sbo=swap, **extra)
    ops.conv_gemm(p)
    torch.cuda.synchronize()
    ref = conv_reference(src0, src1, w, bias, k, stride, up, sc, sh, silu)
    return out, ref, dict(src0=src0, src1=src1, w=w, bias=bias, sc=sc, sh=sh, p=p)


def test_probe_descriptor_convention():
    """Bring-up probe: the simplest GEMM (1x1, 64->64, one tile).  If the documented LBO/SBO roles were
    wrong the swapped variant would match instead -- the message says which."""
    out, ref, _ = _run(1, 16, 8, 64, 0, 64, 1)
    e0 = rel_err(out, ref)
    if e0 > TOL:
        out2, ref2, _ = _run(1, 16, 8, 64, 0, 64, 1, swap=1)
        e1 = rel_err(out2, ref2)
        pytest.fail(f"primary descriptor convention rel_err={e0:.4g}; swapped LBO/SBO rel_err={e1:.4g}\n"
                    + describe(out, ref, "primary") + "\n" + describe(out2, ref2, "swapped"))


CASES = [
    # B, H, W, cin0, cin1, N, k, stride, up, affine, silu
    (1, 16, 8, 64, 0, 64, 1, 1, 0, False, False),
    (2, 32, 32, 64, 0, 64, 3, 1, 0, False, False),
    (2, 32, 32, 64, 0, 64, 3, 1, 0, True, True),
    (1, 32, 16, 128, 0, 128, 3, 1, 0, True, True),
    (1, 16, 16, 256, 0, 256, 3, 1, 0, False, False),
    (2, 32, 32, 64, 64, 64, 3, 1, 0, False, False),       # skip-concat
    (1, 16, 16, 256, 128, 256, 3, 1, 0, False, False),    # 384 -> 256
    (1, 16, 16, 256, 256, 256, 3, 1, 0, False, False),    # 512 -> 256, K = 4608
    (1, 32, 32, 128, 64, 128, 1, 1, 0, False, False),     # 1x1 shortcut on a concat
    (2, 16, 16, 64, 0, 64, 4, 2, 0, False, False),        # 4x4 stride-2 downsample
    (1, 16, 16, 64, 0, 128, 4, 2, 0, False, False),
    (1, 8, 8, 128, 0, 256, 4, 2, 0, False, False),
    (1, 32, 32, 256, 0, 128, 3, 1, 1, False, False),      # nearest x2 upsample + 3x3
    (1, 24, 20, 64, 0, 64, 3, 1, 0, True, True),          # ragged tiles (H % 16, W % 8 != 0)
    (1, 4, 4, 256, 0, 256, 3, 1, 0, False, False),        # image smaller than one tile
    (1, 32, 32, 64, 0, 384, 1, 1, 0, False, False),       # N split over 3 tiles of 128
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_matches_fp32_reference(case):
    B, H, W, c0, c1, N, k, s, up, aff, silu = case
    NT = 128 if N == 384 else None
    out, ref, _ = _run(B, H, W, c0, c1, N, k, s, up, aff, silu, NT=NT)
    assert rel_err(out, ref) <= TOL, describe(out, ref, str(case))


def _run_rowpair(B, H, W, affine, seed=0, gn=False, packed=False, gn_fuse=None, src=None, gen=None):
    """idiff_conv3_rowpair (full-width MMA formulation of the 3x3 64 -> 64 layers) on the same contract."""
    from instancediff_b200 import _lib, ops
    from instancediff_b200.packing import pack_conv3_rowpair
    g = torch.Generator().manual_seed(seed) if gen is None else gen
    src0 = rand_act(B, H, W, 64, g) if src is None else src
    w = (torch.rand(64, 64, 3, 3, generator=g) * 2 - 1).cuda() / math.sqrt(64 * 9)
    bias = (torch.rand(64, generator=g) * 2 - 1).cuda() * 0.1
    sc = sh = None
    if affine:
        sc = (torch.rand(B, 64, generator=g) + 0.5).cuda()
        sh = (torch.rand(B, 64, generator=g) - 0.5).cuda()
    out = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    rows = _lib.lib().idiff_conv3_rowpair_gn_rows(H, W)
    part = torch.full((B, rows, 8, 2), float("nan"), device="cuda") if gn else None
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=3, stride=1, cin0=64, N=64, NT=64, a_silu=int(affine) * (3 if packed else 1),
                             epi=0, out_ld=64, src0=src0, a_scale=sc, a_shift=sh, w=pack_conv3_rowpair(w.cpu()).cuda(), bias=bias,
                             out=out, gn_groups=8 if (gn or gn_fuse is not None) else 0, gn_partial=part,
                             **({"gn_fuse": gn_fuse} if gn_fuse is not None else {}))
    assert _lib.lib().idiff_conv3_rowpair_supported(p)
    ops.conv3_rowpair(p)
    torch.cuda.synchronize()
    ref = conv_reference(src0, None, w, bias, 3, 1, 0, sc, sh, affine)
    return out, ref, part


ROWPAIR_CASES = [
    # B, H, W, affine+SiLU
    (1, 2, 128, False),          # one item: both row pairs are the image border
    (2, 32, 128, False),
    (1, 16, 256, True),          # two strips per row
    (1, 8, 40, True),            # image narrower than a strip (clipped TMA store, zero-filled loads)
    (1, 24, 200, True),          # ragged second strip (72 of 128 pixels)
    (3, 64, 256, False),         # 192 items: every CTA of the grid has work
    (4, 128, 256, True),         # 512 items on 148 CTAs: item ranges start and end in the middle of a strip
    (2, 30, 144, False),         # H/2 odd, ragged strip
]


@pytest.mark.parametrize("case", ROWPAIR_CASES, ids=lambda c: "x".join(map(str, c)))
def test_rowpair_conv_matches_fp32_reference(case):
    B, H, W, aff = case
    out, ref, _ = _run_rowpair(B, H, W, aff, seed=H + W)
    assert not torch.isnan(out.float()).any(), "unwritten output pixels"
    assert rel_err(out, ref) <= TOL, describe(out, ref, "rowpair " + str(case))


def test_rowpair_packed_bf16x2_transform_stays_inside_the_tolerance():
    """a_silu = 3: affine + SiLU of the loader in packed bf16x2 arithmetic (scale / shift rounded to bf16, tanh in bf16).
    The A operand is bf16 on both paths; the packed evaluation adds two bf16 roundings before it."""
    for case in [(1, 16, 256), (2, 64, 128), (1, 24, 200)]:
        out, ref, _ = _run_rowpair(*case, True, seed=3, packed=True)
        exact, _, _ = _run_rowpair(*case, True, seed=3, packed=False)
        assert rel_err(out, ref) <= TOL, describe(out, ref, "rowpair packed " + str(case))
        assert rel_err(out, ref) <= 3 * rel_err(exact, ref) + 2e-3


def test_rowpair_conv_equals_the_generic_engine_and_writes_groupnorm_partials():
    """Same layer through both kernels: identical bf16 inputs, fp32 accumulation in a different order -> equal up to
    the bf16 rounding of the output; GroupNorm partial sums (own row layout) finalize to torch's GroupNorm."""
    from instancediff_b200 import ops
    B, H, W = 2, 32, 256
    out, ref, part = _run_rowpair(B, H, W, True, seed=5, gn=True)
    gen, ref2, _ = _run(B, H, W, 64, 0, 64, 3, affine=True, silu=True, seed=5)
    assert not torch.isnan(part).any(), "unwritten GroupNorm partial rows"
    s = part.sum(dim=1)
    r = ref.reshape(B, H * W, 8, 8)
    assert torch.allclose(s[..., 0], r.sum(dim=(1, 3)), rtol=1e-3, atol=1e-2), (s[..., 0], r.sum(dim=(1, 3)))
    assert torch.allclose(s[..., 1], (r * r).sum(dim=(1, 3)), rtol=1e-3, atol=1e-2)
    gamma, beta = torch.rand(64, device="cuda") + 0.5, torch.rand(64, device="cuda") - 0.5
    sc, sh = ops.gn_finalize(part, gamma, beta, H * W * 8, 1e-5)
    y = ref * sc[:, None, None, :] + sh[:, None, None, :]
    gn = F.group_norm(ref.permute(0, 3, 1, 2), 8, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    assert rel_err(y, gn) <= 1e-3, describe(y, gn, "gn")
    assert rel_err(out, ref) <= TOL, describe(out, ref, "rowpair")


def test_rowpair_rejects_unsupported_layers():
    from instancediff_b200 import _lib, ops
    x = torch.zeros(1, 16, 128, 64, dtype=torch.bfloat16, device="cuda")
    ok = dict(B=1, H=16, W=128, ksize=3, stride=1, cin0=64, N=64, NT=64, out_ld=64, src0=x, w=x, out=x)
    for bad in (dict(cin1=64, src1=x), dict(N=128, NT=128, out_ld=128), dict(ksize=1), dict(res0=x), dict(H=15), dict(up0=1)):
        p = ops.make_gemm_params(**{**ok, **bad})
        assert not _lib.lib().idiff_conv3_rowpair_supported(p)
        with pytest.raises(_lib.IdiffError):
            ops.conv3_rowpair(p)


def test_cuda_core_reference_agrees():
    """idiff_conv_ref (plain CUDA-core loop nest) is an independent check of the same params struct."""
    from instancediff_b200 import ops
    out, ref, d = _run(1, 16, 16, 64, 64, 64, 3, affine=True, silu=True)
    o32 = torch.empty(1, 16, 16, 64, dtype=torch.float32, device="cuda")
    w_nhwc = bf16r(d["w"]).permute(0, 2, 3, 1).contiguous()
    ops.conv_ref(d["p"], w_nhwc, o32)
    torch.cuda.synchronize()
    assert rel_err(o32, ref) <= 2e-3, describe(o32, ref, "conv_ref")
    assert rel_err(out, o32) <= TOL, describe(out, o32, "tc vs conv_ref")


def test_groupnorm_partials():
    from instancediff_b200 import ops
    B, H, W, N = 2, 32, 24, 128
    from instancediff_b200 import _lib
    rows = _lib.lib().idiff_conv_gemm_gn_rows(H, W)
    assert rows == 4 * ((H + 15) // 16) * ((W + 7) // 8)
    part = torch.zeros(B, rows, 8, 2, device="cuda")
    out, ref, _ = _run(B, H, W, 64, 0, N, 3, gn_groups=8, gn_partial=part)
    s = part.sum(dim=1)                                       # [B, 8, 2]
    r = ref.reshape(B, H * W, 8, N // 8)
    assert torch.allclose(s[..., 0], r.sum(dim=(1, 3)), rtol=1e-3, atol=1e-2), (s[..., 0], r.sum(dim=(1, 3)))
    assert torch.allclose(s[..., 1], (r * r).sum(dim=(1, 3)), rtol=1e-3, atol=1e-2)
    # finalize -> per-channel affine equals torch GroupNorm
    gamma, beta = torch.rand(N, device="cuda") + 0.5, torch.rand(N, device="cuda") - 0.5
    sc, sh = ops.gn_finalize(part, gamma, beta, H * W * (N // 8), 1e-5)
    y = ref * sc[:, None, None, :] + sh[:, None, None, :]
    gn = F.group_norm(ref.permute(0, 3, 1, 2), 8, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    assert rel_err(y, gn) <= 1e-3, describe(y, gn, "gn")


@pytest.mark.parametrize("N,NT", [(256, 128), (256, 64), (128, 64)])
def test_groupnorm_partials_with_a_split_n_tile(N, NT):
    """NT < N (small grids split a wide N tile): every N tile fills the entries of its own groups; output and partial
    sums are BIT-identical to the single-tile run (same K sequence per element, same per-group reduction)."""
    from instancediff_b200 import _lib
    B, H, W = 2, 32, 24
    rows = _lib.lib().idiff_conv_gemm_gn_rows(H, W)
    part_a = torch.full((B, rows, 8, 2), float("nan"), device="cuda")
    part_b = torch.full((B, rows, 8, 2), float("nan"), device="cuda")
    out_a, ref, _ = _run(B, H, W, 128, 0, N, 3, seed=4, gn_groups=8, gn_partial=part_a)
    out_b, _, _ = _run(B, H, W, 128, 0, N, 3, seed=4, NT=NT, gn_groups=8, gn_partial=part_b)
    assert torch.equal(out_a, out_b) and rel_err(out_b, ref) <= TOL
    assert not torch.isnan(part_b).any() and torch.equal(part_a, part_b)
    r = ref.reshape(B, H * W, 8, N // 8)
    assert torch.allclose(part_b.sum(dim=1)[..., 0], r.sum(dim=(1, 3)), rtol=1e-3, atol=1e-2)


def _fuse_for(B, N, H, W, seed, with_time):
    import ctypes
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(seed)
    gamma, beta = (torch.rand(N, generator=g) + 0.5).cuda(), (torch.rand(N, generator=g) - 0.5).cuda()
    ts = tb = None
    if with_time:
        ts, tb = (torch.rand(B, N, generator=g) - 0.5).cuda(), (torch.rand(B, N, generator=g) - 0.5).cuda()
    f, sc, sh, state = ops.make_gn_fuse(B, gamma, beta, H * W * (N // 8), 1e-5, t_scale=ts, t_shift=tb, t_ld=N if with_time else 0)
    return f, ctypes.pointer(f), sc, sh, state, gamma, beta, ts, tb


@pytest.mark.parametrize("N,NT,H,W,with_time", [(64, 64, 32, 24, True), (128, 128, 32, 24, False), (256, 128, 16, 40, True),
                                                (256, 256, 32, 32, False)])
def test_groupnorm_finalize_folded_into_the_conv(N, NT, H, W, with_time):
    """gn_fuse: exact integer sums + last-CTA finalize == partial rows + idiff_gn_finalize (up to fp32 rounding of the old
    path's sums), equals torch's GroupNorm, and leaves its state zeroed for the next launch."""
    from instancediff_b200 import _lib, ops
    B = 3
    f, fp, sc, sh, state, gamma, beta, ts, tb = _fuse_for(B, N, H, W, 21, with_time)
    out, ref, _ = _run(B, H, W, 128, 0, N, 3, seed=4, NT=NT, gn_groups=8, gn_fuse=fp)
    rows = _lib.lib().idiff_conv_gemm_gn_rows(H, W)
    part = torch.zeros(B, rows, 8, 2, device="cuda")
    out2, _, _ = _run(B, H, W, 128, 0, N, 3, seed=4, NT=NT, gn_groups=8, gn_partial=part)
    sc2, sh2 = ops.gn_finalize(part, gamma, beta, H * W * (N // 8), 1e-5, t_scale=ts, t_shift=tb, t_ld=N if with_time else 0)
    assert torch.equal(out, out2)
    assert not torch.isnan(sc).any() and not torch.isnan(sh).any(), "affine rows not written"
    assert torch.allclose(sc, sc2, rtol=2e-5, atol=1e-6) and torch.allclose(sh, sh2, rtol=2e-5, atol=2e-6), \
        ((sc - sc2).abs().max(), (sh - sh2).abs().max())
    assert int(state[0].abs().max()) == 0 and int(state[1][0]) == 0, "gn_fuse state not left zeroed"
    gn = F.group_norm(ref.permute(0, 3, 1, 2), 8, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    if with_time:
        gn = gn * (1 + ts[:, None, None, :]) + tb[:, None, None, :]
    y = ref * sc[:, None, None, :] + sh[:, None, None, :]
    assert rel_err(y, gn) <= 1e-3, describe(y, gn, "gn fused")
    # second launch on the same state (what a plan does layer after layer), and a batch shard: rows are per image
    sc_first = sc.clone()
    out3, _, _ = _run(B, H, W, 128, 0, N, 3, seed=4, NT=NT, gn_groups=8, gn_fuse=fp)
    assert torch.equal(sc, sc_first) and int(state[0].abs().max()) == 0


def test_groupnorm_finalize_folded_is_invariant_to_the_batch_shard():
    """exact integer sums: an image's affine rows do not depend on the batch it is processed in (sharding invariance)"""
    from instancediff_b200 import ops
    import ctypes
    N, H, W = 64, 32, 128
    g = torch.Generator().manual_seed(2)
    gamma, beta = (torch.rand(N, generator=g) + 0.5).cuda(), (torch.rand(N, generator=g) - 0.5).cuda()
    res = {}
    src4 = rand_act(4, H, W, 64, g)
    w_state = g.get_state()
    for B in (4, 2):
        g.set_state(w_state)                                  # same weights and bias for both batch sizes
        f, sc, sh, _ = ops.make_gn_fuse(B, gamma, beta, H * W * 8, 1e-5)
        _, ref, _ = _run_rowpair(B, H, W, False, gn_fuse=ctypes.pointer(f), src=src4[:B].contiguous(), gen=g)
        res[B] = (sc.clone(), sh.clone(), ref)
        assert not torch.isnan(sc).any()
    assert torch.equal(res[4][0][:2], res[2][0]) and torch.equal(res[4][1][:2], res[2][1])
    gn = F.group_norm(res[2][2].permute(0, 3, 1, 2), 8, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    y = res[2][2] * res[2][0][:, None, None, :] + res[2][1][:, None, None, :]
    assert rel_err(y, gn) <= 1e-3


@pytest.mark.parametrize("B,H,W,Cc", [(3, 32, 32, 256), (40, 8, 8, 256), (2, 16, 16, 64), (2, 16, 24, 128)])
def test_chan_ln_gn_equals_the_three_launch_chain(B, H, W, Cc):
    """idiff_chan_ln_gn == idiff_chan_ln -> idiff_gn_stats -> idiff_gn_finalize (SpatialTransformer entry); B = 40 has more
    (image, group) pairs than one pass of the finishing CTA's scratch holds"""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(8)
    G = min(32, Cc // 8)                                      # groups of at least 8 channels
    x = rand_act(B, H, W, Cc, g)
    gain = (torch.rand(Cc, generator=g) + 0.5).cuda()
    gamma, beta = (torch.rand(Cc, generator=g) + 0.5).cuda(), (torch.rand(Cc, generator=g) - 0.5).cuda()
    f, sc, sh, state = ops.make_gn_fuse(B, gamma, beta, H * W * (Cc // G), 1e-6, G=G)
    y = ops.chan_ln_gn(x, gain, f, G)
    y_ref = ops.chan_ln(x, gain)
    part = ops.gn_stats(y_ref, G)
    sc2, sh2 = ops.gn_finalize(part, gamma, beta, H * W * (Cc // G), 1e-6)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    assert torch.allclose(sc, sc2, rtol=2e-5, atol=1e-6) and torch.allclose(sh, sh2, rtol=2e-5, atol=2e-6)
    assert int(state[0].abs().max()) == 0 and int(state[1][0]) == 0


def test_layernorm_fold_and_row_stats():
    """acc' = (acc - mean*wsum)*rstd (+bias) equals Linear(LayerNorm(x)); out_row_stats are the LN stats."""
    from instancediff_b200 import ops
    from instancediff_b200.packing import fold_layernorm, pack_conv_weight
    g = torch.Generator().manual_seed(3)
    B, H, W, Cc, N = 2, 16, 16, 256, 256
    x = rand_act(B, H, W, Cc, g)
    gain, beta = torch.rand(Cc, generator=g).cuda() + 0.5, torch.rand(Cc, generator=g).cuda() - 0.5
    w = ((torch.rand(N, Cc, generator=g) * 2 - 1) / 16).cuda()
    xf = x.float().reshape(-1, Cc)
    mean, var = xf.mean(1), xf.var(1, unbiased=False)
    stats = torch.stack([mean, torch.rsqrt(var + 1e-5)], 1).contiguous()
    wf, wsum, extra = fold_layernorm(w, gain, beta)
    out = torch.zeros(B, H, W, N, dtype=torch.bfloat16, device="cuda")
    ostats = torch.zeros(B * H * W, 2, device="cuda")
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=1, stride=1, cin0=Cc, N=N, NT=256, out_ld=N, src0=x,
                             w=pack_conv_weight(wf, 256).cuda(), bias=extra.contiguous(), row_stats=stats, wsum=wsum.contiguous(),
                             out=out, out_row_stats=ostats, ln_eps=1e-5)
    ops.conv_gemm(p)
    torch.cuda.synchronize()
    ref = F.linear(F.layer_norm(xf, (Cc,), gain, beta, 1e-5), w).reshape(B, H, W, N)
    assert rel_err(out, ref) <= 1.5e-2, describe(out, ref, "ln-fold")
    rf = ref.reshape(-1, N)
    assert torch.allclose(ostats[:, 0], rf.mean(1), atol=2e-2)
    assert torch.allclose(ostats[:, 1], torch.rsqrt(rf.var(1, unbiased=False) + 1e-5), rtol=3e-2)


def test_epilogues_qsoftmax_geglu_lnout_residuals():
    from instancediff_b200 import ops
    from instancediff_b200.packing import interleave_geglu, pack_conv_weight
    g = torch.Generator().manual_seed(5)
    B, H, W, Cc = 1, 16, 16, 64
    x = rand_act(B, H, W, Cc, g)
    # q-softmax on columns [0,128)
    w = ((torch.rand(384, Cc, generator=g) * 2 - 1)).cuda()
    out = torch.zeros(B, H, W, 384, dtype=torch.bfloat16, device="cuda")
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=1, stride=1, cin0=Cc, N=384, NT=128, epi=1, out_ld=384, src0=x,
                             w=pack_conv_weight(w, 128).cuda(), out=out, qscale=32 ** -0.5)
    ops.conv_gemm(p)
    raw = F.linear(x.float(), bf16r(w))
    ref = raw.clone()
    ref[..., :128] = raw[..., :128].reshape(B, H, W, 4, 32).softmax(-1).reshape(B, H, W, 128) * 32 ** -0.5
    torch.cuda.synchronize()
    assert rel_err(out[..., :128], ref[..., :128]) <= TOL, describe(out[..., :128], ref[..., :128], "qsoftmax")
    assert rel_err(out[..., 128:], ref[..., 128:]) <= TOL, describe(out[..., 128:], ref[..., 128:], "kv")

    # GEGLU
    wf = ((torch.rand(512, Cc, generator=g) * 2 - 1) / 8).cuda()
    bf = (torch.rand(512, generator=g) - 0.5).cuda()
    wi, bi = interleave_geglu(wf, bf)
    out = torch.zeros(B, H, W, 256, dtype=torch.bfloat16, device="cuda")
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=1, stride=1, cin0=Cc, N=512, NT=256, epi=2, out_ld=256, src0=x,
                             w=pack_conv_weight(wi, 256).cuda(), bias=bi.contiguous(), out=out)
    ops.conv_gemm(p)
    a, gate = (F.linear(x.float(), bf16r(wf)) + bf).chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL, describe(out, ref, "geglu")

    # LN_OUT + residual, per-image weights, strided source (q columns of a [.,384] tensor)
    B = 2
    qkv = rand_act(B, H, W, 384, g)
    xres = rand_act(B, H, W, Cc, g)
    wimg = ((torch.rand(B, Cc, 128, generator=g) * 2 - 1) / 8).cuda()
    wp = torch.stack([pack_conv_weight(wimg[b], Cc) for b in range(B)]).cuda()
    bias = (torch.rand(Cc, generator=g) - 0.5).cuda()
    gain = (torch.rand(Cc, generator=g) + 0.5).cuda()
    out = torch.zeros(B, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=1, stride=1, cin0=128, src0_ld=384, N=Cc, NT=Cc, epi=3, out_ld=Cc,
                             src0=qkv, w=wp, w_image_stride=Cc * 128, bias=bias, ln_g=gain, res0=xres, out=out,
                             ln_eps=1e-5)
    ops.conv_gemm(p)
    y = torch.einsum("bhwk,bck->bhwc", qkv[..., :128].float(), bf16r(wimg)) + bias
    ref = (y - y.mean(-1, keepdim=True)) / (y.var(-1, unbiased=False, keepdim=True) + 1e-5).sqrt() * gain + xres.float()
    torch.cuda.synchronize()
    assert rel_err(out, ref) <= TOL, describe(out, ref, "ln_out")


@pytest.mark.parametrize("N,H,W", [(128, 16, 16), (64, 32, 24), (256, 8, 8)])
def test_resblock_tail_in_shortcut_epilogue_and_image_bias(N, H, W):
    """1x1 shortcut conv whose epilogue adds silu(GN(y2)) (res0 affine), a second residual and a per-image bias.
    N = 64 takes the TMA residual path (ragged tiles: zero-filled loads, clipped stores), N = 256 stages the
    per-image columns with a 128-thread epilogue group."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(7)
    B = 2
    y2 = rand_act(B, H, W, N, g)
    sc2 = (torch.rand(B, N, generator=g) + 0.5).cuda()
    sh2 = (torch.rand(B, N, generator=g) - 0.5).cuda()
    bimg = (torch.rand(B, N, generator=g) - 0.5).cuda()
    r1 = rand_act(B, H, W, N, g)
    out, ref, _ = _run(B, H, W, 128, 64, N, 1, seed=11, res0=y2, res0_scale=sc2, res0_shift=sh2, res1=r1, bias_img=bimg)
    ref = ref + F.silu(y2.float() * sc2[:, None, None, :] + sh2[:, None, None, :]) + r1.float() + bimg[:, None, None, :]
    assert rel_err(out, ref) <= TOL, describe(out, ref, "shortcut+tail")


def test_bad_arguments_return_errors():
    from instancediff_b200 import _lib, ops
    x = torch.zeros(1, 16, 8, 64, dtype=torch.bfloat16, device="cuda")
    p = ops.make_gemm_params(B=1, H=16, W=8, ksize=5, stride=1, cin0=64, N=64, NT=64, out_ld=64, src0=x, w=x, out=x)
    with pytest.raises(_lib.IdiffError):
        ops.conv_gemm(p)
    p = ops.make_gemm_params(B=1, H=16, W=8, ksize=1, stride=1, cin0=48, N=64, NT=64, out_ld=64, src0=x, w=x, out=x)
    with pytest.raises(_lib.IdiffError):
        ops.conv_gemm(p)
