"""Elementwise / normalisation / attention kernels vs fp32 torch references of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16r, describe, no_tf32, rand_act, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    no_tf32()
    yield
    from instancediff_b200 import _lib
    _lib.watchdog()


def test_stem_conv7():
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 40, 24
    x, mu = torch.randn(B, 1, H, W, generator=g).cuda(), torch.randn(B, 1, H, W, generator=g).cuda()
    w = ((torch.rand(64, 2, 7, 7, generator=g) * 2 - 1) / 10).cuda()
    b = (torch.rand(64, generator=g) - 0.5).cuda()
    out = ops.stem_conv7(x, mu, w.permute(0, 2, 3, 1).contiguous(), b)
    ref = F.conv2d(torch.cat([x - mu, mu], 1), w, b, padding=3).permute(0, 2, 3, 1)
    assert rel_err(out, ref) <= 5e-3, describe(out, ref, "stem")


@pytest.mark.parametrize("shape", [(2, 40, 24), (1, 16, 8), (3, 64, 64), (1, 21, 13)])
def test_stem_conv7_tensor_core(shape):
    """tcgen05 stem (in-kernel im2col, bf16 hi/lo activations, bf16 weights, bias folded into the GEMM) vs the fp32
    convolution; ragged and single-tile shapes included.  Tolerance = bf16 rounding of weights and output."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, H, W = shape
    x, mu = torch.randn(B, 1, H, W, generator=g).cuda(), torch.randn(B, 1, H, W, generator=g).cuda()
    w = ((torch.rand(64, 2, 7, 7, generator=g) * 2 - 1) / 10).cuda()
    b = (torch.rand(64, generator=g) - 0.5).cuda()
    out = ops.stem_conv7_tc(x, mu, w, b)
    ref = F.conv2d(torch.cat([x - mu, mu], 1), w, b, padding=3).permute(0, 2, 3, 1)
    assert rel_err(out, ref) <= 5e-3, describe(out, ref, "stem tc")
    # against the same bf16-rounded weights the error is the output rounding only (activations are hi+lo exact)
    ref_q = F.conv2d(torch.cat([x - mu, mu], 1), w.to(torch.bfloat16).float(), b, padding=3).permute(0, 2, 3, 1)
    assert rel_err(out, ref_q) <= 4e-3, describe(out, ref_q, "stem tc vs bf16 weights")   # bf16 half-ulp = 2^-8
    simt = ops.stem_conv7(x, mu, w.permute(0, 2, 3, 1).contiguous(), b)
    assert rel_err(out, simt) <= 8e-3          # two bf16-rounded results may differ by one ulp (2^-7)


@pytest.mark.parametrize("shape", [(2, 24, 40), (1, 16, 16), (1, 21, 13), (3, 64, 64)])
def test_head_conv3(shape):
    """Head 3x3 conv on mma.sync with hi/lo-split weights: fp32-level agreement with the fp32 convolution of the same
    bf16 activations; ragged and single-tile shapes included."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, H, W = shape
    src = rand_act(B, H, W, 64, g)
    w = ((torch.rand(1, 64, 3, 3, generator=g) * 2 - 1) / 24).cuda()
    out = ops.head_conv3(src, w[0].permute(1, 2, 0).contiguous(), 0.25)
    ref = F.conv2d(src.float().permute(0, 3, 1, 2), w, torch.tensor([0.25], device="cuda"), padding=1)
    assert rel_err(out, ref) <= 1e-4, describe(out, ref, "head")


@pytest.mark.parametrize("shared", [True, False])
def test_time_embedding(shared):
    from instancediff_b200 import ops
    from oracle.unet_oracle import SinusoidalPosEmb
    g = torch.Generator().manual_seed(2)
    nf, S, B = 64, 96, 3
    w1, b1 = (torch.randn(256, 64, generator=g) / 8).cuda(), torch.randn(256, generator=g).cuda() * 0.1
    w2, b2 = (torch.randn(256, 256, generator=g) / 16).cuda(), torch.randn(256, generator=g).cuda() * 0.1
    wss, bss = (torch.randn(S, 256, generator=g) / 16).cuda(), torch.randn(S, generator=g).cuda() * 0.1
    if shared:
        t, Bt, tv = None, 1, torch.tensor([37.0], device="cuda")
        _, out = ops.time_embed(None, 37.0, w1.t().contiguous(), b1, w2.t().contiguous(), b2, wss, bss, 1, nf)
    else:
        tv = torch.tensor([1.0, 50.0, 100.0], device="cuda")
        _, out = ops.time_embed(tv, 0.0, w1.t().contiguous(), b1, w2.t().contiguous(), b2, wss, bss, B, nf)
    temb = F.linear(F.gelu(F.linear(SinusoidalPosEmb(64)(tv), w1, b1)), w2, b2)
    ref = F.linear(F.silu(temb), wss, bss)
    assert rel_err(out, ref) <= 1e-4, describe(out, ref, "time_embed")


@pytest.mark.parametrize("C,G", [(64, 8), (128, 8), (256, 8), (256, 32)])
def test_gn_stats_and_finalize(C, G):
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 20, 12
    x = rand_act(B, H, W, C, g)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.rand(C, generator=g) - 0.5).cuda()
    ts, tb = (torch.rand(B, 2 * C, generator=g) - 0.5).cuda(), None
    part = ops.gn_stats(x, G)
    sc, sh = ops.gn_finalize(part, gamma, beta, H * W * (C // G), 1e-6)
    y = x.float() * sc[:, None, None, :] + sh[:, None, None, :]
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), G, gamma, beta, 1e-6).permute(0, 2, 3, 1)
    assert rel_err(y, ref) <= 1e-4, describe(y, ref, "gn")
    # with time modulation: GN(x)*(1+scale)+shift ; row layout [B][2C] = (scale | shift), t_ld = 2C
    ss = ts.contiguous()
    sc2, sh2 = ops.gn_finalize(part, gamma, beta, H * W * (C // G), 1e-6, t_scale=ss, t_shift=ss[:, C:], t_ld=2 * C)
    y2 = x.float() * sc2[:, None, None, :] + sh2[:, None, None, :]
    ref2 = ref * (1 + ss[:, None, None, :C]) + ss[:, None, None, C:]
    assert rel_err(y2, ref2) <= 1e-4, describe(y2, ref2, "gn+time")


@pytest.mark.parametrize("C", [64, 128, 256])
def test_block_tail_add_rows_chan_ln(C):
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 12, 20
    y, res = rand_act(B, H, W, C, g), rand_act(B, H, W, C, g)
    sc, sh = (torch.rand(B, C, generator=g) + 0.5).cuda(), (torch.rand(B, C, generator=g) - 0.5).cuda()
    out, st = ops.block_tail(y, sc, sh, res, want_stats=True)
    ref = F.silu(y.float() * sc[:, None, None, :] + sh[:, None, None, :]) + res.float()
    assert rel_err(out, ref) <= 8e-3, describe(out, ref, "block_tail")
    rf = ref.reshape(-1, C)
    assert torch.allclose(st[:, 0], rf.mean(1), atol=5e-3)
    assert torch.allclose(st[:, 1], torch.rsqrt(rf.var(1, unbiased=False) + 1e-5), rtol=2e-2)
    o2, _ = ops.add_rows(y, res)
    assert rel_err(o2, y.float() + res.float()) <= 8e-3
    gain = (torch.rand(C, generator=g) + 0.5).cuda()
    o3 = ops.chan_ln(y, gain)
    yf = y.float()
    ref3 = (yf - yf.mean(-1, keepdim=True)) / (yf.var(-1, unbiased=False, keepdim=True) + 1e-5).sqrt() * gain
    assert rel_err(o3, ref3) <= 8e-3, describe(o3, ref3, "chan_ln")


@pytest.mark.parametrize("C,HW", [(64, (24, 24)), (128, (16, 8)), (256, (72, 64))])
def test_linear_attention_context(C, HW):
    """Weff[b] = Wout * blockdiag(ctx^T) with ctx = softmax_n(k) v^T / HW  (oracle LinearAttention)."""
    from instancediff_b200 import ops
    from instancediff_b200.packing import unpack_conv_weight
    g = torch.Generator().manual_seed(5)
    B, (H, W) = 2, HW
    qkv = rand_act(B, H, W, 384, g, scale=2.0)
    w_out = ((torch.rand(C, 128, generator=g) * 2 - 1) / 8).cuda()
    weff = ops.linattn_context(qkv, w_out)
    torch.cuda.synchronize()
    k = qkv[..., 128:256].float().reshape(B, H * W, 4, 32).permute(0, 2, 3, 1)      # b h d n
    v = qkv[..., 256:].float().reshape(B, H * W, 4, 32).permute(0, 2, 3, 1) / (H * W)
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(-1), v)
    ref = torch.einsum("che,bhde->bchd", w_out.reshape(C, 4, 32), ctx).reshape(B, C, 128)
    got = torch.stack([unpack_conv_weight(weff[b], C, 128, 1, C)[:, :, 0, 0] for b in range(B)])
    assert rel_err(got, ref) <= 1e-2, describe(got, ref, "weff")


@pytest.mark.parametrize("L", [16, 128, 200, 1024])
def test_self_attention(L):
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(6)
    B, heads = 2, 8
    C = heads * 32
    qkv = (torch.randn(B, L, 3 * C, generator=g)).cuda().to(torch.bfloat16)
    out = ops.self_attention(qkv, heads, 32 ** -0.5)
    torch.cuda.synchronize()
    q, k, v = (t.float().reshape(B, L, heads, 32).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, L, C)
    err = rel_err(out, ref)
    if err > 1.5e-2:
        from instancediff_b200 import _lib
        msgs = [describe(out, ref, "primary")]
        for flags in (2, 4, 6):
            _lib.lib().idiff_set_debug_flags(flags)
            o2 = ops.self_attention(qkv, heads, 32 ** -0.5)
            torch.cuda.synchronize()
            msgs.append(f"debug flags {flags}: rel_err={rel_err(o2, ref):.4g}")
        _lib.lib().idiff_set_debug_flags(0)
        pytest.fail("\n".join(msgs))


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 64, 64, 64), (2, 32, 32, 128), (1, 128, 128, 64)])
def test_linattn_fused_matches_reference(shape):
    """Three-pass fused linear attention (q, k, v never in HBM) vs the fp32 module math of SURVEY App. A:
    x + LN_c(to_out(ctx^T q)) with q = softmax_d * scale, k = softmax_n, v / HW, x^ = ChanLayerNorm(x) * g."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, H, W, Cc = shape
    x = rand_act(B, H, W, Cc, g)
    xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    stats = torch.cat([mean, rstd], -1).reshape(-1, 2).contiguous()
    g_pre = (torch.rand(Cc, generator=g) + 0.5).cuda()
    wqkv = ((torch.rand(384, Cc, generator=g) * 2 - 1) * (2.0 / Cc ** 0.5)).cuda()
    w_out = ((torch.rand(Cc, 128, generator=g) * 2 - 1) / 128 ** 0.5).cuda()
    b_out = (torch.rand(Cc, generator=g) - 0.5).cuda()
    g_out = (torch.rand(Cc, generator=g) + 0.5).cuda()
    out = ops.linattn_fused(x, stats, wqkv, g_pre, w_out, b_out, g_out)
    torch.cuda.synchronize()
    xn = (xf - mean) * rstd * g_pre
    qkv = xn.reshape(B, H * W, Cc) @ wqkv.t()                          # [B, N, 384]
    q, k, v = (t.reshape(B, H * W, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=-1))   # [B, h, d, N]
    q = q.softmax(dim=-2) * 32 ** -0.5
    k = k.softmax(dim=-1)
    v = v / (H * W)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, 128, H * W).permute(0, 2, 1)       # [B, N, 128]
    y = o @ w_out.t() + b_out
    y = (y - y.mean(-1, keepdim=True)) * torch.rsqrt(y.var(-1, unbiased=False, keepdim=True) + 1e-5) * g_out
    ref = xf + y.reshape(B, H, W, Cc)
    assert rel_err(out, ref) <= 1e-2, describe(out, ref, "linattn fused")
    # the attention branch alone (without the dominating residual) must match as well
    assert rel_err(out.float() - xf, ref - xf) <= 3e-2, describe(out.float() - xf, ref - xf, "linattn branch")


@pytest.mark.parametrize("shape", [(2, 64, 64, 64), (1, 32, 64, 128), (3, 16, 8, 64)])
def test_linattn_pass_versions_agree(shape, monkeypatch):
    """Version 2 of the two big passes (one pipelined CTA per SM) against version 1 (two lock-step CTAs per SM): the
    context pass does the same arithmetic in the same order (bit-identical output), the output pass folds the
    LayerNorm into the softmax instead of normalising x on CUDA cores (equal up to bf16 rounding of the operand)."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(31)
    B, H, W, Cc = shape
    x = rand_act(B, H, W, Cc, g)
    xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    stats = torch.cat([mean, torch.rsqrt(var + 1e-5)], -1).reshape(-1, 2).contiguous()
    args = (x, stats, ((torch.rand(384, Cc, generator=g) * 2 - 1) * (2.0 / Cc ** 0.5)).cuda(), (torch.rand(Cc, generator=g) + 0.5).cuda(),
            ((torch.rand(Cc, 128, generator=g) * 2 - 1) / 128 ** 0.5).cuda(), (torch.rand(Cc, generator=g) - 0.5).cuda(),
            (torch.rand(Cc, generator=g) + 0.5).cuda())
    outs = {}
    for ctx_v1, out_v1 in ((0, 0), (1, 0), (1, 1), (0, 1)):
        monkeypatch.setenv("IDIFF_LA_CTX_V2", str(1 - ctx_v1))
        monkeypatch.setenv("IDIFF_LA_OUT_V1", str(out_v1))
        outs[(ctx_v1, out_v1)] = ops.linattn_fused(*args).clone()
    torch.cuda.synchronize()
    assert torch.equal(outs[(0, 0)], outs[(1, 0)]) and torch.equal(outs[(0, 1)], outs[(1, 1)]), "context pass v2 != v1"
    assert rel_err(outs[(0, 0)].float() - xf, outs[(0, 1)].float() - xf) <= 2e-2, "output pass v2 vs v1"


def test_linattn_fused_reference_far_below_the_maximum():
    """The context pass exponentiates against the channel maximum of each chunk's FIRST tile (no k-max pre-pass).
    Adversarial input: every pixel of the first tile is anti-aligned with the k weights and a late pixel is aligned,
    so the true maximum exceeds the reference by ~40 nats; the rescaled merge must still reproduce softmax_n."""
    from instancediff_b200 import ops
    g = torch.Generator().manual_seed(23)
    B, H, W, Cc = 1, 32, 32, 64
    wqkv = ((torch.rand(384, Cc, generator=g) * 2 - 1) * (2.0 / Cc ** 0.5)).cuda()
    wqkv[128:256] *= 2.5                                               # large k logits
    wdir = wqkv[128:256].mean(0)                                       # a direction most k channels like
    wdir = wdir - wdir.mean()
    wdir = wdir / wdir.norm() * Cc ** 0.5
    x = rand_act(B, H, W, Cc, g).float() * 0.05
    x.reshape(-1, Cc)[:128] -= wdir                                    # first tile: anti-aligned
    x.reshape(-1, Cc)[700] += 4 * wdir                                 # one late pixel: strongly aligned
    x = x.to(torch.bfloat16)
    xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    stats = torch.cat([mean, rstd], -1).reshape(-1, 2).contiguous()
    ones, zeros = torch.ones(Cc).cuda(), torch.zeros(Cc).cuda()
    w_out = ((torch.rand(Cc, 128, generator=g) * 2 - 1) * 300.0).cuda()   # large: the channel LayerNorm sees var >> eps
    out = ops.linattn_fused(x, stats, wqkv, ones, w_out, zeros, ones)
    torch.cuda.synchronize()
    xn = ((xf - mean) * rstd).to(torch.bfloat16).float()               # the operand the kernel sees
    qkv = xn.reshape(B, H * W, Cc) @ wqkv.to(torch.bfloat16).float().t()
    q, k, v = (t.reshape(B, H * W, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=-1))
    assert (k.amax(-1) - k[..., :128].amax(-1)).max() > 20             # the reference really is far below the max
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(dim=-1), v / (H * W))
    o = torch.einsum("bhde,bhdn->bhen", ctx, q.softmax(dim=-2) * 32 ** -0.5).reshape(B, 128, H * W).permute(0, 2, 1)
    y = o @ w_out.t()
    assert y.var(-1, unbiased=False).median() > 1e-3
    y = (y - y.mean(-1, keepdim=True)) * torch.rsqrt(y.var(-1, unbiased=False, keepdim=True) + 1e-5)
    ref = xf + y.reshape(B, H, W, Cc)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) <= 1e-2, describe(out, ref, "linattn fused, adversarial reference")
    # the attention branch alone on the undoctored pixels (|x| ~ 0.05 there, so the bf16 rounding of x + branch does
    # not mask it): rows 0-3 hold the anti-aligned tile, row 21 the aligned pixel
    keep = [r for r in range(H) if r >= 4 and r != 700 // W]
    br, br_ref = (out.float() - xf)[:, keep], (ref - xf)[:, keep]
    assert br_ref.abs().max() > 0.5
    assert rel_err(br, br_ref) <= 3e-2, describe(br, br_ref, "branch on undoctored rows")
