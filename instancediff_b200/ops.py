"""Tensor-level wrappers over the C ABI (one function per kernel entry point of include/idiff.h).

Thin plumbing: pointer extraction, output allocation, error checks.  All tensors must be CUDA;
activations are channels-last bf16 ``[B,H,W,C]``.  Used by the tests and available as the low-level API.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GN_SLOTS, GemmParams, GnFuse, check

_PTR_FIELDS = {"src0", "src1", "a_scale", "a_shift", "w", "bias", "bias_img", "row_stats", "wsum", "res0", "res1",
               "res0_scale", "res0_shift", "ln_g", "out", "gn_partial", "out_row_stats"}


def _s(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def make_gemm_params(**kw) -> GemmParams:
    """Fill an ``idiff_gemm_params``; tensor-valued keyword arguments become device pointers."""
    p = GemmParams()
    keep = []
    for k, v in kw.items():
        if k in _PTR_FIELDS:
            if v is not None:
                keep.append(v)
                setattr(p, k, v.data_ptr())
        else:
            setattr(p, k, v)
    p._keepalive = keep
    return p


def conv_gemm(p: GemmParams, stream: Optional[int] = None) -> None:
    s = torch.cuda.current_stream().cuda_stream if stream is None else stream
    check(_lib.lib().idiff_conv_gemm(C.byref(p), s), "conv_gemm")


def conv3_rowpair(p: GemmParams, stream: Optional[int] = None) -> None:
    """Row-pair full-width-MMA kernel for 3x3 64 -> 64 layers (``p.w`` packed by ``pack_conv3_rowpair``)."""
    s = torch.cuda.current_stream().cuda_stream if stream is None else stream
    check(_lib.lib().idiff_conv3_rowpair(C.byref(p), s), "conv3_rowpair")


def conv_ref(p: GemmParams, w_f32: torch.Tensor, out_f32: torch.Tensor) -> None:
    check(_lib.lib().idiff_conv_ref(C.byref(p), w_f32.data_ptr(), out_f32.data_ptr(), _s(out_f32)), "conv_ref")


def sde_step(x, eps, mu, z, coef_row, *, is_score=False, philox=False, seed=0, offset=0, out=None):
    out = torch.empty_like(x) if out is None else out
    check(_lib.lib().idiff_sde_step(out.data_ptr(), x.data_ptr(), eps.data_ptr(),
                                    None if mu is None else mu.data_ptr(), None if z is None else z.data_ptr(),
                                    coef_row.data_ptr(), int(is_score), int(philox), seed, offset, x.numel(), _s(x)),
          "sde_step")
    return out


def philox_normal(n, device, seed=0, offset=0, step=0):
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(_lib.lib().idiff_philox_normal(out.data_ptr(), seed, offset, step, n, _s(out)), "philox_normal")
    return out


def stem_conv7(x, mu, w_nhwc, bias):
    B, _, H, W = x.shape
    N = w_nhwc.shape[0]
    out = torch.empty(B, H, W, N, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().idiff_stem_conv7(x.data_ptr(), mu.data_ptr(), w_nhwc.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                      B, H, W, N, _s(x)), "stem_conv7")
    return out


def stem_conv7_tc(x, mu, w, bias):
    """Production stem (tcgen05): w [64,2,7,7] fp32, bias [64] -> bf16 NHWC [B,H,W,64]."""
    from .packing import pack_stem_weight
    B, _, H, W = x.shape
    wp = pack_stem_weight(w.float(), bias.float()).to(x.device)
    assert wp.numel() * 2 == _lib.lib().idiff_stem_packed_bytes()
    out = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().idiff_stem_conv7_tc(x.data_ptr(), mu.data_ptr(), wp.data_ptr(), out.data_ptr(), B, H, W, _s(x)),
          "stem_conv7_tc")
    return out


def head_conv3(src, w_hwc, bias: float):
    """w_hwc: fp32 [3,3,64] (or [1,64,3,3]); packed into MMA fragments here (the UNet packs once per weight load)."""
    from .packing import pack_head_weight
    B, H, W, Cc = src.shape
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=src.device)
    wp = pack_head_weight(w_hwc.float()).to(src.device)
    check(_lib.lib().idiff_head_conv3(src.data_ptr(), wp.data_ptr(), float(bias), out.data_ptr(), B, H, W, Cc,
                                      _s(src)), "head_conv3")
    return out


def time_embed(t: Optional[torch.Tensor], t_scalar: float, w1t, b1, w2t, b2, wss, bss, B, nf):
    S = wss.shape[0]
    dev = w1t.device
    temb = torch.empty(B, 4 * nf, dtype=torch.float32, device=dev)
    out = torch.empty(B, S, dtype=torch.float32, device=dev)
    check(_lib.lib().idiff_time_embed(None if t is None else t.data_ptr(), float(t_scalar), w1t.data_ptr(),
                                      b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), wss.data_ptr(), bss.data_ptr(),
                                      temb.data_ptr(), out.data_ptr(), B, nf, S, _s(out)), "time_embed")
    return temb, out


def gn_stats(src, G):
    B, H, W, Cc = src.shape
    L = _lib.lib()
    ntile = L.idiff_gn_stats_ntile(H * W)
    part = torch.empty(B, ntile, G, 2, dtype=torch.float32, device=src.device)
    check(L.idiff_gn_stats(src.data_ptr(), part.data_ptr(), B, H * W, Cc, G, _s(src)), "gn_stats")
    return part


def gn_finalize(partial, gamma, beta, count, eps, t_scale=None, t_shift=None, t_ld=0):
    B, ntile, G, _ = partial.shape
    Cc = gamma.numel()
    sc = torch.empty(B, Cc, dtype=torch.float32, device=partial.device)
    sh = torch.empty_like(sc)
    check(_lib.lib().idiff_gn_finalize(partial.data_ptr(), ntile, gamma.data_ptr(), beta.data_ptr(),
                                       None if t_scale is None else t_scale.data_ptr(),
                                       None if t_shift is None else t_shift.data_ptr(), t_ld, sc.data_ptr(),
                                       sh.data_ptr(), B, Cc, G, count, eps, _s(partial)), "gn_finalize")
    return sc, sh


def make_gn_fuse(B, gamma, beta, count, eps, G=8, t_scale=None, t_shift=None, t_ld=0, state=None):
    """``idiff_gn_fuse`` for a producer launch (GroupNorm finalize folded into the kernel).  Returns
    ``(fuse, scale, shift, state)``; ``state = (sums, arrivals)`` is zeroed device memory that the kernel leaves zeroed
    again, so one pair can be passed on from launch to launch."""
    dev, Cc = gamma.device, gamma.numel()
    if state is None:
        state = (torch.zeros(B * GN_SLOTS * G * 2, dtype=torch.int64, device=dev), torch.zeros(4, dtype=torch.int32, device=dev))
    sc = torch.full((B, Cc), float("nan"), dtype=torch.float32, device=dev)
    sh = torch.full_like(sc, float("nan"))
    f = GnFuse()
    f.sums, f.arrivals = state[0].data_ptr(), state[1].data_ptr()
    f.gamma, f.beta = gamma.data_ptr(), beta.data_ptr()
    if t_scale is not None:
        f.t_scale, f.t_shift, f.t_ld = t_scale.data_ptr(), t_shift.data_ptr(), t_ld
    f.scale_out, f.shift_out = sc.data_ptr(), sh.data_ptr()
    f.count_per_group, f.eps = count, eps
    f._keepalive = [gamma, beta, t_scale, t_shift, sc, sh, state]
    return f, sc, sh, state


def chan_ln_gn(x, g, fuse: GnFuse, G=32, eps=1e-5):
    """Channel LayerNorm (y returned) + GroupNorm(G) statistics of y + their finalize (rows of ``fuse``) in one launch."""
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    check(_lib.lib().idiff_chan_ln_gn(x.data_ptr(), g.data_ptr(), y.data_ptr(), eps, B, H * W, Cc, G, C.byref(fuse), _s(x)),
          "chan_ln_gn")
    return y


def block_tail(y, scale, shift, res, want_stats=False, eps=1e-5):
    B, H, W, Cc = y.shape
    out = torch.empty_like(y)
    st = torch.empty(B * H * W, 2, dtype=torch.float32, device=y.device) if want_stats else None
    check(_lib.lib().idiff_block_tail(y.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                      None if res is None else res.data_ptr(), out.data_ptr(),
                                      None if st is None else st.data_ptr(), eps, B, H * W, Cc, _s(y)), "block_tail")
    return out, st


def add_rows(a, b, want_stats=False, eps=1e-5):
    Cc = a.shape[-1]
    rows = a.numel() // Cc
    out = torch.empty_like(a)
    st = torch.empty(rows, 2, dtype=torch.float32, device=a.device) if want_stats else None
    check(_lib.lib().idiff_add_rows(a.data_ptr(), None if b is None else b.data_ptr(), out.data_ptr(),
                                    None if st is None else st.data_ptr(), eps, rows, Cc, _s(a)), "add_rows")
    return out, st


def chan_ln(x, g, eps=1e-5):
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    check(_lib.lib().idiff_chan_ln(x.data_ptr(), g.data_ptr(), y.data_ptr(), eps, x.numel() // Cc, Cc, _s(x)), "chan_ln")
    return y


def linattn_context(qkv, w_out):
    B, H, W, _ = qkv.shape
    Cc = w_out.shape[0]
    L = _lib.lib()
    scratch = torch.empty(L.idiff_linattn_scratch_floats(B, H * W), dtype=torch.float32, device=qkv.device)
    weff = torch.empty(B, Cc * 128, dtype=torch.bfloat16, device=qkv.device)
    check(L.idiff_linattn_context(qkv.data_ptr(), w_out.data_ptr(), weff.data_ptr(), scratch.data_ptr(), B, H * W, Cc,
                                  _s(qkv)), "linattn_context")
    return weff


def self_attention(qkv, heads, scale):
    B, Lq, C3 = qkv.shape
    out = torch.empty(B, Lq, C3 // 3, dtype=torch.bfloat16, device=qkv.device)
    check(_lib.lib().idiff_self_attention(qkv.data_ptr(), out.data_ptr(), B, Lq, heads, scale, _s(qkv)),
          "self_attention")
    return out


def linattn_fused(x, row_stats, wqkv, g_pre, w_out, b_out, g_out, eps=1e-5):
    """Residual(PreNorm(LinearAttention)) for C = 64 / 128 in three fused passes.  x: bf16 [B,H,W,C]; row_stats: fp32
    [B*H*W, 2] (mean, rstd) of x; wqkv [384, C]; g_pre [C]; w_out [C, 128]; b_out, g_out [C]."""
    from .packing import pack_conv_weight
    B, H, W, Cc = x.shape
    L = _lib.lib()
    wg = wqkv.float() * g_pre.float()[None, :]
    wq = pack_conv_weight(wg[0:128], 128).to(x.device)
    wk = pack_conv_weight(wg[128:256], 128).to(x.device)
    wv = wg[256:384].contiguous().to(x.device)
    weff = torch.empty(B, Cc * 128, dtype=torch.bfloat16, device=x.device)
    scratch = torch.empty(L.idiff_linattn_fused_scratch_floats(B, H * W, Cc), dtype=torch.float32, device=x.device)
    out = torch.empty_like(x)
    check(L.idiff_linattn_fused(x.data_ptr(), row_stats.data_ptr(), wq.data_ptr(), wk.data_ptr(), wv.data_ptr(),
                                w_out.float().contiguous().data_ptr(), b_out.float().contiguous().data_ptr(),
                                g_out.float().contiguous().data_ptr(), weff.data_ptr(), out.data_ptr(),
                                scratch.data_ptr(), B, H * W, Cc, 32 ** -0.5, eps, _s(x)), "linattn_fused")
    return out
