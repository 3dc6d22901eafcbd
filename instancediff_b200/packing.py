"""Weight packing for the tcgen05 engine (host side, done once per weight load).

The B operand of ``idiff_conv_gemm`` is streamed by 1-D bulk TMA, so the weights are stored in
global memory as the exact shared-memory image of the UMMA SWIZZLE_NONE K-major layout:

    packed[n_tile][chunk][tap][c8][n_local][e]      bf16
      n_tile  = n // NT,  n_local = n % NT
      chunk   = c // 64   (64 input channels per pipeline stage)
      tap     = ky * k + kx
      c8      = (c % 64) // 8,  e = c % 8           (8 channels = one 16 B core-matrix row)

so one (chunk, tap) stage is ``NT * 64`` contiguous elements with LBO = NT*16 B, SBO = 128 B.
"""
from __future__ import annotations

import torch


def pack_conv_weight(w: torch.Tensor, NT: int) -> torch.Tensor:
    """w: [N, Cin, k, k] (or [N, Cin] for a linear layer) fp32 -> packed bf16 1-D tensor."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    N, Cin, kh, kw = w.shape
    assert kh == kw and N % NT == 0 and Cin % 64 == 0, (w.shape, NT)
    t = w.permute(0, 2, 3, 1).reshape(N // NT, NT, kh * kw, Cin // 64, 8, 8)
    #            n_tile, n_local, tap, chunk, c8, e  ->  n_tile, chunk, tap, c8, n_local, e
    t = t.permute(0, 3, 2, 4, 1, 5).contiguous()
    return t.to(torch.bfloat16).reshape(-1)


def pack_conv3_rowpair(w: torch.Tensor) -> torch.Tensor:
    """w: [64, 64, 3, 3] fp32 -> the resident B operand of ``idiff_conv3_rowpair``:

        packed[dx][c8][blk][n][e]      bf16,   blk = 0, 1, 2  <->  filter row dy = 2, 1, 0

    Per filter column dx and 8-channel plane c8 the three filter rows are stacked along N (64 output channels each),
    so the N = 128 windows [W(dy=2); W(dy=1)] and [W(dy=1); W(dy=0)] and the N = 64 windows W(2), W(0) are the same
    buffer viewed from a shifted start address (LBO = 3072 B between planes, SBO = 128 B)."""
    assert tuple(w.shape) == (64, 64, 3, 3), w.shape
    t = w.permute(3, 1, 2, 0).reshape(3, 8, 8, 3, 64)      # dx, c8, e, dy, n
    t = t.flip(3).permute(0, 1, 3, 4, 2).contiguous()      # dx, c8, blk (dy = 2,1,0), n, e
    return t.to(torch.bfloat16).reshape(-1)


def unpack_conv_weight(p: torch.Tensor, N: int, Cin: int, k: int, NT: int) -> torch.Tensor:
    """Inverse of pack_conv_weight (tests)."""
    t = p.reshape(N // NT, Cin // 64, k * k, 8, NT, 8).permute(0, 4, 2, 1, 3, 5)
    return t.reshape(N, k, k, Cin).permute(0, 3, 1, 2).float()


def bf16_round(w: torch.Tensor) -> torch.Tensor:
    return w.to(torch.bfloat16).float()


def fold_layernorm(w: torch.Tensor, gain: torch.Tensor, beta: torch.Tensor | None):
    """LayerNorm folded into the following linear layer.

    y = W (g * (x - m) * r + beta)  =  r * (W' x - m * wsum) + W beta,   W' = W diag(g)
    Returns (W', wsum, extra_bias) with wsum computed from the bf16-rounded W' the tensor core sees.
    """
    wf = w * gain[None, :]
    wsum = bf16_round(wf).sum(dim=1)
    extra = w @ beta if beta is not None else None
    return wf, wsum, extra


def interleave_geglu(w: torch.Tensor, b: torch.Tensor):
    """[2F, K] (value rows then gate rows) -> rows interleaved (v0, g0, v1, g1, ...)."""
    F = w.shape[0] // 2
    wi = torch.stack([w[:F], w[F:]], dim=1).reshape(2 * F, -1)
    bi = torch.stack([b[:F], b[F:]], dim=1).reshape(2 * F)
    return wi, bi


def pack_stem_weight(w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Stem 7x7 conv (w: [64, 2, 7, 7], bias: [64]) -> the shared-memory image idiff_stem_conv7_tc expects.

    K index k' = tap*2 + ci (tap = ky*7 + kx) for k' < 98; k' = 98 / 99 hold the bias as a bf16 hi / lo pair (the
    kernel's A rows carry the constant 1 there); zero padding up to 112.  The kernel stores its A block twice
    (bf16 hi and lo parts of the fp32 activations), so the 112 weights are repeated for the second half, with
    the bias rows cleared.  Layout: [28 K-groups][64 n][8 k] bf16 (LBO = 1024 B, SBO = 128 B).
    """
    N = w.shape[0]
    assert tuple(w.shape) == (N, 2, 7, 7) and N == 64, w.shape
    wk = torch.zeros(N, 112, dtype=torch.float32, device=w.device)
    wk[:, :98] = w.permute(0, 2, 3, 1).reshape(N, 98)
    lo_half = wk.clone()
    b_hi = bias.to(torch.bfloat16).float()
    wk[:, 98] = b_hi
    wk[:, 99] = bias - b_hi
    full = torch.cat([wk, lo_half], dim=1)                      # [N, 224]
    return full.reshape(N, 28, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16).reshape(-1)


def pack_head_weight(w: torch.Tensor) -> torch.Tensor:
    """Head 3x3 conv weights (w: [1, 64, 3, 3] or [3, 3, 64] fp32) -> the mma.sync.m16n8k16 B fragments
    idiff_head_conv3 loads: [3 n-tiles][4 k-chunks][32 lanes][b0, b1] uint32.  The taps sit on the N side: column
    n = lane//4 of n-tile j is tap 4j + n//2, holding bf16(w) for even n and bf16(w - bf16(w)) for odd n (the fp32
    weights keep ~16 mantissa bits); taps 9..11 are zero.  A lane holds k = 2*(lane%4) + {0,1} (b0) and k + 8 (b1) of
    its column.  Returned as a bf16 tensor of 1536 elements."""
    if w.dim() == 4:
        w = w[0].permute(1, 2, 0)                                   # [3, 3, 64]
    wk = w.reshape(9, 4, 16).float()                                # [tap][k-chunk][k]
    hi = wk.to(torch.bfloat16)
    lo = (wk - hi.float()).to(torch.bfloat16)
    frag = torch.zeros(3, 4, 32, 4, dtype=torch.bfloat16, device=w.device)      # [j][kc][lane][b0.lo, b0.hi, b1.lo, b1.hi]
    for j in range(3):
        for lane in range(32):
            n, k0 = lane // 4, (lane % 4) * 2
            tap = 4 * j + n // 2
            if tap >= 9:
                continue
            src = hi if n % 2 == 0 else lo
            frag[j, :, lane, 0] = src[tap, :, k0]
            frag[j, :, lane, 1] = src[tap, :, k0 + 1]
            frag[j, :, lane, 2] = src[tap, :, k0 + 8]
            frag[j, :, lane, 3] = src[tap, :, k0 + 9]
    return frag.reshape(-1).contiguous()
