"""First piece of the native backward: data gradients of the stride-1 convolutions on
the SAME tcgen05 engine as the forward.

``dgrad`` of a 3x3 / stride 1 / pad 1 convolution is itself a 3x3 / stride 1 / pad 1 convolution of the output gradient
with the filter flipped in both spatial axes and its channel axes swapped:

    gx[b, c, y, x] = sum_{n, ky, kx} w[n, c, ky, kx] * gy[b, n, y + 1 - ky, x + 1 - kx]
                   = conv3x3(gy, w')[b, c, y, x],      w'[c, n, ky, kx] = w[n, c, 2 - ky, 2 - kx]

so it runs on ``idiff_conv_gemm`` (or on ``idiff_conv3_rowpair`` for the 64 -> 64 layers) with a second packing of the
same weights; a 1x1 layer's dgrad is the linear layer with the transposed matrix.  ``Conv3x3`` wraps both directions as a
``torch.autograd.Function`` over channels-last bf16 tensors (the layout of the inference path); the weight gradient still
comes from PyTorch -- wgrad (a GEMM reduced over pixels) and the stride-2 / upsampled layers are not built, and the
training step of ``train.py`` does not use this yet.  Reference: ``loss.backward()``, models/drift_noise_model.py:294.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .packing import pack_conv3_rowpair, pack_conv_weight


def dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """[N, Cin, k, k] -> the forward-shaped filter [Cin, N, k, k] whose convolution with gy is the data gradient."""
    return w.flip(2, 3).transpose(0, 1).contiguous()


def _engine_conv(src: torch.Tensor, w: torch.Tensor, bias=None) -> torch.Tensor:
    """src: bf16 [B,H,W,Cin] CUDA; w: fp32 [N,Cin,k,k] (k = 1 or 3), stride 1."""
    B, H, W, cin = src.shape
    N, k = w.shape[0], w.shape[2]
    if cin % 64 or N % 64 or N > 256 and N % 256:
        raise _lib.IdiffError(f"engine convolution needs channel counts that are multiples of 64 (got {cin} -> {N})")
    out = torch.empty(B, H, W, N, dtype=torch.bfloat16, device=src.device)
    NT = min(N, 256)
    b = None if bias is None else bias.detach().float().contiguous()
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=k, stride=1, cin0=cin, N=N, NT=NT, epi=0, out_ld=N, src0=src.contiguous(), out=out,
                             w=pack_conv_weight(w.detach().float().cpu(), NT).to(src.device), bias=b)
    if k == 3 and cin == 64 and N == 64 and H % 2 == 0 and _lib.lib().idiff_conv3_rowpair_supported(p):
        p.w = (wrp := pack_conv3_rowpair(w.detach().float().cpu()).to(src.device)).data_ptr()
        p._keepalive.append(wrp)
        ops.conv3_rowpair(p)
    else:
        ops.conv_gemm(p)
    return out


def conv_dgrad(gy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Data gradient of ``y = conv(x, w)`` (stride 1, k = 1 or 3, pad k//2); gy, result: bf16 channels-last."""
    return _engine_conv(gy, dgrad_weight(w))


class Conv3x3(torch.autograd.Function):
    """y = conv3x3(x, w) + b on channels-last bf16 activations, forward AND data gradient on the tcgen05 engine."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return _engine_conv(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = conv_dgrad(gy, w) if ctx.needs_input_grad[0] else None
        gw = gb = None
        if ctx.needs_input_grad[1]:                            # wgrad: PyTorch (not built natively)
            gw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).float(), w.shape, gy.permute(0, 3, 1, 2).float(),
                                             padding=w.shape[2] // 2).to(w.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.float().sum(dim=(0, 1, 2))
        return gx, gw, gb
