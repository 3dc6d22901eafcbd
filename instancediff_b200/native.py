"""ctypes view of the C-level network object (``idiff_unet_*``, include/idiff.h, csrc/unet_plan.cu).

What a non-Python host does through the C ABI alone -- create, load the fp32 parameters by state_dict name, finalize,
set the image embedding, forward / whole reverse-SDE loop -- wrapped for tests and for Python callers that want the
plan, the packing and the loop to live in the library.  ``ConditionalUNet`` (unet.py) remains the instrumented Python
plan over the same kernel entry points (per-launch timing, named activations for the layer-wise parity tests).
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import IdiffError, UNetCfg, check


class NativeUNet:
    """``eps = net(x_t, mu, t, image_context=emb)`` with every host-side step inside libidiff_sm100.so."""

    def __init__(self, in_nc=2, out_nc=1, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512, down_kernel=4, device="cuda"):
        self.L = _lib.lib()
        cfg = UNetCfg()
        cfg.in_nc, cfg.out_nc, cfg.nf, cfg.n_levels = in_nc, out_nc, nf, len(ch_mult)
        for i, m in enumerate(ch_mult):
            cfg.ch_mult[i] = m
        cfg.context_dim, cfg.down_kernel = context_dim, down_kernel
        self._h = C.c_void_p()
        check(self.L.idiff_unet_create(C.byref(cfg), C.byref(self._h)), "unet_create")
        self.device = torch.device(device)
        self.context_dim = context_dim
        self._ctx_keep = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.L.idiff_unet_destroy(h)

    def load_state_dict(self, sd) -> "NativeUNet":
        for name, t in sd.items():
            host = t.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(1, host.dim()))(*host.shape)
            check(self.L.idiff_unet_load_weight(self._h, name.encode(), host.data_ptr(), host.dim(), shape), "unet_load_weight")
        with self._guard():
            check(self.L.idiff_unet_finalize(self._h), "unet_finalize")
        return self

    def _guard(self):
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_image_context(self, ctx: torch.Tensor) -> None:
        if ctx.dim() == 3:
            ctx = ctx[:, 0]
        c = ctx.detach().to(self.device, torch.float32).contiguous()
        with self._guard():
            check(self.L.idiff_unet_set_context(self._h, c.data_ptr(), c.shape[0], self._stream()), "unet_set_context")
        self._ctx_keep = c

    def forward(self, xt, cond, time, *unused, image_context=None, **unused_kw):
        if image_context is None:
            raise IdiffError("image_context is required")
        B, _, H, W = xt.shape
        if H % 16 or W % 16:
            raise IdiffError("NativeUNet: H and W must be multiples of 16")
        self.set_image_context(image_context)
        xt, cond = xt.contiguous(), cond.contiguous()
        out = torch.empty_like(xt)
        t_dev, t_scalar = None, 0.0
        if torch.is_tensor(time) and time.numel() > 1:
            t_dev = time.detach().to(self.device, torch.float32).reshape(B).contiguous()
        else:
            t_scalar = float(time.reshape(-1)[0].item()) if torch.is_tensor(time) else float(time)
        with self._guard():
            check(self.L.idiff_unet_forward(self._h, xt.data_ptr(), cond.data_ptr(), None if t_dev is None else t_dev.data_ptr(),
                                            t_scalar, out.data_ptr(), B, H, W, self._stream()), "unet_forward")
        return out

    __call__ = forward

    def reverse_sde(self, xt, mu, image_context, table: torch.Tensor, T: int, sample_scale: float = 1.0, seed: int = 0,
                    offset: int = 0) -> torch.Tensor:
        """The loop of utils/sde_utils.py:244-261 inside the library; ``table`` = IRSDE._coef_table(device)."""
        B, _, H, W = xt.shape
        self.set_image_context(image_context)
        x = xt.clone().contiguous()
        with self._guard():
            check(self.L.idiff_unet_reverse_sde(self._h, x.data_ptr(), mu.contiguous().data_ptr(), table.data_ptr(), int(T),
                                                float(sample_scale), seed, offset, B, H, W, self._stream()), "unet_reverse_sde")
        return x

    def num_launches(self, B, H, W) -> int:
        n = self.L.idiff_unet_num_launches(self._h, B, H, W)
        if n < 0:
            check(n, "unet_num_launches")
        return n
