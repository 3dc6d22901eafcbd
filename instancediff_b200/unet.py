"""Drift/noise-prediction UNet on hand-written sm_100a kernels (host orchestration only).

Drop-in for the network the reference plugs into ``IRSDE.set_model`` (utils/sde_utils.py:164-165)
and calls as ``self.model(x, self.mu, t * scale, **kwargs)`` (utils/sde_utils.py:198); the wrapper
convention ``net(a, b, t[B], names, text_encoder, image_context=A_emb)``
(models/drift_noise_model.py:250-268) is accepted too (extra positionals are ignored).
Architecture: SURVEY.md App. A (``models/modules/MSM_degEmb_Unet.py`` is not in the reference
snapshot); hyper-parameters follow Configurations/config.yml:109-113.

This file contains NO arithmetic on activations: it packs weights once, builds a per-shape launch
plan (a flat list of C-ABI calls with pre-filled argument structs) and replays it.  Every kernel
is in ``csrc/``; if ``libidiff_sm100.so`` is missing the constructor raises (no fallback).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import EPI_GEGLU, EPI_LN_OUT, EPI_PLAIN, EPI_QSOFTMAX, GN_SLOTS, GemmParams, GnFuse, check
from .packing import (fold_layernorm, interleave_geglu, pack_conv3_rowpair, pack_conv_weight, pack_head_weight,
                      pack_stem_weight)

GN_GROUPS = 8
TILE_H, TILE_W = 16, 8


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------
# parameter inventory (names == oracle/unet_oracle.py state_dict keys == the weight contract)
# ----------------------------------------------------------------------------------------------
def _conv(prefix, cout, cin, k, bias=True):
    yield prefix + ".weight", (cout, cin, k, k), cin * k * k
    if bias:
        yield prefix + ".bias", (cout,), cin * k * k


def _linear(prefix, cout, cin, bias=True):
    yield prefix + ".weight", (cout, cin), cin
    if bias:
        yield prefix + ".bias", (cout,), cin


def _norm(prefix, c):
    yield prefix + ".weight", (c,), "ones"
    yield prefix + ".bias", (c,), "zeros"


def _resblock(prefix, cin, cout, td):
    yield from _linear(prefix + ".mlp", 2 * cout, td)
    yield from _conv(prefix + ".conv1", cout, cin, 3)
    yield from _norm(prefix + ".norm1", cout)
    yield from _conv(prefix + ".conv2", cout, cout, 3)
    yield from _norm(prefix + ".norm2", cout)
    if cin != cout:
        yield from _conv(prefix + ".res_conv", cout, cin, 1)


def _attn(prefix, dim, kind, context_dim):
    yield prefix + ".prenorm.g", (1, dim, 1, 1), "ones"
    f = prefix + ".fn"
    if kind == "linear":
        yield from _conv(f + ".to_qkv", 384, dim, 1, bias=False)
        yield from _conv(f + ".to_out", dim, 128, 1)
        yield f + ".out_norm.g", (1, dim, 1, 1), "ones"
    else:
        yield from _norm(f + ".norm", dim)
        yield from _conv(f + ".proj_in", dim, dim, 1)
        yield from _norm(f + ".norm1", dim)
        for n in ("to_q", "to_k", "to_v"):
            yield from _linear(f + ".attn1." + n, dim, dim, bias=False)
        yield from _linear(f + ".attn1.to_out", dim, dim)
        yield from _norm(f + ".norm2", dim)
        yield from _linear(f + ".attn2.to_q", dim, dim, bias=False)
        yield from _linear(f + ".attn2.to_k", dim, context_dim, bias=False)
        yield from _linear(f + ".attn2.to_v", dim, context_dim, bias=False)
        yield from _linear(f + ".attn2.to_out", dim, dim)
        yield from _norm(f + ".norm3", dim)
        yield from _linear(f + ".ff.proj", dim * 8, dim)
        yield from _linear(f + ".ff.out", dim, dim * 4)
        yield from _conv(f + ".proj_out", dim, dim, 1)


def param_specs(in_nc=2, out_nc=1, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512, down_kernel=4):
    td = nf * 4
    dims = [nf] + [nf * m for m in ch_mult]
    io = list(zip(dims[:-1], dims[1:]))
    n = len(io)
    yield from _conv("init_conv", nf, in_nc, 7)
    yield from _linear("time_lin1", td, nf)
    yield from _linear("time_lin2", td, td)
    for i, (di, do) in enumerate(io):
        last = i == n - 1
        yield from _resblock(f"downs.{i}.0", di, di, td)
        yield from _resblock(f"downs.{i}.1", di, di, td)
        yield from _attn(f"downs.{i}.2", di, "spatial" if last else "linear", context_dim)
        yield from _conv(f"downs.{i}.3", do, di, 3 if last else down_kernel)
    mid = dims[-1]
    yield from _resblock("mid_block1", mid, mid, td)
    yield from _attn("mid_attn", mid, "spatial", context_dim)
    yield from _resblock("mid_block2", mid, mid, td)
    for i, (di, do) in enumerate(reversed(io)):
        lvl = n - 1 - i
        yield from _resblock(f"ups.{i}.0", do + di, do, td)
        yield from _resblock(f"ups.{i}.1", do + di, do, td)
        yield from _attn(f"ups.{i}.2", do, "spatial" if lvl == n - 1 else "linear", context_dim)
        yield from _conv(f"ups.{i}.3.conv" if lvl > 0 else f"ups.{i}.3", di, do, 3)
    yield from _resblock("final_res", 2 * nf, nf, td)
    yield from _conv("final_conv", out_nc, nf, 3)


class _Act:
    """A channels-last bf16 activation [B,H,W,C] (+ optional per-pixel LayerNorm stats)."""

    __slots__ = ("t", "H", "W", "C", "stats")

    def __init__(self, t, H, W, C, stats=None):
        self.t, self.H, self.W, self.C, self.stats = t, H, W, C, stats


class ConditionalUNet:
    """``eps = net(x_t, mu, t, image_context=emb)``; fp32 [B,1,H,W] in and out, bf16 inside."""

    def __init__(self, in_nc=2, out_nc=1, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512, down_kernel=4,
                 device="cuda", seed: Optional[int] = None):
        if in_nc != 2 or out_nc != 1 or nf != 64 or tuple(ch_mult) != (1, 2, 4, 4) or down_kernel != 4:
            raise _lib.IdiffError("ConditionalUNet: kernels are built for in_nc=2, out_nc=1, nf=64, "
                                  "ch_mult=[1,2,4,4], down_kernel=4 (Configurations/config.yml:109-113)")
        self.L = _lib.lib()                       # raises if the CUDA library is not built
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, ch_mult=tuple(ch_mult), context_dim=context_dim,
                        down_kernel=down_kernel)
        self.device = torch.device(device)
        self.nf, self.td, self.context_dim = nf, nf * 4, context_dim
        self.dims = [nf] + [nf * m for m in ch_mult]
        self.io = list(zip(self.dims[:-1], self.dims[1:]))
        self.params: Dict[str, torch.Tensor] = {}
        self._plans: Dict[tuple, "_Plan"] = {}
        self._ctx_key = None
        self._ctx_ref = None
        self._crossvec: Dict[tuple, torch.Tensor] = {}
        # 3x3 64 -> 64 layers on the full-width-MMA row-pair kernel (csrc/conv3_rowpair.cu); IDIFF_NO_ROWPAIR=1 keeps
        # them on the generic engine (A/B measurements)
        self.use_rowpair = os.environ.get("IDIFF_NO_ROWPAIR", "0") != "1"
        self.rowpair_packed_silu = os.environ.get("IDIFF_ROWPAIR_F32_SILU", "0") != "1"
        # IDIFF_GN_FUSE=1: GroupNorm finalize folded into the producing conv (exact integer sums + last-CTA finalize, 44
        # launches fewer per step).  Off by default: measured neutral to 0.8 % slower inside the captured step -- the
        # finishing CTA's tail (fence, arrival, L2 round trip: ~2.5 us) costs what a dependent 2 us launch costs in a graph
        # (profiles/README.md); the default keeps partial rows + idiff_gn_finalize.
        self.fuse_gn = os.environ.get("IDIFF_GN_FUSE", "0") == "1"
        # same for the SpatialTransformer entry (channel LayerNorm + GroupNorm(32) statistics + finalize in one launch)
        self.fuse_gn_st = os.environ.get("IDIFF_NO_GN_FUSE_ST", "0") != "1"
        self.pk: Optional[Dict[str, dict]] = None    # packed weights: built lazily, ONCE per weight load
        self._version = 0
        self._init_params(seed)
        self._invalidate()

    # ------------------------------------------------------------------ parameters
    def _init_params(self, seed):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(1 if seed is None else seed)
        for name, shape, init in param_specs(**self.cfg):
            if init == "ones":
                t = torch.ones(shape)
            elif init == "zeros":
                t = torch.zeros(shape)
            else:                                  # nn.Conv2d / nn.Linear default: U(-1/sqrt(fan_in), +)
                bound = 1.0 / math.sqrt(init)
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            self.params[name] = t.to(self.device, torch.float32)

    def state_dict(self):
        return {k: v.clone() for k, v in self.params.items()}

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.params if k not in sd]
        extra = [k for k in sd if k not in self.params]
        if strict and (missing or extra):
            raise KeyError(f"load_state_dict: missing {missing[:4]}..., unexpected {extra[:4]}...")
        for k in self.params:
            if k in sd:
                if tuple(sd[k].shape) != tuple(self.params[k].shape):
                    raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(self.params[k].shape)}")
                self.params[k] = sd[k].detach().to(self.device, torch.float32).contiguous()
        self._invalidate()
        return self

    def eval(self):
        return self

    def to(self, device):
        def norm(d):
            d = torch.device(d)
            if d.type == "cuda" and d.index is None:
                return torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
            return d
        if norm(device) != norm(self.device):
            self.device = torch.device(device)
            self.params = {k: v.to(self.device) for k, v in self.params.items()}
            self._crossvec.clear()                   # context vectors live on the old device
            self._ctx_ref = None
            self._invalidate()
        return self

    # ------------------------------------------------------------------ packing
    def _invalidate(self):
        """The weights changed: drop the packed arena and every launch plan.  Packing itself is deferred to the first
        forward, so ``ConditionalUNet(...).load_state_dict(sd)`` packs once, not twice."""
        self.pk = None
        self._variants: Dict[tuple, torch.Tensor] = {}
        self._variant_src: Dict[str, torch.Tensor] = {}
        self._plans.clear()
        self._ctx_key = None
        self._version += 1                          # captured graphs hold pointers into the old packing

    def _ensure_packed(self):
        if self.pk is None:
            self._pack()

    def packed_variant(self, name: str, NT: int) -> torch.Tensor:
        """The weights of layer ``name`` packed for another N tile than the default one (small grids: a 256-channel
        layer with fewer work items than SMs runs as two 128-channel tiles).  Packed on the host on first use."""
        key = (name, NT)
        if key not in self._variants:
            w = self._variant_src[name]
            self._variants[key] = pack_conv_weight(w, NT).to(self.device)
        return self._variants[key]

    def _pack(self):
        """Host-side packing: every layout transform runs on CPU copies of the parameters, the results are laid out
        in ONE byte arena and reach the device in a single copy (no ATen kernel launches on the GPU)."""
        P = {k: v.detach().to("cpu", torch.float32) for k, v in self.params.items()}
        pk: Dict[str, dict] = {}

        def f32(t):
            return t.detach().to(torch.float32).contiguous()

        def conv_entry(name, NT=None, w=None, b=None, k=None):
            w = P[name + ".weight"] if w is None else w
            b = P.get(name + ".bias") if b is None else b
            N = w.shape[0]
            NT = min(N, 256) if NT is None else NT
            pk[name] = dict(w=pack_conv_weight(w, NT), bias=None if b is None else f32(b), N=N, NT=NT,
                            cin=w.shape[1], k=(w.shape[2] if w.dim() == 4 else 1) if k is None else k)
            if tuple(w.shape) == (64, 64, 3, 3):          # second packing for idiff_conv3_rowpair
                pk[name]["w_rp"] = pack_conv3_rowpair(w)
            if N >= 256:                                   # may be re-packed for NT = 128 (small grids)
                pk[name]["name"] = name
                self._variant_src[name] = w.detach().clone()
            return pk[name]

        def resblock(prefix):
            conv_entry(prefix + ".conv1")
            conv_entry(prefix + ".conv2")
            if (prefix + ".res_conv.weight") in P:
                conv_entry(prefix + ".res_conv")
            for n in ("norm1", "norm2"):
                pk[prefix + "." + n] = dict(g=f32(P[f"{prefix}.{n}.weight"]), b=f32(P[f"{prefix}.{n}.bias"]))

        def attn(prefix, kind, dim):
            f = prefix + ".fn"
            g_pre = P[prefix + ".prenorm.g"].reshape(-1)
            if kind == "linear":
                wq = P[f + ".to_qkv.weight"].reshape(384, dim)
                wf, wsum, _ = fold_layernorm(wq, g_pre, None)
                e = conv_entry(f + ".to_qkv", NT=128, w=wf, b=None)
                e["wsum"] = f32(wsum)
                pk[f + ".to_out"] = dict(w_f32=f32(P[f + ".to_out.weight"].reshape(dim, 128)),
                                         bias=f32(P[f + ".to_out.bias"]), g=f32(P[f + ".out_norm.g"].reshape(-1)))
                if dim in (64, 128):                  # fused three-pass kernel: q / k weights packed separately,
                    wg = wq * g_pre[None, :]          # v folded into the merge step (pre-norm gain in the weights)
                    pk[f + ".fused"] = dict(wq=pack_conv_weight(wg[0:128], 128),
                                            wk=pack_conv_weight(wg[128:256], 128),
                                            wv=f32(wg[256:384]))
            else:
                pk[prefix + ".prenorm"] = dict(g=f32(g_pre))
                pk[f + ".norm"] = dict(g=f32(P[f + ".norm.weight"]), b=f32(P[f + ".norm.bias"]))
                conv_entry(f + ".proj_in")
                wqkv = torch.cat([P[f + ".attn1.to_q.weight"], P[f + ".attn1.to_k.weight"],
                                  P[f + ".attn1.to_v.weight"]], dim=0)
                wf, wsum, extra = fold_layernorm(wqkv, P[f + ".norm1.weight"], P[f + ".norm1.bias"])
                e = conv_entry(f + ".attn1.qkv", NT=256, w=wf, b=extra)
                e["wsum"] = f32(wsum)
                conv_entry(f + ".attn1.to_out")
                wi, bi = interleave_geglu(P[f + ".ff.proj.weight"], P[f + ".ff.proj.bias"])
                wf, wsum, extra = fold_layernorm(wi, P[f + ".norm3.weight"], P[f + ".norm3.bias"])
                e = conv_entry(f + ".ff.proj", NT=256, w=wf, b=bi + extra)
                e["wsum"] = f32(wsum)
                conv_entry(f + ".ff.out")
                conv_entry(f + ".proj_out")

        # stem / head / time
        pk["init_conv"] = dict(w=pack_stem_weight(f32(P["init_conv.weight"]), f32(P["init_conv.bias"])))
        pk["final_conv"] = dict(w=pack_head_weight(f32(P["final_conv.weight"])),
                                bias=float(P["final_conv.bias"].reshape(-1)[0].item()))
        self._res_names: List[str] = []
        n = len(self.io)
        for i, (di, do) in enumerate(self.io):
            kind = "spatial" if i == n - 1 else "linear"
            for j in (0, 1):
                resblock(f"downs.{i}.{j}")
                self._res_names.append(f"downs.{i}.{j}")
            attn(f"downs.{i}.2", kind, di)
            conv_entry(f"downs.{i}.3")
        resblock("mid_block1")
        self._res_names.append("mid_block1")
        attn("mid_attn", "spatial", self.dims[-1])
        resblock("mid_block2")
        self._res_names.append("mid_block2")
        for i, (di, do) in enumerate(reversed(self.io)):
            lvl = n - 1 - i
            kind = "spatial" if lvl == n - 1 else "linear"
            for j in (0, 1):
                resblock(f"ups.{i}.{j}")
                self._res_names.append(f"ups.{i}.{j}")
            attn(f"ups.{i}.2", kind, do)
            conv_entry(f"ups.{i}.3.conv" if lvl > 0 else f"ups.{i}.3")
        resblock("final_res")
        self._res_names.append("final_res")

        # time embedding: transposed MLP weights + every ResBlock's Linear concatenated
        self._ss_off: Dict[str, int] = {}
        ws, bs, off = [], [], 0
        for name in self._res_names:
            w, b = P[name + ".mlp.weight"], P[name + ".mlp.bias"]
            self._ss_off[name] = off
            off += w.shape[0]
            ws.append(w)
            bs.append(b)
        self.S = off
        pk["time"] = dict(w1t=f32(P["time_lin1.weight"].t()), b1=f32(P["time_lin1.bias"]),
                          w2t=f32(P["time_lin2.weight"].t()), b2=f32(P["time_lin2.bias"]),
                          wss=f32(torch.cat(ws, 0)), bss=f32(torch.cat(bs, 0)))
        # one arena, one host-to-device copy: 256 B aligned slices viewed with each tensor's own dtype
        slots, off = [], 0
        for entry in pk.values():
            for key, val in entry.items():
                if torch.is_tensor(val):
                    val = val.contiguous()
                    nb = val.numel() * val.element_size()
                    slots.append((entry, key, val, off, nb))
                    off += (nb + 255) // 256 * 256
        host = torch.empty(max(off, 256), dtype=torch.uint8)
        for _, _, val, o, nb in slots:
            host[o:o + nb] = val.reshape(-1).view(torch.uint8)
        self._arena = host.to(self.device)
        for entry, key, val, o, nb in slots:
            entry[key] = self._arena[o:o + nb].view(val.dtype).reshape(val.shape)
        self.pk = pk

    # ------------------------------------------------------------------ conditioning
    def spatial_layers(self):
        n = len(self.io)
        return [f"downs.{n - 1}.2", "mid_attn", "ups.0.2"]

    def set_image_context(self, ctx: torch.Tensor):
        """Cross-attention to ONE context token: softmax == 1, so CrossAttn(x, ctx) = Wo (Wv ctx) + bo for
        every pixel and every step (App. A) -- computed once per embedding, added as a per-image bias."""
        ctx_arg = ctx                                  # the caller's tensor object (identity of the embedding)
        if ctx.dim() == 3:
            if ctx.shape[1] != 1:
                raise _lib.IdiffError("image_context must hold one token per image ([B,1,D] or [B,D])")
            ctx = ctx[:, 0]
        # identity of the embedding tensor: the tensor itself is kept alive in self._ctx_ref, otherwise the caching
        # allocator may hand its address to the NEXT embedding (same pointer, version and shape -> stale context)
        key = (ctx.data_ptr(), ctx._version, tuple(ctx.shape))
        if key == self._ctx_key and self._ctx_ref is ctx_arg:
            return
        c = ctx.detach().to(self.device, torch.float32).contiguous()
        B, D = c.shape
        s = torch.cuda.current_stream(self.device).cuda_stream
        for name in self.spatial_layers():
            f = name + ".fn.attn2"
            wv, wo, bo = (self.params[f + ".to_v.weight"], self.params[f + ".to_out.weight"],
                          self.params[f + ".to_out.bias"])
            buf = self._crossvec.get((name, B))
            if buf is None:                      # persistent per batch size: captured graphs keep the pointer
                buf = torch.empty(B, wo.shape[0], dtype=torch.float32, device=self.device)
                self._crossvec[(name, B)] = buf
            check(self.L.idiff_cross_vec(c.data_ptr(), wv.data_ptr(), wo.data_ptr(), bo.data_ptr(), buf.data_ptr(),
                                         B, D, wo.shape[0], s), "cross_vec")
        self._ctx_key = key
        self._ctx_ref = ctx_arg
        for plan in self._plans.values():
            if plan.B == B:
                plan.bind_context(self._crossvec)

    # ------------------------------------------------------------------ forward
    def _plan(self, B, H, W, shared_time) -> "_Plan":
        self._ensure_packed()
        key = (B, H, W, shared_time)
        if key not in self._plans:
            self._plans[key] = _Plan(self, B, H, W, shared_time)
            self._plans[key].bind_context(self._crossvec)
        return self._plans[key]

    def time_table(self, times) -> torch.Tensor:
        """[len(times), S] fp32: the time conditioning (every ResBlock's scale / shift projection of the time embedding)
        for each model time -- it depends on t only, so a sampler builds it once and the captured step copies one row
        (idiff_step_select_ss) instead of launching the time MLP."""
        self._ensure_packed()
        tm = self.pk["time"]
        n = len(times)
        table = torch.empty(n, self.S, dtype=torch.float32, device=self.device)
        temb = torch.empty(1, self.td, dtype=torch.float32, device=self.device)
        s = torch.cuda.current_stream(self.device).cuda_stream
        for i, t in enumerate(times):
            check(self.L.idiff_time_embed(None, float(t), _ptr(tm["w1t"]), _ptr(tm["b1"]), _ptr(tm["w2t"]), _ptr(tm["b2"]),
                                          _ptr(tm["wss"]), _ptr(tm["bss"]), _ptr(temb), table[i].data_ptr(), 1, self.nf, self.S, s),
                  "time_embed")
        return table

    def time_slot(self, B, H, W) -> int:
        """Device address of the [S] buffer the shared-time plan of this shape reads its time conditioning from."""
        return self._plan(B, H, W, True).ss.data_ptr()

    def forward_into(self, xt, cond, time, image_context, time_ptr: Optional[int] = None, ss_ready: bool = False):
        """Runs the network; returns the plan-owned fp32 output buffer [B,1,H,W] (overwritten by the
        next call).  ``time_ptr``: device float holding the shared time (CUDA-graph replay); ``ss_ready``: the caller
        has already placed this step's row of ``time_table`` in ``time_slot`` (no time-embedding launches)."""
        if xt.dtype != torch.float32 or cond.dtype != torch.float32 or not xt.is_cuda:
            raise _lib.IdiffError("ConditionalUNet expects fp32 CUDA tensors [B,1,H,W]")
        B, Cx, H, W = xt.shape
        if Cx != 1 or tuple(cond.shape) != tuple(xt.shape):
            raise _lib.IdiffError(f"expected xt and cond of shape [B,1,H,W], got {tuple(xt.shape)} / {tuple(cond.shape)}")
        if H % 16 or W % 16:
            raise _lib.IdiffError("H and W must be multiples of 16 (pad with ConditionalUNet.forward)")
        if image_context is None:
            raise _lib.IdiffError("image_context is required (use_image_context: True, config.yml:132)")
        if image_context.shape[0] != B:
            raise _lib.IdiffError("image_context batch does not match x")
        self.set_image_context(image_context)
        xt, cond = xt.contiguous(), cond.contiguous()
        t_dev, t_scalar = None, 0.0
        if time_ptr is not None:
            shared = True
        elif torch.is_tensor(time) and time.numel() > 1:
            if time.numel() != B:
                raise _lib.IdiffError("time tensor must have 1 or B elements")
            shared = False
            t_dev = time.detach().to(self.device, torch.float32).reshape(B).contiguous()
        else:
            shared = True
            t_scalar = float(time.reshape(-1)[0].item()) if torch.is_tensor(time) else float(time)
        plan = self._plan(B, H, W, shared)
        plan.run(xt, cond, t_dev if time_ptr is None else time_ptr, t_scalar, skip_time=ss_ready and shared)
        return plan.eps

    def forward(self, xt, cond, time, *unused, image_context=None, **unused_kw):
        """``image_context`` is keyword-only like in the reference's call (models/drift_noise_model.py:253) and
        REQUIRED: the network has no unconditioned mode (raises IdiffError when it is missing)."""
        H, W = xt.shape[-2:]
        if xt.shape[0] == 0:                           # empty batch: nothing to launch
            return torch.empty_like(xt)
        ph, pw = (-H) % 16, (-W) % 16
        if ph or pw:                                 # reflect-pad right/bottom, crop after (App. A)
            xt = torch.nn.functional.pad(xt, (0, pw, 0, ph), mode="reflect")
            cond = torch.nn.functional.pad(cond, (0, pw, 0, ph), mode="reflect")
        out = self.forward_into(xt, cond, time, image_context).clone()
        return out[..., :H, :W] if (ph or pw) else out

    __call__ = forward


# ----------------------------------------------------------------------------------------------
class _Plan:
    """Flat launch list for one (B, H, W) shape.  Ops are closures ``f(stream_ptr)``."""

    def __init__(self, net: ConditionalUNet, B, H, W, shared_time):
        self.net, self.B, self.H, self.W, self.shared_time = net, B, H, W, shared_time
        self.L = net.L
        self.dev = net.device
        self.ops: List = []
        self.op_info: List[tuple] = []        # (kernel kind, algorithmic FLOPs, label) parallel to self.ops
        self.op_bytes: Dict[int, float] = {}  # op index -> algorithmic HBM bytes (conv_gemm launches)
        self.keep: List = []
        self.scratch: Dict[tuple, torch.Tensor] = {}
        self.ctx_slots: Dict[str, GemmParams] = {}
        self.named: Dict[str, _Act] = {}      # block outputs by module name (layer-wise parity tests)
        self.n_launch = 0
        self._gn_sums = self._gn_arrivals = None
        self._build()

    # ---- buffers
    def new(self, shape, dtype=torch.bfloat16):
        t = torch.empty(shape, dtype=dtype, device=self.dev)
        self.keep.append(t)
        return t

    def tmp(self, name, shape, dtype=torch.bfloat16):
        key = (name, tuple(shape), dtype)
        if key not in self.scratch:
            self.scratch[key] = torch.empty(shape, dtype=dtype, device=self.dev)
        return self.scratch[key]

    def act(self, H, W, C, stats=False, tmp_name=None):
        t = self.tmp(tmp_name, (self.B, H, W, C)) if tmp_name else self.new((self.B, H, W, C))
        st = self.new((self.B * H * W, 2), torch.float32) if stats else None
        return _Act(t, H, W, C, st)

    # ---- op helpers
    def gemm(self, src0: _Act, src1: Optional[_Act], entry: dict, out: _Act, *, k=1, stride=1, up=0,
             a_scale=None, a_shift=None, a_silu=0, epi=EPI_PLAIN, gn_partial=None, bias=True, bias_img_slot=None,
             row_stats=None, res0=None, res1=None, res0_scale=None, res0_shift=None, ln_g=None, out_stats=None,
             qscale=1.0, ln_eps=1e-5, cin0=None, src0_ld=0, w_override=None, w_image_stride=0, NT=None, out_ld=None,
             gn_fuse=None):
        p = GemmParams()
        p.B, p.H, p.W = self.B, out.H, out.W
        p.ksize, p.stride, p.up0 = k, stride, up
        p.cin0 = src0.C if cin0 is None else cin0
        p.cin1 = src1.C if src1 is not None else 0
        p.src0_ld, p.src1_ld = src0_ld, 0
        p.N = entry["N"]
        p.NT = entry["NT"] if NT is None else NT
        # Small grids: with fewer (pixel tile x N tile) items than ~3/4 of the SMs, a 256-wide N tile is split in two.
        # An M128 x N128 MMA takes about half the time of an N256 one, so the layer's critical path halves while the
        # idle SMs take the extra items (N = 64 would not help: its MMAs are bound by the A-operand read).  Results
        # are bit-identical: every output element sees the same K sequence, GroupNorm entries are per group.
        if (NT is None and p.NT == 256 and "name" in entry and w_override is None and epi == EPI_PLAIN
                and res0 is None and res1 is None and out_stats is None):
            tiles = self.B * (-(-out.H // TILE_H)) * (-(-out.W // TILE_W))
            if tiles * (p.N // 256) <= 111:
                p.NT = 128
                w_override = self.net.packed_variant(entry["name"], 128)
        p.a_silu, p.epi = a_silu, epi
        p.gn_groups = GN_GROUPS if (gn_partial is not None or gn_fuse is not None) else 0
        if gn_fuse is not None:                  # GroupNorm finalize folded into this launch (csrc/gn_fuse.cuh)
            p.gn_fuse = C.pointer(gn_fuse)
        p.out_ld = out.C if out_ld is None else out_ld
        p.src0, p.src1 = _ptr(src0.t), _ptr(src1.t if src1 is not None else None)
        p.a_scale, p.a_shift = _ptr(a_scale), _ptr(a_shift)
        p.w = _ptr(entry["w"] if w_override is None else w_override)
        p.w_image_stride = w_image_stride
        p.bias = _ptr(entry.get("bias")) if bias else None
        p.row_stats = _ptr(row_stats)
        p.wsum = _ptr(entry.get("wsum")) if row_stats is not None else None
        p.res0, p.res1 = _ptr(res0), _ptr(res1)
        p.res0_scale, p.res0_shift = _ptr(res0_scale), _ptr(res0_shift)
        p.ln_g = _ptr(ln_g)
        p.out = _ptr(out.t)
        p.gn_partial = _ptr(gn_partial)
        p.out_row_stats = _ptr(out_stats)
        p.qscale, p.ln_eps = qscale, ln_eps
        if bias_img_slot is not None:
            self.ctx_slots[bias_img_slot] = p
        self.keep.append(p)
        L, ref = self.L, C.byref(p)
        rp = w_override is None and self.rowpair_ok(entry, src0, src1, out, k, stride, up) and bool(
            L.idiff_conv3_rowpair_supported(ref))
        if rp:
            p.w = _ptr(entry["w_rp"])
            if a_silu and self.net.rowpair_packed_silu:
                p.a_silu = 3                      # bit 1: affine + SiLU of the loader in packed bf16x2 arithmetic
            self.ops.append(lambda s: check(L.idiff_conv3_rowpair(ref, s), "conv3_rowpair"))
        else:
            self.ops.append(lambda s: check(L.idiff_conv_gemm(ref, s), "conv_gemm"))
        flops = 2.0 * self.B * out.H * out.W * p.N * (p.cin0 + p.cin1) * k * k
        # algorithmic HBM bytes: every source pixel once, every output element once, residuals once (bf16)
        px_out = self.B * out.H * out.W
        px_in = px_out * (stride * stride) // (4 if up else 1)
        out_cols = p.N // 2 if epi == EPI_GEGLU else p.N
        nbytes = 2.0 * (px_in * (p.cin0 + p.cin1) + px_out * out_cols
                        + px_out * p.N * ((res0 is not None) + (res1 is not None)))
        self.op_bytes[len(self.ops) - 1] = nbytes
        self.op_info.append(("conv_gemm", flops, f"k{k}s{stride}u{up} {p.cin0 + p.cin1}->{p.N} @{out.H}x{out.W}"
                             + (" rowpair" if rp else "")))
        self.n_launch += 1

    def rowpair_ok(self, entry, src0, src1, out, k=3, stride=1, up=0) -> bool:
        """Layer shape the row-pair kernel is built for AND worth it: strips are 128 pixels wide, so narrow or ragged
        rows waste MMA rows (efficiency = W / (strips * 128))."""
        if not self.net.use_rowpair or "w_rp" not in entry or src1 is not None or k != 3 or stride != 1 or up:
            return False
        W = out.W
        return out.H % 2 == 0 and W / (-(-W // 128) * 128) >= 0.74

    def gn_rows(self, entry, src0, src1, out):
        """rows of GroupNorm partial sums per image the kernel chosen for this layer writes"""
        if self.rowpair_ok(entry, src0, src1, out):
            return self.L.idiff_conv3_rowpair_gn_rows(out.H, out.W)
        return self.L.idiff_conv_gemm_gn_rows(out.H, out.W)

    def bind_context(self, crossvec):
        for name, p in self.ctx_slots.items():
            buf = crossvec.get((name, self.B))
            if buf is not None:
                p.bias_img = buf.data_ptr()

    def tiles(self, H, W):
        """rows of GroupNorm partial sums per image written by idiff_conv_gemm (one per epilogue warp and tile)"""
        return self.L.idiff_conv_gemm_gn_rows(H, W)

    def gn_fuse_desc(self, norm, Cc, G, count, eps, t_off=None, tag="gn"):
        """idiff_gn_fuse for a producer launch: the kernel adds its GroupNorm sums as exact integers and its last CTA writes
        the (scale, shift) rows the next layer applies on load -- no idiff_gn_finalize launch."""
        B = self.B
        if self._gn_sums is None:                # one zeroed pair serves every layer of the plan (stream-ordered, self-cleaning)
            self._gn_sums = torch.zeros(B * GN_SLOTS * 32 * 2, dtype=torch.int64, device=self.dev)
            self._gn_arrivals = torch.zeros(4, dtype=torch.int32, device=self.dev)
        sc = self.tmp(tag + "_sc", (B, Cc), torch.float32)
        sh = self.tmp(tag + "_sh", (B, Cc), torch.float32)
        f = GnFuse()
        f.sums, f.arrivals = self._gn_sums.data_ptr(), self._gn_arrivals.data_ptr()
        f.gamma, f.beta = _ptr(norm["g"]), _ptr(norm["b"])
        if t_off is not None:
            f.t_scale = self.ss.data_ptr() + 4 * t_off
            f.t_shift = self.ss.data_ptr() + 4 * (t_off + Cc)
            f.t_ld = 0 if self.shared_time else self.net.S
        f.scale_out, f.shift_out = _ptr(sc), _ptr(sh)
        f.count_per_group, f.eps = count, eps
        self.keep.append(f)
        return f, sc, sh

    def gn_finalize(self, partial, ntile, norm, C, G, count, eps, t_off=None, tag="gn"):
        sc = self.tmp(tag + "_sc", (self.B, C), torch.float32)
        sh = self.tmp(tag + "_sh", (self.B, C), torch.float32)
        L, B = self.L, self.B
        if t_off is not None:
            ts = self.ss.data_ptr() + 4 * t_off
            tb = self.ss.data_ptr() + 4 * (t_off + C)
            t_ld = 0 if self.shared_time else self.net.S
        else:
            ts = tb = None
            t_ld = 0
        args = (_ptr(partial), ntile, _ptr(norm["g"]), _ptr(norm["b"]), ts, tb, t_ld, _ptr(sc), _ptr(sh), B, C, G,
                count, eps)
        self.ops.append(lambda s: check(L.idiff_gn_finalize(*args, s), "gn_finalize"))
        self.op_info.append(("gn_finalize", 0.0, f"C{C}"))
        self.n_launch += 1
        return sc, sh

    def resblock(self, prefix, src0: _Act, src1: Optional[_Act], cout, want_stats=False) -> _Act:
        pk, B, H, W = self.net.pk, self.B, src0.H, src0.W
        cin = src0.C + (src1.C if src1 is not None else 0)
        count = H * W * (cout // GN_GROUPS)
        y1 = self.act(H, W, cout, tmp_name="y1")
        y2 = self.act(H, W, cout, tmp_name="y2")
        if self.net.fuse_gn:
            f1, sc1, sh1 = self.gn_fuse_desc(pk[prefix + ".norm1"], cout, GN_GROUPS, count, 1e-5,
                                             t_off=self.net._ss_off[prefix], tag="gn1")
            self.gemm(src0, src1, pk[prefix + ".conv1"], y1, k=3, gn_fuse=f1)
            f2, sc2, sh2 = self.gn_fuse_desc(pk[prefix + ".norm2"], cout, GN_GROUPS, count, 1e-5, tag="gn2")
            self.gemm(y1, None, pk[prefix + ".conv2"], y2, k=3, a_scale=sc1, a_shift=sh1, a_silu=1, gn_fuse=f2)
        else:
            nt1 = self.gn_rows(pk[prefix + ".conv1"], src0, src1, y1)
            nt2 = self.gn_rows(pk[prefix + ".conv2"], y1, None, y2)
            part1 = self.tmp("gnp1", (B, nt1, GN_GROUPS, 2), torch.float32)
            part2 = self.tmp("gnp2", (B, nt2, GN_GROUPS, 2), torch.float32)
            self.gemm(src0, src1, pk[prefix + ".conv1"], y1, k=3, gn_partial=part1)
            sc1, sh1 = self.gn_finalize(part1, nt1, pk[prefix + ".norm1"], cout, GN_GROUPS, count, 1e-5,
                                        t_off=self.net._ss_off[prefix], tag="gn1")
            self.gemm(y1, None, pk[prefix + ".conv2"], y2, k=3, a_scale=sc1, a_shift=sh1, a_silu=1, gn_partial=part2)
            sc2, sh2 = self.gn_finalize(part2, nt2, pk[prefix + ".norm2"], cout, GN_GROUPS, count, 1e-5, tag="gn2")
        out = self.act(H, W, cout, stats=want_stats)
        if cin == cout:
            L = self.L
            args = (_ptr(y2.t), _ptr(sc2), _ptr(sh2), _ptr(src0.t), _ptr(out.t), _ptr(out.stats), 1e-5, B, H * W, cout)
            self.ops.append(lambda s: check(L.idiff_block_tail(*args, s), "block_tail"))
            self.op_info.append(("block_tail", 0.0, f"C{cout} @{H}x{W}"))
            self.n_launch += 1
        else:
            self.gemm(src0, src1, pk[prefix + ".res_conv"], out, k=1, res0=y2.t, res0_scale=sc2, res0_shift=sh2,
                      out_stats=out.stats, NT=cout)
        self.named[prefix] = out
        return out

    def linear_attn(self, prefix, x: _Act) -> _Act:
        pk, B, H, W, Cc, L = self.net.pk, self.B, x.H, x.W, x.C, self.L
        f = prefix + ".fn"
        to = pk[f + ".to_out"]
        if (f + ".fused") in pk and (H * W) % 128 == 0:
            # q, k, v never touch HBM: the context and output passes read x directly (csrc/linattn_fused.cu)
            fu = pk[f + ".fused"]
            weff = self.tmp("la_weff", (B, Cc * 128))
            nfl = L.idiff_linattn_fused_scratch_floats(B, H * W, Cc)
            scratch = self.tmp("laf_scratch", (nfl,), torch.float32)
            out = self.act(H, W, Cc)
            args = (_ptr(x.t), _ptr(x.stats), _ptr(fu["wq"]), _ptr(fu["wk"]), _ptr(fu["wv"]), _ptr(to["w_f32"]),
                    _ptr(to["bias"]), _ptr(to["g"]), _ptr(weff), _ptr(out.t), _ptr(scratch), B, H * W, Cc, 32 ** -0.5, 1e-5)
            self.ops.append(lambda s: check(L.idiff_linattn_fused(*args, s), "linattn_fused"))
            # algorithmic work: q and k projections, e * x^ context, q * Weff output
            flops = 2.0 * B * H * W * (2 * 128 * Cc + 128 * (Cc + 16) + 128 * Cc)     # k, q, e*x^, q*Weff
            self.op_info.append(("linattn_fused", flops, f"C{Cc} @{H}x{W}"))
            self.op_bytes[len(self.ops) - 1] = 2.0 * B * H * W * Cc * 3 + 8.0 * B * H * W * 2   # 2 reads + 1 write + stats
            self.n_launch += 3
            self.named[prefix] = out
            return out
        qkv = self.act(H, W, 384, tmp_name="la_qkv")
        self.gemm(x, None, pk[f + ".to_qkv"], qkv, k=1, bias=False, row_stats=x.stats, epi=EPI_QSOFTMAX,
                  qscale=32 ** -0.5)
        weff = self.tmp("la_weff", (B, Cc * 128))
        nfl = L.idiff_linattn_scratch_floats(B, H * W)
        scratch = self.tmp("la_scratch", (nfl,), torch.float32)
        args = (_ptr(qkv.t), _ptr(to["w_f32"]), _ptr(weff), _ptr(scratch), B, H * W, Cc)
        self.ops.append(lambda s: check(L.idiff_linattn_context(*args, s), "linattn_context"))
        self.op_info.append(("linattn_context", 2.0 * 2 * B * H * W * 128 * 32, f"C{Cc} @{H}x{W}"))
        self.n_launch += 3
        out = self.act(H, W, Cc)
        entry = dict(N=Cc, NT=Cc, w=weff, bias=to["bias"])
        self.gemm(qkv, None, entry, out, k=1, cin0=128, src0_ld=384, w_image_stride=Cc * 128, epi=EPI_LN_OUT,
                  ln_g=to["g"], res0=x.t, ln_eps=1e-5)
        self.named[prefix] = out
        return out

    def spatial_attn(self, prefix, x: _Act) -> _Act:
        pk, B, H, W, Cc, L = self.net.pk, self.B, x.H, x.W, x.C, self.L
        f = prefix + ".fn"
        rows, HW = B * H * W, H * W
        y = self.act(H, W, Cc, tmp_name="st_y")
        if self.net.fuse_gn_st and HW % 32 == 0:
            # channel LayerNorm + GroupNorm(32) statistics + finalize in one launch
            fz, sc, sh = self.gn_fuse_desc(pk[f + ".norm"], Cc, 32, HW * (Cc // 32), 1e-6, tag="gn32")
            a_lg = (_ptr(x.t), _ptr(pk[prefix + ".prenorm"]["g"]), _ptr(y.t), 1e-5, B, HW, Cc, 32, C.byref(fz))
            self.ops.append(lambda s: check(L.idiff_chan_ln_gn(*a_lg, s), "chan_ln_gn"))
            self.op_info.append(("chan_ln_gn", 0.0, f"C{Cc} @{H}x{W}"))
            self.n_launch += 1
        else:
            a_ln = (_ptr(x.t), _ptr(pk[prefix + ".prenorm"]["g"]), _ptr(y.t), 1e-5, rows, Cc)
            self.ops.append(lambda s: check(L.idiff_chan_ln(*a_ln, s), "chan_ln"))
            self.op_info.append(("chan_ln", 0.0, f"C{Cc} @{H}x{W}"))
            ntile = L.idiff_gn_stats_ntile(HW)
            part = self.tmp("st_gnp", (B, ntile, 32, 2), torch.float32)
            a_gs = (_ptr(y.t), _ptr(part), B, HW, Cc, 32)
            self.ops.append(lambda s: check(L.idiff_gn_stats(*a_gs, s), "gn_stats"))
            self.op_info.append(("gn_stats", 0.0, f"C{Cc} @{H}x{W}"))
            self.n_launch += 2
            sc, sh = self.gn_finalize(part, ntile, pk[f + ".norm"], Cc, 32, HW * (Cc // 32), 1e-6, tag="gn32")
        h0 = self.act(H, W, Cc, stats=True, tmp_name=None)
        self.gemm(y, None, pk[f + ".proj_in"], h0, k=1, a_scale=sc, a_shift=sh, a_silu=0, out_stats=h0.stats)
        qkv = self.act(H, W, 3 * Cc, tmp_name="st_qkv")
        self.gemm(h0, None, pk[f + ".attn1.qkv"], qkv, k=1, row_stats=h0.stats)
        att = self.act(H, W, Cc, tmp_name="st_att")
        a_at = (_ptr(qkv.t), _ptr(att.t), B, HW, Cc // 32, 32 ** -0.5)
        self.ops.append(lambda s: check(L.idiff_self_attention(*a_at, s), "self_attention"))
        self.op_info.append(("self_attention", 2.0 * 2 * B * HW * HW * Cc, f"L{HW}"))
        self.n_launch += 1
        h2 = self.act(H, W, Cc, stats=True)
        self.gemm(att, None, pk[f + ".attn1.to_out"], h2, k=1, res0=h0.t, out_stats=h2.stats, bias_img_slot=prefix)
        ff = self.act(H, W, 4 * Cc, tmp_name="st_ff")
        self.gemm(h2, None, pk[f + ".ff.proj"], ff, k=1, row_stats=h2.stats, epi=EPI_GEGLU, out_ld=4 * Cc)
        h3 = self.act(H, W, Cc, tmp_name="st_h3")
        self.gemm(ff, None, pk[f + ".ff.out"], h3, k=1, res0=h2.t)
        out = self.act(H, W, Cc)
        self.gemm(h3, None, pk[f + ".proj_out"], out, k=1, res0=y.t, res1=x.t)
        self.named[prefix] = out
        return out

    def conv(self, name, x: _Act, cout, k=3, stride=1, up=0, res0=None) -> _Act:
        H = x.H // 2 if stride == 2 else (x.H * 2 if up else x.H)
        W = x.W // 2 if stride == 2 else (x.W * 2 if up else x.W)
        out = self.act(H, W, cout)
        self.gemm(x, None, self.net.pk[name], out, k=k, stride=stride, up=up, res0=res0)
        key = name[:-5] if name.endswith(".conv") else name
        self.named[key + "+init_conv" if res0 is not None else key] = out
        return out

    # ---- whole network
    def _build(self):
        net, B, H, W, L = self.net, self.B, self.H, self.W, self.L
        pk, nf = net.pk, net.nf
        self.x_in = None
        self.eps = self.new((B, 1, H, W), torch.float32)
        Bt = 1 if self.shared_time else B
        self.temb = self.new((Bt, net.td), torch.float32)
        self.ss = self.new((Bt, net.S), torch.float32)
        tm = pk["time"]
        self._time_args = (_ptr(tm["w1t"]), _ptr(tm["b1"]), _ptr(tm["w2t"]), _ptr(tm["b2"]), _ptr(tm["wss"]),
                           _ptr(tm["bss"]), _ptr(self.temb), _ptr(self.ss), Bt, nf, net.S)
        self.n_launch += 2

        x_first = self.act(H, W, nf)
        self.named["init_conv"] = x_first
        self._stem_tail = (_ptr(pk["init_conv"]["w"]), _ptr(x_first.t), B, H, W)
        self.n_launch += 1

        h = x_first
        skips = []
        n = len(net.io)
        for i, (di, do) in enumerate(net.io):
            last = i == n - 1
            h = self.resblock(f"downs.{i}.0", h, None, di)
            skips.append(h)
            h = self.resblock(f"downs.{i}.1", h, None, di, want_stats=not last)
            h = self.spatial_attn(f"downs.{i}.2", h) if last else self.linear_attn(f"downs.{i}.2", h)
            skips.append(h)
            h = self.conv(f"downs.{i}.3", h, do, k=3) if last else self.conv(f"downs.{i}.3", h, do, k=4, stride=2)
        h = self.resblock("mid_block1", h, None, net.dims[-1])
        h = self.spatial_attn("mid_attn", h)
        h = self.resblock("mid_block2", h, None, net.dims[-1])
        for i, (di, do) in enumerate(reversed(net.io)):
            lvl = n - 1 - i
            last = lvl == n - 1
            h = self.resblock(f"ups.{i}.0", h, skips.pop(), do)
            h = self.resblock(f"ups.{i}.1", h, skips.pop(), do, want_stats=not last)
            h = self.spatial_attn(f"ups.{i}.2", h) if last else self.linear_attn(f"ups.{i}.2", h)
            if lvl > 0:
                h = self.conv(f"ups.{i}.3.conv", h, di, k=3, up=1)
            else:       # last up conv: its epilogue also adds the stem output (x + x_ of the final concat, App. A)
                h = self.conv(f"ups.{i}.3", h, di, k=3, res0=x_first.t)
        xa = h
        h = self.resblock("final_res", xa, x_first, nf)
        a_head = (_ptr(h.t), _ptr(pk["final_conv"]["w"]), pk["final_conv"]["bias"], _ptr(self.eps), B, H, W, nf)
        self.ops.append(lambda s: check(L.idiff_head_conv3(*a_head, s), "head_conv3"))
        self.op_info.append(("head_conv3", 2.0 * B * H * W * nf * 9, f"@{H}x{W}"))
        self.n_launch += 1

    def run_timed(self, xt, cond, t_scalar, reps=3):
        """Instrumented replay (bench/profiling): CUDA events around every launch on the launching stream.
        Returns [(kind, label, flops, mean_ms, algorithmic_bytes)] including the time-embedding and stem launches."""
        L = self.L
        stream = torch.cuda.current_stream(self.dev)
        s = stream.cuda_stream
        calls = [("time_embed", "", 0.0, lambda: check(L.idiff_time_embed(None, t_scalar, *self._time_args, s))),
                 ("stem_conv7", f"@{self.H}x{self.W}", 2.0 * self.B * self.H * self.W * 98 * self.net.nf,
                  lambda: check(L.idiff_stem_conv7_tc(xt.data_ptr(), cond.data_ptr(), *self._stem_tail, s)))]
        nbytes = [0.0, 4.0 * self.B * self.H * self.W * 2 + 2.0 * self.B * self.H * self.W * self.net.nf]
        for i, (op, (kind, flops, label)) in enumerate(zip(self.ops, self.op_info)):
            calls.append((kind, label, flops, (lambda op=op: op(s))))
            nbytes.append(self.op_bytes.get(i, 0.0))
        acc = [0.0] * len(calls)
        for _ in range(reps):
            evs = []
            for _, _, _, fn in calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                evs.append((e0, e1))
            torch.cuda.synchronize(self.dev)
            for i, (e0, e1) in enumerate(evs):
                acc[i] += e0.elapsed_time(e1)
        return [(k, lbl, fl, a / reps, nb) for (k, lbl, fl, _), a, nb in zip(calls, acc, nbytes)]

    def run(self, xt, cond, t_dev, t_scalar, skip_time=False):
        L = self.L
        s = torch.cuda.current_stream(self.dev).cuda_stream
        t_ptr = t_dev if (t_dev is None or isinstance(t_dev, int)) else t_dev.data_ptr()
        if not skip_time:
            check(L.idiff_time_embed(t_ptr, t_scalar, *self._time_args, s), "time_embed")
        check(L.idiff_stem_conv7_tc(xt.data_ptr(), cond.data_ptr(), *self._stem_tail, s), "stem_conv7")
        for op in self.ops:
            op(s)
