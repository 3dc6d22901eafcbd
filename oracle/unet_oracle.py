"""TEST INFRASTRUCTURE -- fp32 PyTorch oracle of the drift/noise UNet.

PARITY UNPINNED against upstream: the reference snapshot imports
``models/modules/MSM_degEmb_Unet.py`` (``models/drift_noise_model.py:18-22``) but does
not ship it.  What IS pinned by the reference and followed here:

* call convention ``model(x, mu, t, **kwargs) -> eps``       utils/sde_utils.py:198
* wrapper convention ``net(a, b, t[B], names, text_encoder, image_context=A_emb)``
                                                              models/drift_noise_model.py:250-268
* ``in_nc 2, nf 64, ch_mult [1,2,4,4], context_dim 512``      Configurations/config.yml:109-113
* image embedding ``[B,1,512]``                               data/MedSpeckle.py:46

Topology (App. A): cat([x-mu, mu]) -> 7x7 conv -> 4 down levels of two time-modulated
ResBlocks (GroupNorm-8, scale/shift, SiLU) + attention -> mid -> 4 up levels with skip
concatenation -> final ResBlock -> 3x3 conv.  Levels 0-2 use linear attention, level 3 and
the middle use a SpatialTransformer (self-attn, cross-attn to the image embedding, GEGLU).

The state_dict key names of this module ARE the weight-naming contract of the CUDA path
(``instancediff_b200.unet`` consumes exactly these keys).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32)
                         * (-math.log(10000.0) / (half - 1)))
        arg = t.float()[:, None] * freq[None, :]
        return torch.cat([arg.sin(), arg.cos()], dim=-1)


class ChanLayerNorm(nn.Module):
    """LayerNorm over the channel axis of an NCHW tensor, learned gain only."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))

    def forward(self, x):
        var = torch.var(x, dim=1, unbiased=False, keepdim=True)
        mean = torch.mean(x, dim=1, keepdim=True)
        return (x - mean) / (var + self.eps).sqrt() * self.g


class ResBlock(nn.Module):
    def __init__(self, cin, cout, time_dim, groups=8):
        super().__init__()
        self.mlp = nn.Linear(time_dim, 2 * cout)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm1 = nn.GroupNorm(groups, cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout)
        self.res_conv = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x, temb):
        ss = self.mlp(F.silu(temb))[:, :, None, None]
        scale, shift = ss.chunk(2, dim=1)
        h = self.norm1(self.conv1(x)) * (scale + 1) + shift
        h = F.silu(h)
        h = F.silu(self.norm2(self.conv2(h)))
        return h + self.res_conv(x)


class LinearAttention(nn.Module):
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)
        self.out_norm = ChanLayerNorm(dim)

    def forward(self, x):
        b, c, h, w = x.shape
        q, k, v = self.to_qkv(x).chunk(3, dim=1)
        q, k, v = (t.reshape(b, self.heads, self.dim_head, h * w) for t in (q, k, v))
        q = q.softmax(dim=-2) * self.scale
        k = k.softmax(dim=-1)
        v = v / (h * w)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q)
        out = out.reshape(b, self.heads * self.dim_head, h, w)
        return self.out_norm(self.to_out(out))


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=32):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Linear(inner, query_dim)

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, _ = x.shape
        h = self.heads
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        q, k, v = (t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3) for t in (q, k, v))
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v)
        out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class GEGLUFeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.proj = nn.Linear(dim, dim * mult * 2)
        self.out = nn.Linear(dim * mult, dim)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return self.out(a * F.gelu(gate))


class SpatialTransformer(nn.Module):
    def __init__(self, dim, context_dim, dim_head=32):
        super().__init__()
        heads = dim // dim_head
        self.norm = nn.GroupNorm(32, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = GEGLUFeedForward(dim)
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, context):
        b, c, h, w = x.shape
        y = self.proj_in(self.norm(x))
        y = y.reshape(b, c, h * w).permute(0, 2, 1)
        y = self.attn1(self.norm1(y)) + y
        y = self.attn2(self.norm2(y), context) + y
        y = self.ff(self.norm3(y)) + y
        y = y.permute(0, 2, 1).reshape(b, c, h, w)
        return self.proj_out(y) + x


class AttnBlock(nn.Module):
    """x + Attn(PreNorm(x)); ``kind`` is 'linear' or 'spatial'."""

    def __init__(self, dim, kind, context_dim):
        super().__init__()
        self.kind = kind
        self.prenorm = ChanLayerNorm(dim)
        self.fn = LinearAttention(dim) if kind == "linear" else SpatialTransformer(dim, context_dim)

    def forward(self, x, context):
        y = self.prenorm(x)
        y = self.fn(y) if self.kind == "linear" else self.fn(y, context)
        return x + y


class Upsample(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class OracleUNet(nn.Module):
    def __init__(self, in_nc=2, out_nc=1, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512,
                 down_kernel=4):
        super().__init__()
        assert down_kernel in (3, 4)
        self.depth = len(ch_mult)
        self.out_nc = out_nc
        time_dim = nf * 4
        dims = [nf] + [nf * m for m in ch_mult]
        io = list(zip(dims[:-1], dims[1:]))

        self.init_conv = nn.Conv2d(in_nc, nf, 7, padding=3)
        self.time_pos = SinusoidalPosEmb(nf)
        self.time_lin1 = nn.Linear(nf, time_dim)
        self.time_lin2 = nn.Linear(time_dim, time_dim)

        self.downs = nn.ModuleList()
        for i, (di, do) in enumerate(io):
            last = i == len(io) - 1
            kind = "spatial" if last else "linear"
            down = (nn.Conv2d(di, do, 3, padding=1) if last else
                    nn.Conv2d(di, do, down_kernel, stride=2, padding=1))
            self.downs.append(nn.ModuleList([
                ResBlock(di, di, time_dim), ResBlock(di, di, time_dim),
                AttnBlock(di, kind, context_dim), down]))

        mid = dims[-1]
        self.mid_block1 = ResBlock(mid, mid, time_dim)
        self.mid_attn = AttnBlock(mid, "spatial", context_dim)
        self.mid_block2 = ResBlock(mid, mid, time_dim)

        self.ups = nn.ModuleList()
        for i, (di, do) in enumerate(reversed(io)):
            lvl = len(io) - 1 - i
            kind = "spatial" if lvl == len(io) - 1 else "linear"
            up = Upsample(do, di) if lvl > 0 else nn.Conv2d(do, di, 3, padding=1)
            self.ups.append(nn.ModuleList([
                ResBlock(do + di, do, time_dim), ResBlock(do + di, do, time_dim),
                AttnBlock(do, kind, context_dim), up]))

        self.final_res = ResBlock(nf * 2, nf, time_dim)
        self.final_conv = nn.Conv2d(nf, out_nc, 3, padding=1)

    def time_embedding(self, time, batch, device):
        if not torch.is_tensor(time):
            time = torch.tensor([float(time)], device=device)
        time = time.to(device=device, dtype=torch.float32).reshape(-1)
        if time.numel() == 1 and batch > 1:
            time = time.expand(batch)
        return self.time_lin2(F.gelu(self.time_lin1(self.time_pos(time))))

    def forward(self, xt, cond, time, *unused, image_context=None, **unused_kw):
        b, _, H, W = xt.shape
        x = torch.cat([xt - cond, cond], dim=1)
        mult = 2 ** self.depth
        ph, pw = (-H) % mult, (-W) % mult
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph), mode="reflect")
        ctx = image_context
        if ctx is not None and ctx.dim() == 2:
            ctx = ctx[:, None, :]
        x = self.init_conv(x)
        x_first = x
        temb = self.time_embedding(time, b, x.device)

        skips = []
        for res1, res2, attn, down in self.downs:
            x = res1(x, temb)
            skips.append(x)
            x = res2(x, temb)
            x = attn(x, ctx)
            skips.append(x)
            x = down(x)

        x = self.mid_block1(x, temb)
        x = self.mid_attn(x, ctx)
        x = self.mid_block2(x, temb)

        for res1, res2, attn, up in self.ups:
            x = res1(torch.cat([x, skips.pop()], dim=1), temb)
            x = res2(torch.cat([x, skips.pop()], dim=1), temb)
            x = attn(x, ctx)
            x = up(x)

        x = self.final_res(torch.cat([x + x_first, x_first], dim=1), temb)
        x = self.final_conv(x)
        return x[..., :H, :W]


def make_oracle_unet(seed=1, **cfg) -> OracleUNet:
    """Random-init network under a fixed seed (testUM.py:43 uses seed 1)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = OracleUNet(**cfg)
    torch.random.set_rng_state(gen_state)
    return net.eval().requires_grad_(False)


def algorithmic_flops(H, W, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512):
    """2*MAC count of one forward per image, split (conv, attention) as in App. A."""
    dims = [nf] + [nf * m for m in ch_mult]
    io = list(zip(dims[:-1], dims[1:]))
    td = nf * 4
    conv = attn = 0.0

    def c(hw, cin, cout, k):
        return 2.0 * hw * cin * cout * k * k

    def res(hw, cin, cout):
        f = c(hw, cin, cout, 3) + c(hw, cout, cout, 3) + 2.0 * td * 2 * cout
        if cin != cout:
            f += c(hw, cin, cout, 1)
        return f

    def lin(hw, d):
        return c(hw, d, 384, 1) + 2 * (2.0 * hw * 128 * 32) + c(hw, 128, d, 1)

    def spat(hw, d):
        f = c(hw, d, d, 1) * 2                       # proj_in / proj_out
        f += 4 * c(hw, d, d, 1)                      # self q,k,v,out
        f += 2 * 2.0 * hw * hw * d                   # QK^T, PV
        f += 2 * c(hw, d, d, 1) + 2 * 2.0 * context_dim * d   # cross q,out + k,v of 1 token
        f += 2 * 2.0 * hw * 1 * d                    # cross sim + weighted sum (1 key)
        f += c(hw, d, 8 * d, 1) + c(hw, 4 * d, d, 1)
        return f

    hw = H * W
    conv += c(hw, 2, nf, 7) + 2.0 * nf * td + 2.0 * td * td
    n = len(io)
    for i, (di, do) in enumerate(io):
        conv += 2 * res(hw, di, di)
        attn += spat(hw, di) if i == n - 1 else lin(hw, di)
        if i == n - 1:
            conv += c(hw, di, do, 3)
        else:
            hw //= 4
            conv += c(hw, di, do, 4)
    conv += 2 * res(hw, dims[-1], dims[-1])
    attn += spat(hw, dims[-1])
    for i, (di, do) in enumerate(reversed(io)):
        lvl = n - 1 - i
        conv += 2 * res(hw, do + di, do)
        attn += spat(hw, do) if lvl == n - 1 else lin(hw, do)
        if lvl > 0:
            hw *= 4
        conv += c(hw, do, di, 3)
    conv += res(hw, 2 * nf, nf) + c(hw, nf, 1, 3)
    return conv, attn
