"""Training step v1 for BASELINE config 5: noise-matching loss, UNet forward + backward, gradient all-reduce, Adam.

What the reference does per step (``models/drift_noise_model.py:242-312``): sample ``t`` and the noisy state, run the
network(s), ``MSELoss`` against the standardised noise (``:157,279``), ``zero_grad`` / ``backward`` / ``Adam.step``
(``:292-296``; Adam with ``lr 2e-5, betas (0.9, 0.99), weight_decay 1e-4``, ``Configurations/config.yml:138-143``), with
the gradients averaged across ranks by ``DistributedDataParallel`` (``:144-146``).  With the ``IRSDE`` of this slice the
single-network form is::

    t, x_t = sde.generate_random_states(x0, mu)              # utils/sde_utils.py:322-338
    loss   = mse(net(x_t, mu, t.squeeze(), image_context=emb), noise)     # noise == sde.get_real_noise(x_t, x0, t), :222

SCOPE OF v1: the hand-written sm_100a kernels cover the INPUT side of the step
(``idiff_random_states``: x_t and the noise target in one pass) and the inference network; the differentiable network
below is plain PyTorch autograd over the SAME parameter names (cuDNN / ATen kernels under bf16 autocast), the
optimizer is torch's fused Adam, and the gradient averaging is ``parallel.GradientAllReducer`` (bucketed asynchronous
NCCL all-reduce launched from backward hooks).  Native dgrad / wgrad kernels are the next step and
are NOT built.  ``TrainableUNet.state_dict()`` loads into ``ConditionalUNet`` unchanged, so trained weights sample on
the tcgen05 path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .parallel import GradientAllReducer
from .unet import param_specs


def _chan_ln(x, g, eps=1e-5):
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * g


class TrainableUNet(torch.nn.Module):
    """Differentiable twin of ``ConditionalUNet`` (architecture: SURVEY.md App. A): one flat parameter table with the
    inference network's key names, forward written functionally over it."""

    def __init__(self, in_nc=2, out_nc=1, nf=64, ch_mult=(1, 2, 4, 4), context_dim=512, down_kernel=4, seed: int = 1):
        super().__init__()
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, ch_mult=tuple(ch_mult), context_dim=context_dim,
                        down_kernel=down_kernel)
        self.nf = nf
        self.dims = [nf] + [nf * m for m in ch_mult]
        self.io = list(zip(self.dims[:-1], self.dims[1:]))
        gen = torch.Generator().manual_seed(seed)
        self._names = []
        self._p = torch.nn.ParameterList()
        for name, shape, init in param_specs(**self.cfg):
            if init == "ones":
                t = torch.ones(shape)
            elif init == "zeros":
                t = torch.zeros(shape)
            else:                                  # nn.Conv2d / nn.Linear default: U(-1/sqrt(fan_in), +)
                t = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(init)
            self._names.append(name)
            self._p.append(torch.nn.Parameter(t))

    # ---- the inference network's weight contract -----------------------------------------------------------
    def table(self) -> Dict[str, torch.nn.Parameter]:
        return dict(zip(self._names, self._p))

    def state_dict(self, *a, **kw):                # same keys as ConditionalUNet.state_dict()
        return {k: v.detach().clone() for k, v in self.table().items()}

    def load_state_dict(self, sd, strict=True):
        tab = self.table()
        missing = [k for k in tab if k not in sd]
        extra = [k for k in sd if k not in tab]
        if strict and (missing or extra):
            raise KeyError(f"load_state_dict: missing {missing[:4]}..., unexpected {extra[:4]}...")
        with torch.no_grad():
            for k, p in tab.items():
                if k in sd:
                    p.copy_(sd[k].to(p.device, p.dtype))
        return self

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(self, xt, cond, time, *unused, image_context=None, **unused_kw):
        P = self.table()
        B, _, H, W = xt.shape
        if image_context is None:
            raise ValueError("image_context is required")
        ctx = image_context if image_context.dim() == 3 else image_context[:, None, :]
        if not torch.is_tensor(time):
            time = torch.tensor([float(time)], device=xt.device)
        time = time.to(xt.device, torch.float32).reshape(-1)
        if time.numel() == 1 and B > 1:
            time = time.expand(B)

        def conv(name, x, stride=1, pad=None):
            w = P[name + ".weight"]
            pad = (w.shape[-1] // 2 if stride == 1 else 1) if pad is None else pad
            return F.conv2d(x, w, P.get(name + ".bias"), stride=stride, padding=pad)

        def lin(name, x):
            return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))

        half = self.nf // 2
        freq = torch.exp(torch.arange(half, device=xt.device, dtype=torch.float32) * (-math.log(10000.0) / (half - 1)))
        arg = time[:, None] * freq[None, :]
        temb = lin("time_lin2", F.gelu(lin("time_lin1", torch.cat([arg.sin(), arg.cos()], dim=-1))))
        st = F.silu(temb)

        def resblock(prefix, x):
            scale, shift = lin(prefix + ".mlp", st)[:, :, None, None].chunk(2, dim=1)
            cout = P[prefix + ".conv1.weight"].shape[0]
            h = F.group_norm(conv(prefix + ".conv1", x), 8, P[prefix + ".norm1.weight"], P[prefix + ".norm1.bias"])
            h = F.silu(h * (scale + 1) + shift)
            h = F.silu(F.group_norm(conv(prefix + ".conv2", h), 8, P[prefix + ".norm2.weight"], P[prefix + ".norm2.bias"]))
            return h + (conv(prefix + ".res_conv", x, pad=0) if x.shape[1] != cout else x)

        def linear_attn(prefix, x):
            b, c, h, w = x.shape
            f = prefix + ".fn"
            y = _chan_ln(x, P[prefix + ".prenorm.g"])
            q, k, v = F.conv2d(y, P[f + ".to_qkv.weight"]).chunk(3, dim=1)
            q, k, v = (t.reshape(b, 4, 32, h * w) for t in (q, k, v))
            q = q.softmax(dim=-2) * 32 ** -0.5
            k = k.softmax(dim=-1)
            cmat = torch.einsum("bhdn,bhen->bhde", k, v / (h * w))
            o = torch.einsum("bhde,bhdn->bhen", cmat, q).reshape(b, 128, h, w)
            o = F.conv2d(o, P[f + ".to_out.weight"], P[f + ".to_out.bias"])
            return x + _chan_ln(o, P[f + ".out_norm.g"])

        def mha(q, k, v, heads):
            b, n, c = q.shape
            sp = lambda t: t.reshape(b, t.shape[1], heads, c // heads).transpose(1, 2)
            return F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(b, n, c)

        def spatial_attn(prefix, x):
            b, c, h, w = x.shape
            f = prefix + ".fn"
            x0 = _chan_ln(x, P[prefix + ".prenorm.g"])
            y = conv(f + ".proj_in", F.group_norm(x0, 32, P[f + ".norm.weight"], P[f + ".norm.bias"], eps=1e-6), pad=0)
            y = y.reshape(b, c, h * w).transpose(1, 2)
            ln = lambda n, t: F.layer_norm(t, (c,), P[f + f".{n}.weight"], P[f + f".{n}.bias"])
            z = ln("norm1", y)
            y = lin(f + ".attn1.to_out", mha(lin(f + ".attn1.to_q", z), lin(f + ".attn1.to_k", z), lin(f + ".attn1.to_v", z), c // 32)) + y
            z = ln("norm2", y)
            y = lin(f + ".attn2.to_out", mha(lin(f + ".attn2.to_q", z), lin(f + ".attn2.to_k", ctx), lin(f + ".attn2.to_v", ctx), c // 32)) + y
            a, gate = lin(f + ".ff.proj", ln("norm3", y)).chunk(2, dim=-1)
            y = lin(f + ".ff.out", a * F.gelu(gate)) + y
            y = y.transpose(1, 2).reshape(b, c, h, w)
            return x + conv(f + ".proj_out", y, pad=0) + x0

        x = torch.cat([xt - cond, cond], dim=1)
        ph, pw = (-H) % 16, (-W) % 16
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph), mode="reflect")
        x = conv("init_conv", x)
        x_first = x
        skips = []
        n = len(self.io)
        for i in range(n):
            last = i == n - 1
            x = resblock(f"downs.{i}.0", x)
            skips.append(x)
            x = resblock(f"downs.{i}.1", x)
            x = spatial_attn(f"downs.{i}.2", x) if last else linear_attn(f"downs.{i}.2", x)
            skips.append(x)
            x = conv(f"downs.{i}.3", x) if last else conv(f"downs.{i}.3", x, stride=2)
        x = resblock("mid_block1", x)
        x = spatial_attn("mid_attn", x)
        x = resblock("mid_block2", x)
        for i in range(n):
            lvl = n - 1 - i
            x = resblock(f"ups.{i}.0", torch.cat([x, skips.pop()], dim=1))
            x = resblock(f"ups.{i}.1", torch.cat([x, skips.pop()], dim=1))
            x = spatial_attn(f"ups.{i}.2", x) if lvl == n - 1 else linear_attn(f"ups.{i}.2", x)
            if lvl > 0:
                x = conv(f"ups.{i}.3.conv", F.interpolate(x, scale_factor=2, mode="nearest"))
            else:
                x = conv(f"ups.{i}.3", x)
        x = resblock("final_res", torch.cat([x + x_first, x_first], dim=1))
        x = conv("final_conv", x)
        return x[..., :H, :W]


class NoiseMatchingTrainer:
    """One training step of config 5 on one rank: fused state sampling -> forward -> MSE -> backward (bucketed gradient
    all-reduce launched from backward hooks when ``torch.distributed`` is initialised) -> Adam."""

    def __init__(self, net: TrainableUNet, sde, lr=2e-5, betas=(0.9, 0.99), weight_decay=1e-4, process_group=None,
                 bucket_mb: float = 25.0, comm_dtype: Optional[torch.dtype] = None,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16):
        import torch.distributed as dist
        self.net, self.sde = net, sde
        self.autocast_dtype = autocast_dtype
        dev = next(net.parameters()).device
        self.opt = torch.optim.Adam(net.parameters(), lr=lr, betas=betas, weight_decay=weight_decay,
                                    fused=dev.type == "cuda")
        self.reducer = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            self.reducer = GradientAllReducer(net.parameters(), bucket_mb=bucket_mb, process_group=process_group,
                                              comm_dtype=comm_dtype).attach()
        self.last = {}

    def loss(self, x0, mu, image_context, timesteps=None):
        # x_t and the noise target from ONE fused kernel (idiff_random_states); t ~ U{1..T} per sample (:322-330)
        t, x_t = self.sde.generate_random_states(x0, mu, timesteps=timesteps)
        target = self.sde.last_noises
        dev_type = x_t.device.type
        with torch.autocast(dev_type, dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            pred = self.net(x_t, mu.to(x_t.device), t.reshape(-1).to(x_t.device), image_context=image_context)
        return F.mse_loss(pred.float(), target)               # nn.MSELoss(reduction='mean'), drift_noise_model.py:157

    def step(self, x0, mu, image_context, timesteps=None):
        if self.reducer is not None:
            self.reducer.zero_grad()                          # :292 -- gradients live in the reducer's bucket buffers
        else:
            self.opt.zero_grad(set_to_none=True)
        loss = self.loss(x0, mu, image_context, timesteps)
        loss.backward()                                       # :294 -- the hooks launch the bucket all-reduces
        if self.reducer is not None:
            self.reducer.wait()                               # averaged gradients, what DDP leaves in .grad
        self.opt.step()                                       # :295
        return loss.detach()

    def sync_to(self, unet) -> None:
        """Copy the current weights into a ``ConditionalUNet`` (sampling on the tcgen05 path)."""
        unet.load_state_dict(self.net.state_dict())
