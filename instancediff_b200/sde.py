"""Drop-in ``SDE`` / ``IRSDE`` for the reference's ``utils/sde_utils.py`` on fused sm_100a kernels.

Same constructor, attributes and method names as the reference (``utils/sde_utils.py:10-75`` for the
base class, ``:81-343`` for ``IRSDE``), so ``from instancediff_b200 import IRSDE`` replaces
``from utils import IRSDE``.  What changes underneath:

* ``reverse_sde_step`` / ``reverse_sde_step_mean`` / ``noise_state`` run ONE fused kernel
  (``idiff_sde_step`` / ``idiff_noise_state``) instead of the ~12 ATen launches of
  ``:45-46,178-179,184-185,187-188`` -- same fp32 operation order, bit-identical results.
* ``reverse_sde`` drives that kernel directly from the predicted noise (score never materialised)
  and, when the model is a ``ConditionalUNet``, captures one whole step (schedule-row select ->
  UNet forward -> fused update with in-kernel Philox noise) in a CUDA graph and replays it T times.
* The hot functions have NO CPU path: CPU tensors raise ``IdiffError``.

Noise: by default ``z = torch.randn_like(x)`` from torch's global generator exactly like the
reference (``:185``).  ``sde.noise_source`` may be set to ``"philox"`` (drawn inside the kernel) or
to a callable ``f(t, x) -> z`` (pre-drawn noise for parity tests).
"""
from __future__ import annotations

import abc
import ctypes as C
import math
import os
from typing import Callable, Optional, Union

import torch

from . import _lib
from ._lib import IdiffError, check


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _as_i64(v: int) -> int:
    """uint64 bit pattern as the int64 torch stores (the kernel reads it back as unsigned)."""
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _require_cuda_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not torch.is_tensor(t) or not t.is_cuda:
        raise IdiffError(f"{name}: the fused SDE kernels need CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float32:
        raise IdiffError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


class SDE(abc.ABC):
    """Base surface of utils/sde_utils.py:10-75."""

    def __init__(self, T, device=None):
        self.T = T
        self.dt = 1 / T
        self.device = device

    @abc.abstractmethod
    def drift(self, x, t): ...

    @abc.abstractmethod
    def dispersion(self, x, t): ...

    @abc.abstractmethod
    def sde_reverse_drift(self, x, score, t): ...

    @abc.abstractmethod
    def ode_reverse_drift(self, x, score, t): ...

    @abc.abstractmethod
    def score_fn(self, x, t): ...

    # composed steps (:38-49) -- generic versions; IRSDE overrides the two hot ones with fused kernels
    def forward_step(self, x, t):
        return x + self.drift(x, t) + self.dispersion(x, t)

    def reverse_sde_step_mean(self, x, score, t):
        return x - self.sde_reverse_drift(x, score, t)

    def reverse_sde_step(self, x, score, t):
        return x - self.sde_reverse_drift(x, score, t) - self.dispersion(x, t)

    def reverse_ode_step(self, x, score, t):
        return x - self.ode_reverse_drift(x, score, t)

    def forward(self, x0, T=-1):
        T = self.T if T < 0 else T
        x = x0.clone()
        for t in range(1, T + 1):
            x = self.forward_step(x, t)
        return x

    def reverse_sde(self, xt, T=-1):
        T = self.T if T < 0 else T
        x = xt.clone()
        for t in reversed(range(1, T + 1)):
            x = self.reverse_sde_step(x, self.score_fn(x, t), t)
        return x

    def reverse_ode(self, xt, T=-1):
        T = self.T if T < 0 else T
        x = xt.clone()
        for t in reversed(range(1, T + 1)):
            x = self.reverse_ode_step(x, self.score_fn(x, t), t)
        return x


def _theta_table(schedule: str, n_steps: int) -> torch.Tensor:
    """theta_0..theta_T on the CPU in fp32; arithmetic order of utils/sde_utils.py:94-124."""
    if schedule == "cosine":
        n = n_steps + 2
        grid = torch.linspace(0, n, n + 1, dtype=torch.float32)
        acp = torch.cos(((grid / n) + 0.008) / 1.008 * math.pi * 0.5) ** 2
        acp = acp / acp[0]
        return 1 - acp[1:-1]
    if schedule == "linear":
        n = n_steps + 1
        k = 1000 / n
        return torch.linspace(k * 0.0001, k * 0.02, n, dtype=torch.float32)
    if schedule == "constant":
        return torch.ones(n_steps + 1, dtype=torch.float32)
    print("Not implemented such schedule yet!!!")          # the reference prints, then NameError (:141-144)
    raise NameError(f"name 'thetas' is not defined (schedule {schedule!r})")


class IRSDE(SDE):
    """Mean-reverting SDE sampler; timesteps run 1..T, state 0 is never used (utils/sde_utils.py:81-84)."""

    def __init__(self, max_sigma, T=100, sample_T=-1, schedule="cosine", eps=0.01, device=None):
        super().__init__(T, device)
        self.max_sigma = max_sigma / 255 if max_sigma >= 1 else max_sigma        # :87
        self.sample_T = self.T if sample_T < 0 else sample_T                       # :88
        self.sample_scale = self.T / self.sample_T                                 # :89
        self.noise_source: Union[None, str, Callable] = None
        self.philox_seed = 0
        self.philox_offset = 0          # global element index of this shard's first element
        self.use_cuda_graph = True
        self._graph_cache = {}
        self.graph_cache_size = 4       # captured loops kept alive (one per state shape / model / device)
        self._table_dev = None
        self._initialize(self.max_sigma, self.sample_T, schedule, eps)

    # ------------------------------------------------------------------ schedule (:92-155)
    def _initialize(self, max_sigma, T, schedule, eps=0.01):
        thetas = _theta_table(schedule, T)
        sigmas = torch.sqrt(max_sigma ** 2 * 2 * thetas)
        thetas_cumsum = torch.cumsum(thetas, dim=0) - thetas[0]
        self.dt = -1 / thetas_cumsum[-1] * math.log(eps)              # 0-d CPU tensor, as upstream
        sigma_bars = torch.sqrt(max_sigma ** 2 * (1 - torch.exp(-2 * thetas_cumsum * self.dt)))
        self._host = (thetas, sigmas, sigma_bars)

        self.thetas = thetas.to(self.device)
        self.sigmas = sigmas.to(self.device)
        self.thetas_cumsum = thetas_cumsum.to(self.device)
        self.sigma_bars = sigma_bars.to(self.device)
        self.mu = 0.
        self.model = None

    def _coef_table(self, dev) -> torch.Tensor:
        """[T+1, 8] fp32 device table {theta, sigma, sigma_bar, dt, sqrt(dt), t, 0, 0} (idiff_sde_pack_table)."""
        if self._table_dev is None or self._table_dev.device != dev:
            th, sg, sb = (t.contiguous() for t in self._host)
            host = torch.empty(th.numel(), 8, dtype=torch.float32)
            check(_lib.lib().idiff_sde_pack_table(th.data_ptr(), sg.data_ptr(), sb.data_ptr(), th.numel(),
                                                  float(self.dt), math.sqrt(self.dt), host.data_ptr()),
                  "sde_pack_table")
            self._table_dev = host.to(dev)
        return self._table_dev

    # ------------------------------------------------------------------ setters (:160-165)
    def set_mu(self, mu):
        self.mu = mu

    def set_model(self, model):
        self.model = model

    # ------------------------------------------------------------------ closed forms (:169-188)
    def mu_bar(self, x0, t):
        return self.mu + (x0 - self.mu) * torch.exp(-self.thetas_cumsum[t] * self.dt)

    def sigma_bar(self, t):
        return self.sigma_bars[t]

    def drift(self, x, t):
        return self.thetas[t] * (self.mu - x) * self.dt

    def sde_reverse_drift(self, x, score, t):
        return (self.thetas[t] * (self.mu - x) - self.sigmas[t] ** 2 * score) * self.dt

    def ode_reverse_drift(self, x, score, t):
        return (self.thetas[t] * (self.mu - x) - 0.5 * self.sigmas[t] ** 2 * score) * self.dt

    def dispersion(self, x, t):
        return self.sigmas[t] * (self._draw(t, x) * math.sqrt(self.dt)).to(self.device)

    def get_score_from_noise(self, noise, t):
        return -noise / self.sigma_bar(t)

    # ------------------------------------------------------------------ fused hot functions
    def _draw(self, t, x):
        if callable(self.noise_source):
            return self.noise_source(t, x)
        if self.noise_source == "philox":
            z = torch.empty_like(x)
            check(_lib.lib().idiff_philox_normal(z.data_ptr(), self.philox_seed, self.philox_offset, int(t),
                                                 z.numel(), _stream(x.device)), "philox_normal")
            return z
        return torch.randn_like(x)

    def _mu_ptr(self, x):
        mu = self.mu
        if torch.is_tensor(mu):
            if tuple(mu.shape) != tuple(x.shape):
                mu = mu.expand_as(x)
            mu = _require_cuda_f32("mu", mu)
            return mu, mu.data_ptr()
        if mu == 0:
            return None, None                   # the reference's default mu = 0. (:152)
        mu = torch.full_like(x, float(mu))
        return mu, mu.data_ptr()

    def _row_index(self, t) -> int:
        """Schedule row of step ``t`` with the reference's own indexing rules (``self.thetas[t]``, :179): Python
        negative indices wrap, anything outside the [sample_T + 1]-row tables raises IndexError -- the row becomes a
        raw device pointer below, so it is checked here."""
        n = self._host[0].numel()
        t = int(t)
        if not -n <= t < n:
            raise IndexError(f"index {t} is out of bounds for the schedule tables of size {n}")
        return t + n if t < 0 else t

    def _check_T(self, T) -> int:
        """A loop ``for t in reversed(range(1, T + 1))`` indexes rows 1..T (:248): T beyond the tables is the
        reference's IndexError, not an out-of-bounds device read."""
        T = int(T)
        n = self._host[0].numel()
        if T >= n:
            raise IndexError(f"index {T} is out of bounds for the schedule tables of size {n}")
        return T

    def _fused_step(self, x, e, t, *, is_score, with_noise, out=None, ode=False):
        """x - drift - dispersion in one pass (replaces :45-46 + :178-179 + :184-185 + :187-188); ``ode`` selects the
        probability-flow drift (:48-49 + :181-182, no dispersion)."""
        x = _require_cuda_f32("x", x)
        e = _require_cuda_f32("score/noise", e)
        if tuple(e.shape) != tuple(x.shape):
            raise IdiffError(f"shape mismatch {tuple(e.shape)} vs {tuple(x.shape)}")
        out = torch.empty_like(x) if out is None else out
        mu_keep, mu_ptr = self._mu_ptr(x)
        table = self._coef_table(x.device)
        coef_ptr = table.data_ptr() + 32 * self._row_index(t)
        z_ptr, philox, z = None, 0, None
        if with_noise:
            if self.noise_source == "philox":
                philox = 1
            else:
                z = _require_cuda_f32("z", self._draw(t, x))
                z_ptr = z.data_ptr()
        check(_lib.lib().idiff_sde_step(out.data_ptr(), x.data_ptr(), e.data_ptr(), mu_ptr, z_ptr, coef_ptr,
                                        (1 if is_score else 0) | (2 if ode else 0), philox, self.philox_seed,
                                        self.philox_offset, x.numel(), _stream(x.device)), "sde_step")
        return out

    def reverse_sde_step(self, x, score, t):              # :45-46
        return self._fused_step(x, score, t, is_score=True, with_noise=True)

    def reverse_sde_step_mean(self, x, score, t):         # :41-42
        return self._fused_step(x, score, t, is_score=True, with_noise=False)

    def reverse_ode_step(self, x, score, t):              # :48-49
        return self._fused_step(x, score, t, is_score=True, with_noise=False, ode=True)

    def noise_state(self, tensor):                        # :340-341
        mu = _require_cuda_f32("tensor", tensor)
        out = torch.empty_like(mu)
        L = _lib.lib()
        if self.noise_source == "philox":
            check(L.idiff_noise_state(out.data_ptr(), mu.data_ptr(), None, float(self.max_sigma), 1,
                                      self.philox_seed, self.philox_offset, mu.numel(), _stream(mu.device)),
                  "noise_state")
        else:
            z = _require_cuda_f32("z", self._draw(0, mu))
            check(L.idiff_noise_state(out.data_ptr(), mu.data_ptr(), z.data_ptr(), float(self.max_sigma), 0, 0, 0,
                                      mu.numel(), _stream(mu.device)), "noise_state")
        return out

    # ------------------------------------------------------------------ model calls (:190-203)
    def score_fn_(self, x, t, scale=1.0):
        x0 = self.model(x, self.mu, t * scale)
        return -(x - self.mu_bar(x0, t)) / self.sigma_bar(t) ** 2

    def score_fn(self, x, t, scale=1.0, **kwargs):
        noise = self.model(x, self.mu, t * scale, **kwargs)
        return self.get_score_from_noise(noise, t)

    def noise_fn(self, x, t, scale=1.0, **kwargs):
        return self.model(x, self.mu, t * scale, **kwargs)

    # ------------------------------------------------------------------ analytic helpers (:206-231)
    def reverse_optimum_step(self, xt, x0, t):
        A = torch.exp(-self.thetas[t] * self.dt)
        B = torch.exp(-self.thetas_cumsum[t] * self.dt)
        Cc = torch.exp(-self.thetas_cumsum[t - 1] * self.dt)
        term1 = A * (1 - Cc ** 2) / (1 - B ** 2)
        term2 = Cc * (1 - A ** 2) / (1 - B ** 2)
        return term1 * (xt - self.mu) + term2 * (x0 - self.mu) + self.mu

    def sigma(self, t):
        return self.sigmas[t]

    def theta(self, t):
        return self.thetas[t]

    def get_real_noise(self, xt, x0, t):
        return (xt - self.mu_bar(x0, t)) / self.sigma_bar(t)

    def get_real_score(self, xt, x0, t):
        return -(xt - self.mu_bar(x0, t)) / self.sigma_bar(t) ** 2

    def get_init_state_from_noise(self, xt, noise, t):
        A = torch.exp(self.thetas_cumsum[t] * self.dt)
        return (xt - self.mu - self.sigma_bar(t) * noise) * A + self.mu

    # ------------------------------------------------------------------ loops (:233-316)
    def forward(self, x0, T=-1, save_dir="forward_state"):
        T = self.T if T < 0 else T
        x = x0.clone()
        for t in range(1, T + 1):
            x = self.forward_step(x, t)
            self._save_state(x, save_dir, f"state_{t}.png", dim=0)
        return x

    @staticmethod
    def _save_state(x, save_dir, name, dim):
        import torchvision.utils as tvutils
        os.makedirs(save_dir, exist_ok=True)
        x_L, x_R = x.chunk(2, dim=1)
        tvutils.save_image(torch.cat([x_L, x_R], dim=dim).data, f"{save_dir}/{name}", normalize=False)

    def reverse_sde(self, xt, T=-1, save_states=False, save_dir="sde_state", **kwargs):
        """The hot loop (:244-261): T sequential (model forward, fused update) pairs, no host sync inside."""
        T = self._check_T(self.sample_T if T < 0 else T)
        xt = _require_cuda_f32("xt", xt)
        if T == 0 or xt.numel() == 0:                  # empty loop / empty batch: the reference returns the clone (:247)
            return xt.clone()
        if self._graph_eligible(xt, save_states, kwargs):
            return self._reverse_sde_graph(xt, T, kwargs)
        x = xt.clone()
        fast_model = hasattr(self.model, "forward_into") and set(kwargs) <= {"image_context"}
        for t in reversed(range(1, T + 1)):
            if fast_model:
                noise = self.model.forward_into(x, self._mu_tensor(x), t * self.sample_scale,
                                                kwargs.get("image_context"))
            else:
                noise = self.model(x, self.mu, t * self.sample_scale, **kwargs)
            x = self._fused_step(x, noise, t, is_score=False, with_noise=True, out=x)
            if save_states:
                interval = self.T // 100
                if t % interval == 0:
                    self._save_state(x, save_dir, f"state_{t // interval}.png", dim=3)
        _lib.watchdog()
        return x

    def _mu_tensor(self, x):
        mu = self.mu
        if torch.is_tensor(mu):
            return _require_cuda_f32("mu", mu if tuple(mu.shape) == tuple(x.shape) else mu.expand_as(x))
        return torch.full_like(x, float(mu))

    # ---- CUDA-graph replay of one whole step ---------------------------------------------------
    def _graph_eligible(self, xt, save_states, kwargs):
        # noise_source None = torch.randn_like from torch's global CUDA generator (:185): that generator is
        # capture-aware (its Philox offset is a graph input), so the reference's default noise lives inside the graph
        return (self.use_cuda_graph and not save_states and (self.noise_source is None or self.noise_source == "philox")
                and hasattr(self.model, "forward_into") and set(kwargs) == {"image_context"}
                and torch.is_tensor(self.mu))

    def _reverse_sde_graph(self, xt, T, kwargs, ode=False):
        dev = xt.device
        ctx = kwargs["image_context"]
        # the Philox stream {seed, offset} lives in device memory, so the captured graph does not depend on it:
        # a data-set loop (one reverse process per item, each with its own offset) replays ONE graph
        off4 = int(self.philox_offset) % 4 == 0
        torch_noise = (not ode) and self.noise_source is None
        key = (tuple(xt.shape), id(self.model), getattr(self.model, "_version", 0), str(dev), off4, ode, torch_noise)
        st = self._graph_cache.pop(key, None)
        L = _lib.lib()
        if st is None:
            st = dict(x=torch.empty_like(xt), mu=torch.empty_like(xt), ctx=ctx.detach().clone().to(dev),
                      counter=torch.zeros(1, dtype=torch.int32, device=dev),
                      row=torch.zeros(8, dtype=torch.float32, device=dev),
                      time=torch.zeros(1, dtype=torch.float32, device=dev),
                      rng=torch.zeros(2, dtype=torch.int64, device=dev), graph=None,
                      z=torch.empty_like(xt) if torch_noise else None,
                      model=self.model)                 # keeps id(model) in the key unique while the entry lives
            if hasattr(self.model, "time_table"):
                # the network's time conditioning depends on t only: one row per step, built once per captured loop
                # (model time of step t: float32(t) * float32(sample_scale), what idiff_step_select publishes)
                scale32 = torch.tensor(self.sample_scale, dtype=torch.float32)
                times = [float(torch.tensor(float(t), dtype=torch.float32) * scale32) for t in range(self._host[0].numel())]
                st["ss_table"] = self.model.time_table(times)
                st["ss_slot"] = self.model.time_slot(*(xt.shape[0], xt.shape[2], xt.shape[3]))
            while len(self._graph_cache) >= self.graph_cache_size:      # oldest entry first (dicts keep order)
                self._graph_cache.pop(next(iter(self._graph_cache)))
        self._graph_cache[key] = st                                     # most recently used last
        st["rng"].copy_(torch.tensor([_as_i64(self.philox_seed), _as_i64(self.philox_offset)], dtype=torch.int64))
        st["x"].copy_(xt)
        st["mu"].copy_(self._mu_tensor(xt))
        st["ctx"].copy_(ctx)
        self.model.set_image_context(st["ctx"])      # refresh the persistent cross-attention biases (eager)
        st["counter"].fill_(T)
        table = self._coef_table(dev)

        def one_step():
            s = _stream(dev)
            if "ss_table" in st:
                check(L.idiff_step_select_ss(table.data_ptr(), st["counter"].data_ptr(), st["row"].data_ptr(),
                                             st["time"].data_ptr(), float(self.sample_scale), st["ss_table"].data_ptr(),
                                             st["ss_slot"], st["ss_table"].shape[1], s), "step_select_ss")
                eps = self.model.forward_into(st["x"], st["mu"], None, st["ctx"], time_ptr=st["time"].data_ptr(), ss_ready=True)
            else:
                check(L.idiff_step_select(table.data_ptr(), st["counter"].data_ptr(), st["row"].data_ptr(),
                                          st["time"].data_ptr(), float(self.sample_scale), s), "step_select")
                eps = self.model.forward_into(st["x"], st["mu"], None, st["ctx"], time_ptr=st["time"].data_ptr())
            if ode:                                     # :48-49 -- no dispersion, half the sigma^2 * score term
                check(L.idiff_sde_step(st["x"].data_ptr(), st["x"].data_ptr(), eps.data_ptr(), st["mu"].data_ptr(), None,
                                       st["row"].data_ptr(), 2, 0, 0, 0, st["x"].numel(), s), "sde_step")
            elif torch_noise:                           # z = randn_like(x) (:185), drawn by torch inside the graph
                st["z"].normal_()
                check(L.idiff_sde_step(st["x"].data_ptr(), st["x"].data_ptr(), eps.data_ptr(), st["mu"].data_ptr(),
                                       st["z"].data_ptr(), st["row"].data_ptr(), 0, 0, 0, 0, st["x"].numel(), s), "sde_step")
            else:
                check(L.idiff_sde_step_rng(st["x"].data_ptr(), st["x"].data_ptr(), eps.data_ptr(), st["mu"].data_ptr(),
                                           st["row"].data_ptr(), 0, st["rng"].data_ptr(), 1 if off4 else 0,
                                           st["x"].numel(), s), "sde_step")

        done = 0
        if st["graph"] is None:
            one_step()                                   # eager warm-up: sets kernel attributes, binds context
            done = 1
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            st["graph"] = g
        for _ in range(T - done):
            st["graph"].replay()
        _lib.watchdog()
        return st["x"].clone()

    def reverse_ode(self, xt, T=-1, save_states=False, save_dir="ode_state", **kwargs):
        """Probability-flow Euler loop (:263-280): model forward + ONE fused update per step.  The reference does not
        forward model kwargs here (:267); they are accepted so a conditioned network can be sampled this way too."""
        T = self._check_T(self.sample_T if T < 0 else T)
        xt = _require_cuda_f32("xt", xt)
        if T == 0 or xt.numel() == 0:
            return xt.clone()
        if (self.use_cuda_graph and not save_states and hasattr(self.model, "forward_into")
                and set(kwargs) == {"image_context"} and torch.is_tensor(self.mu)):
            return self._reverse_sde_graph(xt, T, kwargs, ode=True)
        x = xt.clone()
        fast_model = hasattr(self.model, "forward_into") and set(kwargs) <= {"image_context"}
        for t in reversed(range(1, T + 1)):
            if fast_model:
                noise = self.model.forward_into(x, self._mu_tensor(x), t * self.sample_scale, kwargs.get("image_context"))
            else:
                noise = self.model(x, self.mu, t * self.sample_scale, **kwargs)
            x = self._fused_step(x, noise, t, is_score=False, with_noise=False, out=x, ode=True)
            if save_states:
                interval = self.T // 100
                if t % interval == 0:
                    self._save_state(x, save_dir, f"state_{t // interval}.png", dim=3)
        _lib.watchdog()
        return x

    def ode_sampler(self, xt, rtol=1e-5, atol=1e-5, method="RK45", eps=1e-3):
        """Black-box probability-flow ODE solve (:282-306); host-driven, not on the hot path."""
        from scipy import integrate
        shape = xt.shape

        def rhs(t, flat):
            x = torch.from_numpy(flat.reshape(shape)).to(self.device).type(torch.float32)
            score = self.score_fn(x, int(t))
            return self.ode_reverse_drift(x, score, int(t)).detach().cpu().numpy().reshape(-1)

        sol = integrate.solve_ivp(rhs, (self.T, eps), xt.detach().cpu().numpy().reshape(-1), rtol=rtol, atol=atol,
                                  method=method)
        return torch.tensor(sol.y[:, -1]).reshape(shape).to(self.device).type(torch.float32)

    def optimal_reverse(self, xt, x0, T=-1):
        T = self.T if T < 0 else T
        x = xt.clone()
        for t in reversed(range(1, T + 1)):
            x = self.reverse_optimum_step(x, x0, t)
        return x

    # ------------------------------------------------------------------ training states (:318-338)
    def weights(self, t):
        return torch.exp(-self.thetas_cumsum[t] * self.dt)

    def generate_random_states(self, x0, mu, timesteps=None, T_start=1, T_end=-1):
        x0 = x0.to(self.device)
        mu = mu.to(self.device)
        self.set_mu(mu)
        if timesteps is None:
            batch = x0.shape[0]
            T_end = self.T + 1 if T_end <= 1 else T_end + 1
            timesteps = torch.randint(T_start, T_end, (batch, 1, 1, 1)).long()
        # :331-336 in one fused pass (idiff_random_states); the per-sample coefficients are gathered from the
        # schedule tables exactly as the reference indexes them (`thetas_cumsum[t]`, `sigma_bars[t]`)
        x0 = _require_cuda_f32("x0", x0)
        mu = _require_cuda_f32("mu", mu if tuple(mu.shape) == tuple(x0.shape) else mu.expand_as(x0))
        dev = x0.device
        B = x0.shape[0]
        t_idx = torch.as_tensor(timesteps).to(dev).long().reshape(-1)
        if t_idx.numel() == 1 and B != 1:
            t_idx = t_idx.expand(B)
        if t_idx.numel() != B:
            raise IdiffError(f"generate_random_states: {t_idx.numel()} timesteps for a batch of {B}")
        decay = torch.exp(-self.thetas_cumsum.to(dev)[t_idx] * self.dt).to(torch.float32).contiguous()   # :170
        sbar = self.sigma_bars.to(dev)[t_idx].to(torch.float32).contiguous()                                      # :335
        noisy_states = torch.empty_like(x0)
        philox = self.noise_source == "philox"
        z = None if philox else _require_cuda_f32("noise", self._draw(-1, x0))
        self.last_noises = torch.empty_like(x0) if philox else z      # the noise-matching target (:222-223)
        if B:
            check(_lib.lib().idiff_random_states(
                noisy_states.data_ptr(), self.last_noises.data_ptr() if philox else None, x0.data_ptr(), mu.data_ptr(),
                None if philox else z.data_ptr(), decay.data_ptr(), sbar.data_ptr(), 1 if philox else 0,
                self.philox_seed, self.philox_offset, 0xFFFFFFFE, x0.numel() // B, x0.numel(), _stream(dev)),
                "random_states")
        return timesteps, noisy_states.to(torch.float32)
