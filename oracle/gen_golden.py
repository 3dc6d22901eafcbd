"""Generate tests/golden/*.npz by IMPORTING THE REFERENCE ITSELF (run in the dev container).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

* irsde_tables.npz  -- theta/sigma/cumsum/sigma_bar/dt of ``utils.sde_utils.IRSDE`` for
                       several constructor settings (reference: utils/sde_utils.py:85-155)
* irsde_loop.npz    -- SURVEY.md App. B KAT-3: a 100-step ``reverse_sde`` with a queue of
                       pre-drawn Gaussians replacing ``torch.randn_like`` and the analytic
                       model ``get_real_noise`` (reference: :244-261), plus a single
                       ``reverse_sde_step`` and ``generate_random_states`` / ``noise_state``.
* irsde_ode.npz     -- the reference's ``reverse_ode`` / ``reverse_ode_step`` (:48-49, :263-280) on the same
                       inputs (`--ode-only` regenerates just this file).
* unet_oracle.npz   -- output of oracle/unet_oracle.py (seed-1 weights) on a fixed input.
                       The reference ships no network source, so this one pins the oracle to
                       itself only (see the module header).

/root/reference is read-only and absent on the GPU box; only the fixtures travel.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("IDIFF_REFERENCE", "/root/reference")


def import_reference_irsde():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from utils.sde_utils import IRSDE  # noqa: the reference's own class
    sys.path.pop(0)
    return IRSDE


TABLE_CASES = {
    "cos_s0p4_T100_e0p01": dict(max_sigma=0.4, T=100, schedule="cosine", eps=0.01),
    "cos_s50_T100_e0p005": dict(max_sigma=50, T=100, schedule="cosine", eps=0.005),
    "lin_s0p4_T100_e0p01": dict(max_sigma=0.4, T=100, schedule="linear", eps=0.01),
    "const_s0p4_T50_e0p01": dict(max_sigma=0.4, T=50, schedule="constant", eps=0.01),
    "cos_s0p4_T100_sT25_e0p01": dict(max_sigma=0.4, T=100, sample_T=25, schedule="cosine", eps=0.01),
}


def gen_tables(IRSDE):
    out = {}
    for name, kw in TABLE_CASES.items():
        sde = IRSDE(device="cpu", **kw)
        out[f"{name}/thetas"] = sde.thetas.numpy()
        out[f"{name}/sigmas"] = sde.sigmas.numpy()
        out[f"{name}/thetas_cumsum"] = sde.thetas_cumsum.numpy()
        out[f"{name}/sigma_bars"] = sde.sigma_bars.numpy()
        out[f"{name}/dt"] = np.asarray(sde.dt.numpy())
        out[f"{name}/max_sigma"] = np.asarray(sde.max_sigma, dtype=np.float64)
        out[f"{name}/sample_scale"] = np.asarray(sde.sample_scale, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "irsde_tables.npz"), **out)


def gen_loop(IRSDE):
    g = torch.Generator().manual_seed(1234)
    B, H, W = 2, 16, 16
    mu = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    x0t = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    zs = [torch.randn(B, 1, H, W, generator=g) for _ in range(101)]
    sde = IRSDE(max_sigma=0.4, T=100, schedule="cosine", eps=0.01, device="cpu")
    sde.set_mu(mu)
    sde.set_model(lambda x, m, t, **kw: sde.get_real_noise(x, x0t, int(t)))

    queue = [zs[0]] + [zs[t] for t in range(100, 0, -1)]
    real_randn_like = torch.randn_like
    torch.randn_like = lambda x, **kw: queue.pop(0)
    try:
        xT = sde.noise_state(mu)
        # per-step trace: run the reference's own loop one step at a time so the
        # intermediate states are exactly what reverse_sde would hold
        states = []
        x = xT.clone()
        for t in reversed(range(1, 101)):
            score = sde.score_fn(x, t, sde.sample_scale)
            x = sde.reverse_sde_step(x, score, t)
            states.append(x.clone())
        # and the untouched driver loop end to end
        queue.extend([zs[t] for t in range(100, 0, -1)])
        x_end = sde.reverse_sde(xT, T=-1)
        assert len(queue) == 0
        # one isolated step with an arbitrary eps
        eps_any = zs[7] * 0.5
        queue.append(zs[9])
        x_step = sde.reverse_sde_step(xT, sde.get_score_from_noise(eps_any, 63), 63)
        # training-state sampler with fixed timesteps
        ts = torch.tensor([5, 97]).reshape(B, 1, 1, 1)
        queue.append(zs[11])
        _, xt_train = sde.generate_random_states(x0t, mu, timesteps=ts)
        real_eps = sde.get_real_noise(xt_train, x0t, ts)
    finally:
        torch.randn_like = real_randn_like
    assert torch.equal(x_end, states[-1])
    np.savez_compressed(
        os.path.join(OUT, "irsde_loop.npz"),
        mu=mu.numpy(), x0t=x0t.numpy(), zs=torch.stack(zs).numpy(), xT=xT.numpy(),
        states=torch.stack(states).numpy(), x_end=x_end.numpy(),
        step_eps=eps_any.numpy(), step_t=np.asarray(63), step_z=zs[9].numpy(), step_out=x_step.numpy(),
        train_t=ts.numpy(), train_z=zs[11].numpy(), train_xt=xt_train.numpy(), train_eps=real_eps.numpy())
    print("KAT-3 sums:", mu.double().sum().item(), x0t.double().sum().item(),
          xT.double().sum().item(), x_end.double().sum().item())


def gen_ode(IRSDE):
    """irsde_ode.npz -- the reference's own ``reverse_ode`` (:263-280) on the KAT-3 inputs with the analytic model,
    every intermediate state, plus one isolated ``reverse_ode_step`` (:48-49)."""
    loop = np.load(os.path.join(OUT, "irsde_loop.npz"))
    mu, x0t, xT = (torch.from_numpy(loop[k]) for k in ("mu", "x0t", "xT"))
    sde = IRSDE(max_sigma=0.4, T=100, schedule="cosine", eps=0.01, device="cpu")
    sde.set_mu(mu)
    sde.set_model(lambda x, m, t, **kw: sde.get_real_noise(x, x0t, int(t)))
    states = []
    x = xT.clone()
    for t in reversed(range(1, 101)):
        x = sde.reverse_ode_step(x, sde.score_fn(x, t, sde.sample_scale), t)
        states.append(x.clone())
    x_end = sde.reverse_ode(xT, T=-1)
    assert torch.equal(x_end, states[-1])
    eps_any = torch.from_numpy(loop["step_eps"])
    x_step = sde.reverse_ode_step(xT, sde.get_score_from_noise(eps_any, 63), 63)
    np.savez_compressed(os.path.join(OUT, "irsde_ode.npz"), states=torch.stack(states).numpy(), x_end=x_end.numpy(),
                        step_out=x_step.numpy())
    print("ODE sums:", x_end.double().sum().item(), x_step.double().sum().item())


def gen_unet():
    sys.path.insert(0, ROOT)
    from oracle.unet_oracle import make_oracle_unet
    net = make_oracle_unet(seed=1)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 1, 32, 32, generator=g)
    mu = torch.rand(1, 1, 32, 32, generator=g) * 2 - 1
    ctx = torch.nn.functional.normalize(torch.randn(1, 1, 512, generator=g), dim=-1)
    with torch.no_grad():
        y = net(x, mu, 37.0, image_context=ctx)
        y2 = net(x, mu, 3.0, image_context=ctx)
    np.savez_compressed(os.path.join(OUT, "unet_oracle.npz"), x=x.numpy(), mu=mu.numpy(),
                        ctx=ctx.numpy(), eps_t37=y.numpy(), eps_t3=y2.numpy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    IRSDE = import_reference_irsde()
    if "--ode-only" not in sys.argv:
        gen_tables(IRSDE)
    if "--ode-only" not in sys.argv:
        gen_loop(IRSDE)
        gen_unet()
    gen_ode(IRSDE)
    print("wrote", sorted(os.listdir(OUT)))
