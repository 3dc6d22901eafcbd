"""Pin the oracle restatement (oracle/irsde_oracle.py) to the reference.

The fixtures were produced by importing /root/reference/utils/sde_utils.py::IRSDE
(oracle/gen_golden.py); equality is bit-exact because both run the same fp32 torch ops
in the same order.  KAT numbers are SURVEY.md App. B.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import irsde_oracle as O
from oracle.gen_golden import TABLE_CASES


@pytest.fixture(scope="module")
def tables(golden_dir):
    return np.load(os.path.join(golden_dir, "irsde_tables.npz"))


@pytest.fixture(scope="module")
def loop(golden_dir):
    return np.load(os.path.join(golden_dir, "irsde_loop.npz"))


@pytest.mark.parametrize("name", sorted(TABLE_CASES))
def test_tables_bit_exact(tables, name):
    s = O.make_schedule(**TABLE_CASES[name])
    for field in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
        got = getattr(s, field).numpy()
        ref = tables[f"{name}/{field}"]
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.array_equal(got, ref), field
    assert np.array_equal(np.asarray(s.dt.numpy()), tables[f"{name}/dt"])
    assert s.max_sigma == float(tables[f"{name}/max_sigma"])
    assert s.sample_scale == float(tables[f"{name}/sample_scale"])


def test_kat1_values():
    s = O.make_schedule(0.4, 100, schedule="cosine", eps=0.01)
    assert float(s.dt) == pytest.approx(0.09047583490610123, abs=0)
    assert math.sqrt(s.dt) == pytest.approx(0.3007920127032984, abs=1e-15)
    assert float(s.thetas[50]) == 0.506156325340271
    assert float(s.sigmas[100]) == 0.5656194090843201
    assert float(s.sigma_bars[1]) == 0.007004070561379194
    assert float(s.sigma_bars[0]) == 0.0
    a, b, c = O.step_coefficients(s)
    assert float(a[100]) == pytest.approx(0.09045472, rel=1e-6)
    assert float(b[100]) == pytest.approx(0.07236740, rel=1e-6)
    assert float(c[100]) == pytest.approx(0.17013380, rel=1e-6)
    assert float(b[1]) == pytest.approx(0.0070046913, rel=1e-6)
    assert float(torch.prod(1 + a[1:])) == pytest.approx(86.211, rel=1e-4)
    assert float(a[1:].sum()) == pytest.approx(-math.log(0.01), rel=1e-6)


def test_kat2_values():
    s = O.make_schedule(50, 100, schedule="cosine", eps=0.005)
    assert s.max_sigma == 0.19607843137254902
    assert float(s.dt) == 0.10409380495548248
    assert float(s.sigma_bars[100]) == 0.19607597589492798


def test_unknown_schedule_raises_nameerror():
    with pytest.raises(NameError):
        O.make_schedule(0.4, schedule="sigmoid")


def test_loop_bit_exact(loop):
    s = O.make_schedule(0.4, 100, schedule="cosine", eps=0.01)
    mu, x0t = torch.from_numpy(loop["mu"]), torch.from_numpy(loop["x0t"])
    zs = torch.from_numpy(loop["zs"])
    xT = O.noise_state(s, mu, zs[0])
    assert torch.equal(xT, torch.from_numpy(loop["xT"]))
    trace = []
    x_end = O.reverse_sde(s, lambda x, m, t: O.real_noise(s, x, x0t, m, int(t)), xT, mu,
                          lambda t, x: zs[t], trace=trace)
    assert torch.equal(torch.stack(trace), torch.from_numpy(loop["states"]))
    assert torch.equal(x_end, torch.from_numpy(loop["x_end"]))
    assert float(x_end.double().sum()) == -13.649582519428805
    assert float((x_end - x0t).abs().max()) == pytest.approx(0.02373816817998886, abs=1e-9)


def test_single_step_and_training_states(loop):
    s = O.make_schedule(0.4, 100, schedule="cosine", eps=0.01)
    mu = torch.from_numpy(loop["mu"])
    out = O.reverse_step(s, torch.from_numpy(loop["xT"]), mu, torch.from_numpy(loop["step_eps"]),
                         torch.from_numpy(loop["step_z"]), int(loop["step_t"]))
    assert torch.equal(out, torch.from_numpy(loop["step_out"]))
    ts = torch.from_numpy(loop["train_t"])
    x0t = torch.from_numpy(loop["x0t"])
    xt = O.random_states(s, x0t, mu, ts, torch.from_numpy(loop["train_z"]))
    assert torch.equal(xt, torch.from_numpy(loop["train_xt"]))
    assert torch.equal(O.real_noise(s, xt, x0t, mu, ts), torch.from_numpy(loop["train_eps"]))


def test_closed_form_matches_reference_order(loop):
    """x' = x - a(mu-x) - b eps - c z agrees with the reference op order to ~1 ulp."""
    s = O.make_schedule(0.4, 100, schedule="cosine", eps=0.01)
    a, b, c = O.step_coefficients(s)
    t = int(loop["step_t"])
    x, mu = torch.from_numpy(loop["xT"]), torch.from_numpy(loop["mu"])
    e, z = torch.from_numpy(loop["step_eps"]), torch.from_numpy(loop["step_z"])
    closed = x.double() - a[t] * (mu.double() - x.double()) - b[t] * e.double() - c[t] * z.double()
    ref = torch.from_numpy(loop["step_out"]).double()
    assert (closed - ref).abs().max() < 1e-6
