"""CPU checks of two contracts: (1) the JSON line of bench.py as committed under profiles/ carries every key the
driver reads; (2) the LayerNorm fold of la_out2 (csrc/linattn_fused.cu): feeding the tensor core the RAW bf16 x and
correcting the accumulator, rstd * (Wq x - mean * rowsum(Wq)), equals Wq applied to the normalised bf16 x^ of version 1
up to the bf16 rounding of the operand."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_lines_carry_the_contract_keys():
    for name, impl in (("r02b_bench.json", None), ("r02b_bench_reference.json", "reference")):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
            assert k in d, (name, k)
        assert "workload" in d["config"] and d["unit"] == "images/s" and d["higher_is_better"] is True
        for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert k in d["e2e"], (name, k)
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in d["cpu_baseline"], (name, k)
        if impl:
            assert d.get("impl") == impl and d["e2e"]["h2d_bytes_per_step"] == 0
        else:
            assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
            for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
                assert k in d["roofline"], k
            assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
            assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
            assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    a = json.load(open(os.path.join(ROOT, "profiles", "r02b_bench.json")))["config"]
    b = json.load(open(os.path.join(ROOT, "profiles", "r02b_bench_reference.json")))["config"]
    assert a == b, "both arms must describe the same workload"


def test_layernorm_fold_of_the_output_pass_equals_normalise_then_multiply():
    g = torch.Generator().manual_seed(0)
    bf = lambda t: t.to(torch.bfloat16).float()
    for Cc, offset in ((64, 0.0), (128, 3.0)):                     # a per-pixel mean of 3 std stresses the cancellation
        x = bf(torch.randn(512, Cc, generator=g) + offset)           # activations as stored (bf16)
        wq = bf((torch.rand(128, Cc, generator=g) * 2 - 1) * (2.0 / Cc ** 0.5))   # bf16 image of Wq (gain folded in)
        mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + 1e-5)
        exact = ((x - mean) * rstd).double() @ wq.double().t()       # fp64 reference of the projection
        v1 = bf((x - mean) * rstd) @ wq.t()                          # version 1: x^ rounded to bf16 feeds the MMA
        v2 = rstd * (x @ wq.t() - mean * wq.sum(dim=1)[None, :])     # version 2: raw x, fold in fp32
        e1 = (v1.double() - exact).abs().max().item()
        e2 = (v2.double() - exact).abs().max().item()
        scale = exact.abs().max().item()
        assert e2 <= 2e-5 * scale * (1 + abs(offset)), (Cc, e2 / scale)     # the fold only adds fp32 cancellation error
        assert e2 <= e1, (Cc, e1 / scale, e2 / scale)                         # and it spares the bf16 rounding of x^
        assert e1 <= 1e-2 * scale
