"""CPU-side checks: the C ABI exports every symbol include/idiff.h declares, host logic (packing,
sharding, schedule tables, drop-in surface) and the N>1 gather path on gloo with world_size 2."""
import inspect
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "idiff.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(idiff_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from instancediff_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/idiff.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names)
    assert L.idiff_abi_version() == 1
    import ctypes
    assert L.idiff_sizeof_gemm_params() == ctypes.sizeof(_lib.GemmParams)


def _struct_fields(name):
    """member names of `typedef struct <name> { ... } <name>;` in include/idiff.h, in declaration order"""
    text = open(os.path.join(ROOT, "include", "idiff.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):                      # `int32_t B, H, W` declares three members
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    return fields


def test_ctypes_mirrors_follow_the_header_field_by_field():
    """_lib.GemmParams / _lib.GnFuse against include/idiff.h: same members in the same order, same total size"""
    import ctypes
    from instancediff_b200 import _lib
    L = _lib.lib()
    for cname, mirror, sizeof in (("idiff_gemm_params", _lib.GemmParams, L.idiff_sizeof_gemm_params()),
                                  ("idiff_gn_fuse", _lib.GnFuse, L.idiff_sizeof_gn_fuse())):
        assert [f[0] for f in mirror._fields_] == _struct_fields(cname), cname
        assert ctypes.sizeof(mirror) == sizeof, cname
    assert _lib.GN_SLOTS == int(re.search(r"#define IDIFF_GN_SLOTS (\d+)", open(os.path.join(ROOT, "include", "idiff.h")).read()).group(1))


def test_missing_library_fails_loudly(monkeypatch):
    from instancediff_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libidiff_sm100.so")
    with pytest.raises(_lib.IdiffError, match="no CPU fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "instancediff_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
            assert "/root/reference" not in src, fn


def test_pack_table_host_entry_point():
    import ctypes
    from instancediff_b200 import IRSDE, _lib
    s = IRSDE(0.4, device="cpu")
    th, sg, sb = (t.contiguous() for t in s._host)
    out = torch.zeros(101, 8)
    assert _lib.lib().idiff_sde_pack_table(th.data_ptr(), sg.data_ptr(), sb.data_ptr(), 101, float(s.dt),
                                           float(np.sqrt(float(s.dt))), out.data_ptr()) == 0
    assert torch.equal(out[:, 0], th) and torch.equal(out[:, 2], sb)
    assert out[7, 5].view(torch.int32).item() == 7
    assert _lib.lib().idiff_sde_pack_table(None, None, None, 0, 0.0, 0.0, None) < 0
    assert b"sde_pack_table" in _lib.lib().idiff_last_error()


def test_weight_packing_roundtrip_and_layout():
    from instancediff_b200.packing import fold_layernorm, interleave_geglu, pack_conv_weight, unpack_conv_weight
    g = torch.Generator().manual_seed(0)
    w = torch.randn(256, 192, 3, 3, generator=g)
    for NT in (64, 128, 256):
        p = pack_conv_weight(w, NT)
        assert p.dtype == torch.bfloat16 and p.numel() == w.numel()
        assert torch.equal(unpack_conv_weight(p, 256, 192, 3, NT), w.to(torch.bfloat16).float())
    # explicit element check of the documented layout
    NT, n, c, ky, kx = 128, 200, 77, 2, 1
    p = pack_conv_weight(w, NT).reshape(2, 3, 9, 8, NT, 8)
    assert p[n // NT, c // 64, ky * 3 + kx, (c % 64) // 8, n % NT, c % 8].float() == w[n, c, ky, kx].to(torch.bfloat16).float()
    # LayerNorm fold algebra
    wl, gain, beta, x = torch.randn(32, 64, generator=g), torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g), torch.randn(5, 64, generator=g)
    wf, wsum, extra = fold_layernorm(wl, gain, beta)
    m, r = x.mean(1, keepdim=True), torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (64,), gain, beta, 1e-5), wl)
    got = (x @ wf.t() - m * wf.sum(1)) * r + extra
    assert torch.allclose(got, ref, atol=1e-4)
    wi, bi = interleave_geglu(torch.arange(8.).reshape(8, 1), torch.arange(8.))
    assert bi.tolist() == [0, 4, 1, 5, 2, 6, 3, 7]


def test_param_inventory_matches_oracle_contract():
    from instancediff_b200 import param_specs
    from oracle.unet_oracle import make_oracle_unet
    sd = make_oracle_unet().state_dict()
    specs = {n: tuple(s) for n, s, _ in param_specs()}
    assert set(specs) == set(sd)
    assert all(tuple(sd[k].shape) == specs[k] for k in sd)


def test_irsde_surface_and_tables_on_cpu(golden_dir):
    """Drop-in surface of utils/sde_utils.py: same names/signatures; tables equal the reference goldens."""
    from instancediff_b200 import IRSDE, SDE
    from oracle.gen_golden import TABLE_CASES
    tab = np.load(os.path.join(golden_dir, "irsde_tables.npz"))
    for name, kw in TABLE_CASES.items():
        s = IRSDE(device="cpu", **kw)
        for f in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
            assert np.array_equal(getattr(s, f).numpy(), tab[f"{name}/{f}"])
        assert float(s.dt) == float(tab[f"{name}/dt"]) and s.max_sigma == float(tab[f"{name}/max_sigma"])
    s = IRSDE(0.4)
    for m in ("set_mu set_model mu_bar sigma_bar drift sde_reverse_drift ode_reverse_drift dispersion "
              "get_score_from_noise score_fn_ score_fn noise_fn reverse_optimum_step sigma theta get_real_noise "
              "get_real_score get_init_state_from_noise forward reverse_sde reverse_ode ode_sampler optimal_reverse "
              "weights generate_random_states noise_state forward_step reverse_sde_step_mean reverse_sde_step "
              "reverse_ode_step").split():
        assert callable(getattr(s, m)), m
    assert list(inspect.signature(IRSDE.__init__).parameters) == ["self", "max_sigma", "T", "sample_T", "schedule", "eps", "device"]
    sig = inspect.signature(IRSDE.reverse_sde)
    assert list(sig.parameters) == ["self", "xt", "T", "save_states", "save_dir", "kwargs"]
    assert sig.parameters["save_dir"].default == "sde_state" and sig.parameters["T"].default == -1
    assert list(inspect.signature(IRSDE.generate_random_states).parameters) == ["self", "x0", "mu", "timesteps", "T_start", "T_end"]
    assert issubclass(IRSDE, SDE) and s.mu == 0. and s.model is None and s.sample_scale == 1.0
    with pytest.raises(NameError):
        IRSDE(0.4, schedule="sigmoid")


def test_hot_functions_reject_cpu_tensors():
    from instancediff_b200 import IRSDE, IdiffError
    s = IRSDE(0.4, device="cpu")
    x = torch.zeros(1, 1, 4, 4)
    for call in (lambda: s.reverse_sde_step(x, x, 5), lambda: s.reverse_sde_step_mean(x, x, 5),
                 lambda: s.noise_state(x), lambda: s.reverse_sde(x)):
        with pytest.raises(IdiffError, match="no CPU path"):
            call()


def test_shard_bounds_cover_batch_exactly():
    from instancediff_b200 import shard_bounds
    for total in (0, 1, 7, 16, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(256, w, 0)[1] for w in (2, 4, 8)] == [128, 64, 32]     # BASELINE config 3
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from instancediff_b200 import gather_shards, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    total = 7
    full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 1, 2, 3)
    lo, hi = shard_bounds(total, world, rank)
    out = gather_shards(full[lo:hi] * 2, total, rank, world)        # stand-in for the per-rank sampler result
    dist.barrier()
    if rank == 0:
        q.put(bool(torch.equal(out, full * 2)))
    else:
        q.put(out is None)
    dist.destroy_process_group()


def test_two_rank_gloo_gather_reassembles_batch_in_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(results) and all(p.exitcode == 0 for p in procs)


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from instancediff_b200.parallel import GradientAllReducer
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ok = []

    def make():
        torch.manual_seed(0)                                      # identical replicas, as DDP starts them
        return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.SiLU(), torch.nn.Linear(16, 16), torch.nn.SiLU(),
                                   torch.nn.Linear(16, 3))
    xs = [torch.randn(5, 6, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    # expected: average over ranks of the per-rank gradients, computed locally on every rank
    want = None
    for r in range(world):
        m = make()
        m(xs[r]).square().mean().backward()
        g = [p.grad.clone() for p in m.parameters()]
        want = g if want is None else [a + b for a, b in zip(want, g)]
    want = [w / world for w in want]
    # (a) hooks fire the buckets from inside backward; tiny buckets force several of them
    net = make()
    red = GradientAllReducer(net.parameters(), bucket_mb=0.0005).attach()
    ok.append(len(red.buckets) > 2)
    for _ in range(2):                                            # two steps: bucket state is re-armed by wait()
        net.zero_grad(set_to_none=False)
        net(xs[rank]).square().mean().backward()
        red.wait()
        ok.append(all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(net.parameters(), want)))
    red.detach()
    # (b) manual mode, bf16 wire format, a parameter without a gradient counts as zeros
    net2 = make()
    net2(xs[rank]).square().mean().backward()
    net2[4].bias.grad = None
    red2 = GradientAllReducer(net2.parameters(), bucket_mb=25, comm_dtype=torch.bfloat16)
    red2.reduce().wait()
    ok.append(len(red2.buckets) == 1 and red2.bytes_per_step == 2 * sum(p.numel() for p in net2.parameters()))
    ok.append(all(torch.allclose(p.grad, w, atol=2e-2, rtol=2e-2) for p, w in list(zip(net2.parameters(), want))[:-1]))
    ok.append(bool((net2[4].bias.grad == 0).all()))
    # every rank ends with the same bits
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok.append(all(torch.equal(gathered[0], t) for t in gathered))
    dist.barrier()
    q.put(all(ok))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_matches_ddp_averaging():
    """Config 5's only collective (models/drift_noise_model.py:145-146 wraps the nets in DDP): bucketed, asynchronous,
    launched from backward hooks; world size 2 on gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(results) and all(p.exitcode == 0 for p in procs)


def test_head_weight_fragments_reproduce_the_convolution():
    """packing.pack_head_weight lays the 3x3 x 64 -> 1 weights out as mma.sync.m16n8k16 B fragments with the taps on
    the N side (csrc/conv_simt.cu).  Consuming them the way the kernel does -- per-pixel tap partials, then the
    9-point gather -- must give the fp32 convolution (weights are split hi + lo: ~2^-16 relative)."""
    from instancediff_b200.packing import pack_head_weight
    g = torch.Generator().manual_seed(5)
    w = (torch.rand(1, 64, 3, 3, generator=g) * 2 - 1) / 24
    frag = pack_head_weight(w).float().reshape(3, 4, 32, 4)                 # [n-tile][k-chunk][lane][b0.lo b0.hi b1.lo b1.hi]
    assert frag.numel() == 1536
    # B[j][kc] as a 16 x 8 matrix: lane (n = lane // 4, q = lane % 4) holds k = 2q, 2q+1 (b0) and 2q+8, 2q+9 (b1) of column n
    Bm = torch.zeros(3, 4, 16, 8)
    for lane in range(32):
        n, q = lane // 4, lane % 4
        for slot, k in enumerate((2 * q, 2 * q + 1, 2 * q + 8, 2 * q + 9)):
            Bm[:, :, k, n] = frag[:, :, lane, slot]
    H, W = 7, 9
    x = torch.randn(1, 64, H, W, generator=g).to(torch.bfloat16).float()
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))[0].permute(1, 2, 0)       # [H+2][W+2][64] patch with halo
    part = torch.zeros(H + 2, W + 2, 9)
    for j in range(3):
        acc = torch.zeros(H + 2, W + 2, 8)
        for kc in range(4):
            acc += xp[:, :, kc * 16:(kc + 1) * 16] @ Bm[j, kc]              # the m16n8k16 MMAs of n-tile j
        for qd in range(4):
            tap = 4 * j + qd
            if tap < 9:
                part[:, :, tap] = acc[:, :, 2 * qd] + acc[:, :, 2 * qd + 1]   # (hi, lo) column pair of the tap
    out = torch.zeros(H, W)
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        out += part[dy:dy + H, dx:dx + W, tap]
    ref = torch.nn.functional.conv2d(x, w, padding=1)[0, 0]
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
    # columns of taps 9..11 (n-tile 2, n >= 2) are zero
    assert float(Bm[2, :, :, 2:].abs().max()) == 0.0


def test_shard_bounds_properties():
    """Shards are contiguous, ordered, cover [0, total) exactly once and differ in size by at most one."""
    from hypothesis import given, settings, strategies as st
    from instancediff_b200 import shard_bounds

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 5000), st.integers(1, 64))
    def check(total, world):
        bounds = [shard_bounds(total, world, r) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        sizes = [hi - lo for lo, hi in bounds]
        assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    check()


def test_c_level_network_object_reports_errors_without_a_gpu():
    """idiff_unet_* (csrc/unet_plan.cu): configuration and missing-parameter errors are host-side checks."""
    import torch
    from instancediff_b200 import _lib
    from instancediff_b200.native import NativeUNet
    with pytest.raises(_lib.IdiffError, match="built for"):
        NativeUNet(nf=32, device="cpu")
    net = NativeUNet(device="cpu")
    with pytest.raises(_lib.IdiffError, match="'init_conv.bias' was not loaded"):
        net.load_state_dict({"init_conv.weight": torch.zeros(64, 2, 7, 7)})
