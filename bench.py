#!/usr/bin/env python
"""Benchmark of the reverse-SDE hot path (BASELINE.json metric: images/sec of a 100-step reverse SDE at
256x256, bf16 UNet on B200).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = one complete T=100 sampling (100 x [UNet forward + fused Euler-Maruyama update]) of one batch
of 32 synthetic 1x256x256 images per GPU (BASELINE.json configs[1]); N GPUs = N independent shards of 32
images, no collective on the data path (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_STEPS = 100
BATCH_PER_GPU = 32
RES = 256
SCALING = "weak"
METRIC = "images/sec, 100-step reverse-SDE @256^2"
UNIT = "images/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def conv_traffic_from_profile(labels):
    """(mean DRAM bytes per conv launch of one forward, provenance) from profiles/conv_dram_traffic.json -- written by
    tools/ncu_conv_traffic.py from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` capture together
    with the list of layer labels it saw.  DRAM bytes cannot be measured without the profiler, so the figure is only
    reported when that list equals the launch list of THIS run (same kernels, same order); otherwise null."""
    path = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")
    if not os.path.exists(path):
        return None, None
    d = json.load(open(path))
    if d.get("labels") != list(labels):
        return None, f"{os.path.basename(path)} is stale (launch list differs): not reported"
    return d["bytes_total"] / len(labels), f"{os.path.basename(path)}: {d.get('source', '')}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm_sorted = sorted(sm)
            out.update(sm_mhz=sm_sorted[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_config(world):
    """The SAME `config` object in both arms (the driver compares them): what is computed, not how."""
    return {"workload": f"reverse-SDE sampling T={T_STEPS}, batch {BATCH_PER_GPU}/GPU, 1x{RES}x{RES}, drift/noise UNet of "
                        "SURVEY App. A (22.75M parameters, random init) + Euler-Maruyama update per step",
            "sde": "IRSDE(max_sigma=0.4,T=100,cosine,eps=0.01)", "global_batch": BATCH_PER_GPU * world,
            "parallelism": f"batch-shard x{world}, no collective"}


def synthetic_inputs(B, seed, pinned):
    g = torch.Generator().manual_seed(seed)
    mu = torch.rand(B, 1, RES, RES, generator=g) * 2 - 1                      # data/MedSpeckle.py:69-70 range
    ctx = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1)
    if pinned:
        mu, ctx = mu.pin_memory(), ctx.pin_memory()
    return mu, ctx


# ------------------------------------------------------------------------------------------------
class CpuReference:
    """Oracle port (reference IRSDE restatement + fp32 oracle UNet) on the host cores.  The network is built once and
    warmed with one step; every timed sample continues the same reverse process on ONE image (BASELINE config 1)."""

    def __init__(self, threads=None):
        from oracle import irsde_oracle as O
        from oracle.unet_oracle import make_oracle_unet
        torch.set_num_threads(threads or host_threads())    # torchrun exports OMP_NUM_THREADS=1: claim the cores
        self.O = O
        self.net = make_oracle_unet(seed=1)
        self.s = O.make_schedule(0.4, T_STEPS, schedule="cosine", eps=0.01)
        self.mu, self.ctx = synthetic_inputs(1, 1, False)
        self.g = torch.Generator().manual_seed(2)
        self.x = O.noise_state(self.s, self.mu, torch.randn(self.mu.shape, generator=self.g))
        self.t = T_STEPS
        self.sample(1)                                      # warm-up step (allocator, thread pool, oneDNN primitives)

    def sample(self, n_sde_steps):
        """Runs `n_sde_steps` further SDE steps; returns (images/s of a full T=100 sampling at this rate, seconds)."""
        O, s = self.O, self.s
        with torch.no_grad():
            t0 = time.perf_counter()
            for _ in range(n_sde_steps):
                if self.t < 1:
                    self.t = T_STEPS
                eps = self.net(self.x, self.mu, self.t * s.sample_scale, image_context=self.ctx)
                self.x = O.reverse_step(s, self.x, self.mu, eps, torch.randn(self.x.shape, generator=self.g), self.t)
                self.t -= 1
            dt = time.perf_counter() - t0
        return 1.0 / (dt / n_sde_steps * T_STEPS), dt


REF_STEPS_PER_SAMPLE = 3


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = host_threads()
    ref = CpuReference()
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt = ref.sample(REF_STEPS_PER_SAMPLE)
        if i >= args.warmup:
            rates.append((r, dt))
    val = sum(r for r, _ in rates) / len(rates)
    sample = (f"each bench step = {REF_STEPS_PER_SAMPLE} consecutive SDE steps (UNet forward + update) of one reverse process on 1 image "
              f"@{RES}x{RES} fp32 (BASELINE config 1), warm network, extrapolated to T={T_STEPS}; oracle port of "
              "utils/sde_utils.py + App. A UNet (the reference's own network source is not in the snapshot)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(d for _, d in rates) / len(rates),
            "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def eager_gpu_baseline(steps=3):
    """PyTorch eager on the SAME GPU (the bar a GPU user starts from): tests/bench_torch_eager.py (fp32
    oracle network + unfused SDE chain, fp32 and bf16 autocast) run as a subprocess, its JSON embedded."""
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_torch_eager.py"), "--steps", str(steps),
                              "--batch", str(BATCH_PER_GPU), "--res", str(RES)], capture_output=True, text=True, timeout=600)
        d = json.loads(res.stdout.strip().splitlines()[-1])
        return {"unit": UNIT, "fp32": d["fp32"]["images_per_s"], "bf16_autocast": d["bf16_autocast"]["images_per_s"],
                "sample": d["workload"], "how": "tests/bench_torch_eager.py: oracle UNet + oracle IRSDE chain, torch eager / cuDNN"}
    except Exception as exc:                                  # a reported baseline only: never fails the bench
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch.distributed as dist
    from instancediff_b200 import ConditionalUNet, IRSDE, _lib

    rank, world, local = dist_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created; stdout must carry ONE
        # JSON line, so fd 1 points at stderr while the communicator is set up.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = BATCH_PER_GPU
    pk = peaks()

    net = ConditionalUNet(device=dev, seed=1)                # random-init weights of the App. A architecture
    sde = IRSDE(max_sigma=0.4, T=T_STEPS, schedule="cosine", eps=0.01, device=dev)
    sde.set_model(net)
    sde.noise_source = "philox"
    sde.philox_seed = 1
    sde.philox_offset = rank * B * RES * RES
    mu_h, ctx_h = synthetic_inputs(B, 1 + rank, pinned=True)
    out_h = torch.empty(B, 1, RES, RES).pin_memory()
    mu_d, ctx_d = mu_h.to(dev), ctx_h.to(dev)
    sde.set_mu(mu_d)
    xT = sde.noise_state(mu_d)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():                                     # inputs already in HBM
        return sde.reverse_sde(xT, T=-1, image_context=ctx_d)

    def step_e2e():                                          # public API with HOST buffers
        mu = mu_h.to(dev, non_blocking=True)
        ctx = ctx_h.to(dev, non_blocking=True)
        sde.set_mu(mu)
        x = sde.noise_state(mu)
        x0 = sde.reverse_sde(x, T=-1, image_context=ctx)
        out_h.copy_(x0, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_h

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, cs.summary()

    def step_e2e_default_noise():                            # the drop-in default: z = torch.randn_like (sde_utils.py:185)
        sde.noise_source = None
        try:
            return step_e2e()
        finally:
            sde.noise_source = "philox"

    ms_res, clocks = timed(step_resident)
    ms_e2e, clocks_e2e = timed(step_e2e)
    ms_e2e_dn, _ = timed(step_e2e_default_noise)
    total_images = B * world * args.steps
    value = total_images / (ms_res / 1e3)
    e2e_value = total_images / (ms_e2e / 1e3)
    sde.set_mu(mu_d)

    # ---- live per-kernel timing (CUDA events on the launching stream) for the roofline ------------
    plan = net._plan(B, RES, RES, True)
    rows = plan.run_timed(xT, mu_d, 50.0, reps=3)
    by_kind = {}
    split = {"kxk": dict(ms=0.0, flops=0.0, bytes=0.0, n=0), "1x1": dict(ms=0.0, flops=0.0, bytes=0.0, n=0)}
    for kind, label, flops, ms, nbytes in rows:
        k = by_kind.setdefault(kind, dict(ms=0.0, flops=0.0, n=0))
        k["ms"] += ms
        k["flops"] += flops
        k["n"] += 1
        if kind == "conv_gemm":                              # 3x3 / 4x4 convolutions vs 1x1 (linear) layers
            c = split["1x1" if label.startswith("k1") else "kxk"]
            c["ms"] += ms
            c["flops"] += flops
            c["bytes"] += nbytes
            c["n"] += 1
    fwd_ms = sum(k["ms"] for k in by_kind.values())
    conv = by_kind["conv_gemm"]
    rp = dict(ms=sum(ms for kind, label, _, ms, _ in rows if kind == "conv_gemm" and label.endswith("rowpair")),
              flops=sum(fl for kind, label, fl, _, _ in rows if kind == "conv_gemm" and label.endswith("rowpair")),
              n=sum(1 for kind, label, _, _, _ in rows if kind == "conv_gemm" and label.endswith("rowpair")))
    conv_tflops = conv["flops"] / (conv["ms"] / 1e3) / 1e12
    # fused SDE step, timed alone with events (16 B/element with the in-kernel Philox draw)
    n_el = xT.numel()
    row = sde._coef_table(dev)[50]
    s_ptr = torch.cuda.current_stream(dev).cuda_stream

    def time_sde(n, nsets, reps):
        """(us per launch inside a captured graph of 20 launches -- how the sampling loop runs it --, us per launch of
        `reps` eager back-to-back launches, which adds the host-side launch path).  Launches rotate over `nsets`
        buffer sets (x, eps, mu) so that one cycle moves more bytes than the 126 MB L2 holds: every launch reads
        its inputs from HBM, as in the loop, where 2 GB of activations pass through L2 between two updates."""
        sets = [tuple(torch.randn(n, device=dev) for _ in range(3)) for _ in range(nsets)]

        def launch(stream, i):
            xs_, eps_, mu_ = sets[i % nsets]
            _lib.check(_lib.lib().idiff_sde_step(xs_.data_ptr(), xs_.data_ptr(), eps_.data_ptr(), mu_.data_ptr(), None,
                                                 row.data_ptr(), 0, 1, 1, 0, n, stream), "sde_step")
        for i in range(nsets):
            launch(s_ptr, i)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        gr = torch.cuda.CUDAGraph()
        nl = 20 if 20 % nsets == 0 else 2 * nsets
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for i in range(nl):
                    launch(side.cuda_stream, i)
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(2):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        us_graph = e0.elapsed_time(e1) / (10 * nl) * 1e3
        e0.record()
        for i in range(reps):
            launch(s_ptr, i)
        e1.record()
        torch.cuda.synchronize(dev)
        del gr, sets
        return us_graph, e0.elapsed_time(e1) / reps * 1e3

    # 16 B per element per launch; rotate over enough sets that a cycle exceeds 2 x L2
    sde_sets = max(2, -(-(2 * 126 * 1024 * 1024) // (16 * n_el)))
    sde_us, sde_us_eager = time_sde(n_el, sde_sets, 48)
    sde_gbs = 16.0 * n_el / (sde_us * 1e-6) / 1e9
    # the same kernel on the B=256 state of BASELINE config 3 (268 MB of traffic per launch)
    n_big = 8 * n_el
    sde_big_us, sde_big_us_eager = time_sde(n_big, 2, 20)
    sde_big_gbs = 16.0 * n_big / (sde_big_us * 1e-6) / 1e9

    # DRAM bytes of the conv_gemm launches of one forward from the committed ncu capture of this configuration
    # (profiles/r01s2_final_conv_gemm_dram_bytes_ncu.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch)
    conv_labels = [label for kind, label, _, _, _ in rows if kind == "conv_gemm"]
    conv_traffic, traffic_src = conv_traffic_from_profile(conv_labels) if B == 32 and RES == 256 else (None, None)
    conv_algo_bytes = sum(nb for kind, _, _, _, nb in rows if kind == "conv_gemm")
    la_rows = [(fl, ms, nb) for kind, _, fl, ms, nb in rows if kind == "linattn_fused"]
    la_n, la_flops, la_ms, la_bytes = len(la_rows), sum(r[0] for r in la_rows), sum(r[1] for r in la_rows), sum(r[2] for r in la_rows)
    launches_per_fwd = plan.n_launch
    # per captured step: the plan's launches minus the two time-embedding kernels (their output is a table row copied by
    # the step-select kernel) plus step select and the fused SDE update
    gpu_launches = args.steps * T_STEPS * (launches_per_fwd - 2 + 2)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world),
        "impl_detail": {"path": "bf16 tcgen05 UNet + fused SDE step, one CUDA graph per step replayed T times, in-kernel Philox noise",
                        "l2": "working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush",
                        "peaks": pk["src"]},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "power_w_max": clocks.get("power_w_max")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (mu_h.numel() + ctx_h.numel()) * 4,
                "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": ms_e2e / args.steps},
        "e2e_default_noise": {"value": total_images / (ms_e2e_dn / 1e3), "unit": UNIT, "ms_per_step": ms_e2e_dn / args.steps,
                              "note": "same end-to-end call with noise_source=None: z = torch.randn_like(x) drawn by torch's "
                                      "CUDA generator INSIDE the captured step (the reference's default, utils/sde_utils.py:185); "
                                      "20 instead of 16 bytes per element and one more launch per step"},
        "gpu_launches": gpu_launches,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel + conv3_rowpair_kernel (tcgen05 implicit GEMM, all conv/linear layers)",
                     "achieved": conv_tflops, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / pk["tf_sustained"], "frac_of_burst": conv_tflops / pk["tf_burst"],
                     "traffic": conv_traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": conv_algo_bytes / conv["n"],
                     "launches_per_forward": conv["n"], "ms_per_forward": conv["ms"],
                     "share_of_forward": conv["ms"] / fwd_ms,
                     "note": "achieved = sum of algorithmic 2*M*N*K over the conv launches of one forward / sum of their CUDA-event "
                             "durations (3 instrumented replays right after the timed loops, chip under the same power cap); peak = "
                             "MEASURED_PEAKS bf16 sustained, frac_of_burst = against the burst figure"},
        # the same launches split by what bounds them (tensor roofline for the convolutions,
        # HBM roofline for the 1x1 / linear layers, whose arithmetic intensity is below the ridge)
        "roofline_conv_kxk": {"bound": "tensor", "kernel": "conv_gemm_kernel, 3x3 and 4x4/s2 convolutions",
                              "achieved": split["kxk"]["flops"] / (split["kxk"]["ms"] / 1e3) / 1e12, "peak": pk["tf_sustained"],
                              "unit": "TFLOP/s",
                              "frac": split["kxk"]["flops"] / (split["kxk"]["ms"] / 1e3) / 1e12 / pk["tf_sustained"],
                              "frac_of_burst": split["kxk"]["flops"] / (split["kxk"]["ms"] / 1e3) / 1e12 / pk["tf_burst"],
                              "launches_per_forward": split["kxk"]["n"], "ms_per_forward": split["kxk"]["ms"],
                              "rowpair": {"launches": rp["n"], "ms_per_forward": rp["ms"],
                                          "achieved": (rp["flops"] / (rp["ms"] / 1e3) / 1e12) if rp["ms"] > 0 else None},
                              "note": "3x3 64->64 layers run on the row-pair kernel (N-stacked output rows, shared-memory-bandwidth "
                                      "bound: 14 KB of operands per 96 tensor cycles); the other Cout = 64 layers on M128xN64 MMAs "
                                      "of the generic engine; see DESIGN.md section 4"},
        "roofline_linear_1x1": {"bound": "hbm", "kernel": "conv_gemm_kernel, 1x1 / linear layers",
                                "achieved": split["1x1"]["bytes"] / (split["1x1"]["ms"] / 1e3) / 1e9, "peak": pk["hbm"],
                                "unit": "GB/s", "frac": split["1x1"]["bytes"] / (split["1x1"]["ms"] / 1e3) / 1e9 / pk["hbm"],
                                "tflops": split["1x1"]["flops"] / (split["1x1"]["ms"] / 1e3) / 1e12,
                                "launches_per_forward": split["1x1"]["n"], "ms_per_forward": split["1x1"]["ms"]},
        "roofline_linattn": {"bound": "hbm", "kernel": "la_ctx + la_merge + la_out2 (fused linear attention blocks of the C = 64 / 128 levels)",
                             "achieved": la_bytes / (la_ms / 1e3) / 1e9 if la_ms > 0 else None, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": (la_bytes / (la_ms / 1e3) / 1e9 / pk["hbm"]) if la_ms > 0 else None,
                             "tflops": (la_flops / (la_ms / 1e3) / 1e12) if la_ms > 0 else None,
                             "blocks_per_forward": la_n, "ms_per_forward": la_ms,
                             "note": "algorithmic bytes = x read by both passes + the result + the row statistics; the measured bound is "
                                     "the TMEM read port (16 B/clk per lane quadrant) and the MUFU, not HBM: DESIGN.md section 4"},
        "roofline_sde": {"bound": "hbm", "kernel": "sde_step_kernel", "achieved": sde_gbs, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": sde_gbs / pk["hbm"], "us_per_launch": sde_us, "us_per_launch_eager": sde_us_eager,
                         "bytes_per_element": 16, "timing": f"CUDA events around 10 replays of a captured graph of launches rotating over {sde_sets} buffer sets "
                         "(one cycle > 2x the 126 MB L2, inputs come from HBM as in the loop); the eager figure adds the host launch path",
                         "at_batch_256": {"elements": n_big, "us_per_launch": sde_big_us, "us_per_launch_eager": sde_big_us_eager, "achieved": sde_big_gbs,
                                          "frac": sde_big_gbs / pk["hbm"],
                                          "note": "same kernel on the B=256 state (268 MB per launch, exceeds L2)"}},
        "forward_breakdown_ms": {k: round(v["ms"], 4) for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1]["ms"])},
        "forward_ms_sum_of_kernels": fwd_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        val, dt = CpuReference().sample(args.cpu_steps)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_steps} of {T_STEPS} SDE steps on 1 image @{RES}x{RES} fp32 (BASELINE config 1) after one "
                                          f"warm-up step, {dt:.1f} s of CPU work, extrapolated to T={T_STEPS}"}
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        del xT
        torch.cuda.empty_cache()
        line["eager_gpu_baseline"] = eager_gpu_baseline()
    if args.dump_kernels and rank == 0:
        with open(args.dump_kernels, "w") as f:
            f.write("kind,label,gflop,ms,tflops,algo_mbytes,algo_gbs\n")
            for kind, label, flops, ms, nbytes in rows:
                f.write(f"{kind},{label},{flops / 1e9:.3f},{ms:.4f},{(flops / (ms / 1e3) / 1e12) if ms > 0 else 0:.2f},"
                        f"{nbytes / 1e6:.1f},{(nbytes / (ms / 1e3) / 1e9) if ms > 0 else 0:.0f}\n")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """BASELINE configs[4]: one training step = fused state sampling (idiff_random_states) -> UNet forward + backward
    (PyTorch autograd / cuDNN under bf16 autocast: v1) -> bucketed gradient all-reduce over NCCL,
    launched from backward hooks -> fused Adam.  16 crops of 256x256 per GPU (weak scaling)."""
    import torch.distributed as dist
    from instancediff_b200 import IRSDE
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet

    rank, world, local = dist_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                                       # NCCL's banner must not reach stdout (ONE JSON line)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.batch
    torch.backends.cudnn.benchmark = True
    net = TrainableUNet(seed=1).to(dev)
    sde = IRSDE(max_sigma=0.4, T=T_STEPS, schedule="cosine", eps=0.01, device=dev)
    sde.noise_source, sde.philox_seed, sde.philox_offset = "philox", 1, rank * B * RES * RES
    tr = NoiseMatchingTrainer(net, sde, comm_dtype=torch.bfloat16 if args.comm_bf16 else None, bucket_mb=args.bucket_mb)
    g = torch.Generator().manual_seed(10 + rank)
    x0_h = (torch.rand(B, 1, RES, RES, generator=g) * 2 - 1).pin_memory()
    mu_h = (x0_h + 0.1 * torch.randn(B, 1, RES, RES, generator=g)).clamp(-1, 1).pin_memory()
    ctx_h = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1).pin_memory()
    x0_d, mu_d, ctx_d = x0_h.to(dev), mu_h.to(dev), ctx_h.to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return tr.step(x0_d, mu_d, ctx_d)

    def step_e2e():                                          # host batch in, loss value out (trainUM.py:255-277 logs it)
        loss = tr.step(x0_h.to(dev, non_blocking=True), mu_h.to(dev, non_blocking=True), ctx_h.to(dev, non_blocking=True))
        return float(loss.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, cs.summary()

    ms_res, clocks = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, args.steps, 1)
    comm = None
    if tr.reducer is not None:
        red = tr.reducer
        ms_comm, _ = timed(lambda: red.reduce().wait(), 5, 2)               # the bucket all-reduces alone
        red.detach()
        saved_red, tr.reducer = tr.reducer, None
        ms_nocomm, _ = timed(step_resident, args.steps, 1)                   # same step without any collective
        tr.reducer = saved_red.attach()
        t_comm, nbytes = ms_comm / 5, red.bytes_per_step
        exposed = max(0.0, ms_res / args.steps - ms_nocomm / args.steps)
        comm = {"allreduce_bytes_per_step": nbytes, "buckets": len(red.buckets), "wire_dtype": "bf16" if args.comm_bf16 else "f32",
                "allreduce_alone_ms": t_comm, "bus_gbs": 2.0 * (world - 1) / world * nbytes / (t_comm * 1e-3) / 1e9,
                "step_ms_without_collective": ms_nocomm / args.steps, "exposed_ms": exposed,
                # the two step times differ by run-to-run noise of ~1 %; an all-reduce shorter than that cannot be resolved
                "overlap_fraction": (max(0.0, min(1.0, 1.0 - exposed / t_comm)) if t_comm > 0.02 * ms_nocomm / args.steps else None),
                "overlap_note": "overlap = 1 - (step with collective - step without) / all-reduce alone; null when the all-reduce "
                                "alone is < 2 % of a step (below the run-to-run spread of the step time): its cost is then bounded "
                                "by allreduce_alone_ms / ms_per_step",
                "reference_bus_gbs": "725 GB/s: measured 8-rank all-reduce bus bandwidth at 1 GiB on this pool (B200_PROFILING.md)"}
    total = B * world * args.steps
    line = {"metric": "images/sec, training step (UNet fwd+bwd, noise-matching loss, gradient all-reduce, Adam) @256^2",
            "value": total / (ms_res / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 autocast (fp32 master weights, fp32 Adam)", "data": "synthetic",
            "config": {"workload": f"training step, batch {B} x 1x{RES}x{RES} crops per GPU, noise-matching MSE loss (BASELINE configs[4])",
                       "sde": "IRSDE(max_sigma=0.4,T=100,cosine,eps=0.01)", "global_batch": B * world,
                       "parallelism": f"data parallel x{world}, bucketed NCCL all-reduce of {sum(p.numel() for p in net.parameters())} gradients"},
            "impl_detail": "v1: idiff_random_states (fused x_t + target) -> torch autograd / cuDNN forward+backward -> "
                           "GradientAllReducer (async NCCL buckets from backward hooks) -> torch fused Adam; native dgrad/wgrad not built",
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": (x0_h.numel() + mu_h.numel() + ctx_h.numel()) * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps, "gpu_launches_note": "own kernels per step: 1 (idiff_random_states); the rest is cuDNN / ATen / NCCL",
            "collective": comm}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global RES, BATCH_PER_GPU, METRIC, SCALING
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--mode", default="sample", choices=["sample", "train"], help="train = BASELINE configs[4] (training step)")
    ap.add_argument("--comm-bf16", action="store_true", help="train mode: bf16 gradient buckets on the wire")
    ap.add_argument("--bucket-mb", type=float, default=25.0, help="train mode: all-reduce bucket size (DDP default 25 MB)")
    ap.add_argument("--res", type=int, default=RES, help="image side; 256 = the BASELINE metric's configuration, "
                    "512 = the LoDoPaB-CT-shaped configs[3] (extra bench lines, not the headline)")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU (32 = configs[1])")
    ap.add_argument("--cpu-steps", type=int, default=48, help="SDE steps of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-this-GPU comparison leg")
    ap.add_argument("--global-batch", type=int, default=0, help="STRONG scaling: this many images in total, split evenly over "
                    "the GPUs (BASELINE configs[2]: --global-batch 256 --gpus 2/4/8)")
    ap.add_argument("--dump-kernels", default=None, help="write the per-launch timing table (CSV) here")
    args = ap.parse_args()
    if args.global_batch:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_batch % world:
            ap.error("--global-batch must be divisible by the number of GPUs")
        args.batch, SCALING = args.global_batch // world, "strong"
    if (args.res, args.batch) != (RES, BATCH_PER_GPU):
        if args.res % 16 or args.res < 32 or args.batch < 1:
            ap.error("--res must be a multiple of 16 (>= 32) and --batch >= 1")
        METRIC = METRIC.replace(f"@{RES}^2", f"@{args.res}^2")
        RES, BATCH_PER_GPU = args.res, args.batch
    if args.mode == "train":
        if args.batch == 32 and "--batch" not in " ".join(sys.argv):
            args.batch = 16                                  # configs[4]: 16 crops per GPU
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference's training step (trainUM.py) needs its unshipped "
                              "models/modules and clip/ema_pytorch; no CPU arm for --mode train"}), flush=True)
            return
        RES = args.res
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
