"""Training step v1 (instancediff_b200/train.py, BASELINE config 5) on the CPU: the differentiable network equals the
fp32 oracle network (values AND gradients), and a two-rank gloo step equals the single-process step on the
concatenated batch -- the effect of the reference's DistributedDataParallel wrap (models/drift_noise_model.py:144-146)."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

from oracle import irsde_oracle as O
from oracle.unet_oracle import make_oracle_unet


class _CpuStates:
    """Stand-in for IRSDE.generate_random_states on the CPU (the product's version is a CUDA kernel): the oracle's
    restatement of utils/sde_utils.py:322-338 with a seeded generator."""

    def __init__(self, seed):
        self.s = O.make_schedule(0.4, 100, schedule="cosine", eps=0.01)
        self.g = torch.Generator().manual_seed(seed)
        self.last_noises = None

    def generate_random_states(self, x0, mu, timesteps=None):
        B = x0.shape[0]
        t = torch.randint(1, 101, (B, 1, 1, 1), generator=self.g) if timesteps is None else timesteps
        z = torch.randn(x0.shape, generator=self.g)
        self.last_noises = z
        decay = torch.exp(-self.s.thetas_cumsum[t] * self.s.dt)
        return t, z * self.s.sigma_bars[t] + (mu + (x0 - mu) * decay)


def _batch(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    mu = (x0 + 0.1 * torch.randn(B, 1, H, W, generator=g)).clamp(-1, 1)
    ctx = F.normalize(torch.randn(B, 1, 512, generator=g), dim=-1)
    return x0, mu, ctx


def test_trainable_network_equals_the_oracle_in_value_and_gradient():
    from instancediff_b200.train import TrainableUNet
    oracle = make_oracle_unet(seed=1).requires_grad_(True)
    net = TrainableUNet(seed=7)
    assert list(net.state_dict()) == list(oracle.state_dict())          # same keys, same order: the weight contract
    net.load_state_dict(oracle.state_dict())
    x0, mu, ctx = _batch(2, 32, 32, 3)
    t = torch.tensor([17, 93])
    out, ref = net(x0, mu, t, image_context=ctx), oracle(x0, mu, t, image_context=ctx)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), (out - ref).abs().max()
    tgt = torch.randn(ref.shape, generator=torch.Generator().manual_seed(4))
    F.mse_loss(out, tgt).backward()
    F.mse_loss(ref, tgt).backward()
    got = net.table()
    # a conv bias in front of a GroupNorm has an exactly-zero gradient in exact arithmetic (numerical noise on both
    # sides), so the tolerance is relative to the largest gradient of the whole network as well as to the tensor's own
    gmax = max(p.grad.abs().max().item() for p in oracle.parameters())
    worst = ("", 0.0)
    for name, p in oracle.named_parameters():
        assert got[name].grad is not None, name
        tol = 2e-3 * p.grad.abs().max().item() + 2e-6 * gmax
        r = (got[name].grad - p.grad).abs().max().item() / tol
        worst = max(worst, (name, r), key=lambda t: t[1])
    assert worst[1] < 1.0, worst
    # 224 (not a multiple of 16) takes the reflect-pad path like the oracle
    x0, mu, ctx = _batch(1, 24, 40, 5)
    with torch.no_grad():
        assert torch.allclose(net(x0, mu, 9.0, image_context=ctx), oracle(x0, mu, 9.0, image_context=ctx), rtol=1e-4, atol=1e-5)


def test_trainer_step_reduces_the_noise_matching_loss():
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet
    net = TrainableUNet(seed=2)
    tr = NoiseMatchingTrainer(net, _CpuStates(5), lr=1e-4, autocast_dtype=None)
    x0, mu, ctx = _batch(2, 16, 16, 8)
    ts = torch.tensor([40, 70]).reshape(2, 1, 1, 1)
    losses = [tr.step(x0, mu, ctx, timesteps=ts).item() for _ in range(8)]
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, sd, out_q):
    import torch.distributed as dist
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net = TrainableUNet(seed=1).load_state_dict(sd)
    tr = NoiseMatchingTrainer(net, _CpuStates(100 + rank), lr=1e-3, autocast_dtype=None, bucket_mb=8.0)
    assert tr.reducer is not None and len(tr.reducer.buckets) > 3
    x0, mu, ctx = _batch(4, 16, 16, 8)
    lo, hi = rank * 2, rank * 2 + 2
    ts = torch.tensor([10, 30, 60, 90]).reshape(4, 1, 1, 1)
    loss = tr.step(x0[lo:hi], mu[lo:hi], ctx[lo:hi], timesteps=ts[lo:hi])
    keys = ("init_conv.weight", "final_conv.bias", "mid_attn.fn.attn2.to_v.weight", "downs.1.2.fn.to_out.weight")
    # numpy arrays travel through the queue BY VALUE (torch tensors are passed as shared-memory handles that die with
    # this process -- the parent may read them after the rank has exited)
    out_q.put((rank, float(loss), {k: v.numpy().copy() for k, v in net.state_dict().items() if k in keys},
               {k: net.table()[k].grad.numpy().copy() for k in keys}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_step_equals_the_single_process_step_on_the_whole_batch():
    import torch.multiprocessing as mp
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet
    sd = TrainableUNet(seed=1).state_dict()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_rank_main, args=(r, 2, port, sd, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda g: g[0])
    got = [(r, l, {k: torch.from_numpy(v) for k, v in w.items()}, {k: torch.from_numpy(v) for k, v in gr.items()}) for r, l, w, gr in got]

    # single process, whole batch, SAME per-sample noise: rank r drew its noise from generator 100 + r
    class _Joined(_CpuStates):
        def generate_random_states(self, x0, mu, timesteps=None):
            parts, noises = [], []
            for r in range(2):
                st = _CpuStates(100 + r)
                _, xt = st.generate_random_states(x0[2 * r:2 * r + 2], mu[2 * r:2 * r + 2], timesteps[2 * r:2 * r + 2])
                parts.append(xt)
                noises.append(st.last_noises)
            self.last_noises = torch.cat(noises)
            return timesteps, torch.cat(parts)

    net = TrainableUNet(seed=1).load_state_dict(sd)
    tr = NoiseMatchingTrainer(net, _Joined(0), lr=1e-3, autocast_dtype=None)
    x0, mu, ctx = _batch(4, 16, 16, 8)
    ts = torch.tensor([10, 30, 60, 90]).reshape(4, 1, 1, 1)
    loss = tr.step(x0, mu, ctx, timesteps=ts)
    assert abs(0.5 * (got[0][1] + got[1][1]) - float(loss)) < 1e-5
    for k in got[0][2]:
        assert torch.equal(got[0][2][k], got[1][2][k]), k                      # ranks agree bit for bit after the step
        # averaged gradients == gradients of the whole-batch loss.  (The weights themselves are not compared across
        # the two runs: Adam's first update is lr * g / (|g| + eps), so a rounding-level sign flip of a near-zero
        # gradient moves a weight by 2 lr.)
        g_ref, g_avg = net.table()[k].grad, got[0][3][k]
        assert torch.equal(g_avg, got[1][3][k]), k
        assert torch.allclose(g_avg, g_ref, rtol=1e-3, atol=1e-5 * g_ref.abs().max().item()), (k, (g_avg - g_ref).abs().max())
