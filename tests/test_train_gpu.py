"""Training step v1 on the GPU (BASELINE config 5): fused state sampling kernel on the input side, bf16-autocast
autograd network, weights handed to the tcgen05 inference network; NCCL gradient averaging when 2 GPUs are visible."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")


def _batch(B, H, W, seed, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    mu = (x0 + 0.1 * torch.randn(B, 1, H, W, generator=g)).clamp(-1, 1)
    ctx = F.normalize(torch.randn(B, 1, 512, generator=g), dim=-1)
    return x0.to(dev), mu.to(dev), ctx.to(dev)


def test_training_steps_on_the_fused_input_kernel_and_hand_over_to_the_inference_network():
    from instancediff_b200 import ConditionalUNet, IRSDE
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet
    dev = torch.device("cuda")
    net = TrainableUNet(seed=3).to(dev)
    sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=dev)
    sde.noise_source, sde.philox_seed = "philox", 11
    tr = NoiseMatchingTrainer(net, sde, lr=1e-4)
    x0, mu, ctx = _batch(4, 64, 64, 5)
    ts = torch.tensor([20, 45, 70, 95]).reshape(4, 1, 1, 1)
    losses = [tr.step(x0, mu, ctx, timesteps=ts).item() for _ in range(8)]
    assert all(l == l and l < 1e4 for l in losses) and losses[-1] < losses[0], losses
    # the noise target of the step is the kernel's own draw: x_t - mu_bar == sigma_bar * noise (utils/sde_utils.py:222)
    t, xt = sde.generate_random_states(x0, mu, timesteps=ts)
    noises = sde.last_noises.clone()
    real = []
    for i in range(4):
        sde.set_mu(mu[i:i + 1])
        real.append(sde.get_real_noise(xt[i:i + 1], x0[i:i + 1], int(ts[i])).squeeze(0))
    sde.set_mu(mu)
    assert torch.allclose(torch.stack(real), noises, atol=2e-4), (torch.stack(real) - noises).abs().max()
    # trained weights run on the tcgen05 path: same output within the bf16 tolerance of the forward tests
    infer = ConditionalUNet(device=dev)
    tr.sync_to(infer)
    with torch.no_grad():
        ref = net(xt, mu, ts.reshape(-1).to(dev), image_context=ctx)
    out = infer(xt, mu, ts.reshape(-1).to(dev).float(), image_context=ctx)
    rel = ((out - ref).abs().max() / ref.abs().max()).item()
    assert rel <= 2e-2, rel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_rank(rank, world, port, sd, q):
    import torch.distributed as dist
    from instancediff_b200 import IRSDE
    from instancediff_b200.train import NoiseMatchingTrainer, TrainableUNet
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net = TrainableUNet(seed=1).load_state_dict(sd).to(dev)
    sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=dev)
    sde.noise_source, sde.philox_seed, sde.philox_offset = "philox", 2, rank * 2 * 64 * 64
    tr = NoiseMatchingTrainer(net, sde, lr=1e-4, autocast_dtype=None)
    x0, mu, ctx = _batch(4, 64, 64, 9, dev)
    ts = torch.tensor([10, 30, 60, 90]).reshape(4, 1, 1, 1)
    lo = rank * 2
    loss = tr.step(x0[lo:lo + 2], mu[lo:lo + 2], ctx[lo:lo + 2], timesteps=ts[lo:lo + 2])
    q.put((rank, float(loss), {k: v.cpu().numpy().copy() for k, v in net.state_dict().items() if k in ("init_conv.weight", "final_conv.bias")}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_nccl_step_keeps_the_ranks_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from instancediff_b200.train import TrainableUNet
    sd = TrainableUNet(seed=1).state_dict()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_nccl_rank, args=(r, 2, port, sd, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(2)], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in got[0][2]:
        a, b = torch.from_numpy(got[0][2][k]), torch.from_numpy(got[1][2][k])
        assert torch.equal(a, b), k
        assert not torch.equal(a, sd[k])
