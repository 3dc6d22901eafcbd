"""NCCL leg of the gradient all-reducer (config 5's only collective): needs two GPUs, skipped otherwise."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from instancediff_b200.parallel import GradientAllReducer
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)

    def make():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Conv2d(2, 64, 3, padding=1), torch.nn.SiLU(), torch.nn.Conv2d(64, 64, 3, padding=1),
                                   torch.nn.SiLU(), torch.nn.Conv2d(64, 1, 3, padding=1)).to(dev)
    xs = [torch.randn(4, 2, 32, 32, generator=torch.Generator().manual_seed(10 + r)).to(dev) for r in range(world)]
    want = None
    for r in range(world):
        m = make()
        m(xs[r]).abs().mean().backward()                      # noise-matching L1 loss shape
        g = [p.grad.clone() for p in m.parameters()]
        want = g if want is None else [a + b for a, b in zip(want, g)]
    want = [w / world for w in want]
    net = make()
    red = GradientAllReducer(net.parameters(), bucket_mb=0.05).attach()
    net(xs[rank]).abs().mean().backward()
    red.wait()
    torch.cuda.synchronize(dev)
    ok = len(red.buckets) >= 2 and all(torch.allclose(p.grad, w, atol=1e-6, rtol=1e-5) for p, w in zip(net.parameters(), want))
    dist.barrier()
    q.put(bool(ok))
    dist.destroy_process_group()


def test_two_gpu_nccl_gradient_allreduce():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(results) and all(p.exitcode == 0 for p in procs)
