"""PyTorch-eager-on-B200 comparison point ("the real bar to beat"): the fp32 oracle UNet +
the oracle IRSDE update chain on the GPU, in fp32 and under bf16 autocast, at the bench workload (batch 32,
256x256).  A bounded number of SDE steps is timed with CUDA events and extrapolated to T = 100.

Not a pytest module (no test_ prefix) and not part of the product: it lives under tests/ because it executes
oracle/.  Usage:  python tests/bench_torch_eager.py [--steps 5] > profiles/<name>.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import irsde_oracle as O  # noqa: E402
from oracle.unet_oracle import make_oracle_unet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--res", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    net = make_oracle_unet(seed=1).to(dev).eval()
    B, R, T = args.batch, args.res, 100
    g = torch.Generator().manual_seed(1)
    mu = (torch.rand(B, 1, R, R, generator=g) * 2 - 1).to(dev)
    ctx = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1).to(dev)
    s = O.make_schedule(0.4, T, schedule="cosine", eps=0.01)
    s = O.Schedule(s.T, s.max_sigma, s.sample_T, s.sample_scale, s.dt, s.thetas.to(dev), s.sigmas.to(dev),
                   s.thetas_cumsum.to(dev), s.sigma_bars.to(dev))
    x0 = O.noise_state(s, mu, torch.randn_like(mu))
    out = {"workload": f"reverse SDE, batch {B}, {R}x{R}, {args.steps} of {T} steps timed, extrapolated", "device": torch.cuda.get_device_name(0)}
    for name, ctxmgr in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with torch.no_grad(), ctxmgr:
            O.reverse_sde(s, net, x0, mu, lambda t, x: torch.randn_like(x), T=2, image_context=ctx)      # warm-up / cuDNN autotune
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            O.reverse_sde(s, net, x0, mu, lambda t, x: torch.randn_like(x), T=args.steps, image_context=ctx)
            e1.record()
            torch.cuda.synchronize()
        ms_step = e0.elapsed_time(e1) / args.steps
        out[name] = {"ms_per_sde_step": ms_step, "images_per_s": B / (ms_step * T / 1e3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
