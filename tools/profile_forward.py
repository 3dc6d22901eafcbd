"""One eager (no CUDA graph) reverse-SDE step at the bench shape, for ncu captures.
    ncu --set full -k regex:conv_gemm -c 4 python tools/profile_forward.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import ConditionalUNet, IRSDE  # noqa: E402

B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
RES = int(os.environ.get("IDIFF_PROFILE_RES", "256"))
STEPS = int(os.environ.get("IDIFF_PROFILE_STEPS", "2"))
dev = torch.device("cuda:0")
net = ConditionalUNet(device=dev, seed=1)
sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=dev)
sde.set_model(net)
sde.noise_source = "philox"
sde.use_cuda_graph = False
g = torch.Generator().manual_seed(1)
mu = (torch.rand(B, 1, RES, RES, generator=g) * 2 - 1).to(dev)
ctx = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1).to(dev)
sde.set_mu(mu)
x = sde.noise_state(mu)
x = sde.reverse_sde(x, T=STEPS, image_context=ctx)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
# training-side state sampler: one fused launch, captured with `-k regex:random_states`
sde.noise_source = "philox"
ts = torch.randint(1, 101, (B, 1, 1, 1), generator=g)
for _ in range(3):
    _, xt = sde.generate_random_states(x, mu, timesteps=ts)
torch.cuda.synchronize()
print("random_states ok", float(xt.abs().mean()))
