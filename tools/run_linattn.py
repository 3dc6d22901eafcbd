"""Run the fused linear-attention block alone (target for ncu launch lists / full captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import ops  # noqa: E402

B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
cases = [(256, 64), (128, 64), (128, 128), (64, 128)]
sel = os.environ.get("IDIFF_LA_CASE")
if sel:
    cases = [cases[int(sel)]]
g = torch.Generator().manual_seed(0)
for HWs, Cc in cases:
    x = torch.randn(B, HWs, HWs, Cc, generator=g).cuda().to(torch.bfloat16)
    xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    stats = torch.cat([mean, torch.rsqrt(var + 1e-5)], -1).reshape(-1, 2).contiguous()
    del xf
    wqkv = ((torch.rand(384, Cc, generator=g) * 2 - 1) / Cc ** 0.5).cuda()
    args = (x, stats, wqkv, torch.ones(Cc).cuda(), (torch.rand(Cc, 128, generator=g) / 11).cuda(), torch.zeros(Cc).cuda(),
            torch.ones(Cc).cuda())
    for _ in range(2):
        ops.linattn_fused(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.linattn_fused(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f"linattn_fused C{Cc} @{HWs}x{HWs} B{B}: {e0.elapsed_time(e1) / 5:.3f} ms")
