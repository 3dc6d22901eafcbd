"""Per-role cycle breakdown of la_out2 (CTA (0, 0)) -- needs the profiling build:
make -C instancediff_b200/csrc OBJDIR=build_prof OUT=../libidiff_prof.so EXTRA=-DIDIFF_PROF, then
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so IDIFF_LA_PROF=1 python tools/prof_linattn.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import _lib, ops  # noqa: E402

NAMES = ["kernel", "tiles", "sm:wait q_full", "sm:wait s_empty", "sm:work", "ln:wait o_full", "ln:work", "mma:wait x_full",
         "mma:wait q_empty", "mma:wait s_full", "mma:wait o_empty", "mma:issue", "tma:wait x_empty", "prologue", "mma:wait weights"]
B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
g = torch.Generator().manual_seed(0)
for HWs, Cc in [(256, 64), (128, 128)]:
    x = torch.randn(B, HWs, HWs, Cc, generator=g).cuda().to(torch.bfloat16)
    xf = x.float()
    mean, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    stats = torch.cat([mean, torch.rsqrt(var + 1e-5)], -1).reshape(-1, 2).contiguous()
    del xf
    wqkv = ((torch.rand(384, Cc, generator=g) * 2 - 1) / Cc ** 0.5).cuda()
    args = (x, stats, wqkv, torch.ones(Cc).cuda(), (torch.rand(Cc, 128, generator=g) / 11).cuda(), torch.zeros(Cc).cuda(),
            torch.ones(Cc).cuda())
    for _ in range(3):
        ops.linattn_fused(*args)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    _lib.check(_lib.lib().idiff_debug_read_prof(buf), "read_prof")
    n = max(1, buf[1])
    print(f"la_out2 C{Cc} @{HWs}x{HWs} B{B}: CTA (0,0), {buf[1]} tiles, {buf[0]} cycles = {buf[0] / n:.0f} per tile")
    for i, name in enumerate(NAMES):
        if i >= 2:
            print(f"   {name:20s} {buf[i]:10d}  {buf[i] / n:8.0f} / tile")
