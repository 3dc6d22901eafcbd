"""Per-role cycle breakdown of conv_gemm on the bench's dominant layer shapes (CTA 0 counters).

Needs the profiling build:  make -C instancediff_b200/csrc OBJDIR=build_prof OUT=../libidiff_prof.so EXTRA=-DIDIFF_PROF
and IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so (without it the counters read 0 and only the timings are valid)."""
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import _lib, ops  # noqa: E402
from instancediff_b200.packing import pack_conv_weight  # noqa: E402

NAMES = ["kernel", "items", "ld:wait_empty", "ld:issue", "ld:data+store", "mma:wait_tmem", "mma:wait_A", "mma:wait_B",
         "mma:issue", "epi:wait_full", "epi:work"]


def run(B, H, W, cin0, cin1, N, k, NT=None, affine=False, gn=False, res=False, stats=False, epi=0, label=""):
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    src0 = torch.randn(B, H, W, cin0, generator=g).to(dev).to(torch.bfloat16)
    src1 = torch.randn(B, H, W, cin1, generator=g).to(dev).to(torch.bfloat16) if cin1 else None
    cin = cin0 + cin1
    w = (torch.rand(N, cin, k, k, generator=g) * 2 - 1) / math.sqrt(cin * k * k)
    NT = min(N, 256) if NT is None else NT
    out = torch.empty(B, H, W, N, dtype=torch.bfloat16, device=dev)
    kw = {}
    if affine:
        kw.update(a_scale=torch.ones(B, cin, device=dev), a_shift=torch.zeros(B, cin, device=dev), a_silu=1)
    if gn:
        tiles = 4 * ((H + 15) // 16) * ((W + 7) // 8)
        kw.update(gn_groups=8, gn_partial=torch.zeros(B, tiles, 8, 2, device=dev))
    if res:
        kw.update(res0=torch.randn(B, H, W, N, generator=g).to(dev).to(torch.bfloat16),
                  res0_scale=torch.ones(B, N, device=dev), res0_shift=torch.zeros(B, N, device=dev))
    if stats:
        kw.update(row_stats=torch.ones(B * H * W, 2, device=dev), wsum=torch.zeros(N, device=dev))
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=k, stride=1, cin0=cin0, cin1=cin1, N=N, NT=NT, epi=epi, out_ld=N,
                             src0=src0, src1=src1, w=pack_conv_weight(w, NT).to(dev), bias=torch.zeros(N, device=dev),
                             out=out, qscale=1.0, reserved0=1, **kw)
    for _ in range(3):
        ops.conv_gemm(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_gemm(p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    buf = (ctypes.c_ulonglong * 16)()
    if _lib.lib().idiff_debug_read_prof(buf) != 0:      # production build: no counters, timings only
        buf = (ctypes.c_ulonglong * 16)()
    v = list(buf)
    items = max(v[1], 1)
    flops = 2.0 * B * H * W * N * cin * k * k
    print(f"== {label}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  kernel {v[0]} cyc, {items} items on CTA0, "
          f"{v[0] / items:.0f} cyc/item")
    print("   " + "  ".join(f"{n}={x / items:.0f}" for n, x in zip(NAMES[2:], v[2:11])))


if __name__ == "__main__":
    B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
    run(B, 256, 256, 64, 0, 64, 3, gn=True, label="3x3 64->64 plain+gn")
    run(B, 256, 256, 64, 0, 64, 3, gn=True, affine=True, label="3x3 64->64 affine+silu+gn")
    run(B, 256, 256, 64, 64, 64, 3, gn=True, label="3x3 128->64 concat+gn")
    run(B, 256, 256, 64, 64, 64, 1, res=True, label="1x1 128->64 shortcut+tail")
    run(B, 256, 256, 64, 0, 384, 1, NT=128, stats=True, epi=1, label="1x1 64->384 qkv (ln-fold, qsoftmax)")
    run(B, 128, 128, 128, 64, 128, 3, gn=True, label="3x3 192->128 @128")
    run(B, 32, 32, 256, 0, 256, 3, gn=True, label="3x3 256->256 @32")
    run(B, 32, 32, 256, 0, 2048, 1, NT=256, stats=True, epi=2, label="1x1 256->2048 geglu @32")
