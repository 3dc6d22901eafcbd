"""Per-role cycle breakdown (CTA 0) and timing of idiff_conv3_rowpair on the bench's 64 -> 64 layer shapes.

Counters need the profiling build (make ... OBJDIR=build_prof OUT=../libidiff_prof.so EXTRA=-DIDIFF_PROF and
IDIFF_LIB_PATH=instancediff_b200/libidiff_prof.so); with the production library only the timings are valid.
IDIFF_RP_ONLY=plain|affine restricts the run to one variant (ncu target)."""
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from instancediff_b200 import _lib, ops  # noqa: E402
from instancediff_b200.packing import pack_conv3_rowpair  # noqa: E402

NAMES = ["ld:wait_empty", "ld:issue", "ld:data+store", "mma:wait_tmem", "mma:wait_top", "mma:wait_bottom", "mma:issue",
         "epi:wait_full", "epi:work"]


def run(B, H, W, affine, gn=True, label="", reps=5, dbg=1, packed=False):
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    src = torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / math.sqrt(576)
    out = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=dev)
    kw = {}
    if affine:
        kw.update(a_scale=torch.ones(B, 64, device=dev), a_shift=torch.zeros(B, 64, device=dev), a_silu=3 if packed else 1)
    if gn:
        kw.update(gn_groups=8, gn_partial=torch.zeros(B, _lib.lib().idiff_conv3_rowpair_gn_rows(H, W), 8, 2, device=dev))
    p = ops.make_gemm_params(B=B, H=H, W=W, ksize=3, stride=1, cin0=64, N=64, NT=64, epi=0, out_ld=64, src0=src,
                             w=pack_conv3_rowpair(w).to(dev), bias=torch.zeros(64, device=dev), out=out, reserved0=dbg, **kw)
    for _ in range(3):
        ops.conv3_rowpair(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv3_rowpair(p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    buf = (ctypes.c_ulonglong * 16)()
    if _lib.lib().idiff_debug_read_prof(buf) != 0:
        buf = (ctypes.c_ulonglong * 16)()
    v = list(buf)
    items = max(v[1], 1)
    flops = 2.0 * B * H * W * 64 * 576
    print(f"== {label}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {2.0 * 2 * B * H * W * 64 / ms / 1e6:.0f} GB/s algorithmic; "
          f"kernel {v[0]} cyc, {items} items on CTA0, {v[0] / items:.0f} cyc/item")
    print("   " + "  ".join(f"{n}={x / items:.0f}" for n, x in zip(NAMES, v[2:11])))


if __name__ == "__main__":
    B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
    only = os.environ.get("IDIFF_RP_ONLY", "")
    if only in ("", "plain"):
        run(B, 256, 256, False, label="rowpair 3x3 64->64 plain+gn @256")
    if only in ("", "affine"):
        run(B, 256, 256, True, label="rowpair 3x3 64->64 affine+silu+gn @256")
    if only in ("", "packed"):
        run(B, 256, 256, True, label="rowpair 3x3 64->64 affine+silu (bf16x2)+gn @256", packed=True)
    if only == "":
        run(B, 128, 128, False, label="rowpair plain+gn @128")
        run(B, 128, 128, True, label="rowpair affine+silu+gn @128")
        run(2, 512, 512, True, label="rowpair affine+silu+gn @512 B=2")
        run(1, 224, 224, True, label="rowpair affine+silu+gn @224 B=1")
