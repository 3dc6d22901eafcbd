"""Run single conv_gemm layer configs (no instrumentation) -- target for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import prof_layers as P  # noqa: E402
from instancediff_b200 import ops  # noqa: E402

_orig = ops.make_gemm_params


def _no_prof(**kw):
    kw["reserved0"] = 0
    return _orig(**kw)


ops.make_gemm_params = _no_prof
B = int(os.environ.get("IDIFF_PROFILE_B", "32"))
which = os.environ.get("IDIFF_LAYER", "a")
if "a" in which:
    P.run(B, 256, 256, 64, 0, 64, 3, gn=True, affine=True, label="3x3 64->64 affine+silu+gn")
if "b" in which:
    P.run(B, 256, 256, 64, 64, 64, 1, res=True, label="1x1 128->64 shortcut+tail")
if "c" in which:
    P.run(B, 256, 256, 64, 0, 64, 3, gn=True, label="3x3 64->64 plain+gn")
if "d" in which:
    P.run(B, 256, 256, 64, 0, 384, 1, NT=128, stats=True, epi=1, label="1x1 64->384 qkv (ln-fold, qsoftmax)")
if "e" in which:
    P.run(B, 256, 256, 64, 64, 64, 3, gn=True, label="3x3 128->64 concat+gn")
torch.cuda.synchronize()
