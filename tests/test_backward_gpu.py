"""Native data gradients (instancediff_b200/backward.py): dgrad of the stride-1 convolutions on the tcgen05 engine vs
torch autograd of the same layer on the same bf16-rounded operands."""
import math

import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16r, describe, no_tf32, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    no_tf32()
    yield
    from instancediff_b200 import _lib
    _lib.watchdog()


@pytest.mark.parametrize("case", [(2, 32, 128, 64, 64, 3), (1, 16, 24, 128, 64, 3), (1, 16, 16, 256, 256, 3), (2, 32, 32, 128, 256, 1),
                                  (1, 64, 256, 64, 64, 3)], ids=lambda c: "x".join(map(str, c)))
def test_dgrad_matches_autograd(case):
    from instancediff_b200.backward import conv_dgrad
    B, H, W, cin, N, k = case
    g = torch.Generator().manual_seed(H + N)
    w = ((torch.rand(N, cin, k, k, generator=g) * 2 - 1) / math.sqrt(cin * k * k)).cuda()
    gy = torch.randn(B, H, W, N, generator=g).cuda().to(torch.bfloat16)
    x = torch.zeros(B, cin, H, W, device="cuda", requires_grad=True)
    y = F.conv2d(x, bf16r(w), padding=k // 2)
    (ref,) = torch.autograd.grad(y, x, gy.float().permute(0, 3, 1, 2))
    got = conv_dgrad(gy, w)
    torch.cuda.synchronize()
    assert rel_err(got, ref.permute(0, 2, 3, 1)) <= 1e-2, describe(got, ref.permute(0, 2, 3, 1), "dgrad " + str(case))


def test_autograd_function_trains_a_two_layer_stack():
    """Forward and dgrad on the engine, wgrad from PyTorch: gradients equal those of the fp32 torch stack on the same
    (bf16-rounded) operands, and a few SGD steps reduce the loss."""
    from instancediff_b200.backward import Conv3x3
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 32, 128
    x = torch.randn(B, H, W, 64, generator=g).cuda().to(torch.bfloat16)
    tgt = torch.randn(B, H, W, 64, generator=g).cuda()
    w1 = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).cuda().requires_grad_(True)
    w2 = ((torch.rand(64, 64, 3, 3, generator=g) * 2 - 1) / 24).cuda().requires_grad_(True)
    b1 = torch.zeros(64, device="cuda", requires_grad=True)

    def loss_engine():
        h = F.silu(Conv3x3.apply(x, w1, b1).float()).to(torch.bfloat16)
        return F.mse_loss(Conv3x3.apply(h, w2, None).float(), tgt)

    def loss_torch():
        xn = x.float().permute(0, 3, 1, 2)
        h = bf16r(F.silu(bf16r(F.conv2d(xn, bf16r(w1), b1, padding=1))))
        return F.mse_loss(bf16r(F.conv2d(h, bf16r(w2), padding=1)).permute(0, 2, 3, 1), tgt)

    ge = torch.autograd.grad(loss_engine(), (w1, w2, b1))
    gt = torch.autograd.grad(loss_torch(), (w1, w2, b1))
    for a, b, name in zip(ge, gt, ("w1", "w2", "b1")):
        assert rel_err(a, b) <= 3e-2, describe(a, b, name)
    opt = torch.optim.SGD([w1, w2, b1], lr=0.5)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        l = loss_engine()
        l.backward()
        opt.step()
        losses.append(l.item())
    assert losses[-1] < losses[0], losses
