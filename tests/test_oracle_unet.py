"""Oracle UNet: pinned to its own committed output + SURVEY.md App. A invariants."""
import os

import numpy as np
import pytest
import torch

from oracle.unet_oracle import algorithmic_flops, make_oracle_unet


@pytest.fixture(scope="module")
def net():
    return make_oracle_unet(seed=1)


def test_param_count_and_flops(net):
    n = sum(p.numel() for p in net.parameters())
    assert n == 22_753_217                      # App. A: ~22.75 M
    conv, attn = algorithmic_flops(256, 256)
    assert (conv + attn) / 1e9 == pytest.approx(207.6, abs=0.05)
    assert conv / 1e9 == pytest.approx(180.1, abs=0.05)
    conv, attn = algorithmic_flops(512, 512)
    assert (conv + attn) / 1e9 == pytest.approx(869.0, abs=0.05)


def test_golden_output(net, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_oracle.npz"))
    x, mu, ctx = (torch.from_numpy(g[k]) for k in ("x", "mu", "ctx"))
    with torch.no_grad():
        y = net(x, mu, 37.0, image_context=ctx)
    assert torch.allclose(y, torch.from_numpy(g["eps_t37"]), rtol=1e-4, atol=1e-5)


def test_call_conventions_agree(net):
    x, mu, ctx = torch.randn(2, 1, 32, 32), torch.randn(2, 1, 32, 32), torch.randn(2, 1, 512)
    with torch.no_grad():
        a = net(x, mu, 5.0, image_context=ctx)                       # IRSDE convention (:198)
        b = net(x, mu, torch.tensor([5, 5]), ["n0", "n1"], None, image_context=ctx)  # wrapper
        c = net(x, mu, 5, image_context=ctx[:, 0])                   # [B,512] embedding
    assert torch.equal(a, b) and torch.equal(a, c)


def test_cross_attention_is_step_and_x_independent(net):
    """App. A: one context token => softmax == 1 => CrossAttn(x, ctx) = W_o W_v ctx + b_o."""
    st = net.mid_attn.fn
    ctx = torch.randn(2, 1, 512)
    with torch.no_grad():
        y1 = st.attn2(torch.randn(2, 16, 256), ctx)
        y2 = st.attn2(torch.randn(2, 16, 256), ctx)
        closed = st.attn2.to_out(st.attn2.to_v(ctx))
    assert torch.allclose(y1, y2, atol=1e-6)
    assert torch.allclose(y1, closed.expand_as(y1), atol=1e-6)


def test_non_multiple_of_16_is_padded_and_cropped(net):
    x, mu = torch.randn(1, 1, 24, 40), torch.randn(1, 1, 24, 40)
    with torch.no_grad():
        y = net(x, mu, 1.0, image_context=torch.randn(1, 1, 512))
    assert y.shape == (1, 1, 24, 40)
