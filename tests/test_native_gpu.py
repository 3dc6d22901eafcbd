"""The C-level network object (idiff_unet_*, csrc/unet_plan.cu): one forward and a 4-step loop driven through those
symbols only, against the Python plan (same kernels, same parameters) and against the fp32 oracle."""
import pytest
import torch

from gpu_util import describe, no_tf32, rel_err
from oracle.unet_oracle import make_oracle_unet

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _setup():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    no_tf32()
    yield
    from instancediff_b200 import _lib
    _lib.watchdog()


@pytest.fixture(scope="module")
def nets():
    from instancediff_b200 import ConditionalUNet
    from instancediff_b200.native import NativeUNet
    oracle = make_oracle_unet(seed=1).cuda()
    py = ConditionalUNet(device="cuda")
    py.load_state_dict(oracle.state_dict())
    nat = NativeUNet(device="cuda").load_state_dict(oracle.state_dict())
    return oracle, py, nat


def _inputs(B, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    mu = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).cuda()
    x = mu + 0.4 * torch.randn(B, 1, H, W, generator=g).cuda()
    ctx = torch.nn.functional.normalize(torch.randn(B, 1, 512, generator=g), dim=-1).cuda()
    return x, mu, ctx


@pytest.mark.parametrize("shape,t", [((2, 64, 64), 37.0), ((1, 32, 48), 100.0), ((2, 256, 256), 3.0), ((1, 224, 224), 50.0)])
def test_forward_through_the_c_plan(nets, shape, t):
    oracle, py, nat = nets
    x, mu, ctx = _inputs(*shape)
    with torch.no_grad():
        ref = oracle(x, mu, t, image_context=ctx)
    a = nat(x, mu, t, image_context=ctx)
    b = py(x, mu, t, image_context=ctx)
    # same kernels with the same parameters; the two packers sum `wsum` / folded biases in a different order
    assert rel_err(a, b) <= 1e-3, describe(a, b, "C plan vs Python plan")
    assert rel_err(a, ref) <= 2e-2, describe(a, ref, "C plan vs oracle")
    assert nat.num_launches(*shape) == py._plan(shape[0], shape[1], shape[2], True).n_launch


def test_per_sample_times_and_repeated_calls(nets):
    oracle, py, nat = nets
    x, mu, ctx = _inputs(2, 32, 32, seed=5)
    tt = torch.tensor([5.0, 77.0], device="cuda")
    with torch.no_grad():
        ref = oracle(x, mu, tt, image_context=ctx)
    a = nat(x, mu, tt, image_context=ctx)
    assert rel_err(a, ref) <= 2e-2, describe(a, ref, "eps[t tensor]")
    assert torch.equal(a, nat(x, mu, tt, image_context=ctx))


def test_reverse_sde_loop_inside_the_library(nets):
    """idiff_unet_reverse_sde == the Python IRSDE loop (same Philox stream) up to the packers' rounding."""
    from instancediff_b200 import IRSDE
    _, py, nat = nets
    B, H, W, T = 2, 32, 32, 4
    x, mu, ctx = _inputs(B, H, W, seed=7)
    sde = IRSDE(0.4, T=100, schedule="cosine", eps=0.01, device=torch.device("cuda"))
    sde.set_model(py)
    sde.set_mu(mu)
    sde.noise_source, sde.philox_seed, sde.philox_offset = "philox", 9, 4096
    want = sde.reverse_sde(x, T=T, image_context=ctx)
    got = nat.reverse_sde(x, mu, ctx, sde._coef_table(x.device), T, sde.sample_scale, seed=9, offset=4096)
    torch.cuda.synchronize()
    assert rel_err(got, want) <= 1e-3, describe(got, want, "C loop vs Python loop")
    assert torch.equal(got, nat.reverse_sde(x, mu, ctx, sde._coef_table(x.device), T, sde.sample_scale, seed=9, offset=4096))
    assert torch.equal(nat.reverse_sde(x, mu, ctx, sde._coef_table(x.device), 0), x)


def test_errors_come_back_as_status_codes(nets):
    from instancediff_b200 import _lib
    from instancediff_b200.native import NativeUNet
    _, _, nat = nets
    x, mu, ctx = _inputs(1, 24, 40)
    with pytest.raises(_lib.IdiffError):
        nat(x, mu, 1.0, image_context=ctx)                      # not a multiple of 16
    with pytest.raises(_lib.IdiffError):
        NativeUNet(nf=32)                                       # configuration the kernels are not built for
    bare = NativeUNet()
    with pytest.raises(_lib.IdiffError, match="not loaded"):
        bare.load_state_dict({"init_conv.weight": torch.zeros(64, 2, 7, 7)})
