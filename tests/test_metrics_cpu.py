"""Result-side utilities (utils/img_utils.py:136-238, testUM.py:151-164) against the reference's own functions run on
seeded images (oracle/gen_golden_metrics.py -> tests/golden/metrics.npz).  Host logic: runs without a GPU."""
import math
import os

import numpy as np
import pytest
import torch

from instancediff_b200 import metrics as M


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics.npz"))


def test_tensor2img_bit_exact(g):
    t = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    assert np.array_equal(M.tensor2img(t["smooth"]), g["img_a8"]) and M.tensor2img(t["smooth"]).dtype == np.uint8
    assert np.array_equal(M.tensor2img(t["noisy"]), g["img_b8"])                      # clamp to [0, 1] first
    assert np.array_equal(M.tensor2img(t["smooth"], out_type=np.float32), g["img_af"])
    assert np.array_equal(M.tensor2img(t["noisy"], out_type=np.float32, min_max=(-0.5, 1.5)), g["img_bf"])
    assert np.array_equal(M.tensor2img(t["rgb"]), g["img_c8"])                         # CHW RGB -> HWC BGR
    assert np.array_equal(M.tensor2img(t["batch"]), g["img_grid"])                     # 4-D -> make_grid(nrow=2)
    x = t["smooth"].clone()
    M.tensor2img(x)
    assert torch.equal(x, t["smooth"])                                                 # the caller's tensor is not clamped in place
    with pytest.raises(TypeError):
        M.tensor2img(torch.zeros(2, 2, 2, 2, 2))


def test_psnr_matches_reference(g):
    assert M.calculate_psnr(g["img_a8"], g["img_b8"]) == float(g["psnr_gray"])
    assert M.calculate_psnr(g["img_c8"], g["img_d8"]) == float(g["psnr_rgb"])
    assert math.isinf(M.calculate_psnr(g["img_a8"], g["img_a8"]))


def test_ssim_matches_reference(g):
    assert M.calculate_ssim(g["img_a8"], g["img_b8"]) == pytest.approx(float(g["ssim_gray"]), abs=1e-12)
    assert M.calculate_ssim(g["img_c8"], g["img_d8"]) == pytest.approx(float(g["ssim_rgb"]), abs=1e-12)
    assert M.calculate_ssim(g["img_a8"][:, :, None], g["img_b8"][:, :, None]) == pytest.approx(float(g["ssim_hw1"]), abs=1e-12)
    assert M.ssim_unit_range(g["unit_a"], g["unit_b"]) == pytest.approx(float(g["ssim_unit"]), abs=1e-12)
    assert M.ssim_unit_range(g["unit_a"][None, None], g["unit_a"][None, None]) == pytest.approx(1.0, abs=1e-12)
    with pytest.raises(ValueError):
        M.calculate_ssim(g["img_a8"], g["img_b8"][:-1])
    with pytest.raises(ValueError):
        M.calculate_ssim(np.zeros((2, 2, 2, 2)), np.zeros((2, 2, 2, 2)))
    assert M.calculate_ssim(np.zeros((16, 16, 2)), np.zeros((16, 16, 2))) is None      # reference falls through (:228-235)
