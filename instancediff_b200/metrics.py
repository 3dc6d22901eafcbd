"""Result-side image utilities of the reference, restated (host side, numpy): ``utils/img_utils.py`` is exported next to
``IRSDE`` by ``utils/__init__.py:3-4`` and ``testUM.py:151-164`` computes RMSE / PSNR / SSIM on ``x/2+0.5``.

* ``tensor2img``      ``utils/img_utils.py:136-163``  tensor (2/3/4-D, any range) -> numpy image, ``[0,255]`` uint8 by default
* ``calculate_psnr``  ``:182-189``                    PSNR of two ``[0,255]`` images, float64, ``inf`` for identical inputs
* ``ssim`` / ``calculate_ssim``  ``:192-238``         Wang et al. SSIM, 11x11 Gaussian window (sigma 1.5), "valid" region
* ``ssim_unit_range``  the same index for ``[0,1]`` images: ``testUM.py:158-161`` asks ``skimage`` for exactly this
  definition (``gaussian_weights=True, sigma=1.5, win_size=11, use_sample_covariance=False, K1=0.01, K2=0.03,
  data_range=1.0``); ``skimage`` is not installed here, so the driver reports it through this function.

Pinned against the reference's own functions on seeded images (``oracle/gen_golden_metrics.py`` ->
``tests/golden/metrics.npz``).  The window filter is evaluated separably (two 1-D passes) instead of the reference's
dense ``cv2.filter2D``, so values agree to rounding (1e-12), not bit for bit.
"""
from __future__ import annotations

import math

import numpy as np
import torch

_WIN, _SIGMA = 11, 1.5


def _gauss_window() -> np.ndarray:
    """Normalised 1-D Gaussian of ``cv2.getGaussianKernel(11, 1.5)`` (``:198``): exp(-(i-5)^2 / (2 sigma^2)) / sum."""
    x = np.arange(_WIN, dtype=np.float64) - (_WIN - 1) / 2
    w = np.exp(-(x * x) / (2.0 * _SIGMA * _SIGMA))
    return w / w.sum()


def _valid_blur(img: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Window mean over the positions where the 11x11 window fits entirely (the ``[5:-5, 5:-5]`` crop of ``:201-208``;
    the border mode of the reference's filter never reaches it)."""
    view = np.lib.stride_tricks.sliding_window_view
    rows = view(img, _WIN, axis=1) @ w                       # [H, W-10]
    return view(rows, _WIN, axis=0) @ w                      # [H-10, W-10]


def _ssim_index(a: np.ndarray, b: np.ndarray, data_range: float) -> float:
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    a, b = a.astype(np.float64), b.astype(np.float64)
    w = _gauss_window()
    mu_a, mu_b = _valid_blur(a, w), _valid_blur(b, w)
    var_a = _valid_blur(a * a, w) - mu_a * mu_a
    var_b = _valid_blur(b * b, w) - mu_b * mu_b
    cov = _valid_blur(a * b, w) - mu_a * mu_b
    index = ((2 * mu_a * mu_b + c1) * (2 * cov + c2)) / ((mu_a * mu_a + mu_b * mu_b + c1) * (var_a + var_b + c2))
    return float(index.mean())


def ssim(img1: np.ndarray, img2: np.ndarray) -> float:
    """``utils/img_utils.py:192-213``: 2-D images in [0, 255]."""
    return _ssim_index(img1, img2, 255.0)


def ssim_unit_range(img1: np.ndarray, img2: np.ndarray) -> float:
    """The SSIM ``testUM.py:158-161`` requests, for 2-D images in [0, 1]."""
    return _ssim_index(np.squeeze(img1), np.squeeze(img2), 1.0)


def calculate_ssim(img1: np.ndarray, img2: np.ndarray):
    """``utils/img_utils.py:216-238``.  Faithful to two quirks: a 3-channel pair is scored three times on the WHOLE
    H x W x 3 array (``:229-231`` passes ``img1, img2``, not the channel), and an unsupported 3-D channel count falls
    through and returns ``None``."""
    if not img1.shape == img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.ndim == 2:
        return ssim(img1, img2)
    if img1.ndim == 3:
        if img1.shape[2] == 3:
            return np.array([_ssim_hwc(img1, img2) for _ in range(3)]).mean()
        if img1.shape[2] == 1:
            return ssim(np.squeeze(img1), np.squeeze(img2))
        return None
    raise ValueError("Wrong input image dimensions.")


def _ssim_hwc(a: np.ndarray, b: np.ndarray) -> float:
    """The reference's ``ssim`` applied to an H x W x 3 array: ``cv2.filter2D`` filters every channel plane, and the
    ``[5:-5, 5:-5]`` crop acts on H and W only; the mean runs over all three planes."""
    return float(np.mean([_ssim_index(a[:, :, c], b[:, :, c], 255.0) for c in range(a.shape[2])]))


def calculate_psnr(img1: np.ndarray, img2: np.ndarray) -> float:
    """``utils/img_utils.py:182-189``: images in [0, 255]."""
    mse = np.mean((img1.astype(np.float64) - img2.astype(np.float64)) ** 2)
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def tensor2img(tensor: torch.Tensor, out_type=np.uint8, min_max=(0, 1)) -> np.ndarray:
    """``utils/img_utils.py:136-163``: squeeze, clamp to ``min_max``, rescale to [0, 1]; 4-D -> ``make_grid`` with
    ``nrow = int(sqrt(n))`` then HWC in BGR order, 3-D -> HWC BGR, 2-D as is; uint8 output is ``round(255 x)``."""
    t = tensor.squeeze().float().cpu().clamp(*min_max)
    t = (t - min_max[0]) / (min_max[1] - min_max[0])
    if t.dim() == 4:
        from torchvision.utils import make_grid
        arr = make_grid(t, nrow=int(math.sqrt(len(t))), normalize=False).numpy()
        arr = np.transpose(arr[[2, 1, 0], :, :], (1, 2, 0))
    elif t.dim() == 3:
        arr = np.transpose(t.numpy()[[2, 1, 0], :, :], (1, 2, 0))
    elif t.dim() == 2:
        arr = t.numpy()
    else:
        raise TypeError("Only support 4D, 3D and 2D tensor. But received with dimension: {:d}".format(t.dim()))
    if out_type == np.uint8:
        arr = (arr * 255.0).round()
    return arr.astype(out_type)
