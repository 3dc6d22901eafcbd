"""TEST INFRASTRUCTURE (never imported by the product).  tests/golden/metrics.npz = outputs of the reference's own
`utils/img_utils.py` (`tensor2img`, `calculate_psnr`, `ssim`, `calculate_ssim`) on seeded images.  Dev container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_metrics.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "metrics.npz")
REF = os.environ.get("IDIFF_REFERENCE", "/root/reference")


def cases():
    g = torch.Generator().manual_seed(2024)
    clean = torch.rand(1, 1, 64, 48, generator=g)
    smooth = torch.nn.functional.avg_pool2d(clean, 5, 1, 2)
    noisy = (smooth + 0.1 * torch.randn(smooth.shape, generator=g))
    rgb = torch.rand(3, 40, 40, generator=g)
    rgb2 = (rgb + 0.05 * torch.randn(rgb.shape, generator=g))
    batch = torch.rand(4, 3, 8, 8, generator=g) * 1.4 - 0.2
    return dict(smooth=smooth, noisy=noisy, rgb=rgb, rgb2=rgb2, batch=batch)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from utils import img_utils as R   # the reference's own functions
    sys.path.pop(0)
    c = cases()
    a8, b8 = R.tensor2img(c["smooth"]), R.tensor2img(c["noisy"])
    af, bf = R.tensor2img(c["smooth"], out_type=np.float32), R.tensor2img(c["noisy"], out_type=np.float32, min_max=(-0.5, 1.5))
    c8, d8 = R.tensor2img(c["rgb"]), R.tensor2img(c["rgb2"])
    grid = R.tensor2img(c["batch"])
    unit_a = c["smooth"].squeeze().numpy().astype(np.float64)
    unit_b = c["noisy"].squeeze().clamp(0, 1).numpy().astype(np.float64)
    out = dict(
        img_a8=a8, img_b8=b8, img_af=af, img_bf=bf, img_c8=c8, img_d8=d8, img_grid=grid,
        psnr_gray=R.calculate_psnr(a8, b8), psnr_rgb=R.calculate_psnr(c8, d8), psnr_same=R.calculate_psnr(a8, a8),
        ssim_gray=R.calculate_ssim(a8, b8), ssim_rgb=R.calculate_ssim(c8, d8),
        ssim_hw1=R.calculate_ssim(a8[:, :, None], b8[:, :, None]),
        # data_range 1 (testUM.py:158-161): the reference formula on [0,255]-scaled inputs is the same index
        ssim_unit=R.ssim(unit_a * 255.0, unit_b * 255.0), unit_a=unit_a, unit_b=unit_b)
    np.savez_compressed(OUT, **{k: np.asarray(v) for k, v in out.items()}, **{"in_" + k: v.numpy() for k, v in c.items()})
    print({k: float(v) for k, v in out.items() if np.ndim(v) == 0})


if __name__ == "__main__":
    main()
