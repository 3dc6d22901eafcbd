"""Numeric contract of the integer GroupNorm sums of csrc/gn_fuse.cuh, restated in numpy (no GPU): warp-tile partials
are rounded to 2^-20 (sums) / 2^-16 (sums of squares) units and added as int64 -- exact, hence order independent --
and the (mean, rstd) the last CTA derives from them equals the fp64 statistics of the data far inside fp32 rounding."""
import numpy as np

U1, U2 = 2.0 ** 20, 2.0 ** 16


def _partials(x, tile=256):
    """fp32 (sum, sum of squares) per warp tile of `tile` values, as the epilogue warps form them"""
    t = x.astype(np.float32).reshape(-1, tile)
    return t.sum(axis=1, dtype=np.float32), (t * t).sum(axis=1, dtype=np.float32)


def _fixed(s1, s2):
    return np.rint(s1.astype(np.float64) * U1).astype(np.int64), np.rint(s2.astype(np.float64) * U2).astype(np.int64)


def test_integer_sums_are_order_independent_and_accurate():
    rng = np.random.default_rng(0)
    for scale, offset in ((1.0, 0.0), (0.03, 2.5), (40.0, -7.0), (1e-3, 0.0)):
        x = (rng.standard_normal(256 * 2048) * scale + offset).astype(np.float32)     # one group of a 256^2 image, 8 channels
        q1, q2 = _fixed(*_partials(x))
        perm = rng.permutation(q1.size)
        assert q1.sum() == q1[perm].sum() and q2.sum() == q2[perm].sum()              # any arrival order, any slot split
        n = x.size
        mean = q1.sum() / U1 / n
        var = max(q2.sum() / U2 / n - mean * mean, 0.0)
        ref_mean, ref_var = x.astype(np.float64).mean(), x.astype(np.float64).var()
        # the fp32 partials themselves carry ~1e-7 relative error; the quantisation must not add to it visibly
        assert abs(mean - ref_mean) <= 1e-6 * max(1.0, abs(ref_mean)) + 1e-7 * scale
        # sums of squares are quantised to 2^-16 per tile: for a near-constant group (var 1e-6, far below eps) that is the
        # visible term -- 5e-10 absolute on the variance, 3e-5 relative on rstd
        assert abs(var - ref_var) <= 1e-4 * max(ref_var, 1e-5) + 2e-6 * (abs(ref_mean) * scale + ref_mean ** 2), (scale, offset, var, ref_var)
        rstd, ref_rstd = 1.0 / np.sqrt(var + 1e-5), 1.0 / np.sqrt(ref_var + 1e-5)
        assert abs(rstd - ref_rstd) <= 1e-4 * ref_rstd


def test_overflow_margin():
    """int64 totals of one (image, group): the sum of squares is the first to overflow, at a group mean square of
    2^63 / 2^16 / count -- 2.7e8 for the largest group of the network (256^2 pixels x 8 channels), i.e. activations with
    an rms above 16 000; bf16 feature maps of a normalised UNet stay orders of magnitude below."""
    count = 256 * 256 * 8
    limit_mean_square = (2.0 ** 63) / U2 / count
    assert limit_mean_square > 2.5e8
    # a single warp-tile partial converts without overflow for |x| up to 1e6 (fp32 partial 256e12 -> 1.7e19 > 2^63? no: guard)
    big = np.float64(256) * 1e4 ** 2                       # tile of 256 values of magnitude 1e4
    assert big * U2 < 2.0 ** 63
