"""Numerical check (CPU, torch) of the row-pair formulation of the Cout = 64 3x3 convolutions sketched in DESIGN.md
section 4: M = 128 = 64 output channels x 2 vertically adjacent output rows, N = pixels of the row pair, and the four
row views of the stacked weights taken as 128-row WINDOWS of one buffer  [0 ; W(dy=2) ; W(dy=1) ; W(dy=0) ; 0]  per
filter column.  Not part of the product; it pins the index conventions a future kernel has to follow.

    python tools/design_rowpair_check.py
"""
import torch


def rowpair_conv(x, w):
    """x: [Cin, H, W] (H even), w: [64, Cin, 3, 3] -> [64, H, W] through 12 'MMAs' per row pair."""
    co, cin, H, W = w.shape[0], x.shape[0], x.shape[1], x.shape[2]
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))                       # patch with halo: row y-1+vy, col x-1+vx
    zero = torch.zeros(co, cin)
    out = torch.zeros(co, H, W)
    # per filter column dx: blocks of 64 rows, contiguous: [0 ; W2 ; W1 ; W0 ; 0]
    bufs = [torch.cat([zero, w[:, :, 2, dx], w[:, :, 1, dx], w[:, :, 0, dx], zero]) for dx in range(3)]   # [320, Cin]
    for y in range(0, H, 2):
        acc = torch.zeros(2 * co, W)                                    # accumulator: lanes = (row r, channel), columns = pixels
        for vy in range(4):                                             # input row y - 1 + vy
            start = (3 - vy) * co                                       # window [start, start + 128): rows r=0 -> W(dy=vy), r=1 -> W(dy=vy-1)
            for vx in range(3):
                a = bufs[vx][start:start + 2 * co]                      # A operand: 128 x Cin, a shifted view of ONE buffer
                b = xp[:, y + vy, vx:vx + W]                            # B operand: Cin x W pixels, a shifted view of the patch
                acc += a @ b
        out[:, y] = acc[:co]
        out[:, y + 1] = acc[co:]
    return out


if __name__ == "__main__":
    g = torch.Generator().manual_seed(0)
    for cin, H, W in ((64, 8, 16), (128, 6, 10)):
        x = torch.randn(cin, H, W, generator=g, dtype=torch.float64)
        w = torch.randn(64, cin, 3, 3, generator=g, dtype=torch.float64) / 24
        ref = torch.nn.functional.conv2d(x[None], w, padding=1)[0]
        got = rowpair_conv(x.float(), w.float()).double()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        print(f"Cin={cin} {H}x{W}: max rel err {err:.2e}")
        assert err < 1e-5
    print("12 full-width views per row pair reproduce the 3x3 convolution; window start = (3 - vy) * 64 rows")
