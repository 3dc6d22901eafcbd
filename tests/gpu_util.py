"""Helpers for the GPU parity tests (fp32 torch references of the individual kernels)."""
import torch
import torch.nn.functional as F


def no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def bf16r(t):
    return t.to(torch.bfloat16).float()


def rand_act(B, H, W, C, gen, scale=1.0, dev="cuda"):
    return (torch.randn(B, H, W, C, generator=gen) * scale).to(dev).to(torch.bfloat16)


def conv_reference(src0, src1, w, bias, k, stride=1, up=0, a_scale=None, a_shift=None, a_silu=False):
    """fp32 NHWC result of the engine's contract, given bf16 sources and fp32 weights [N,Cin,k,k]."""
    x = src0.float() if src1 is None else torch.cat([src0.float(), src1.float()], dim=-1)
    if a_scale is not None:
        x = x * a_scale[:, None, None, :] + a_shift[:, None, None, :]
        if a_silu:
            x = F.silu(x)
        x = bf16r(x)
    x = x.permute(0, 3, 1, 2)
    if up:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, bf16r(w), bias, stride=stride, padding=0 if k == 1 else 1)
    return y.permute(0, 2, 3, 1).contiguous()


def describe(got, ref, name=""):
    """Human-readable mismatch summary (printed on failure to localise layout bugs from one GPU run)."""
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    msg = [f"{name}: max_abs_err={err.max().item():.4g} ref_max={scale:.4g} rel={err.max().item() / scale:.4g} "
           f"mean_abs_err={err.mean().item():.4g} got_max={got.abs().max().item():.4g} "
           f"nan={int(torch.isnan(got).sum())}"]
    if err.dim() == 4:
        bad = err > 0.02 * scale
        msg.append(f"  bad fraction={bad.float().mean().item():.4f}")
        msg.append("  bad by channel%16: " + str([round(v, 3) for v in bad.float().mean(dim=(0, 1, 2)).reshape(-1, 16).mean(0).tolist()]))
        msg.append("  bad by x%8: " + str([round(bad[:, :, i::8].float().mean().item(), 3) for i in range(min(8, bad.shape[2]))]))
        msg.append("  bad by y%16: " + str([round(bad[:, i::16].float().mean().item(), 3) for i in range(min(16, bad.shape[1]))]))
    return "\n".join(msg)


def rel_err(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()
