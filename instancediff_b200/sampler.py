"""testUM.py-style driver helpers: SDE factory and batch-sharded multi-GPU sampling.

The reference's ``create_sde`` / ``driftSDE`` modules are not in the snapshot (``testUM.py:17,27``
import them); this factory keeps the call shape ``create_sde(nets, sde_opt)``
(``testUM.py:91``) and returns the drop-in ``IRSDE`` wired to ``nets['noise_net']``
(``models/drift_noise_model.py:657-668`` defines the dict keys).

Multi-GPU: every image's chain is independent, so the batch is cut into contiguous shards, one
process per GPU, NO collective inside the loop.  Noise is drawn from a Philox
stream indexed by the GLOBAL element index, so the result of a sample does not depend on how the
batch was sharded.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .sde import IRSDE


def create_sde(nets: Dict[str, object], sde_opt: dict, device=None) -> IRSDE:
    opt = dict(sde_opt)
    sde = IRSDE(max_sigma=opt.get("max_sigma", 0.4), T=opt.get("T", 100), sample_T=opt.get("sample_T", -1),
                schedule=opt.get("schedule", "cosine"), eps=opt.get("eps", 0.01), device=device)
    net = nets["noise_net"] if isinstance(nets, dict) else nets
    sde.set_model(net)
    return sde


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``total`` samples for ``rank``; the first ``total % world`` ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_sharded(sde: IRSDE, mu_full: torch.Tensor, ctx_full: torch.Tensor, rank: int, world_size: int,
                   seed: int = 1, T: int = -1, x_T: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """Runs ``reverse_sde`` on this rank's shard of the batch and returns (x0_shard, (lo, hi)).

    ``mu_full`` [B,1,H,W] / ``ctx_full`` [B,1,D] may live on the host; only the shard is moved.
    """
    B = mu_full.shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    per_sample = mu_full[0].numel()
    dev = sde.device if sde.device is not None else torch.device("cuda")
    mu = mu_full[lo:hi].to(dev, torch.float32).contiguous()
    ctx = ctx_full[lo:hi].to(dev, torch.float32).contiguous()
    sde.noise_source = "philox"
    sde.philox_seed = int(seed)
    sde.philox_offset = lo * per_sample           # global element index -> sharding-invariant noise
    sde.set_mu(mu)
    xT = sde.noise_state(mu) if x_T is None else x_T[lo:hi].to(dev, torch.float32).contiguous()
    x0 = sde.reverse_sde(xT, T=T, image_context=ctx)
    return x0, (lo, hi)


def gather_shards(x_shard: torch.Tensor, total: int, rank: int, world_size: int):
    """Host-side gather of the per-rank results to rank 0 AFTER the loop (the only communication of the
    sampling path; works on any backend because it moves CPU tensors).  Returns [total, ...] on rank 0,
    None elsewhere."""
    import torch.distributed as dist
    piece = x_shard.detach().cpu().contiguous()
    if world_size == 1:
        return piece
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object((shard_bounds(total, world_size, rank), piece), gathered, dst=0)
    if rank != 0:
        return None
    out = torch.empty((total,) + tuple(piece.shape[1:]), dtype=piece.dtype)
    for (lo, hi), t in gathered:
        out[lo:hi] = t
    return out
