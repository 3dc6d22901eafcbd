"""Gradient averaging for the training step (BASELINE config 5): the one collective of this code base.

The sampling loop shards by batch and issues NO collective.  The reference trains both networks
under ``DistributedDataParallel`` (``models/drift_noise_model.py:145-146``), i.e. every step ends with a sum
all-reduce of the gradients divided by the world size (~91 MB fp32 per network).  ``GradientAllReducer`` reproduces
that effect without wrapping the module: parameters are grouped into size-bounded buckets in REVERSE registration order
(the order backward produces gradients), a bucket's flat buffer is all-reduced asynchronously as soon as all of its
gradients exist (``register_post_accumulate_grad_hook``), so on NCCL the transfers over NVLink/NVSwitch overlap the rest
of backward on the communicator's own stream; ``wait()`` joins the handles, divides by the world size and scatters the
averages back into ``.grad``.  Works on any ``torch.distributed`` backend (gloo in the CPU tests, NCCL on the GPUs).

Plumbing only: the hand-written backward kernels of SURVEY.md section 8f rank 3 are not built yet.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat: Optional[torch.Tensor] = None
        self.pending = len(params)
        self.handle = None


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0, process_group=None,
                 comm_dtype: Optional[torch.dtype] = None):
        """``comm_dtype``: dtype of the flat communication buffers (``torch.bfloat16`` halves the bytes on the wire,
        46 MB instead of 91 MB per network; the division and the write-back stay in the gradient's dtype)."""
        self.group = process_group
        self.comm_dtype = comm_dtype
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("GradientAllReducer: no trainable parameters")
        cap = max(1, int(bucket_mb * 1024 * 1024))
        self.buckets: List[_Bucket] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(plist):                            # backward reaches the last layers first
            nbytes = p.numel() * (torch.finfo(comm_dtype).bits // 8 if comm_dtype else p.element_size())
            if cur and (size + nbytes > cap or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        self.buckets.append(_Bucket(cur))
        self._owner = {id(p): b for b in self.buckets for p in b.params}
        self._hooks = []

    # ---- automatic mode: fire buckets from inside backward --------------------------------------------------
    def attach(self) -> "GradientAllReducer":
        for b in self.buckets:
            for p in b.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def detach(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self._owner[id(p)]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    # ---- manual mode -------------------------------------------------------------------------------------------
    def reduce(self) -> "GradientAllReducer":
        """Launch every bucket that has not been launched yet (gradients that do not exist count as zeros)."""
        for b in self.buckets:
            if b.handle is None:
                self._launch(b)
        return self

    def _launch(self, b: _Bucket) -> None:
        ref = b.params[0]
        dtype = self.comm_dtype or ref.dtype
        if b.flat is None or b.flat.dtype != dtype or b.flat.device != ref.device:
            b.flat = torch.empty(b.numel, dtype=dtype, device=ref.device)
        off = 0
        for p in b.params:
            n = p.numel()
            if p.grad is None:
                b.flat[off:off + n].zero_()
            else:
                b.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        b.handle = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self) -> None:
        """Join all transfers and write the AVERAGED gradients back (what DDP leaves in ``.grad``)."""
        self.reduce()
        world = dist.get_world_size(self.group)
        for b in self.buckets:
            b.handle.wait()
            off = 0
            for p in b.params:
                n = p.numel()
                avg = (b.flat[off:off + n].to(p.dtype) / world).view_as(p)
                if p.grad is None:
                    p.grad = avg.clone()
                else:
                    p.grad.copy_(avg)
                off += n
            b.handle = None
            b.pending = len(b.params)

    @property
    def bytes_per_step(self) -> int:
        return sum(b.numel * (torch.finfo(self.comm_dtype).bits // 8 if self.comm_dtype else b.params[0].element_size())
                   for b in self.buckets)
