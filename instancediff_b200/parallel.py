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
        self.flat: Optional[torch.Tensor] = None          # the gradients of this bucket LIVE here (p.grad are views)
        self.wire: Optional[torch.Tensor] = None          # communication copy when comm_dtype differs
        self.pending = len(params)
        self.handle = None


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 25.0, process_group=None,
                 comm_dtype: Optional[torch.dtype] = None):
        """``comm_dtype``: dtype of the buffers on the wire (``torch.bfloat16`` halves the bytes, 46 MB instead of 91 MB
        per network; the gradients themselves stay in the parameter dtype).

        Every bucket owns ONE flat buffer and the ``.grad`` of its parameters are views into it (what DDP calls
        ``gradient_as_bucket_view``): backward accumulates straight into the communication buffer, the all-reduce runs
        in place and nothing is copied per parameter -- with 343 parameter tensors the copy-in / copy-out variant spent
        8 ms per step in ~1000 tiny kernels for a transfer that takes well under a millisecond on NVLink."""
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

    def _materialize(self, b: _Bucket) -> None:
        """(Re)build the flat buffer and point every ``.grad`` of the bucket into it, keeping existing gradients."""
        ref = b.params[0]
        intact = b.flat is not None and b.flat.device == ref.device and b.flat.dtype == ref.dtype
        if intact:
            off = 0
            for p in b.params:
                n = p.numel()
                if p.grad is None or p.grad.data_ptr() != b.flat.data_ptr() + off * b.flat.element_size():
                    intact = False
                    break
                off += n
        if intact:
            return
        flat = torch.zeros(b.numel, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in b.params:
            n = p.numel()
            view = flat[off:off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += n
        b.flat = flat

    # ---- automatic mode: fire buckets from inside backward --------------------------------------------------
    def attach(self) -> "GradientAllReducer":
        for b in self.buckets:
            self._materialize(b)
            for p in b.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def detach(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def zero_grad(self) -> None:
        """Start of a step (replaces ``optimizer.zero_grad``): one fill per bucket, the views stay in place."""
        for b in self.buckets:
            self._materialize(b)
            b.flat.zero_()
            b.pending = len(b.params)

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
        self._materialize(b)                                 # someone may have replaced a .grad (set_to_none, clipping)
        buf = b.flat
        if self.comm_dtype is not None and self.comm_dtype != b.flat.dtype:
            b.wire = b.flat.to(self.comm_dtype)
            buf = b.wire
        b.handle = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self) -> None:
        """Join all transfers; afterwards ``.grad`` holds the AVERAGED gradients (what DDP leaves there)."""
        self.reduce()
        world = dist.get_world_size(self.group)
        for b in self.buckets:
            b.handle.wait()
            if b.wire is not None:
                b.flat.copy_(b.wire)
                b.wire = None
            b.flat.div_(world)
            b.handle = None
            b.pending = len(b.params)

    @property
    def bytes_per_step(self) -> int:
        return sum(b.numel * (torch.finfo(self.comm_dtype).bits // 8 if self.comm_dtype else b.params[0].element_size())
                   for b in self.buckets)
