"""Bucketed gradient all-reduce over the FLAT gradient buffer, overlapped with backward.

torch DDP keeps per-parameter hooks, 25 MiB buckets and copies gradients into bucket storage.  Here the gradients already live
in one contiguous fp32 buffer laid out in forward order, and the hand-written backward announces finished ranges
(``ParamStore.grad_ready_hook``: CTC head, decoder, encoder layers from last to first, front end).  Ranges are merged into the
contiguous ready tail of the buffer and every time the tail holds >= ``bucket_bytes`` it is all-reduced (SUM) on a side stream
while backward keeps running.  The 1/world average is folded into the fused optimizer (``grad_mult``), so no extra pass
touches the gradients.  BatchNorm running statistics are views of one small flat buffer and are broadcast from rank 0 before
each forward like DDP's ``broadcast_buffers=True`` (C2); the constant ``pe`` tables are not broadcast."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


class FlatDDP:
    def __init__(self, model: nn.Module, store, process_group=None, bucket_bytes: int = 64 << 20, broadcast_buffers: bool = True):
        self.model, self.store, self.pg = model, store, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.cuda = store.gflat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=store.device) if self.cuda else None
        self.ready: List[Tuple[int, int]] = []
        self.tail = store.numel  # everything in [tail, numel) has been handed to NCCL
        self.sync_grads = True   # False == the reference's ``no_sync`` (trainer.py:142-145)
        self.launched: List[Tuple[int, int]] = []
        store.grad_ready_hook = self._on_ready
        self.bn_flat = self._flatten_bn_buffers() if broadcast_buffers else None
        if self.world > 1:  # start from identical weights (DDP broadcasts rank 0's at construction)
            dist.broadcast(store.flat, src=0, group=self.pg)

    # -------------------------------------------------------------- buffers
    def _flatten_bn_buffers(self) -> Optional[torch.Tensor]:
        mods = [m for m in self.model.modules() if isinstance(m, nn.BatchNorm1d)]
        if not mods:
            return None
        n = sum(m.running_mean.numel() + m.running_var.numel() for m in mods)
        flat = torch.empty(n, dtype=torch.float32, device=self.store.device)
        o = 0
        for m in mods:
            for name in ("running_mean", "running_var"):
                b = getattr(m, name)
                v = flat[o:o + b.numel()]
                v.copy_(b)
                m._buffers[name] = v
                o += b.numel()
        return flat

    def broadcast_buffers(self) -> None:
        if self.world > 1 and self.bn_flat is not None:
            dist.broadcast(self.bn_flat, src=0, group=self.pg)

    # -------------------------------------------------------------- gradient buckets
    def begin_backward(self) -> None:
        self.ready.clear()
        self.launched.clear()
        self.tail = self.store.numel

    def _on_ready(self, lo: int, hi: int) -> None:
        if self.world <= 1 or not self.sync_grads:
            return
        self.ready.append((lo, hi))
        new_tail = self.tail
        progressed = True
        while progressed:  # grow the contiguous ready tail
            progressed = False
            for (a, b) in self.ready:
                if b >= new_tail > a:
                    new_tail = a
                    progressed = True
        if self.tail - new_tail >= self.bucket_elems or new_tail == 0:
            self._launch(new_tail, self.tail)
            self.tail = new_tail

    def _launch(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        buf = self.store.gflat[lo:hi]
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
        self.launched.append((lo, hi))

    def finish_backward(self) -> float:
        """Flush what is left, make the compute stream wait for NCCL; returns the gradient multiplier (1/world)."""
        if self.world > 1 and self.sync_grads:
            if self.tail > 0:
                self._launch(0, self.tail)
                self.tail = 0
            if self.cuda:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        return 1.0 / self.world if self.sync_grads else 1.0
