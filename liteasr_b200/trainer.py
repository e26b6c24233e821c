"""Training step and loop for the hot path (the reference's ``liteasr/trainer.py``), B200-first.

``TrainStep`` is one optimizer step of ``Trainer.run`` (trainer.py:140-171) restated for a flat-store model:
    zero grads -> [micro-steps: criterion(model, batch) -> backward (bucketed all-reduce overlapped)] ->
    fused global-norm clip + non-finite skip + (Noam) Adam.
Host->device copies excepted, the whole step can be captured per input shape into a CUDA graph and replayed, so the ~950
kernel launches of a 12-layer Conformer step cost one ``cudaGraphLaunch``.  Real length-bucketed batches (``utils/batchify``)
have a different (Tmax, Lmax) almost every step, and padding to a bucket is not parity-neutral here (BatchNorm statistics
include padded frames, quirk Q2), so the graph cache is a bounded LRU that only captures shapes it has SEEN BEFORE
(``graph_min_hits``), evicts by count and by pool memory, and runs every other step eagerly -- the eager step is within a few
percent of a replay at the bench shape (programmatic dependent launch hides the launch gaps).  Under DDP the capture is safe
with different shapes on different ranks: the warm-up steps of a capture issue NO collectives (their results are thrown away
anyway), so every rank issues exactly one set of NCCL calls per optimizer step whether it replays, captures or runs eagerly.
Semantics kept from the reference: losses of the micro-steps are summed un-scaled (quirk Q9); the loss is normalised by the
batch size inside the criterion; a non-finite gradient norm skips the update (decided on the device, identically on all ranks
because it is evaluated after the all-reduce); ``grad = None``-style zeroing becomes one memset of the flat buffer.
"""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import Callable, Dict, Iterable, Optional, Tuple

import torch
import torch.distributed as dist

from . import functions as F
from . import ops
from .distributed.flat_ddp import FlatDDP
from .optims import FusedAdam, FusedNoam, NoamConfig


class TrainStep:
    def __init__(self, model, criterion, optimizer=None, *, clip_grad_norm: float = 5.0, accum_grad: int = 1,
                 use_graph: bool = True, ddp: Optional[bool] = None, bucket_bytes: int = 64 << 20, device=None,
                 graph_min_hits: int = 1, max_graphs: int = 8, graph_mem_fraction: float = 0.5):
        """use_graph: replay CUDA graphs where one exists.  graph_min_hits: capture a shape at its n-th occurrence (1 = at
        once: fixed-shape training; >= 2: only shapes that repeat).  max_graphs / graph_mem_fraction: LRU bounds of the cache
        (entries; their private activation pools as a fraction of the device memory)."""
        self.model, self.criterion = model, criterion
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.store, _, _ = F.bind(model, self.device)
        self.store.enable_direct_grads()
        self.optimizer = optimizer if optimizer is not None else FusedNoam(self.store, NoamConfig())
        self.clip = clip_grad_norm
        self.accum = accum_grad
        self.use_graph = use_graph
        if ddp is None:
            ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.ddp = FlatDDP(model, self.store, bucket_bytes=bucket_bytes) if ddp else None
        self.graphs: "OrderedDict[Tuple, Tuple]" = OrderedDict()  # key -> (graph, static inputs, loss, pool bytes); LRU order
        self.hits: Dict[Tuple, int] = {}
        self.graph_min_hits, self.max_graphs = max(1, int(graph_min_hits)), max(1, int(max_graphs))
        self.graph_mem_cap = int(graph_mem_fraction * torch.cuda.get_device_properties(self.device).total_memory)
        self.stats = dict(replays=0, eager=0, captures=0, evictions=0, capture_s=0.0)
        self._warm = False
        self._pool_seen = 0  # largest private graph pool captured so far (bytes)
        self._evicted = set()  # shape keys the LRU has dropped
        self.loss_out = torch.zeros(3, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ one optimizer step, eager
    def _body(self, batches) -> torch.Tensor:
        self.store.zero_grads()
        if self.ddp is not None and not self._warm:
            self.ddp.broadcast_buffers()
        total = None
        for i, (xs, xlens, ys, ylens) in enumerate(batches):
            last = i == len(batches) - 1
            if self.ddp is not None:
                # no_sync on all but the last micro-step (trainer.py:142-145); the throw-away warm-up steps of a graph capture
                # issue no collective at all (ranks with other shapes are not capturing)
                self.ddp.sync_grads = last and not self._warm
                self.ddp.begin_backward()
            direct = getattr(self.criterion, "direct_step", None)
            loss = direct(self.model, xs, xlens, ys, ylens) if direct is not None else None
            if loss is None:  # generic criterion / model: through autograd
                loss = self.criterion(self.model, xs, xlens, ys, ylens)
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        mult = self.ddp.finish_backward() if self.ddp is not None else 1.0
        self.optimizer.step(self.clip, grad_mult=mult)
        return total

    def step_eager(self, *batch) -> torch.Tensor:
        return self._body(self._split(batch))

    def _split(self, batch):
        if self.accum == 1:
            return [tuple(batch)]
        xs, xlens, ys, ylens = batch
        return [tuple(t.chunk(self.accum)[i] for t in (xs, xlens, ys, ylens)) for i in range(self.accum)]

    # ------------------------------------------------------------------ CUDA-graph replay per shape bucket
    def __call__(self, xs, xlens, ys, ylens) -> torch.Tensor:
        """Device tensors in, device loss out (sum over micro-steps).  Copies the batch into the graph's static buffers."""
        if not self.use_graph:
            self.stats["eager"] += 1
            return self.step_eager(xs, xlens, ys, ylens)
        key = (tuple(xs.shape), tuple(ys.shape), self.model.training)
        entry = self.graphs.get(key)
        if entry is None:
            n = self.hits[key] = self.hits.get(key, 0) + 1
            if len(self.hits) > 4096:  # the hit counters themselves stay bounded
                self.hits = {key: n}
            if n < self.graph_min_hits:
                self.stats["eager"] += 1
                return self.step_eager(xs, xlens, ys, ylens)
            # a shape the cache has already had to drop, and no room without dropping another: the working set of shapes does not
            # fit (a cyclic epoch over more shapes than pools would re-capture every step, 1.4 - 3 s each) -> this shape stays eager
            entry = None if (key in self._evicted and self._full()) else self._capture(key, xs, xlens, ys, ylens)
            if entry is None:  # no device memory for another graph pool: this shape stays eager
                self.stats["eager"] += 1
                return self.step_eager(xs, xlens, ys, ylens)
        else:
            self.graphs.move_to_end(key)
        graph, static, loss, _ = entry
        for dst, src in zip(static, (xs, xlens, ys, ylens)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        self.stats["replays"] += 1
        return loss

    def static_inputs(self, xs, xlens, ys, ylens):
        """The static device buffers for this shape (capture on first use) -- fill them in place to skip the D2D copy."""
        key = (tuple(xs.shape), tuple(ys.shape), self.model.training)
        if key not in self.graphs and self._capture(key, xs, xlens, ys, ylens) is None:
            raise RuntimeError("not enough free device memory to capture the training step for this shape")
        return self.graphs[key][1]

    def _full(self) -> bool:
        """Would another capture have to evict a resident graph?"""
        est = max(max((e[3] for e in self.graphs.values()), default=0), self._pool_seen)
        return bool(self.graphs) and (len(self.graphs) >= self.max_graphs or sum(e[3] for e in self.graphs.values()) + est > self.graph_mem_cap)

    def _evict_for(self, need_bytes: int) -> bool:
        """LRU eviction: keep at most max_graphs entries and at most graph_mem_cap bytes of private graph pools -- and, whatever the
        bookkeeping says, enough FREE device memory for another pool of the expected size (other graphs, models or processes share
        the GPU; a capture that runs out of memory cannot be resumed).  False: no room even with an empty cache -> run eagerly."""
        def used():
            return sum(e[3] for e in self.graphs.values())

        def pop():
            k, (g, static, loss, nbytes) = self.graphs.popitem(last=False)
            del g, static, loss
            self.stats["evictions"] += 1
            if len(self._evicted) > 4096:
                self._evicted.clear()
            self._evicted.add(k)
        while self.graphs and (len(self.graphs) >= self.max_graphs or used() + need_bytes > self.graph_mem_cap):
            pop()
        torch.cuda.empty_cache()
        want = int(1.5 * need_bytes)  # the pool plus the eager warm-up steps' transient activations
        while need_bytes > 0 and torch.cuda.mem_get_info(self.device)[0] < want:
            if not self.graphs:
                return False
            pop()
            torch.cuda.empty_cache()
        return True

    def _capture(self, key, xs, xlens, ys, ylens):
        t0 = time.perf_counter()
        est = max(max((e[3] for e in self.graphs.values()), default=0), self._pool_seen)  # a new pool is about as large as the largest so far
        if not self._evict_for(est):
            return None
        static = tuple(t.clone() for t in (xs, xlens, ys, ylens))
        # snapshot state mutated by the warm-up steps so capture does not change training semantics
        snap = (self.store.flat.clone(), self.optimizer.exp_avg.clone(), self.optimizer.exp_avg_sq.clone(),
                self.optimizer.state.clone(), {k: v.clone() for k, v in self.model.state_dict().items() if "running" in k or "num_batches" in k})
        snap_rng = self.store.rng.state.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        self._warm = True  # throw-away steps: no NCCL calls (ranks see different shapes and capture at different steps)
        try:
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._body(self._split(static))
        finally:
            self._warm = False
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # torch.cuda.graph() empties the allocator cache on entry: do it BEFORE the baseline reading, or the warm-up steps' cached
        # blocks (about one pool's worth) leave the default pool while the private pool grows and the difference reads as ~0
        torch.cuda.empty_cache()
        mem0 = torch.cuda.memory_reserved(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self._body(self._split(static))
        torch.cuda.synchronize()
        pool_bytes = max(0, torch.cuda.memory_reserved(self.device) - mem0)
        self._pool_seen = max(self._pool_seen, pool_bytes)
        with torch.no_grad():
            self.store.flat.copy_(snap[0])
            self.optimizer.exp_avg.copy_(snap[1])
            self.optimizer.exp_avg_sq.copy_(snap[2])
            self.optimizer.state.copy_(snap[3])
            sd = self.model.state_dict()
            for k, v in snap[4].items():
                sd[k].copy_(v)
            self.store.rng.state.copy_(snap_rng)
        self.graphs[key] = (graph, static, loss, pool_bytes)
        self.stats["captures"] += 1
        self.stats["capture_s"] += time.perf_counter() - t0
        return self.graphs[key]


class Trigger:
    """utils/trigger.py:6-34: fires every ``interval`` iterations."""

    def __init__(self, interval: int):
        self.interval = max(1, int(interval))

    def __call__(self, it: int) -> bool:
        return it % self.interval == 0


class Prefetcher:
    """Pinned-host -> device double buffering on a copy stream: the batch of step i+1 crosses PCIe while step i computes
    (replaces the synchronous ``to_device`` of trainer.py:140).  ``put`` starts the copy of a host batch into the
    idle staging slot, ``get`` makes the current stream wait for the oldest pending copy and returns its device tensors; the
    consumer (``TrainStep.__call__``) copies them device-to-device into the CUDA graph's static inputs."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.head = self.tail = 0  # put writes slot head % 2, get reads slot tail % 2

    def put(self, host_batch) -> None:
        if self.head - self.tail >= 2:
            raise RuntimeError("Prefetcher: both staging slots are pending (call get first)")
        k = self.head % 2
        # the slot's previous contents were consumed by work already enqueued on the current stream
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            slot = self.slots[k]
            if slot is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot, host_batch)):
                slot = self.slots[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_batch)
            for d, h in zip(slot, host_batch):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.stream)
        self.head += 1

    def get(self):
        if self.tail == self.head:
            raise RuntimeError("Prefetcher: nothing pending")
        k = self.tail % 2
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.tail += 1
        return self.slots[k]


class Trainer:
    """Loop mirror of trainer.py:28-227 over an iterable of collated batches ``(xs, xlens, ys, ylens)`` (CPU or GPU
    tensors, the collator contract of dataset/asr_dataset.py:115-126).  Data loading itself is out of scope."""

    def __init__(self, model, criterion, optimizer=None, *, clip_grad_norm=5.0, accum_grad=1, report_interval=100,
                 use_graph=True, device=None, log: Callable[[str], None] = print, max_iter: int = 0, max_epoch: int = 0,
                 graph_min_hits: int = 50, max_graphs: int = 8):
        # a capture costs ~1.4 s (two warm-up steps + instantiation of a ~950-node graph) and a replay is only ~2 % faster than the eager
        # step, so a shape must come back about 50 times before its graph pays off: real length-bucketed batches mostly run eagerly
        self.step_fn = TrainStep(model, criterion, optimizer, clip_grad_norm=clip_grad_norm, accum_grad=accum_grad,
                                 use_graph=use_graph, device=device, graph_min_hits=graph_min_hits, max_graphs=max_graphs)
        self.model, self.criterion = model, criterion
        self.device = self.step_fn.device
        self.report = Trigger(report_interval)
        self.iter = 0
        self.epoch = 0
        self.max_iter, self.max_epoch = max_iter, max_epoch
        self.log = log
        self.loss = torch.zeros((), device=self.device)  # loss of the CURRENT step (trainer.py:148-149,171: reset after every step)

    def run(self, batches: Iterable, max_iters: Optional[int] = None) -> None:
        self.model.train()
        pf = Prefetcher(self.device)
        it = iter(batches)

        def stage():  # start the next batch's host-to-device copy (pinned host tensors make it asynchronous)
            nxt = next(it, None)
            if nxt is not None:
                pf.put(tuple(nxt))
            return nxt is not None

        more = stage()
        while more:
            batch = pf.get()
            more = stage()  # trainer.py:140 moved one step ahead: this copy overlaps the step below
            loss = self.step_fn(*batch)
            self.loss = loss / self.step_fn.accum
            self.iter += 1
            if self.report(self.iter):
                self.report_loss()
            if max_iters is not None and self.iter >= max_iters:
                break

    def _progress(self) -> str:
        return "{} / {} iters, {} / {} epochs".format(self.iter, self.max_iter, self.epoch, self.max_epoch)

    def report_loss(self) -> None:
        """trainer.py:174-186: the loss of the current step, reduced to rank 0 and divided by the world size; same log line."""
        loss = self.loss.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.reduce(loss, dst=0)
            loss /= dist.get_world_size()
        opt = self.step_fn.optimizer
        if not dist.is_initialized() or dist.get_rank() == 0:
            self.log("{} - current loss: {:.2f}".format(self._progress(), loss.item()) +
                     f" (lr {opt.rate():.3e}, grad_norm {opt.last_grad_norm():.3f}, updates {opt.num_updates()})")

    @torch.no_grad()
    def valid(self, batches: Iterable) -> float:
        """trainer.py:188-209: criterion under eval() + no_grad (BatchNorm running statistics), every batch loss reduced to rank 0
        and divided by the world size, mean over batches, and the ``... - valid loss: {:.2f}`` line that
        ``utils/checkpoint.load_ckpt(avg_policy=<log>)`` parses (utils/checkpoint.py:55-60)."""
        self.model.eval()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        losses = []
        for batch in batches:
            batch = tuple(t.to(self.device) for t in batch)
            loss = self.criterion(self.model, *batch).detach().clone()
            if world > 1:
                dist.reduce(loss, dst=0)
                loss /= world
            losses.append(loss)
        reduced = float(torch.stack(losses).mean()) if losses else 0.0
        if not dist.is_initialized() or dist.get_rank() == 0:
            self.log("{} - valid loss: {:.2f}".format(self._progress(), reduced))
        self.model.train()
        return reduced

    def save_model(self, path: str) -> None:
        if not dist.is_initialized() or dist.get_rank() == 0:
            torch.save(self.model.state_dict(), path)  # models/__init__.py:31-32 (model-only checkpoint)
