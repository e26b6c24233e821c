"""Length-bucketed mini-batch formation and the batch collator (host side of the input pipeline).

Mirrors the reference's interface and results:

  * ``SeqBatch`` / ``FrameBatch`` (``utils/batchify.py:76-159`` of the reference) -- policy objects with
    ``batchify(indices, samples)``, ``__getitem__`` and ``__len__``; ``samples[i]`` only needs ``xlen`` / ``ylen``.
  * ``length_sorted_indices`` -- the stable descending-``xlen`` order the dataset feeds them
    (``dataset/asr_dataset.py:107-110``).
  * ``collate`` -- ``dataset/asr_dataset.py:115-126``: zero-padded features, ``-1``-padded labels, int64 lengths; written
    straight into (optionally pinned) host buffers so that ``trainer.Prefetcher`` can start the H2D copy without a staging copy.

The partitions are computed as closed-form scans over the length arrays instead of the reference's push / pop / refresh
state machine, but are identical to it element for element, including its two quirks (kept on purpose, see
``tests/test_batchify_cpu.py`` and the golden fixture generated from the unmodified reference classes):

  Q-a  ``seq``: the batch size is fixed by the FIRST (longest) sample of a batch:
       ``max(min_batch_size, int(batch_size / (1 + max(int(xlen / max_len_in), int(ylen / max_len_out)))))``; a size of 0
       never "fills", so the batch swallows every remaining sample.
  Q-b  ``frame``: a sample that alone exceeds a frame budget closes the (possibly EMPTY) current batch first, so an empty
       batch can be emitted in front of it.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def length_sorted_indices(xlens: Sequence[int]) -> List[int]:
    """Indices by descending input length, ties in original order (``sorted(..., reverse=True)`` is stable)."""
    return [i for i, _ in sorted(enumerate(xlens), key=lambda t: t[1], reverse=True)]


def seq_batches(order: Sequence[int], xlens: Sequence[int], ylens: Sequence[int], batch_size: int, min_batch_size: int,
                max_len_in: int, max_len_out: int) -> List[List[int]]:
    """Sequence-count batching (``batch_count: seq``)."""
    if batch_size is None or batch_size < 1:
        raise ValueError("seq batching needs batch_size >= 1")
    out: List[List[int]] = []
    pos, n = 0, len(order)
    while pos < n:
        head = order[pos]
        shrink = max(int(xlens[head] / max_len_in), int(ylens[head] / max_len_out))
        size = max(min_batch_size, int(batch_size / (1 + shrink)))
        end = n if size < 1 else min(n, pos + size)  # Q-a
        out.append(list(order[pos:end]))
        pos = end
    return out


def frame_batches(order: Sequence[int], xlens: Sequence[int], ylens: Sequence[int], max_frame_in: Optional[int],
                  max_frame_out: Optional[int], max_frame_inout: Optional[int]) -> List[List[int]]:
    """Frame-budget batching (``batch_count: frame``): a batch is closed when adding the next sample would push
    (longest input | longest output | their sum) x (count + 1) over the corresponding budget (0 / None = no budget)."""
    out: List[List[int]] = []
    cur: List[int] = []
    top_in = top_out = 0
    for idx in order:
        xi, yo = xlens[idx], ylens[idx]
        nin, nout, cnt = max(top_in, xi), max(top_out, yo), len(cur) + 1
        over = bool((max_frame_in and nin * cnt > max_frame_in) or (max_frame_out and nout * cnt > max_frame_out)
                    or (max_frame_inout and (nin + nout) * cnt > max_frame_inout))
        if over:
            out.append(cur)  # Q-b: may be empty
            cur, top_in, top_out = [], 0, 0
        cur.append(idx)
        top_in, top_out = max(top_in, xi), max(top_out, yo)
    if cur:
        out.append(cur)
    return out


class _Policy:
    """Common surface of the reference's ``BatchifyPolicy`` (``utils/batchify.py:11-72``)."""

    def __init__(self, dataset_cfg):
        self.dataset_cfg = dataset_cfg
        self.data: List[List[int]] = []

    def _partition(self, order, xlens, ylens) -> List[List[int]]:
        raise NotImplementedError

    def batchify(self, indices, samples) -> None:
        assert len(indices) == len(samples), f"{len(samples)}"
        xlens = [s.xlen for s in samples]
        ylens = [s.ylen for s in samples]
        self.data.extend(self._partition(list(indices), xlens, ylens))

    def __getitem__(self, index):
        return self.data[index]

    def __len__(self):
        return len(self.data)


class SeqBatch(_Policy):
    def _partition(self, order, xlens, ylens):
        c = self.dataset_cfg
        return seq_batches(order, xlens, ylens, c.batch_size, c.min_batch_size, c.max_len_in, c.max_len_out)


class FrameBatch(_Policy):
    def _partition(self, order, xlens, ylens):
        c = self.dataset_cfg
        return frame_batches(order, xlens, ylens, c.max_frame_in, c.max_frame_out, c.max_frame_inout)


def make_policy(dataset_cfg) -> _Policy:
    """``dataset/asr_dataset.py:96-105``: ``batch_count`` selects the policy; anything else is a ValueError."""
    if dataset_cfg.batch_count == "seq":
        return SeqBatch(dataset_cfg)
    if dataset_cfg.batch_count == "frame":
        return FrameBatch(dataset_cfg)
    raise ValueError(f"unsupport strategy {dataset_cfg.batch_count}")


def collate(xs: Sequence[torch.Tensor], ys: Sequence[torch.Tensor], pin_memory: bool = False):
    """``dataset/asr_dataset.py:115-126``: (B,Tmax,F) features padded with 0, (B,Lmax) labels padded with -1, int64 lengths."""
    if len(xs) == 0 or len(xs) != len(ys):
        raise ValueError("collate needs a non-empty batch with one label sequence per utterance")
    B = len(xs)
    xlens = torch.tensor([int(x.shape[0]) for x in xs], dtype=torch.long)
    ylens = torch.tensor([int(y.shape[0]) for y in ys], dtype=torch.long)
    tmax, lmax = int(xlens.max()), int(ylens.max())
    pin = pin_memory and torch.cuda.is_available()
    padded_xs = torch.zeros((B, tmax) + tuple(xs[0].shape[1:]), dtype=xs[0].dtype, pin_memory=pin)
    padded_ys = torch.full((B, lmax), -1, dtype=ys[0].dtype, pin_memory=pin)
    for b in range(B):
        padded_xs[b, : xlens[b]] = xs[b]
        padded_ys[b, : ylens[b]] = ys[b]
    return padded_xs, xlens, padded_ys, ylens
