"""Flat parameter store: one contiguous fp32 buffer (+ fp32 gradient buffer, + bf16 operand copy) behind the
``nn.Parameter`` objects of a module tree.

Why flat (B200-first): 180 GB of HBM makes duplication free, and a flat layout turns the per-step tail into three
streaming passes instead of ~600 small kernels -- one cast (fp32 master -> bf16 GEMM operands), one bucketed NCCL
all-reduce over contiguous gradient ranges, one fused clip+Adam (``lasr_clip_adam_step``).  q/k/v projection
weights of every attention module are laid out adjacently so the three Linears run as ONE (3d x d) GEMM.

The ``nn.Parameter`` objects keep the reference names/shapes (state_dict compatible); their ``.data`` are views into
``flat`` and, in direct-gradient mode, their ``.grad`` are views into ``gflat``.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import ops

ALIGN = 64  # elements: every parameter starts on a 256-byte (fp32) / 128-byte (bf16) boundary


def _ordered_named_parameters(root: nn.Module) -> Tuple[List[Tuple[str, nn.Parameter]], set]:
    """Parameters in layout order + the names that are packed TIGHTLY against their successor (no alignment padding): the
    q/k/v weights of an attention module form one (3d, d) matrix and their biases one (3d,) vector, for any d (the fused
    projections read them as single spans, engine.py::rel_mha_fwd / self_mha_fwd / src_mha_fwd)."""
    named = list(root.named_parameters())
    by_name = dict(named)
    out, done, tight = [], set(), set()
    pat = re.compile(r"^(.*)\.linear_q\.weight$")
    for name, p in named:
        if name in done:
            continue
        m = pat.match(name)
        if m:
            pre = m.group(1)
            group = [f"{pre}.linear_{x}.weight" for x in "qkv"] + [f"{pre}.linear_{x}.bias" for x in "qkv"]
            d = by_name[group[0]].shape[0]
            if d % 4 != 0:  # 16-byte alignment of the k / v sub-matrices (bf16 operands) and of the k / v bias vectors (fp32)
                raise NotImplementedError(f"{pre}: attention width {d} must be a multiple of 4 for the fused q/k/v projection")
            for g in group:
                out.append((g, by_name[g]))
                done.add(g)
            tight.update(group[0:2] + group[3:5])
            continue
        out.append((name, p))
        done.add(name)
    return out, tight


class ParamStore:
    def __init__(self, root: nn.Module, device: torch.device, precision: str):
        assert precision in ("fp32", "bf16")
        self.root = root
        self.device = device
        self.precision = precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.named, self.tight = _ordered_named_parameters(root)
        self.off: Dict[str, Tuple[int, torch.Size]] = {}
        self.span: Dict[str, int] = {}  # elements a parameter occupies including its padding
        n = 0
        for name, p in self.named:
            self.off[name] = (n, p.shape)
            self.span[name] = p.numel() if name in self.tight else (n + p.numel() + ALIGN - 1) // ALIGN * ALIGN - n
            n += self.span[name]
        n = (n + ALIGN - 1) // ALIGN * ALIGN
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.gflat = torch.zeros(n, dtype=torch.float32, device=device)
        self.wflat = torch.zeros(n, dtype=torch.bfloat16, device=device) if precision == "bf16" else None
        with torch.no_grad():
            for name, p in self.named:
                o, shp = self.off[name]
                view = self.flat[o:o + p.numel()].view(shp)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
            # buffers (BatchNorm running statistics, the sinusoid tables) are read by the kernels through raw pointers: they must
            # live on the store's device as well (module.to(device) normally did that already)
            for mod in root.modules():
                for bname, buf in list(mod._buffers.items()):
                    if buf is not None and buf.device != device:
                        mod._buffers[bname] = buf.to(device)
        from .dropout import RngState
        self.rng = RngState(device)  # {seed, step} of the dropout stream (seeded from torch.initial_seed())
        self.direct_grads = False
        self._cast_version = None
        self.derived: Dict[str, torch.Tensor] = {}
        self.anchor = torch.zeros((), device=device, requires_grad=True)
        self.grad_ready_hook = None  # callable(lo, hi) invoked when gflat[lo:hi] is final for this backward
        for m in root.modules():
            m._lasr_store = self

    # ------------------------------------------------------------------ views
    def valid(self) -> bool:
        name, p = self.named[0]
        name2, p2 = self.named[-1]
        o2, _ = self.off[name2]
        return (p.data_ptr() == self.flat.data_ptr() and
                p2.data_ptr() == self.flat.data_ptr() + 4 * o2 and p.device == self.flat.device)

    def p(self, name: str) -> torch.Tensor:
        o, shp = self.off[name]
        return self.flat[o:o + shp.numel()].view(shp)

    def g(self, name: str) -> torch.Tensor:
        o, shp = self.off[name]
        return self.gflat[o:o + shp.numel()].view(shp)

    def w(self, name: str, rows: int = None, cols: int = None) -> torch.Tensor:
        """GEMM-operand view (bf16 copy or the fp32 master) as a 2-D (rows, cols) matrix; ``rows`` may span several
        adjacent parameters (fused q/k/v)."""
        o, shp = self.off[name]
        src = self.wflat if self.wflat is not None else self.flat
        if rows is None:
            rows, cols = shp[0], shp.numel() // shp[0]
        return src[o:o + rows * cols].view(rows, cols)

    def p_span(self, name: str, numel: int) -> torch.Tensor:
        o, _ = self.off[name]
        return self.flat[o:o + numel]

    def g_span(self, name: str, numel: int) -> torch.Tensor:
        o, _ = self.off[name]
        return self.gflat[o:o + numel]

    def gw(self, name: str, rows: int = None, cols: int = None) -> torch.Tensor:
        o, shp = self.off[name]
        if rows is None:
            rows, cols = shp[0], shp.numel() // shp[0]
        return self.gflat[o:o + rows * cols].view(rows, cols)

    def range_of(self, prefix: str) -> Tuple[int, int]:
        offs = [(o, o + self.span[n]) for n, (o, s) in self.off.items() if n.startswith(prefix)]
        return min(a for a, _ in offs), max(b for _, b in offs)

    # ------------------------------------------------------------------ per-step maintenance
    def refresh_operands(self, force: bool = False) -> None:
        """fp32 master -> bf16 operand copy (one streaming kernel) + derived relayouts; cached on the params' versions."""
        ver = self._version()
        if not force and ver == self._cast_version:
            return
        if self.wflat is not None:
            ops.cast_bf16(self.flat, self.wflat)
        self.derived.clear()
        self._cast_version = ver

    def _version(self):
        # parameters are views of one buffer, so any in-place optimizer update bumps flat's version counter
        return (self.flat._version, self.flat.data_ptr())

    def zero_grads(self) -> None:
        ops.zero_(self.gflat)

    def enable_direct_grads(self) -> None:
        """p.grad become permanent views of gflat; backward accumulates in place and returns no parameter grads."""
        self.direct_grads = True
        for name, p in self.named:
            p.grad = self.g(name)

    def grads_for_autograd(self, names: List[str]) -> List[torch.Tensor]:
        # references are retained by the caller so AccumulateGrad clones instead of aliasing gflat
        return [self.g(n) for n in names]


def get_store(module: nn.Module, device: torch.device, precision: str) -> ParamStore:
    st = getattr(module, "_lasr_store", None)
    if st is not None and st.device == device and st.precision == precision and st.valid():
        return st
    return ParamStore(module, device, precision)
