"""``torch.autograd.Function`` wrappers binding the engine's hand-written forward/backward pipelines into autograd.

Parameter gradients: every backward writes into the flat gradient buffer of the ``ParamStore``.
  * autograd mode (default; works under the reference's own ``Trainer`` / torch DDP): the net's gradient range is zeroed,
    filled, and views of it are returned to autograd (references are retained, so AccumulateGrad clones them);
  * direct mode (``store.enable_direct_grads()``, used by ``liteasr_b200.trainer``): ``p.grad`` are permanent views of the
    flat buffer, backward accumulates in place and returns no parameter gradients -- no copies, and the flat buffer is
    what the bucketed NCCL all-reduce and the fused Adam consume.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import dropout as DR
from . import ops
from .engine import Engine
from .store import ParamStore, get_store

DEFAULT_PRECISION = "bf16"


def precision_of(module: nn.Module) -> str:
    return getattr(module, "lasr_precision", None) or DEFAULT_PRECISION


def set_precision(module: nn.Module, precision: str) -> nn.Module:
    """'bf16': tcgen05 tensor-core GEMMs with bf16 operands / fp32 accumulate.  'fp32': SIMT fp32 GEMMs (parity mode)."""
    assert precision in ("bf16", "fp32")
    for m in module.modules():
        m.lasr_precision = precision
    return module


def bind(module: nn.Module, device: torch.device) -> Tuple[ParamStore, Engine, str]:
    """Find (or build) the flat store covering ``module`` and the module's name prefix inside it."""
    if device.type != "cuda":
        raise RuntimeError("liteasr_b200 modules run on CUDA only (there is no CPU fallback); got device " + str(device))
    st = get_store(module, device, precision_of(module))
    eng = getattr(st, "_engine", None)
    if eng is None:
        eng = st._engine = Engine(st)
    pfx = getattr(module, "_lasr_prefix_cache", None)
    if pfx is None or pfx[0] is not st:
        name = ""
        if st.root is not module:
            for n, m in st.root.named_modules():
                if m is module:
                    name = n + "."
                    break
        module._lasr_prefix_cache = (st, name)
        pfx = module._lasr_prefix_cache
    module._lasr_prefix = pfx[1]
    return st, eng, pfx[1]


def _param_names(st: ParamStore, prefix: str) -> List[str]:
    return [n for n, _ in st.named if n.startswith(prefix)]


def _begin_backward(st: ParamStore, prefix: str) -> None:
    if not st.direct_grads:
        lo, hi = st.range_of(prefix)
        ops.zero_(st.gflat[lo:hi])


def _param_grads(st: ParamStore, prefix: str, n: int):
    if st.direct_grads:
        return (None,) * n
    views = st.grads_for_autograd(_param_names(st, prefix))
    keep = getattr(st, "_keepalive", None)
    if keep is None:
        keep = st._keepalive = {}
    keep[prefix] = views  # retained -> autograd clones instead of aliasing the flat buffer
    return tuple(views)


def net_params(module: nn.Module, st: ParamStore, prefix: str) -> List[torch.Tensor]:
    by = dict(st.named)
    return [by[n] for n in _param_names(st, prefix)]


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, xs, xlens, anchor, *params):
        st, eng, pfx = bind(module, xs.device)
        st.refresh_operands(force=True)
        rng = st.rng.begin_pass() if module.training and DR.has_dropout(module) else None
        c = eng.encoder_fwd(module, xs.contiguous().float(), xlens, module.training, rng)
        ctx.c, ctx.st, ctx.eng, ctx.pfx, ctx.np = c, st, eng, pfx, len(params)
        return c.out

    @staticmethod
    def backward(ctx, dh):
        _begin_backward(ctx.st, ctx.pfx)
        ctx.eng.encoder_bwd(ctx.c, dh.float())
        out = (None, None, None, None) + _param_grads(ctx.st, ctx.pfx, ctx.np)
        ctx.c = None
        return out


class CTCHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, h_enc, anchor, *params):
        st, eng, pfx = bind(module, h_enc.device)
        rng = st.rng.begin_pass() if DR.has_dropout(module) else None  # train AND eval (nets/ctc.py:29, quirk Q3)
        c = eng.ctc_head_fwd(module, h_enc.contiguous().float(), rng)
        ctx.c, ctx.st, ctx.eng, ctx.pfx, ctx.np = c, st, eng, pfx, len(params)
        return c.out.view(c.B, c.Tp, -1) if c.out.is_contiguous() else c.out.unflatten(0, (c.B, c.Tp))

    @staticmethod
    def backward(ctx, dlogits):
        c = ctx.c
        _begin_backward(ctx.st, ctx.pfx)
        dl = _as_padded_2d(dlogits, c.B * c.Tp, c.V, ctx.eng.adt)
        dh = ctx.eng.ctc_head_bwd(c, dl)
        out = (None, dh.view(c.B, c.Tp, c.d), None) + _param_grads(ctx.st, ctx.pfx, ctx.np)
        ctx.c = None
        return out


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, tokens, ylens, memory, xlens, anchor, *params):
        st, eng, pfx = bind(module, memory.device)
        rng = st.rng.begin_pass() if module.training and DR.has_dropout(module) else None
        c = eng.decoder_fwd(module, tokens.contiguous(), ylens, memory.contiguous().float(), xlens, training=module.training, rng=rng)
        ctx.c, ctx.st, ctx.eng, ctx.pfx, ctx.np = c, st, eng, pfx, len(params)
        return c.out.unflatten(0, (c.B, c.L))

    @staticmethod
    def backward(ctx, dlogits):
        c = ctx.c
        _begin_backward(ctx.st, ctx.pfx)
        dl = _as_padded_2d(dlogits, c.B * c.L, c.V, ctx.eng.adt)
        dmem = torch.empty((c.B * c.Tp, c.d), dtype=torch.float32, device=dl.device)
        ops.zero_(dmem)
        ctx.eng.decoder_bwd(c, dl, dmem)
        out = (None, None, None, dmem.view(c.B, c.Tp, c.d), None, None) + _param_grads(ctx.st, ctx.pfx, ctx.np)
        ctx.c = None
        return out


def _as_padded_2d(g: torch.Tensor, rows: int, V: int, adt) -> torch.Tensor:
    """(.., V) gradient from an arbitrary criterion -> (rows, V) view of a row-padded buffer in the operand dtype."""
    ld = (V + 7) // 8 * 8
    g2 = g.reshape(rows, V)
    if g2.dtype == adt and g2.stride(1) == 1 and g2.stride(0) % 8 == 0 and g2.data_ptr() % 16 == 0:
        return g2
    buf = torch.zeros((rows, ld), dtype=adt, device=g.device)
    buf[:, :V].copy_(g2)  # plumbing copy for third-party criterions; the fused criterion never takes this path
    return buf[:, :V]


# --------------------------------------------------------------------------------------------------
# losses on logits (usable with any model that yields logits)
# --------------------------------------------------------------------------------------------------
class CTCLossFn(torch.autograd.Function):
    """sum_b nll_b of CTC on (B,T',V) logits (log-softmax fused); gradient fused with the log-softmax backward."""

    @staticmethod
    def forward(ctx, logits, targets, in_len, tgt_len, blank):
        nll, grad = ops.ctc_fwdbwd(logits, targets.clamp(min=0), in_len, tgt_len, time_major=False, blank=blank)
        ctx.save_for_backward(grad)
        return nll

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout.to(grad.dtype).view(-1, 1, 1), None, None, None, None


class LabelSmoothingFn(torch.autograd.Function):
    """per-row label-smoothed KL on (B,L,V) logits with targets built from ys/ylens (models/u2.py:323-328)."""

    @staticmethod
    def forward(ctx, logits, ys, ylens, smoothing):
        B, L, V = logits.shape
        l2 = logits.reshape(B * L, V)
        if l2.stride(1) != 1:
            l2 = l2.contiguous()
        grad = torch.empty_like(l2)
        row = torch.empty(B * L, dtype=torch.float32, device=logits.device)
        ops.lsmooth_kl_fwdbwd(l2, ys.contiguous(), ylens, V, smoothing, 1.0, None, row, grad)
        ctx.save_for_backward(grad)
        ctx.shape = (B, L, V)
        return row

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return (grad * gout.to(grad.dtype).view(-1, 1)).view(ctx.shape), None, None, None


# --------------------------------------------------------------------------------------------------
# the fused training criterion: U2 forward + hybrid CTC/attention loss + full hand-written backward
# --------------------------------------------------------------------------------------------------
class HybridLossFn(torch.autograd.Function):
    """criterions/hybrid_ctc_attn.py:39-79 on a liteasr_b200 ``U2``: one autograd node for the whole step.

    forward : encoder -> CTC head -> decoder -> fused CTC fwd+bwd (grad wrt logits, scale w/B) -> fused label-smoothed KL
              (grad scale (1-w)/B) -> deterministic combine.  The logits gradients are produced BY the loss kernels.
    backward: CTC-head dgrad/wgrad -> decoder backward (memory gradient accumulated into the same fp32 buffer) ->
              encoder backward; parameter gradients land in the flat buffer.
    """

    @staticmethod
    def forward(ctx, model, ctc_weight, smoothing, xs, xlens, ys, ylens, anchor, *params):
        st, eng, _ = bind(model, xs.device)
        for sub in (model.encoder, model.decoder, model.ctc):
            bind(sub, xs.device)
        st.refresh_operands(force=True)
        B = xs.shape[0]
        # one RNG snapshot for the whole step; the CTC-head site is live in eval mode too (nets/ctc.py:29, quirk Q3)
        need_rng = (model.training and DR.has_dropout(model)) or DR.has_dropout(model.ctc)
        rng = st.rng.begin_pass() if need_rng else None
        ce = eng.encoder_fwd(model.encoder, xs.contiguous().float(), xlens, model.training, rng)
        cc = eng.ctc_head_fwd(model.ctc, ce.out, rng)
        tokens = model.decoder_tokens(ys)
        cd = eng.decoder_fwd(model.decoder, tokens, ylens, ce.out, xlens, training=model.training, rng=rng)
        V, Tp, L = cc.V, ce.Tp, cd.L
        ld = (V + 7) // 8 * 8
        dl_ctc = torch.empty((B * Tp, ld), dtype=eng.adt, device=xs.device)
        dl_att = torch.empty((B * L, ld), dtype=eng.adt, device=xs.device)
        nll, _ = ops.ctc_fwdbwd(cc.out.unflatten(0, (B, Tp)), ys.contiguous(), model.get_pred_len(xlens), ylens, time_major=False,
                                grad=dl_ctc[:, :V].unflatten(0, (B, Tp)), grad_scale=ctc_weight / B, blank=model.blank)
        row_kl = torch.empty(B * L, dtype=torch.float32, device=xs.device)
        ops.lsmooth_kl_fwdbwd(cd.out, ys.contiguous(), ylens, V, smoothing, (1.0 - ctc_weight) / B, None, row_kl, dl_att[:, :V])
        out = torch.empty(3, dtype=torch.float32, device=xs.device)
        ops.hybrid_combine(nll, row_kl, ctc_weight, out)
        ctx.saved = (ce, cc, cd, dl_ctc, dl_att)
        ctx.st, ctx.eng, ctx.np, ctx.V = st, eng, len(params), V
        model.last_losses = out  # [loss, ctc term, attention term] (device tensor; read lazily for logging)
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        ce, cc, cd, dl_ctc, dl_att = ctx.saved
        st, eng, V = ctx.st, ctx.eng, ctx.V
        if not st.direct_grads:
            ops.zero_(st.gflat)
            g = gout.reshape(1).float().contiguous()
            ops.scale_by_scalar(dl_ctc, g)
            ops.scale_by_scalar(dl_att, g)
        dh = eng.ctc_head_bwd(cc, dl_ctc[:, :V])
        if st.grad_ready_hook is not None:
            st.grad_ready_hook(*st.range_of(cc.pfx))
        eng.decoder_bwd(cd, dl_att[:, :V], dh)
        eng.encoder_bwd(ce, dh)
        ctx.saved = None
        if st.direct_grads:
            pg = (None,) * ctx.np
        else:
            views = st.grads_for_autograd([n for n, _ in st.named])
            st._keepalive_all = views
            pg = tuple(views)
        return (None,) * 8 + pg


class _PlainCtx:
    """Attribute bag standing in for the autograd context when HybridLossFn's forward/backward are called directly."""


def hybrid_direct_step(model, ctc_weight: float, smoothing: float, xs, xlens, ys, ylens) -> torch.Tensor:
    """HybridLossFn forward + backward with no autograd graph (direct-gradient stores only): parameter gradients are written
    to the store's flat gradient buffer by the kernels; returns the detached loss (model.last_losses holds the parts)."""
    ctx = _PlainCtx()
    with torch.no_grad():
        loss = HybridLossFn.forward(ctx, model, ctc_weight, smoothing, xs, xlens, ys, ylens, None)
        if not ctx.st.direct_grads:
            raise RuntimeError("hybrid_direct_step needs a store in direct-gradient mode (TrainStep enables it)")
        HybridLossFn.backward(ctx, None)
    return loss
