"""Inference side of the path (models/u2.py:221-317 of the reference, driven per utterance by infer.py:97-120).

* ``greedy_ctc``              argmax_v log_softmax(ctc_lo(encoder(x))) -> collapse repeats -> drop blank (batched; defined from
                              nets/ctc.py:25-26, the reference itself ships no greedy decoder).
* ``ctc_prefix_beam_search``  models/u2.py:221-261: maskless eval-mode encoder on ONE utterance, per-frame top-``beam`` prune on
                              the GPU (``lasr_logsoftmax_topk``), the dictionary search itself on the host in float64 with the
                              reference's operation order (``lasr_ctc_prefix_beam_search``).
* ``attention_rescore``       models/u2.py:269-317: one decoder pass over the padded n-best, the (beam, L, V) log-softmax is never
                              materialised (row log-sum-exp + ``lasr_gather_logp``), scores summed left to right in fp32.
* ``inference_batch``         the same result for a list of utterances: encoders run one by one (bit-identical to the batch-1
                              reference call: BatchNorm running statistics, no padding), ONE decoder pass rescoring every n-best
                              of every utterance (memory padded + masked by its true length).

Everything runs under ``torch.no_grad()`` on the hand-written CUDA kernels; there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import functions as F
from . import ops

BEAM = 10          # models/u2.py:223 (hard-coded in the reference)
CTC_WEIGHT = 0.5   # models/u2.py:312


def log_softmax(logits: Tensor) -> Tensor:
    """(…, V) logits (fp32 or bf16, last-dim stride 1) -> fp32 log-probabilities (nets/ctc.py:25-26)."""
    V = logits.shape[-1]
    x2 = logits.reshape(-1, V) if logits.is_contiguous() else logits.flatten(0, -2)
    _, _, _, full = ops.logsoftmax_topk(x2, 0, vocab=V, want_full=True)
    return full.view(*logits.shape[:-1], V)


def _encode(model, xs: Tensor, xlens: Optional[Tensor]):
    """Eval-mode encoder + CTC head -> (h_enc (B,T',d) fp32, logits 2-D view (B*T', >=V), T')."""
    st, eng, _ = F.bind(model, xs.device)
    F.bind(model.encoder, xs.device)
    F.bind(model.ctc, xs.device)
    st.refresh_operands(force=True)
    enc = eng.encoder_fwd(model.encoder, xs.contiguous().float(), xlens, False)
    head = eng.ctc_head_fwd(model.ctc, enc.out)
    return enc.out, head.out, enc.Tp


@torch.no_grad()
def greedy_ctc(model, xs: Tensor, xlens: Optional[Tensor] = None) -> Tuple[List[List[int]], Tensor]:
    """xs (B,T,F); xlens (B,) or None (maskless, every frame decoded).  -> (token lists, frame ids (B,T') int32)."""
    if model.training:
        raise RuntimeError("greedy_ctc runs on the BatchNorm running statistics: call model.eval() first")
    h, logits, Tp = _encode(model, xs, xlens)
    B = xs.shape[0]
    _, idx, _, _ = ops.logsoftmax_topk(logits, 1, vocab=model.vocab_size)
    ids = idx.view(B, Tp)
    host = ids.cpu().numpy()
    n = [Tp] * B if xlens is None else [int(v) for v in model.get_pred_len(xlens.cpu())]
    out = []
    for b in range(B):
        row = host[b, : n[b]]
        keep = np.ones(len(row), dtype=bool)
        keep[1:] = row[1:] != row[:-1]
        keep &= row != model.blank
        out.append([int(v) for v in row[keep]])
    return out, ids


@torch.no_grad()
def _prefix_search(model, x: Tensor, beam: int = BEAM):
    """-> (hyps [(prefix tuple, score)], h_enc (1,T',d))."""
    if model.training:
        raise RuntimeError("inference runs on the BatchNorm running statistics: call model.eval() first")
    assert x.dim() == 3 and x.size(0) == 1, "the reference decodes one utterance at a time (infer.py:97-120)"
    h, logits, Tp = _encode(model, x, None)
    k = min(beam, model.vocab_size)
    tv, ti, _, _ = ops.logsoftmax_topk(logits, k, vocab=model.vocab_size)
    hyps = ops.ctc_prefix_beam_search_host(tv.cpu().numpy(), ti.cpu().numpy(), beam=beam, blank=model.blank)
    return hyps, h


def ctc_prefix_beam_search(model, x: Tensor, beam: int = BEAM):
    """models/u2.py:221-261 -> (hyps, h) like the reference's ``_ctc_prefix_beam_search``."""
    return _prefix_search(model, x, beam)


@torch.no_grad()
def _rescore(model, nbest: Sequence[Sequence[Tuple[tuple, float]]], mems: Sequence[Tensor]):
    """One decoder pass over every hypothesis of every utterance.  nbest[u] = [(prefix, ctc score)], mems[u] = (1,T'_u,d).
    -> per utterance (best index, [scores])."""
    dev = mems[0].device
    st, eng, _ = F.bind(model, dev)
    F.bind(model.decoder, dev)
    flat = [(u, hy) for u, hyps in enumerate(nbest) for hy in hyps]
    n = len(flat)
    lmax = max(len(hy[0]) for _, hy in flat)
    toks = np.full((n, lmax + 1), model.eos, dtype=np.int64)      # [sos | hyp | eos padding]  (models/u2.py:346-353)
    toks[:, 0] = model.sos
    look = np.full((n, lmax + 1), -1, dtype=np.int64)             # class looked up at each position (-1 = none)
    ylens = np.zeros(n, dtype=np.int64)
    for i, (_, hy) in enumerate(flat):
        L = len(hy[0])
        ylens[i] = L
        if L:
            toks[i, 1:L + 1] = hy[0]
            look[i, :L] = hy[0]
        look[i, L] = model.eos
    tmax = max(m.shape[1] for m in mems)
    d = mems[0].shape[2]
    single = len(mems) == 1
    if single:
        mem = mems[0].expand(n, tmax, d).contiguous()            # h.repeat(len(hyps), 1, 1)  (models/u2.py:273)
        mlens = None                                              # memory_mask=None (:297)
    else:
        mem = torch.zeros((n, tmax, d), dtype=torch.float32, device=dev)
        ml = np.zeros(n, dtype=np.int64)
        for i, (u, _) in enumerate(flat):
            t = mems[u].shape[1]
            mem[i, :t] = mems[u][0]
            ml[i] = t
        mlens = torch.from_numpy(ml).to(dev)
    model.decoder.pe.ensure(lmax + 1, dev)
    c = eng.decoder_fwd(model.decoder, torch.from_numpy(toks).to(dev), torch.from_numpy(ylens).to(dev), mem, mlens, mem_mask_mode=1)
    V = model.vocab_size
    _, _, lse, _ = ops.logsoftmax_topk(c.out, 0, vocab=V, want_lse=True)
    lp = ops.gather_logp(c.out, lse, torch.from_numpy(look.reshape(-1)).to(dev), V).view(n, lmax + 1).cpu().numpy()
    out, i = [], 0
    for hyps in nbest:
        best_score, best_index, scores = -float("inf"), 0, []
        for j, hy in enumerate(hyps):
            s = np.float32(0.0)
            for t in range(len(hy[0]) + 1):                       # left-to-right fp32 sum like `score += attn_score[i][j][w]`
                s = np.float32(s + lp[i, t])
            s = np.float32(s + np.float32(hy[1] * CTC_WEIGHT))
            scores.append(float(s))
            if s > best_score:
                best_score, best_index = s, j
            i += 1
        out.append((best_index, scores))
    return out


def attention_rescore(model, x: Tensor, beam: int = BEAM, return_details: bool = False):
    """models/u2.py:269-317 for one utterance x (1,T,F) -> best hypothesis (list of token ids)."""
    hyps, h = _prefix_search(model, x, beam)
    (best, scores), = _rescore(model, [hyps], [h])
    if return_details:
        return dict(best=list(hyps[best][0]), hyps=hyps, scores=scores)
    return list(hyps[best][0])


def inference_batch(model, xs_list: Sequence[Tensor], beam: int = BEAM, return_details: bool = False):
    """Decode a list of utterances (each (T_u,F) or (1,T_u,F)); token-identical to calling ``attention_rescore`` one by one."""
    nbest, mems = [], []
    for x in xs_list:
        x = x if x.dim() == 3 else x.unsqueeze(0)
        hyps, h = _prefix_search(model, x, beam)
        nbest.append(hyps)
        mems.append(h)
    res = _rescore(model, nbest, mems)
    if return_details:
        return [dict(best=list(h[b][0]), hyps=h, scores=s) for h, (b, s) in zip(nbest, res)]
    return [list(h[b][0]) for h, (b, _) in zip(nbest, res)]
