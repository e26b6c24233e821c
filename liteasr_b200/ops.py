"""Thin tensor-level wrappers over the C ABI (one Python function per ``lasr_*`` entry point).

All wrappers take CUDA tensors, pass raw device pointers + the current torch stream, and raise on any
non-zero return code.  Nothing here computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SWISH, BF16, F32  # noqa: F401

_P = C.c_void_p


def _ptr(t: Optional[torch.Tensor]):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _stream():
    return _P(torch.cuda.current_stream().cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("liteasr_b200 ops need CUDA tensors (there is no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, m: int, n: int, k: int, *, lda: int, ldb: int, ldc: int,
         ta: bool = False, tb: bool = False, bias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
         ldres: int = 0, aux: Optional[torch.Tensor] = None, alpha: float = 1.0, act: int = ACT_NONE,
         accumulate: bool = False, split_k: int = 1, batch: Tuple[int, int] = (1, 1), sa: Tuple[int, int] = (0, 0),
         sb: Tuple[int, int] = (0, 0), sc: Tuple[int, int] = (0, 0)) -> None:
    """C[b1,b2] = alpha * act(A.B^T + bias) (+ res); see include/lasr.h ``lasr_gemm``."""
    _require_cuda(a, b, c, bias, res, aux)
    if a.dtype != b.dtype:
        raise TypeError("gemm operands must share a dtype")
    g = _lib.GemmArgs()
    g.a, g.b, g.c = a.data_ptr(), b.data_ptr(), c.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.res = res.data_ptr() if res is not None else None
    g.aux = aux.data_ptr() if aux is not None else None
    g.m, g.n, g.k = m, n, k
    g.ab_dtype, g.c_dtype = dtype_code(a), dtype_code(c)
    g.trans_a, g.trans_b = int(ta), int(tb)
    g.lda, g.ldb, g.ldc, g.ldres = lda, ldb, ldc, (ldres if res is not None else 0)
    g.batch1, g.batch2 = batch
    g.sa1, g.sa2 = sa
    g.sb1, g.sb2 = sb
    g.sc1, g.sc2 = sc
    g.alpha, g.act, g.accumulate, g.split_k = alpha, act, int(accumulate), split_k
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("bias must be fp32")
    if res is not None and res.dtype != torch.float32:
        raise TypeError("res must be fp32")
    if aux is not None and aux.dtype != c.dtype:
        raise TypeError("aux must have C's dtype")
    _lib.check(_lib.lib().lasr_gemm(C.byref(g), _stream()), "gemm")


def linear(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias=None, res=None, aux=None, alpha=1.0,
           act=ACT_NONE) -> torch.Tensor:
    """out (M,N) = alpha * act(x (M,K) @ w (N,K)^T + bias) (+ res); rows may be strided (stride(0))."""
    m, k = x.shape
    n = w.shape[0]
    gemm(x, w, out, m, n, k, lda=x.stride(0), ldb=w.stride(0), ldc=out.stride(0), bias=bias, res=res,
         ldres=(res.stride(0) if res is not None else 0), aux=aux, alpha=alpha, act=act)
    return out


def ctc_workspace_bytes(T: int, B: int, lmax: int) -> int:
    return int(_lib.lib().lasr_ctc_workspace_bytes(C.c_int(T), C.c_int(B), C.c_int(lmax)))


def ctc_fwdbwd(logits: torch.Tensor, targets: torch.Tensor, in_len: torch.Tensor, tgt_len: torch.Tensor, *,
               time_major: bool, grad: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
               upstream: Optional[torch.Tensor] = None, blank: int = 0, vocab: Optional[int] = None,
               workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused CTC fwd+bwd.  logits (T,B,V) if time_major else (B,T,V) (last dim contiguous; may be a
    narrowed view of a padded buffer).  Returns (nll (B,) fp32, grad wrt logits with logits' layout)."""
    _require_cuda(logits, targets, in_len, tgt_len, grad, upstream, workspace)
    assert logits.dim() == 3 and logits.stride(2) == 1
    if time_major:
        T, B, V = logits.shape
        st, sb = logits.stride(0), logits.stride(1)
    else:
        B, T, V = logits.shape
        sb, st = logits.stride(0), logits.stride(1)
    if vocab is not None:
        V = vocab
    if grad is None:
        grad = torch.empty_like(logits)
    gst, gsb = (grad.stride(0), grad.stride(1)) if time_major else (grad.stride(1), grad.stride(0))
    assert grad.stride(2) == 1 and grad.dtype == logits.dtype
    assert targets.dtype == torch.int64 and in_len.dtype == torch.int64 and tgt_len.dtype == torch.int64
    targets = targets.contiguous()
    lmax = max(1, targets.shape[1])
    if targets.shape[1] == 0:
        targets = torch.zeros(B, 1, dtype=torch.int64, device=logits.device)
    nbytes = ctc_workspace_bytes(T, B, lmax)
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
    nll = torch.empty(B, dtype=torch.float32, device=logits.device)
    rc = _lib.lib().lasr_ctc_fwdbwd(
        _ptr(logits), C.c_int(dtype_code(logits)), C.c_int64(st), C.c_int64(sb), _ptr(targets), _ptr(in_len),
        _ptr(tgt_len), C.c_int(T), C.c_int(B), C.c_int(V), C.c_int(lmax), C.c_int(blank), C.c_float(grad_scale),
        _ptr(upstream), _ptr(nll), _ptr(grad), C.c_int64(gst), C.c_int64(gsb), _ptr(workspace),
        C.c_size_t(workspace.numel() * workspace.element_size()), _stream())
    _lib.check(rc, "ctc_fwdbwd")
    return nll, grad
