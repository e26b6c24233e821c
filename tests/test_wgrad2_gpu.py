"""GPU: the CTA-pair weight-gradient kernel (csrc/gemm2_wgrad.cu, include/lasr.h ``lasr_wgrad2``: tcgen05.mma.cta_group::2) against a
torch fp32 matmul of the same bf16 operands -- dW += alpha * dy^T x, the autograd backward of nn.Linear -- and against the
single-CTA split-K path of lasr_gemm that it replaces."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,m,n,sk", [(1000, 256, 256, 4), (1001, 256, 256, 1), (64, 256, 512, 1), (5166, 768, 256, 10), (37674, 256, 2048, 9),
                                      (37674, 2048, 256, 9), (3000, 512, 512, 74), (130, 256, 256, 7)])
def test_wgrad2_matches_fp32_reference(K, m, n, sk):
    from liteasr_b200 import ops
    assert ops.wgrad2_supported(m, n) and not ops.wgrad2_supported(m + 128, n) and not ops.wgrad2_supported(m, 4233)
    g = torch.Generator(device=DEV).manual_seed(K + m + n)
    wide = (torch.randn(K, m + 64, generator=g, device=DEV) * 0.1).bfloat16()   # dy = a column slice of a wider matrix
    dy = wide[:, 64:]
    x = (torch.randn(K, n, generator=g, device=DEV) * 0.1).bfloat16()
    base = torch.randn(m, n, generator=g, device=DEV)
    gw = base.clone()
    ops.wgrad2(dy, x, gw, alpha=0.5, split_k=sk)
    torch.cuda.synchronize()
    want = base + 0.5 * (dy.float().t() @ x.float())
    assert float((gw - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-5 * (K ** 0.5)


def test_wgrad2_equals_the_single_cta_path():
    from liteasr_b200 import ops
    K, m, n = 9568, 256, 2048
    g = torch.Generator(device=DEV).manual_seed(7)
    dy = (torch.randn(K, m, generator=g, device=DEV) * 0.1).bfloat16()
    x = (torch.randn(K, n, generator=g, device=DEV) * 0.1).bfloat16()
    a, b = torch.zeros(m, n, device=DEV), torch.zeros(m, n, device=DEV)
    ops.wgrad2(dy, x, a, split_k=9)
    ops.gemm(dy, x, b, m, n, K, lda=m, ldb=n, ldc=n, ta=True, tb=True, accumulate=True, split_k=9)
    torch.cuda.synchronize()
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())   # same products, another summation order across splits


def test_wgrad2_rejects_unsupported_shapes():
    from liteasr_b200 import ops
    dy = torch.zeros(512, 384, device=DEV, dtype=torch.bfloat16)
    x = torch.zeros(512, 256, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.wgrad2(dy, x, torch.zeros(384, 256, device=DEV))
