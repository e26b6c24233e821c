"""CPU: the HOST half of the inference path -- `lasr_ctc_prefix_beam_search` (C++ in liblasr.so, no device work) against the
oracle restatement of models/u2.py:224-261, which test_oracle_golden.py pins to the unmodified reference's n-best lists."""
import math

import numpy as np
import pytest
import torch

from liteasr_b200 import ops
from oracle import u2_oracle as O


def _pruned(lp: torch.Tensor, k: int):
    """Top-k in (log-prob descending, index ascending) order, the order `lasr_logsoftmax_topk` emits."""
    order = torch.sort(-lp, dim=-1, stable=True).indices[:, :k]
    return torch.gather(lp, 1, order).numpy().astype(np.float32), order.numpy().astype(np.int32)


@pytest.mark.parametrize("frames,V,beam,seed", [(1, 12, 10, 0), (25, 12, 10, 1), (60, 50, 10, 2), (40, 300, 10, 3), (30, 40, 4, 4), (12, 6, 5, 5)])
def test_host_prefix_search_matches_oracle(frames, V, beam, seed):
    g = torch.Generator().manual_seed(seed)
    # peaky posteriors (like a trained CTC head) in float32, so that the oracle and the C++ see identical inputs
    lp = torch.log_softmax(3.0 * torch.randn(frames, V, generator=g), dim=-1).float()
    ref = O.prefix_beam_search_logp(lp, beam)
    tv, ti = _pruned(lp, min(beam, V))
    got = ops.ctc_prefix_beam_search_host(tv, ti, beam=beam, blank=0)
    assert [p for p, _ in got] == [p for p, _ in ref]
    for (_, a), (_, b) in zip(got, ref):
        assert a == b or math.isclose(a, b, rel_tol=0, abs_tol=1e-12)  # same float64 operation order: bit-identical in practice


def test_host_prefix_search_known_answer_and_empty():
    lp = torch.log(torch.tensor([[0.6, 0.3, 0.1], [0.5, 0.4, 0.1]], dtype=torch.float64)).float()
    tv, ti = _pruned(lp, 3)
    hyps = dict(ops.ctc_prefix_beam_search_host(tv, ti, beam=3, blank=0))
    assert math.isclose(math.exp(hyps[()]), 0.30, rel_tol=1e-6)
    assert math.isclose(math.exp(hyps[(1,)]), 0.51, rel_tol=1e-6)
    assert math.isclose(math.exp(hyps[(2,)]), 0.12, rel_tol=1e-6)
    # zero frames: the initial hypothesis survives (models/u2.py:226, 259)
    got = ops.ctc_prefix_beam_search_host(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), beam=3, blank=0)
    assert got == [((), 0.0)]
