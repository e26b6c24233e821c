"""Developer tool (GPU): a few fused-CTC calls at one (T, L, V) point (for ncu captures of the three CTC kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402

T, L, V = (int(a) for a in (sys.argv[1:4] + ["1600", "200", "5000"][len(sys.argv) - 1:]))
B = 64
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, B, V, generator=g, device="cuda")
il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda"); il[0] = T
tl = torch.randint(L // 2, L + 1, (B,), generator=g, device="cuda"); tl[0] = L
tg = torch.randint(1, V, (B, L), generator=g, device="cuda")
grad = torch.empty_like(x)
ws = torch.empty(ops.ctc_workspace_bytes(T, B, L), dtype=torch.uint8, device="cuda")
for _ in range(3):
    nll, _ = ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, grad=grad, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, grad=grad, workspace=ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"T={T} L={L} V={V}: {ms:.3f} ms, {T * B * V * 8 / ms / 1e6:.0f} GB/s algorithmic, nll[0]={float(nll[0]):.3f}")
