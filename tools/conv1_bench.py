"""conv1 weight-gradient kernel at the C2 bench shape (B=126, T=1200, F=80, d=256): tcgen05 path vs LASR_CONV1_TC=0 (SIMT)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from liteasr_b200 import ops  # noqa: E402

B, T, F, d = 126, 1200, 80, 256
T1, F1, U, V, T2, F2 = ops.plane_dims(T, F)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, T, F, generator=g, device="cuda")
dh = (torch.randn(B, 4, U * V, d, generator=g, device="cuda") * 0.1 + 0.02).bfloat16()  # padding slots non-zero on purpose
res = {}
for flag in ("1", "0"):
    os.environ["LASR_CONV1_TC"] = flag
    dw = torch.zeros(d, 9, device="cuda")
    db = torch.zeros(d, device="cuda")
    for _ in range(3):
        ops.conv1_bwd_planes(x, dh, dw, db)
    torch.cuda.synchronize()
    dw.zero_(); db.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        ops.conv1_bwd_planes(x, dh, dw, db)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    gb = dh.numel() * 2 / 1e9
    res[flag] = (dw / n, db / n)
    print(f"LASR_CONV1_TC={flag}: {us:.0f} us per call, {gb / (us * 1e-6):.0f} GB/s of the gradient planes ({gb:.2f} GB)")
# forward
w1 = torch.randn(d, 9, generator=g, device="cuda") * 0.3
b1 = torch.randn(d, generator=g, device="cuda") * 0.1
outs = {}
for flag in ("1", "0"):
    os.environ["LASR_CONV1_TC"] = flag
    h1p = torch.empty(B, 4, U * V, d, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv1_fwd_planes(x, w1, b1, h1p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv1_fwd_planes(x, w1, b1, h1p)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    outs[flag] = h1p.float()
    print(f"forward LASR_CONV1_TC={flag}: {us:.0f} us per call, {h1p.numel() * 2 / 1e9 / (us * 1e-6):.0f} GB/s written")
print(f"forward tensor-core vs SIMT: max abs diff {(outs['1'] - outs['0']).abs().max().item():.3e} (bf16 outputs), "
      f"mismatching zeros {int(((outs['1'] == 0) != (outs['0'] == 0)).sum())}")
rel = ((res["1"][0] - res["0"][0]).norm() / res["0"][0].norm()).item()
relb = ((res["1"][1] - res["0"][1]).norm() / res["0"][1].norm()).item()
print(f"tensor-core vs SIMT: dW rel-L2 {rel:.2e}, dbias rel-L2 {relb:.2e}")
