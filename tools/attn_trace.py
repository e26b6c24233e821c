"""Phase timeline of the fused attention kernel (clock64 stamps of one epilogue warp per CTA)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from liteasr_b200 import _lib, ops  # noqa: E402
from tools.attn_bench import make_inputs  # noqa: E402

B, H, T, dk = 126, 4, 299, 64
d, ld = H * dk, (T + 7) // 8 * 8
qu, qv, k, v, pos = make_inputs(B, H, T, dk, 0, 1.0)
lens = torch.randint(int(0.6 * 4 * T), 4 * T, (B,), device="cuda", dtype=torch.int64)
probs = torch.empty((B, H, T, ld), device="cuda", dtype=torch.bfloat16)
o = torch.empty((B * T, d), device="cuda", dtype=torch.bfloat16)
n = B * H * ((T + 126) // 127)
tr = torch.zeros((n, 16), device="cuda", dtype=torch.int64)
for _ in range(3):
    ops.rel_attn_fwd(qu, qv, k, v, pos, probs, o, lens, 3, dk ** -0.5, B, H, T, dk)
_lib.lib().lasr_rel_attn_fwd_set_trace(C.c_void_p(tr.data_ptr()))
ops.rel_attn_fwd(qu, qv, k, v, pos, probs, o, lens, 3, dk ** -0.5, B, H, T, dk)
torch.cuda.synchronize()
_lib.lib().lasr_rel_attn_fwd_set_trace(C.c_void_p(0))
t = tr.cpu().double()
order = [0, 1, 12, 2, 3, 4, 5, 13, 6, 7, 8, 9, 10, 11]
names = ["wait bd (issued in the previous tile)", "TMEM read of bd", "slab-free barrier", "shift stores", "barrier (flat complete)",
         "wait ac", "TMEM read of ac", "pass A (scores)", "pass B (ex2, sums)", "barrier (max, sum exchange)", "pass C (bf16 tile)",
         "wait O", "O store"]
tot = t[:, 11] - t[:, 0]
print(f"tiles {n}, mean cycles per tile {tot.mean():.0f} (min {tot.min():.0f}, max {tot.max():.0f})")
for i, nm in enumerate(names):
    d = t[:, order[i + 1]] - t[:, order[i]]
    print(f"  {nm:40s} {d.mean():8.0f}  ({100 * d.mean() / tot.mean():4.1f} %)   first tile of a CTA {d[:148].mean():8.0f}  later {d[148:].mean():8.0f}")
