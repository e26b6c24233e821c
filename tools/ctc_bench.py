"""Developer tool (GPU): BASELINE config 4 -- standalone CTC fwd+bwd sweep, B = 64, (T, L) x V grid, fused lasr_ctc_fwdbwd vs the
reference's torch path (log_softmax + nn.CTCLoss(sum) forward + backward on the same GPU, the call of
criterions/hybrid_ctc_attn.py:67-75).  Algorithmic GB/s = T*B*V*(e_in + e_out) / time.

    python tools/ctc_bench.py [--bf16] [--no-torch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    bf16 = "--bf16" in sys.argv
    do_torch = "--no-torch" not in sys.argv
    B = 64
    dt = torch.bfloat16 if bf16 else torch.float32
    es = 2 if bf16 else 4
    print(f"{'T':>5} {'L':>4} {'V':>5} | {'fused ms':>9} {'GB/s':>8} | {'torch ms':>9} {'speed-up':>8} | max|grad - f64|: fused, torch-f32 | nll rel err fused")
    for T, L in ((200, 20), (400, 50), (800, 100), (1600, 200)):
        for V in (500, 1000, 2000, 5000):
            g = torch.Generator(device="cuda").manual_seed(T + V)
            x = torch.randn(T, B, V, generator=g, device="cuda").to(dt)
            il = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda"); il[0] = T
            tl = torch.randint(L // 2, L + 1, (B,), generator=g, device="cuda"); tl[0] = L
            tg = torch.randint(1, V, (B, L), generator=g, device="cuda")
            tg[1, 1] = tg[1, 0]  # at least one repeated label
            grad = torch.empty_like(x)
            ws = torch.empty(ops.ctc_workspace_bytes(T, B, L), dtype=torch.uint8, device="cuda")
            out = {}

            def fused():
                out["nll"], _ = ops.ctc_fwdbwd(x, tg, il, tl, time_major=True, grad=grad, workspace=ws)

            ms = timed(fused, 10 if T * V < 4e6 else 5)
            gbs = T * B * V * 2 * es / (ms * 1e-3) / 1e9
            line = f"{T:5d} {L:4d} {V:5d} | {ms:9.3f} {gbs:8.1f} |"
            if do_torch:
                xr = x.float().clone().requires_grad_(True)

                def ref():
                    xr.grad = None
                    lp = xr.log_softmax(-1)
                    loss = torch.nn.functional.ctc_loss(lp, tg, il, tl, blank=0, reduction="sum", zero_infinity=False)
                    loss.backward()
                    out["ref"] = loss

                ms_t = timed(ref, 5)
                g32 = xr.grad.clone()
                del xr
                x64 = x.double().requires_grad_(True)   # float64 run of the same torch path = the accuracy yardstick
                l64 = torch.nn.functional.ctc_loss(x64.log_softmax(-1), tg, il, tl, blank=0, reduction="sum", zero_infinity=False)
                l64.backward()
                gd = (grad.double() - x64.grad).abs().max().item()
                gt = (g32.double() - x64.grad).abs().max().item()
                nd = abs(float(out["nll"].double().sum()) - float(l64)) / abs(float(l64))
                line += f" {ms_t:9.3f} {ms_t / ms:8.1f} | {gd:.2e} {gt:.2e} | {nd:.2e}"
                del x64, g32
            print(line, flush=True)
            del x, grad, ws


if __name__ == "__main__":
    main()
