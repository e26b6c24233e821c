"""Developer tool (GPU): run ONE eager training step of a bench workload inside a cudaProfiler range (for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from liteasr_b200.criterions.hybrid_ctc_attn import HybridCTCLoss, HybridCTCLossConfig  # noqa: E402
from liteasr_b200.models.u2 import U2, U2Config  # noqa: E402
from liteasr_b200.schema import U2Dims  # noqa: E402
from liteasr_b200.trainer import TrainStep  # noqa: E402
from liteasr_b200.utils.synthetic import synth_batch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dropout = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
wl = dict(bench.WORKLOADS[workload])
if len(sys.argv) > 4:
    wl["batch"] = int(sys.argv[4])
dims = U2Dims(*wl["dims"])
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = U2(U2Config(**dims.__dict__, precision=precision, **bench.my_u2_rates(dropout))).to(dev).train()
crit = HybridCTCLoss(HybridCTCLossConfig(vocab_size=dims.vocab_size, smoothing=wl["smoothing"], ctc_weight=wl["ctc_weight"]))
step = TrainStep(model, crit, use_graph=False, device=dev)
batch = tuple(t.to(dev) for t in synth_batch(wl["batch"], wl["tmax"], wl["lmax"], dims.vocab_size, seed=42))
for _ in range(2):
    step.step_eager(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step.step_eager(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
