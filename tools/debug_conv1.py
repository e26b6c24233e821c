"""Developer tool (GPU): conv1 fwd/bwd kernels in isolation vs torch conv2d (fp64) at the C1 geometry."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liteasr_b200 import ops
torch.manual_seed(0)
B, T, F, d = 8, 500, 80, 256
x = torch.randn(B, T, F, device="cuda")
w = (torch.randn(d, 1, 3, 3, device="cuda") * 0.3)
b = torch.randn(d, device="cuda") * 0.1
T1, F1 = (T - 3) // 2 + 1, (F - 3) // 2 + 1
h1 = torch.empty(B, T1, F1, d, device="cuda")
ops.conv1_fwd(x, w.view(d, 9), b, h1)
ref = torch.relu(torch.nn.functional.conv2d(x.double().unsqueeze(1), w.double(), b.double(), stride=2)).permute(0, 2, 3, 1)
print("fwd max err", float((h1.double() - ref).abs().max()), "max", float(ref.abs().max()))
dh = torch.randn(B, T1, F1, d, device="cuda") * (h1 > 0)
dw = torch.zeros(d, 9, device="cuda"); db = torch.zeros(d, device="cuda")
ops.conv1_bwd(x, dh, dw, db)
xx = x.double().unsqueeze(1).requires_grad_(False)
w64 = w.double().requires_grad_(True); b64 = b.double().requires_grad_(True)
y = torch.nn.functional.conv2d(xx, w64, b64, stride=2)
y.backward(dh.double().permute(0, 3, 1, 2))
print("dw max err", float((dw.double().view_as(w64) - w64.grad).abs().max()), "max", float(w64.grad.abs().max()))
print("db max err", float((db.double() - b64.grad).abs().max()), "max", float(b64.grad.abs().max()))
