"""Developer tool: run one fused conv shape a few times (for ncu captures).
Usage: python tools/ncu_conv.py Cin Cout H W gn [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r2dm_b200 import ops

Cin, Cout, H, W, gn = (int(a) for a in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
B = 8
x = torch.randn(B, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
g = torch.ones(Cin, device="cuda"); b = torch.zeros(Cin, device="cuda")
for _ in range(reps):
    if gn:
        ops.gn_conv2d(x, w, None, gamma=g, beta=b, dtype="bf16")
    else:
        ops.conv2d(x, w, None, dtype="bf16")
torch.cuda.synchronize()
