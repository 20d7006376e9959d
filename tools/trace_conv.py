"""Developer tool: per-role timeline of CTA 0 of one fused conv launch (config-H shapes).
Usage: python tools/trace_conv.py [Cin Cout H W]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r2dm_b200 import _lib as L, ops

Cin, Cout, H, W = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 64, 64, 1024)
B = 8
x = torch.randn(B, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
g = torch.ones(Cin, device="cuda"); b = torch.zeros(Cin, device="cuda")
for _ in range(2):
    ops.gn_conv2d(x, w, None, gamma=g, beta=b, dtype="bf16")
cap = 4096
buf = torch.zeros(4, cap, dtype=torch.int64, device="cuda")
L.lib().r2dm_debug_set_trace(buf.data_ptr(), cap)
ops.gn_conv2d(x, w, None, gamma=g, beta=b, dtype="bf16")
torch.cuda.synchronize()
L.lib().r2dm_debug_set_trace(None, 0)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = ["producer(issue)", "mma(wait,commit)", "xform(wait,arrive)", "epilogue(full,release,-)"]
for r in range(4):
    ev = [(int(v) - t0) / 1e3 for v in t[r].tolist() if v > 0]
    print(names[r], len(ev), "events, us:", " ".join(f"{e:.1f}" for e in ev[:64]))
