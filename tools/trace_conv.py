"""[needs a library built with EXTRA_FLAGS="-DR2DM_DEV=1" ./build.sh - the role traces and the
R2DM_CONV_DEBUG / R2DM_XF_DEBUG ablation knobs are compiled out of the product build]
Developer tool: per-role timeline of CTA 0 of one fused conv launch (config-H shapes).
Usage: python tools/trace_conv.py [Cin Cout H W [gn]]   (gn=0: no fused GroupNorm transform)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r2dm_b200 import _lib as L, ops

Cin, Cout, H, W = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 64, 64, 1024)
gn = int(sys.argv[5]) if len(sys.argv) >= 6 else 1
B = 8
x = torch.randn(B, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
g = torch.ones(Cin, device="cuda"); b = torch.zeros(Cin, device="cuda")


def run():
    if gn:
        ops.gn_conv2d(x, w, None, gamma=g, beta=b, dtype="bf16")
    else:
        ops.conv2d(x, w, None, dtype="bf16")


for _ in range(2):
    run()
cap = 4096
buf = torch.zeros(5, cap, dtype=torch.int64, device="cuda")
L.lib().r2dm_debug_set_trace(buf.data_ptr(), cap)
run()
torch.cuda.synchronize()
L.lib().r2dm_debug_set_trace(None, 0)
t = buf.cpu()
c0, g0, c1, g1 = (int(v) for v in t[0, cap - 4:].tolist())
t[0, cap - 4:] = 0
first = int(t[t > 0].min())
print(f"kernel start -> first traced event: {(first - c0) / ((c1 - c0) / max(g1 - g0, 1) * 1e3):.2f} us; last event -> kernel end: "
      f"{(c1 - int(t.max())) / ((c1 - c0) / max(g1 - g0, 1) * 1e3):.2f} us")
print(f"shape {Cin}->{Cout} @{H}x{W} gn={gn}: CTA0 lifetime {(g1 - g0) / 1e3:.1f} us, {c1 - c0} SM cycles -> "
      f"{(c1 - c0) / max(g1 - g0, 1) * 1e3:.0f} MHz effective SM clock")
mhz = (c1 - c0) / max(g1 - g0, 1) * 1e3
t0 = int(t[t > 0].min())
names = ["producer(issue)", "mma(wait,commit)", "xform0(wait,arrive)", "epilogue(full,release,-)", "xform1(wait,arrive)"]
for r in range(5):
    ev = [(int(v) - t0) / mhz for v in t[r].tolist() if v > 0]   # SM cycles -> us
    print(names[r], len(ev), "events, us:", " ".join(f"{e:.2f}" for e in ev[:48]))
    if r in (1, 2, 4) and len(ev) >= 4:
        busy = [ev[i + 1] - ev[i] for i in range(0, len(ev) - 1, 2)]
        wait = [ev[i + 2] - ev[i + 1] for i in range(0, len(ev) - 2, 2)]
        print(f"   mean busy (wait-done -> signal) {sum(busy) / len(busy):.3f} us, mean gap (signal -> next wait-done) "
              f"{sum(wait) / len(wait):.3f} us over {len(busy)} stages")
