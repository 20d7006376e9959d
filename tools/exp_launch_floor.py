"""Experiment: fixed cost of one conv_umma launch (tiny problem: one tile per CTA or less) back to back in
a CUDA graph, with and without the fused GroupNorm, vs a trivial elementwise kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r2dm_b200 import ops  # noqa: E402


def bench(fn, name, n=50):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / (5 * n) * 1e3:.2f} us per call", flush=True)


for (B, Cin, Cout, H, W) in [(1, 64, 64, 4, 128), (8, 64, 64, 8, 512), (8, 128, 128, 4, 512)]:
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    g_ = torch.ones(Cin, device="cuda"); b_ = torch.zeros(Cin, device="cuda")
    # op-level calls = pack + conv + unpack (3 launches); report per call
    bench(lambda: ops.conv2d(x, w, None, dtype="bf16"), f"conv2d    B={B} {Cin}->{Cout} @{H}x{W} (pack+conv+unpack)")
    bench(lambda: ops.gn_conv2d(x, w, None, gamma=g_, beta=b_, dtype="bf16"), f"gn_conv2d B={B} {Cin}->{Cout} @{H}x{W} (pack+stats+conv+unpack)")
y = torch.randn(1 << 16, device="cuda")
bench(lambda: y.mul_(1.0001), "trivial elementwise kernel")
