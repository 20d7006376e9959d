"""Developer probe: RePaint (BASELINE config 5 shape, fewer steps) wall time per U-Net call."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import r2dm_b200 as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ddpm, _, _ = R.synthetic_model(device="cuda", precision="bf16", seed=0)
known = torch.rand(4, 2, 64, 1024, device="cuda") * 2 - 1
mask = torch.zeros(4, 2, 64, 1024, device="cuda"); mask[:, :, ::4] = 1
def run():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ddpm.repaint(known, mask, num_steps=n, num_resample_steps=10, jump_length=1, progress=False, rng=R.setup_rng(range(4), "cuda"))
    torch.cuda.synchronize(); return time.perf_counter() - t0
run()
ts = [run() for _ in range(3)]
calls = (n - 1) * 10 + 1
print(f"repaint {n}x10 B=4: {min(ts):.3f} s (runs {[round(t, 3) for t in ts]}), {min(ts) / calls * 1e3:.3f} ms per U-Net call")
