"""Developer tool: U-Net forward time (config H, bf16, B=8) inside a CUDA graph + per-kind event profile.
Options are read from the environment (R2DM_OPT_<NAME>), one process per configuration:
    R2DM_OPT_SERPENTINE=0 python tools/fwd_time.py [precision] [batch] [--ops]
Prints one line:  TAG fwd_graph_ms=<min over replays> conv3x3=<ms by events> ... """
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
prec = args[0] if len(args) > 0 else "bf16"
B = int(args[1]) if len(args) > 1 else 8
ddpm = make_ddpm(H_CFG, O.random_state_dict(H_CFG, 0), precision=prec)
eng = ddpm.model.engine(prec)
x = torch.randn(B, 2, 64, 1024, device="cuda")
cond = torch.full((B,), 0.5, device="cuda")
film = eng.cond_embed(cond)
pred = torch.empty_like(x)
for _ in range(3):
    eng.forward_film(x, film, pred)
torch.cuda.synchronize()
NF = 10
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(NF):
        eng.forward_film(x, film, pred)
times = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / NF)
agg = {}
prof = None
for _ in range(3):
    prof = eng.profile_forward(x, cond)
    for kind, ms, fl, by in prof:
        a = agg.setdefault(kind, [0.0, 0])
        a[0] += ms / 3
        a[1] += 1
tag = " ".join(f"{k[9:].lower()}={v}" for k, v in sorted(os.environ.items()) if k.startswith("R2DM_OPT_")) or "default"
kinds = " ".join(f"{k}={a[0]:.3f}" for k, a in agg.items())
print(f"FWD [{tag}] {prec} B={B} graph_ms min={min(times):.4f} med={sorted(times)[len(times) // 2]:.4f} | events: {kinds} "
      f"sum={sum(a[0] for a in agg.values()):.3f}", flush=True)
if "--ops" in sys.argv:
    for i, (kind, ms, fl, by) in enumerate(prof):
        print(f"  #{i:3d} {kind:10s} {ms * 1e3:8.1f} us  {fl / ms / 1e9 if fl else 0:7.1f} TF  {by / ms / 1e6:7.1f} GB/s")
