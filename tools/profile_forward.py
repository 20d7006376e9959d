"""Developer tool: run N eager U-Net forwards (config H, B=8, bf16) for ncu captures.
Usage: python tools/profile_forward.py [n_forwards] [precision] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ddpm = make_ddpm(H_CFG, O.random_state_dict(H_CFG, 0), precision=prec)
eng = ddpm.model.engine(prec)
x = torch.randn(B, 2, 64, 1024, device="cuda")
cond = torch.full((B,), 0.5, device="cuda")
film = eng.cond_embed(cond)
pred = torch.empty_like(x)
torch.cuda.synchronize()
print("MARK forwards begin", flush=True)
for _ in range(n):
    eng.forward_film(x, film, pred)
torch.cuda.synchronize()
if os.environ.get("R2DM_PRINT_PROFILE"):
    agg = {}
    for kind, ms, fl, by in eng.profile_forward(x, cond):
        a = agg.setdefault(kind, [0.0, 0, 0.0, 0.0]); a[0] += ms; a[1] += 1; a[2] += fl; a[3] += by
    for k, a in agg.items():
        print(f"{k:12s} {a[1]:3d} launches {a[0]:8.3f} ms  {a[2]/a[0]/1e9 if a[2] else 0:8.1f} TFLOP/s {a[3]/a[0]/1e6:8.1f} GB/s")
    for i, (kind, ms, fl, by) in enumerate(eng.profile_forward(x, cond)):
        print(f"  #{i:3d} {kind:10s} {ms*1e3:8.1f} us  {fl/ms/1e9 if fl else 0:7.1f} TF  {by/ms/1e6:7.1f} GB/s")
