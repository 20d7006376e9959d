"""Developer tool: per-layer error of the CUDA U-Net against the CPU oracle (needs a GPU).
Usage: R2DM_KEEP_ACTIVATIONS=1 python tools/debug_taps.py [small|H] [fp32|bf16]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("R2DM_KEEP_ACTIVATIONS", "1")

import r2dm_b200 as R  # noqa: E402
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG, SMALL_CFG, rel_l2  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "small"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
cfg = SMALL_CFG if which == "small" else H_CFG
B = 2 if which == "small" else 1
sd = O.random_state_dict(cfg, 1234)
ddpm = make_ddpm(cfg, sd, precision=prec)
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 2, *cfg.resolution, generator=g)
cond = O.log_snr(torch.tensor([0.3, 0.85][:B]))
taps = {}
ref = O.unet_forward(sd, cfg, x, cond, taps)
y = ddpm.model(x.cuda(), cond.cuda())
torch.cuda.synchronize()
eng = ddpm.model.engine()
for name, t in taps.items():
    try:
        got = eng.debug_tensor(name)[:, : t.shape[1]]
        print(f"{name:24s} l2-rel={rel_l2(got, t):.3e}  max={float((got.cpu() - t).abs().max()):.3e}")
    except Exception as e:  # noqa: BLE001
        print(name, "unavailable:", e)
print(f"{'output':24s} l2-rel={rel_l2(y, ref):.3e}  max={float((y.cpu() - ref).abs().max()):.3e}")
