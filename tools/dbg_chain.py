"""Developer probe: U-Net forward (config H or small) with the current R2DM_OPT_CHAIN setting; saves / compares."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O
from tests.helpers import H_CFG, SMALL_CFG
from tests.util_model import make_ddpm

cfgname, prec, B, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
cfg = H_CFG if cfgname == "H" else SMALL_CFG
ddpm = make_ddpm(cfg, O.random_state_dict(cfg, 0), precision=prec)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 2, *cfg.resolution, generator=g).cuda()
cond = torch.linspace(-3, 3, B).cuda()
y = ddpm.model(x, cond)
torch.cuda.synchronize()
y2 = ddpm.model(x, cond)
torch.cuda.synchronize()
print("launches", ddpm.model.engine(prec).launches_per_forward, "repeatable", torch.equal(y, y2), "absmax", y.abs().max().item(),
      "nan", torch.isnan(y).sum().item())
if os.path.exists(out):
    ref = torch.load(out)
    print("bitwise equal to", out, ":", torch.equal(ref, y.cpu()), "max abs diff", (ref - y.cpu()).abs().max().item())
else:
    torch.save(y.cpu(), out)
