import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import r2dm_oracle as O
from tests.helpers import H_CFG, SMALL_CFG, rel_l2
from tests.util_model import make_ddpm
prec, B, cfgname = sys.argv[1], int(sys.argv[2]), sys.argv[3]
cfg = H_CFG if cfgname == "H" else SMALL_CFG
sd = O.random_state_dict(cfg, 1234)
ddpm = make_ddpm(cfg, sd, precision=prec)
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 2, *cfg.resolution, generator=g)
cond = torch.linspace(-3, 3, B)
y = ddpm.model(x.cuda(), cond.cuda())
torch.cuda.synchronize()
ref = O.unet_forward(sd, cfg, x[:1], cond[:1])
print("OK", prec, B, cfgname, "err vs oracle (sample 0):", rel_l2(y[:1], ref))
