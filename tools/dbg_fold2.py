import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import r2dm_oracle as O
from tests.helpers import H_CFG, SMALL_CFG
from tests.util_model import make_ddpm
import r2dm_b200 as R
prec, B, cfgname, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
cfg = H_CFG if cfgname == "H" else SMALL_CFG
ddpm = make_ddpm(cfg, O.random_state_dict(cfg, 1234), precision=prec)
y = ddpm.sample(batch_size=B, num_steps=steps, progress=False, rng=R.setup_rng(range(B), "cuda"), mode="ddim", return_all=len(sys.argv) > 5)
torch.cuda.synchronize()
print("OK", prec, B, cfgname, steps, float(y.abs().mean()))
