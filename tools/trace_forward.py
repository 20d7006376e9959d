"""[needs a library built with EXTRA_FLAGS="-DR2DM_DEV=1" ./build.sh - the role traces and the
R2DM_CONV_DEBUG / R2DM_XF_DEBUG ablation knobs are compiled out of the product build]
Developer tool: role timeline of ONE conv launch inside steady-state forwards (warm clocks, real
predecessors).  Usage: R2DM_TRACE_SKIP=<k> python tools/trace_forward.py   traces the (k+1)-th conv_umma launch
after 5 warm-up forwards (conv launch order within a forward: see profiles/r01_conv_dram_v4.txt; 64 per forward)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402
from r2dm_b200 import _lib as L  # noqa: E402

ddpm = make_ddpm(H_CFG, O.random_state_dict(H_CFG, 0), precision="bf16")
eng = ddpm.model.engine("bf16")
B = 8
x = torch.randn(B, 2, 64, 1024, device="cuda")
cond = torch.full((B,), 0.5, device="cuda")
film = eng.cond_embed(cond)
pred = torch.empty_like(x)
for _ in range(5):
    eng.forward_film(x, film, pred)
cap = 4096
buf = torch.zeros(5, cap, dtype=torch.int64, device="cuda")
L.lib().r2dm_debug_set_trace(buf.data_ptr(), cap)
for _ in range(2):
    eng.forward_film(x, film, pred)
torch.cuda.synchronize()
L.lib().r2dm_debug_set_trace(None, 0)
t = buf.cpu()
c0, g0, c1, g1 = (int(v) for v in t[0, cap - 4:].tolist())
t[0, cap - 4:] = 0
mhz = (c1 - c0) / max(g1 - g0, 1) * 1e3
print(f"traced launch: CTA lifetime {(g1 - g0) / 1e3:.1f} us, {c1 - c0} cycles -> {mhz:.0f} MHz")
names = ["producer(issue)", "mma(wait,commit)", "xform0(wait,arrive)", "epilogue(full,release)", "xform1 / fold(begin,end at idx 0,1)"]
for r in range(5):
    ev = [(int(v) - c0) / mhz for v in t[r].tolist() if v > 0]
    print(names[r], len(ev), "events, us since CTA start:", " ".join(f"{e:.2f}" for e in ev[:40]), "..." if len(ev) > 40 else "",
          f"last {ev[-1]:.2f}" if ev else "")
