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
# steady state: the part is power capped, so warm up for ~1.5 s before timing
import time
t0 = time.time()
while time.time() - t0 < 1.5:
    g.replay()
    torch.cuda.synchronize()
times = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / NF)
# per-launch durations INSIDE a graph replay, recorded by the conv kernels themselves (globaltimer)
from r2dm_b200 import _lib as L  # noqa: E402
nl = 256   # program entries (>= launches)
kt = torch.zeros(nl, 2, dtype=torch.int64, device="cuda")
L.check(L.lib().r2dm_debug_set_ktime(eng.h, L.ptr(kt)))
g1 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g1):
    eng.forward_film(x, film, pred)
L.check(L.lib().r2dm_debug_set_ktime(eng.h, None))
kt[:, 0] = -1
kt[:, 1] = 0
torch.cuda.synchronize()
g.replay()
g1.replay()
torch.cuda.synchronize()
ktc = kt.cpu()
agg = {}
prof = None
for _ in range(3):
    prof = eng.profile_forward(x, cond)
    for kind, ms, fl, by in prof:
        a = agg.setdefault(kind, [0.0, 0])
        a[0] += ms / 3
        a[1] += 1
conv_idx = [i for i, (kind, ms, fl, by) in enumerate(prof) if kind in ("conv3x3", "conv1x1")]
dur = {i: (ktc[i, 1] - ktc[i, 0]).item() / 1e3 for i in conv_idx}
c3 = sum(dur[i] for i in conv_idx if prof[i][0] == "conv3x3") / 1e3
c1 = sum(dur[i] for i in conv_idx if prof[i][0] == "conv1x1") / 1e3
span = (ktc[conv_idx[-1], 1] - ktc[conv_idx[0], 0]).item() / 1e6
fl3 = sum(fl for kind, ms, fl, by in prof if kind == "conv3x3")
tag = " ".join(f"{k[9:].lower()}={v}" for k, v in sorted(os.environ.items()) if k.startswith("R2DM_OPT_")) or "default"
if os.environ.get("R2DM_LIB_PATH"):
    tag += " lib=" + os.path.basename(os.environ["R2DM_LIB_PATH"])
kinds = " ".join(f"{k}={a[0]:.3f}" for k, a in agg.items())
print(f"FWD [{tag}] {prec} B={B} graph_ms mean={sum(times) / len(times):.4f} min={min(times):.4f} | in-graph: conv3x3={c3:.3f} "
      f"({fl3 / c3 / 1e9:.0f} TF) conv1x1={c1:.3f} first->last conv={span:.3f} | events: {kinds} "
      f"sum={sum(a[0] for a in agg.values()):.3f}", flush=True)
if "--ops" in sys.argv:
    t00 = ktc[conv_idx[0], 0].item()
    for i, (kind, ms, fl, by) in enumerate(prof):
        if i in dur:
            print(f"  #{i:3d} {kind:10s} event {ms * 1e3:7.1f} us | in-graph start {(ktc[i, 0].item() - t00) / 1e3:8.1f} dur {dur[i]:7.1f} us "
                  f"{fl / dur[i] / 1e6 if fl else 0:7.1f} TF")
        else:
            print(f"  #{i:3d} {kind:10s} event {ms * 1e3:7.1f} us")
