"""Experiment: how much of a sampler step is NOT the U-Net forward?  Times ddpm.sample(B=8, 256 DDIM
steps) as is, with the per-step noise draws disabled, and with different graph_steps."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402
import r2dm_b200 as R  # noqa: E402

ddpm = make_ddpm(H_CFG, O.random_state_dict(H_CFG, 0), precision="bf16")
B, N = 8, 256


def run(label):
    rng = R.setup_rng(list(range(B)), "cuda")
    for _ in range(2):
        ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=rng, mode="ddim")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=rng, mode="ddim")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label}: {dt * 1e3:.1f} ms per sample() = {dt / N * 1e3:.3f} ms per step, {B / dt:.2f} img/s", flush=True)


for k in (8, 32):
    ddpm.graph_steps = k
    ddpm._loop_state = None
    run(f"graph_steps={k}")
orig = type(ddpm).randn_like
type(ddpm).randn_like = lambda self, x, rng=None, out=None: out if out is not None else orig(self, x, rng=rng)
ddpm._loop_state = None
run("graph_steps=32, no per-step noise draws")
