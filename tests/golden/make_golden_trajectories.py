"""Generate config-H trajectory goldens (tests/golden/traj_H.pt) from the REFERENCE implementation.

Runs only in the build container (needs /root/reference, imported read-only; ~15 CPU-minutes on 8
cores).  These pin the regime that bench.py measures (BASELINE.json configs 2/3/5): the real
architecture (config H, 2x64x1024), 32- and 256-step trajectories, RePaint, with per-sample CPU
generators so that the GPU engines can replay the identical noise stream
(models/diffusion/base.py:71-94).  Stored: final samples in full, strided sub-samples of
intermediate states (error-vs-step curves), and a few (x_k -> x_{k+1}) pairs for teacher-forced
single-step checks along the trajectory.  To keep the fixture small (~8 MB) reference OUTPUTS are
stored as fp16 (quantisation 2.4e-4 relative, far below every tolerance that uses them) and sub-sampled
where the tests compare sub-samples; states that are fed back INTO the engines (x_k) stay fp32, and
x_0 = x_T is not stored at all (the tests re-draw it from the seed).

Usage:  python tests/golden/make_golden_trajectories.py
"""
import os
import sys
import time
import warnings

import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.diffusion import ContinuousTimeGaussianDiffusion  # noqa: E402

from oracle import r2dm_oracle as O  # noqa: E402
from tests.golden.make_golden import build_ref_unet, sub  # noqa: E402
from tests.helpers import H_CFG, repaint_masks  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())

WEIGHT_SEED = 1234
CHECKPOINTS_256 = [1, 2, 4, 8, 16, 32, 64, 96, 128, 160, 192, 224, 240, 248, 252, 254, 255, 256]
CHECKPOINTS_32 = [1, 2, 4, 8, 16, 24, 28, 30, 31, 32]
PAIRS_256 = [0, 128, 255]      # (x_k, x_{k+1}) pairs for teacher-forced p_steps


def compact(out):
    """fp16 for reference outputs, sub-sampled x_{k+1}, no stored x_T (see the module docstring)."""
    for key, g in out.items():
        if not isinstance(g, dict):
            continue
        g["final"] = g["final"].half()
        g["subs"] = g["subs"].half()
        if "pairs" in g:
            g["pairs"] = {k: (None if k == 0 else xk.float().clone(), (xk1 if xk1.shape[-1] < 1024 else sub(xk1)).half())
                          for k, (xk, xk1) in g["pairs"].items() if k in PAIRS_256}
    return out


def main():
    cfg = H_CFG
    sd = O.random_state_dict(cfg, seed=WEIGHT_SEED)
    m = build_ref_unet(cfg, sd)
    ddpm = ContinuousTimeGaussianDiffusion(model=m, prediction_type="eps", noise_schedule="cosine")
    ddpm.eval()
    out = {"weight_seed": WEIGHT_SEED}

    # (a) BASELINE config 2: 32-step DDIM (eta = 0), B = 4
    t0 = time.time()
    seeds = [2000 + i for i in range(4)]
    rng = [torch.Generator().manual_seed(s) for s in seeds]
    ys = ddpm.sample(batch_size=4, num_steps=32, progress=False, rng=rng, return_all=True,
                     mode="ddim", ddim_eta=0.0)
    out["ddim32_b4"] = dict(seeds=seeds, steps=32, mode="ddim", eta=0.0, final=ys[-1].clone(),
                            checkpoints=CHECKPOINTS_32,
                            subs=torch.stack([sub(ys[k]) for k in CHECKPOINTS_32]))
    print(f"ddim32_b4 done in {time.time() - t0:.0f}s  final rms={ys[-1].pow(2).mean().sqrt():.4f}", flush=True)

    # (b) BASELINE config 3/4 regime: 256 steps, B = 1, DDPM and DDIM
    for mode in ("ddpm", "ddim"):
        t0 = time.time()
        seeds = [3000 if mode == "ddpm" else 3100]
        rng = [torch.Generator().manual_seed(s) for s in seeds]
        ys = ddpm.sample(batch_size=1, num_steps=256, progress=False, rng=rng, return_all=True,
                         mode=mode, ddim_eta=0.0)
        out[f"{mode}256_b1"] = dict(
            seeds=seeds, steps=256, mode=mode, eta=0.0, final=ys[-1].clone(), checkpoints=CHECKPOINTS_256,
            subs=torch.stack([sub(ys[k]) for k in CHECKPOINTS_256]),
            pairs={k: (ys[k].clone(), ys[k + 1].clone()) for k in PAIRS_256})
        print(f"{mode}256_b1 done in {time.time() - t0:.0f}s  final rms={ys[-1].pow(2).mean().sqrt():.4f}", flush=True)

    # (c) BASELINE config 5 in miniature: RePaint (8 steps, 3 resamplings, jump 1), B = 4
    t0 = time.time()
    g = torch.Generator().manual_seed(10)
    known = torch.randn(4, 2, *cfg.resolution, generator=g).clamp(-1, 1)
    mask = repaint_masks(4, cfg)
    seeds = [4000 + i for i in range(4)]
    rng = [torch.Generator().manual_seed(s) for s in seeds]
    y = ddpm.repaint(known, mask, num_steps=8, num_resample_steps=3, jump_length=1, progress=False,
                     rng=rng, return_all=True)
    out["repaint_8_3_1_b4"] = dict(seeds=seeds, known_seed=10, final=y[-1].clone(), n_states=y.shape[0],
                                   subs=torch.stack([sub(s) for s in y]))
    print(f"repaint done in {time.time() - t0:.0f}s  states={y.shape[0]}", flush=True)

    torch.save(compact(out), os.path.join(HERE, "traj_H.pt"))
    print("written", os.path.join(HERE, "traj_H.pt"), os.path.getsize(os.path.join(HERE, "traj_H.pt")) / 1e6, "MB")


if __name__ == "__main__":
    if "--recompact" in sys.argv:      # re-apply compact() to an existing (fuller) fixture
        f = os.path.join(HERE, "traj_H.pt")
        torch.save(compact(torch.load(f)), f)
        print("recompacted", os.path.getsize(f) / 1e6, "MB")
    else:
        main()
