"""Informational timings of the other BASELINE.json configs (not bench lines): config 2 (32-step DDIM,
B=4, fp32/tf32 engine), config 3 (256-step DDPM, B=8, bf16), config 5 (RePaint 256x10, B=4, bf16), and
a PyTorch-eager GPU run of the oracle restatement as a stand-in for "the reference on the same GPU".
Usage: python tools/bench_configs.py [repaint_steps]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import r2dm_b200 as R  # noqa: E402
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG, repaint_masks  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402

sd = O.random_state_dict(H_CFG, 0)
res = {}


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


d32 = make_ddpm(H_CFG, sd, precision="fp32")
t = timed(lambda: d32.sample(batch_size=4, num_steps=32, progress=False, rng=R.setup_rng(range(4), "cuda"), mode="ddim"))
res["config2_ddim32_b4_fp32(tf32)"] = {"seconds": t, "images_per_s": 4 / t, "images_per_s_at_256_steps": 4 / (t * 8)}
del d32
dbf = make_ddpm(H_CFG, sd, precision="bf16")
t = timed(lambda: dbf.sample(batch_size=8, num_steps=256, progress=False, rng=R.setup_rng(range(8), "cuda"), mode="ddpm"))
res["config3_ddpm256_b8_bf16"] = {"seconds": t, "images_per_s": 8 / t}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(1)
known = torch.randn(4, 2, 64, 1024, generator=g).clamp(-1, 1).cuda()
mask = repaint_masks(4, H_CFG).cuda()
t = timed(lambda: dbf.repaint(known, mask, num_steps=n, num_resample_steps=10, jump_length=1, progress=False,
                              rng=R.setup_rng(range(4), "cuda")))
calls = (n - 1) * 10 + 1
res[f"config5_repaint{n}x10_b4_bf16"] = {"seconds": t, "unet_calls": calls, "images_per_s": 4 / t,
                                        "ms_per_unet_call": 1e3 * t / calls}
# PyTorch eager on the same GPU (oracle restatement, TF32 convs like the reference's default; bf16 autocast)
sdc = {k: v.cuda() for k, v in sd.items()}
x = torch.randn(8, 2, 64, 1024, device="cuda")
cond = torch.full((8,), 0.5, device="cuda")
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
with torch.inference_mode():
    t = timed(lambda: O.unet_forward(sdc, H_CFG, x, cond), reps=5)
    res["torch_eager_gpu_forward_b8_tf32"] = {"ms": 1e3 * t, "images_per_s_at_256_steps": 8 / (t * 256)}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t = timed(lambda: O.unet_forward(sdc, H_CFG, x, cond), reps=5)
    res["torch_eager_gpu_forward_b8_bf16_autocast"] = {"ms": 1e3 * t, "images_per_s_at_256_steps": 8 / (t * 256)}
eng = dbf.model.engine("bf16")
film = eng.cond_embed(cond)
pred = torch.empty_like(x)
t = timed(lambda: eng.forward_film(x, film, pred), reps=20)
res["r2dm_b200_forward_b8_bf16_eager_launches"] = {"ms": 1e3 * t}
print(json.dumps(res, indent=1))
