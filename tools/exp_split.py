"""Experiment: does running the batch as two half-batches on two streams (one CUDA graph, two parallel
branches) hide the per-layer startup / tail bubbles?  Compares ms per forward of B=8 on one stream with
2 x B=4 on two streams.  Usage: python tools/exp_split.py [n_forwards]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2dm_oracle as O  # noqa: E402
from tests.helpers import H_CFG  # noqa: E402
from tests.util_model import make_ddpm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sd = O.random_state_dict(H_CFG, 0)
B = 8


def engine():
    d = make_ddpm(H_CFG, sd, precision="bf16")
    return d, d.model.engine("bf16")


d0, e0 = engine()
d1, e1 = engine()
d2, e2 = engine()
x = torch.randn(B, 2, 64, 1024, device="cuda")
cond = torch.full((B,), 0.5, device="cuda")
film = e0.cond_embed(cond)
pred = torch.empty_like(x)
xa, xb = x[:4].contiguous(), x[4:].contiguous()
pa, pb = torch.empty_like(xa), torch.empty_like(xb)
fa, fb = e1.cond_embed(cond[:4]), e2.cond_embed(cond[4:])


def run_single():
    e0.forward_film(x, film, pred)


s1 = torch.cuda.Stream()


def run_split():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    e1.forward_film(xa, fa, pa)
    with torch.cuda.stream(s1):
        e2.forward_film(xb, fb, pb)
    cur.wait_stream(s1)


def bench(fn, name):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
        g.replay(); torch.cuda.synchronize()
        e_0, e_1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_0.record()
        for _ in range(3):
            g.replay()
        e_1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e_0.elapsed_time(e_1) / (3 * n):.3f} ms per forward of {B} images", flush=True)


bench(run_single, "single stream, B=8")
bench(run_split, "two streams, 2 x B=4")
torch.cuda.synchronize()
ref = pred.clone()
run_single(); run_split(); torch.cuda.synchronize()
print("max |single - split| =", float((pred - torch.cat([pa, pb])).abs().max()), " (|pred| max", float(pred.abs().max()), ")")
