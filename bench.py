#!/usr/bin/env python
"""Benchmark of the R2DM sampling hot path (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--quick]

One "step" = one full `ddpm.sample()` call: 256 DDIM steps over a batch of 8 LiDAR images
(2x64x1024) per GPU, config-H EfficientUNet, bf16 tensor-core path, synthetic weights / noise.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference's
p_step on the host cores (a bounded sample, extrapolated to 256 steps).

The product path (model construction, sampling, post-processing) imports nothing from oracle/ or tests/;
the oracle is used only by the two BASELINE legs (`cpu_baseline`, `gpu_eager_baseline`), which time the
reference algorithm's restatement beside the product - never instead of it.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_STEPS = 256
BATCH_PER_GPU = 8
GFLOP_PER_IMAGE_STEP = 235.26   # algorithmic work of one U-Net forward per image
WORKLOAD = "ddim256_b8_per_gpu_bf16 (BASELINE configs[3] per-GPU shard: 256-step DDIM, 8 images of 2x64x1024 per GPU)"
METRIC = "LiDAR range images/sec @256 DDIM steps, 2x64x1024"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def oracle_h_cfg():
    """Config H (utils/option.py defaults) as the oracle's UNetCfg - baseline legs only."""
    from oracle import r2dm_oracle as O
    return O, O.UNetCfg(in_channels=2, resolution=(64, 1024), base_channels=64, channel_multiplier=(1, 2, 4, 8),
                        num_residual_blocks=(3, 3, 3, 3))


# --------------------------------------------------------------------------------------- CPU arm
def cpu_pstep_seconds(n_steps: int, warmup: int = 1):
    """Time the oracle's p_step (U-Net forward + DDIM update) at B=1, config H, fp32, all host cores."""
    O, H_CFG = oracle_h_cfg()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.random_state_dict(H_CFG, 0)
    orc = O.OracleDiffusion(sd, H_CFG)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, 64, 1024, generator=g)
    ts = torch.linspace(1.0, 0.0, NUM_STEPS + 1)
    times = []
    with torch.inference_mode():
        for i in range(warmup + n_steps):
            nz = torch.randn(1, 2, 64, 1024, generator=g)
            t0 = time.perf_counter()
            x = orc.p_step(x, ts[i:i + 1], ts[i + 1:i + 2], nz, "ddim", 0.0)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 2  # p_steps per bench "step" (bounded sample of the 256-step workload)
    times, cores = cpu_pstep_seconds(per_step * args.steps, warmup=max(1, min(args.warmup, 2)))
    t = sum(times) / len(times)
    value = 1.0 / (NUM_STEPS * t)
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t * NUM_STEPS, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} p_steps (U-Net fwd + DDIM update) at B=1, config H fp32, "
                                   f"mean {t:.3f} s/p_step, extrapolated x{NUM_STEPS} per image"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        self.power = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w": statistics.median(self.power) if self.power else None}


# --------------------------------------------------------------------------------------- GPU arm
def cuda_ms(fn, iters=1):
    """fn() `iters` times between two CUDA events on the current (launching) stream."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def conv_roofline(eng, B, dev, pk, pk_src, value, world):
    """Roofline of the dominant kernel (3x3 ring conv on tcgen05), measured live with CUDA events on the
    launching stream.  `achieved` times the 56 conv launches of a forward the way the product issues them -
    back to back inside a CUDA graph - after a power-cap warm-up; the eager per-launch event timing (one
    event pair per launch, which adds launch gaps) is reported beside it."""
    x = torch.randn(B, 2, 64, 1024, device=dev)
    cond = torch.full((B,), 0.5, device=dev)
    film = eng.cond_embed(cond)
    pred = torch.empty_like(x)
    eng.profile_forward(x, cond)                       # a real forward: every buffer the convs read is valid
    agg = {}
    reps = 3
    for _ in range(reps):
        for kind, kms, fl, by in eng.profile_forward(x, cond):
            a = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
            a[0] += kms; a[1] += fl; a[2] += by; a[3] += 1
    conv = agg["conv3x3"]
    conv_flops = conv[1] / reps
    fwd_ms_eager = sum(a[0] for a in agg.values()) / reps
    kernels = {k: {"ms_per_forward_eager_events": a[0] / reps, "launches": a[3] // reps,
                   "tflops": a[1] / (a[0] * 1e-3) / 1e12 if a[1] else None,
                   "gbs": a[2] / (a[0] * 1e-3) / 1e9} for k, a in agg.items()}

    def graph_of(fn, n):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        return g

    NF = 8
    g_conv = graph_of(lambda: eng.forward_kinds(x, film, pred, ["conv3x3"]), NF)
    g_full = graph_of(lambda: eng.forward_film(x, film, pred), NF)
    t0 = time.time()
    while time.time() - t0 < 1.0:                      # settle under the power cap, like the long timed step
        g_full.replay()
        torch.cuda.synchronize()
    conv_ms = min(cuda_ms(g_conv.replay, 3) for _ in range(3)) / NF
    full_ms = min(cuda_ms(g_full.replay, 3) for _ in range(3)) / NF
    conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12
    eager_tflops = conv[1] / (conv[0] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv3x3_dram_bytes.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    peak = pk["bf16_tflops_sustained"]
    roofline = {
        "bound": "tensor", "kernel": "conv_umma_kernel (3x3 ring conv, 56 launches/forward)",
        "achieved": conv_tflops, "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
        "peak_source": f"{pk_src} bf16 sustained",
        "how": f"algorithmic FLOPs of the 56 launches / their duration replayed back to back in a CUDA graph "
               f"({NF} forwards' worth per replay, CUDA events on the launching stream, after a 1 s power-cap warm-up)",
        "ms_per_forward": conv_ms, "traffic": traffic,
        "achieved_eager_per_launch_events": eager_tflops, "frac_eager_per_launch_events": eager_tflops / peak,
        "forward_ms_in_graph": full_ms, "share_of_forward": conv_ms / full_ms,
        "whole_step_frac_of_tensor_peak": (value / world) * NUM_STEPS * GFLOP_PER_IMAGE_STEP / 1e3 / peak,
    }
    return roofline, kernels, fwd_ms_eager


def other_configs(R, ddpm_bf16, dev):
    """The other BASELINE.json configs on one GPU (informational; one timed run each after a warm-up run)."""
    out = {}
    ddpm_fp32, _, _ = R.synthetic_model(device=dev, precision="fp32", seed=0)

    def timed(fn, runs=1):
        fn()                                            # warm-up: graph capture, lazy kernel set-up
        return min(cuda_ms(fn) for _ in range(runs))

    ms = timed(lambda: ddpm_fp32.sample(batch_size=4, num_steps=32, progress=False, rng=R.setup_rng(range(4), dev),
                                        mode="ddim"), runs=3)
    out["config2_ddim32_b4_fp32(tf32)"] = {"seconds": ms / 1e3, "images_per_s": 4 / (ms / 1e3)}
    del ddpm_fp32
    ms = timed(lambda: ddpm_bf16.sample(batch_size=8, num_steps=256, progress=False, rng=R.setup_rng(range(8), dev),
                                        mode="ddpm"))
    out["config3_ddpm256_b8_bf16"] = {"seconds": ms / 1e3, "images_per_s": 8 / (ms / 1e3)}
    known = torch.rand(4, 2, 64, 1024, device=dev) * 2 - 1
    mask = torch.zeros(4, 2, 64, 1024, device=dev)
    mask[:, :, ::4] = 1
    ms = timed(lambda: ddpm_bf16.repaint(known, mask, num_steps=256, num_resample_steps=10, jump_length=1,
                                         progress=False, rng=R.setup_rng(range(4), dev)), runs=2)   # host-paced: noisy
    out["config5_repaint256x10_b4_bf16"] = {"seconds": ms / 1e3, "images_per_s": 4 / (ms / 1e3),
                                            "unet_calls": 255 * 10 + 1}
    return out


def gpu_eager_baseline(dev, B):
    """The practical bar: the reference algorithm as plain PyTorch ops on the SAME GPU -
    the oracle restatement moved to CUDA, TF32 allowed (the reference's GPU default) and bf16 autocast,
    cudnn.benchmark on, eager and (when capturable) replayed from a CUDA graph.  Baseline leg only."""
    O, H_CFG = oracle_h_cfg()
    sd = {k: v.to(dev) for k, v in O.random_state_dict(H_CFG, 0).items()}
    x = torch.randn(B, 2, 64, 1024, device=dev)
    cond = torch.full((B,), 0.5, device=dev)
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    res = {"what": "oracle restatement of EfficientUNet.forward on cuda (torch ops, cudnn), B=%d" % B}
    try:
        with torch.inference_mode():
            def fwd():
                return O.unet_forward(sd, H_CFG, x, cond)

            def fwd_bf16():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.unet_forward(sd, H_CFG, x, cond)

            for name, fn in (("tf32", fwd), ("bf16_autocast", fwd_bf16)):
                for _ in range(3):
                    fn()
                ms = min(cuda_ms(fn, 3) for _ in range(2))
                res[f"{name}_eager_ms_per_forward"] = ms
                res[f"{name}_eager_images_per_s_at_256"] = B / (NUM_STEPS * ms / 1e3)
                try:
                    side = torch.cuda.Stream()          # warm up on a side stream (cudnn autotuning must not
                    side.wait_stream(torch.cuda.current_stream())   # happen inside the capture)
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            fn()
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        fn()
                    g.replay()
                    ms = min(cuda_ms(g.replay, 3) for _ in range(2))
                    res[f"{name}_cuda_graph_ms_per_forward"] = ms
                    res[f"{name}_cuda_graph_images_per_s_at_256"] = B / (NUM_STEPS * ms / 1e3)
                except Exception as e:      # not capturable on this torch build: say so, keep the eager number
                    res[f"{name}_cuda_graph"] = f"not capturable: {type(e).__name__}: {str(e)[:160]}"
                    torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def run_ours(args):
    import torch.distributed as dist
    import r2dm_b200 as R
    from r2dm_b200 import parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ddpm, lidar, _ = R.synthetic_model(device=dev, precision="bf16", seed=0)     # config H, random weights
    B = BATCH_PER_GPU
    all_seeds = list(range(B * world))

    def local_sample(seeds):
        return ddpm.sample(batch_size=len(seeds), num_steps=NUM_STEPS, progress=False,
                           rng=R.setup_rng(seeds, dev), mode="ddim")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident arm: sample + the final gather, nothing leaves the GPU
    def step_device():
        parallel.sample_sharded(local_sample, all_seeds)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(step_device, args.steps)
    sampler.stop_flag = True
    value = B * world * args.steps / (ms / 1e3)

    # ---- end-to-end arm through the public API with host buffers (seeds in, 5-channel clouds out)
    host_out = torch.empty(B, 5, 64, 1024, dtype=torch.float32).pin_memory()
    seeds_host = torch.tensor(parallel.shard_seeds(all_seeds, world, rank), dtype=torch.int64).pin_memory()

    def step_e2e():
        seeds_dev = seeds_host.to(dev, non_blocking=True)          # H2D: this step's inputs
        seeds = seeds_dev.tolist()
        x = local_sample(seeds).clamp(-1, 1)
        host_out.copy_(lidar.postprocess(x), non_blocking=True)    # D2H: this step's result
        torch.cuda.current_stream().synchronize()

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    eng = ddpm.model.engine("bf16")
    roofline, kernels, fwd_ms_eager = conv_roofline(eng, B, dev, pk, pk_src, value, world)
    launches = args.steps * (NUM_STEPS * (eng.launches_per_forward + 2) + 2)
    extra = {}
    if not args.quick and world == 1:    # informational legs: single-GPU runs only (keeps the scaling runs short)
        extra["other_configs"] = other_configs(R, ddpm, dev)
        extra["gpu_eager_baseline"] = gpu_eager_baseline(dev, B)

    # ---- CPU baseline (oracle port of the reference path) on this box's host cores, bounded sample
    times, cores = cpu_pstep_seconds(6, warmup=1)
    t = sum(times) / len(times)
    cpu = {"value": 1.0 / (NUM_STEPS * t), "unit": "images/s", "cores": cores, "kind": "port",
           "sample": f"{len(times)} p_steps at B=1, config H fp32, mean {t:.3f} s/p_step, extrapolated x{NUM_STEPS}"}

    line = {
        "metric": METRIC, "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": B * world, "denoising_steps": NUM_STEPS,
                   "sampler": "ddim eta=0", "l2": "working set per step >> L2 (each of the 256 forwards streams "
                   "~1.4 GB of activations; no explicit flush)", "cuda_graph": True,
                   "graph_steps": int(getattr(ddpm, "graph_steps_device_noise", 1)),
                   "noise": "drawn inside the update kernel (Philox4x32-10, bit-identical to the reference's "
                            "per-sample torch.randn on CUDA generators); no host work between steps"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": 8 * B * world,
                "d2h_bytes_per_step": B * world * 5 * 64 * 1024 * 4},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "kernels": kernels,
        "forward_ms_eager_profiled": fwd_ms_eager,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--quick", action="store_true", help="skip the informational other_configs / gpu_eager_baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (r2dm_b200 has no CPU fallback); "
                             "use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
