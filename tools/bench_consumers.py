"""Developer tool: device timings (CUDA events, steady state) of the caller-side consumers on B=8 generated-size
inputs: render splat 800 x 800, BEV histogram, surface normals, PointNet features, LiDAR post-processing.
Usage: python tools/bench_consumers.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import r2dm_b200 as R  # noqa: E402
from r2dm_b200 import pointnet, render  # noqa: E402

B, H, W = 8, 64, 1024
g = torch.Generator(device="cuda").manual_seed(0)
lu = R.LiDARUtility((H, W), "log_depth", 1.45, 80.0).cuda()
sample = (torch.rand(B, 2, H, W, device="cuda", generator=g) * 2 - 1) * 0.9
cloud = lu.postprocess(sample)                       # [B,5,H,W]
xyz = cloud[:, 1:4].contiguous()
pts = (xyz / 80.0).flatten(2).transpose(1, 2).contiguous()
colors = torch.rand(B, H * W, 3, device="cuda", generator=g)
Rm, t = render.make_Rt(pitch=torch.pi / 3, yaw=torch.pi / 4, z=0.8, device="cuda")
net = pointnet.PointNet1(k=16).eval().cuda()
net_bf16 = pointnet.PointNet1(k=16, precision="bf16").eval().cuda()
metric = xyz.flatten(2).transpose(1, 2).contiguous()
pc = (xyz / 80.0).flatten(2).contiguous()


def timeit(name, fn, bytes_moved=None, flops=None, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    extra = ""
    if bytes_moved:
        extra += f"  {bytes_moved / ms / 1e6:8.1f} GB/s"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    print(f"{name:48s} {ms * 1e3:9.1f} us{extra}", flush=True)


N = H * W
timeit("lidar postprocess [8,2,64,1024] -> [8,5,...]", lambda: lu.postprocess(sample), bytes_moved=B * N * 4 * 7)
timeit("render_point_clouds 8 x 65536 pts -> 800x800", lambda: render.render_point_clouds(pts, colors, R=Rm, t=t),
       bytes_moved=B * (N * 24 + 800 * 800 * (16 * 2 + 12)))
timeit("render_point_clouds 8 x 65536 pts -> 256x256", lambda: render.render_point_clouds(pts, colors, size=256, R=Rm, t=t))
timeit("point_clouds_to_histograms 8 x 65536", lambda: render.point_clouds_to_histograms(metric), bytes_moved=B * N * 12)
timeit("estimate_surface_normal closest [8,3,64,1024]", lambda: render.estimate_surface_normal(xyz), bytes_moved=B * N * 24)
fl = 2 * 2 * B * N * (3 * 64 + 64 * 128 + 128 * 1024)
timeit("PointNet1 features tf32 8 x 65536", lambda: net(pc), flops=fl, n=5)
timeit("PointNet1 features bf16 8 x 65536", lambda: net_bf16(pc), flops=fl, n=5)
