"""Caller-side consumers of the sampled range images — host mirror of the reference's `utils/render.py`
(make_Rt :10-29, render_point_clouds :32-80, bilinear_rasterizer :83-142, estimate_surface_normal :145-234,
colorize :237-246) and of `metrics/bev.py:5-24` (point_cloud_to_histogram), with the arithmetic in CUDA kernels
(`csrc/render.cu`) behind the C ABI (`r2dm_render_point_clouds`, `r2dm_bilinear_rasterize`, `r2dm_surface_normal`,
`r2dm_bev_histogram`).  Same names, argument meaning and shapes as the reference; tensors must live on a CUDA
device (there is no CPU path here - the CPU restatement is `oracle/render_oracle.py`, test infrastructure only).

`make_Rt` restates kornia 0.7.0's `axis_angle_to_rotation_matrix` (the reference's pinned dependency,
environment.yaml:14; not installed in this image) for the three single-axis rotations the reference composes.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L


def _axis_angle_to_rotation_matrix(v: torch.Tensor) -> torch.Tensor:
    """kornia.geometry.conversions.axis_angle_to_rotation_matrix (0.7.0) for v [N, 3]: Rodrigues with the axis
    normalised as v / (|v| + 1e-6), first-order Taylor form when |v|^2 <= 1e-6."""
    theta2 = (v * v).sum(dim=1, keepdim=True)
    theta = theta2.sqrt()
    w = v / (theta + 1e-6)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = theta.cos(), theta.sin()
    normal = torch.cat([
        c + wx * wx * (1 - c), wx * wy * (1 - c) - wz * s, wy * s + wx * wz * (1 - c),
        wz * s + wx * wy * (1 - c), c + wy * wy * (1 - c), -wx * s + wy * wz * (1 - c),
        -wy * s + wx * wz * (1 - c), wx * s + wy * wz * (1 - c), c + wz * wz * (1 - c)], dim=1).view(-1, 3, 3)
    rx, ry, rz = v[:, 0:1], v[:, 1:2], v[:, 2:3]
    one = torch.ones_like(rx)
    taylor = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)
    big = (theta2 > 1e-6).view(-1, 1, 1).to(v.dtype)
    return big * normal + (1 - big) * taylor


def make_Rt(roll: float = 0.0, pitch: float = 0.0, yaw: float = 0.0, x: float = 0.0, y: float = 0.0,
            z: float = 0.0, device="cpu"):
    """Extrinsics of the virtual camera: R = Rz(yaw) Ry(pitch) Rx(roll) [1,3,3], t [1,3]; render.py:10-29."""
    def rot(axis: int, angle: float) -> torch.Tensor:
        v = torch.zeros(1, 3, device=device)
        v[0, axis] = angle
        return _axis_angle_to_rotation_matrix(v)
    R = rot(2, yaw) @ rot(1, pitch) @ rot(0, roll)
    t = torch.tensor([[x, y, z]], device=device)
    return R, t


def _cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise L.R2dmError(f"{what}: expected a CUDA tensor (r2dm_b200 has no CPU path)")
    return L.f32c(t)


def _one_matrix(R, shape, what):
    """The reference broadcasts R / t over the batch; its callers pass one camera ([1,3,3] / [1,3])."""
    R = R.reshape(-1, *shape)
    if R.shape[0] != 1:
        raise NotImplementedError(f"{what}: one camera per call (got a batch of {R.shape[0]})")
    return R[0]


@torch.no_grad()
def render_point_clouds(points: torch.Tensor, colors: torch.Tensor | None = None, size: int = 800,
                        R: torch.Tensor | None = None, t: torch.Tensor | None = None,
                        focal_length: float = 1.0) -> torch.Tensor:
    """points [B,N,3] (+ colors [B,N,3]) -> [B,3,size,size] soft z-buffered splat; render.py:32-80."""
    assert points.dim() == 3 and points.shape[-1] == 3, f"expected (B,N,3), but got {tuple(points.shape)}"
    p = _cuda_f32(points, "render_point_clouds")
    B, N, _ = p.shape
    c = None
    if colors is not None:
        c = _cuda_f32(colors.expand(B, N, 3), "render_point_clouds")
    Rm = tm = None
    if R is not None:
        assert R.shape[-2:] == (3, 3)
        Rm = L.f32c(_one_matrix(R, (3, 3), "render_point_clouds").to(p.device))
    if t is not None:
        assert t.shape[-1:] == (3,)
        tm = L.f32c(_one_matrix(t, (3,), "render_point_clouds").to(p.device))
    out = torch.empty(B, 3, size, size, device=p.device, dtype=torch.float32)
    acc = torch.empty(B, size * size, 4, device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        L.check(L.lib().r2dm_render_point_clouds(L.ptr(p), L.ptr(c), L.ptr(Rm), L.ptr(tm), L.ptr(acc), L.ptr(out),
                                                 B, N, int(size), float(focal_length), L.stream_ptr()),
                "r2dm_render_point_clouds")
    return out


@torch.no_grad()
def bilinear_rasterizer(coords: torch.Tensor, values: torch.Tensor, out_shape) -> torch.Tensor:
    """coords [B,N,2] (row, column), values [B,N,C] -> [B,C,H,W]; render.py:83-142."""
    B, N, C = values.shape
    H, W = out_shape
    assert coords.shape == (B, N, 2)
    co, va = _cuda_f32(coords, "bilinear_rasterizer"), _cuda_f32(values, "bilinear_rasterizer")
    out = torch.empty(B, C, H, W, device=co.device, dtype=torch.float32)
    with torch.cuda.device(co.device):
        L.check(L.lib().r2dm_bilinear_rasterize(L.ptr(co), L.ptr(va), L.ptr(out), B, N, C, int(H), int(W),
                                                L.stream_ptr()), "r2dm_bilinear_rasterize")
    return out


@torch.no_grad()
def estimate_surface_normal(points: torch.Tensor, d: int = 2, mode: str = "closest") -> torch.Tensor:
    """points [B,3,H,W] -> unit normals [B,3,H,W]; render.py:145-234."""
    assert points.dim() == 4, f"expected (B,3,H,W), but got {points.shape}"
    B, C, H, W = points.shape
    assert C == 3, f"expected C==3, but got {C}"
    if mode not in ("closest", "mean"):
        raise NotImplementedError(mode)
    p = _cuda_f32(points, "estimate_surface_normal")
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        L.check(L.lib().r2dm_surface_normal(L.ptr(p), L.ptr(out), B, H, W, int(d), 0 if mode == "closest" else 1,
                                            L.stream_ptr()), "r2dm_surface_normal")
    return out


@torch.no_grad()
def colorize(tensor: torch.Tensor, cmap_fn=None) -> torch.Tensor:
    """[B,1,H,W] or [B,H,W] in [0,1] -> uint8 [B,3,H,W] through a 256-entry colour map; render.py:237-246.
    `cmap_fn` is a matplotlib colormap (callable on an array of 256 positions) or a [256, >=3] table; the
    reference's default `cm.turbo` needs matplotlib."""
    if cmap_fn is None:
        try:
            import matplotlib.cm as cm
        except ImportError as e:  # pragma: no cover - depends on the image
            raise L.R2dmError("colorize: matplotlib is not installed; pass cmap_fn (callable or [256,3] table)") from e
        cmap_fn = cm.turbo
    table = cmap_fn(np.linspace(0, 1, 256)) if callable(cmap_fn) else cmap_fn
    table = torch.as_tensor(np.asarray(table)[:, :3]).to(tensor)
    tensor = tensor.squeeze(1) if tensor.ndim == 4 else tensor
    ids = (tensor * 256).clamp(0, 255).long()
    out = F.embedding(ids, table).permute(0, 3, 1, 2)
    return out.mul(255).clamp(0, 255).byte()


@torch.no_grad()
def point_clouds_to_histograms(point_clouds: torch.Tensor, field_size: float = 160.0, bins: int = 100,
                               min_depth: float = 3.0, max_depth: float = 70.0) -> torch.Tensor:
    """Batched metrics/bev.py:5-24: point_clouds [B,N,3] -> BEV occupancy counts [B,bins,bins] (fp32, on the
    input's device; the reference moves each cloud to the CPU for torch.histogramdd)."""
    assert point_clouds.ndim == 3 and point_clouds.shape[-1] == 3, "must be (B, N, 3)"
    assert bins % 2 == 0
    p = _cuda_f32(point_clouds, "point_cloud_to_histogram")
    B, N, _ = p.shape
    bound = field_size / 2
    edges = torch.linspace(-bound, bound, bins + 1, dtype=torch.float32).to(p.device)   # histogramdd's bin edges
    hist = torch.empty(B, bins, bins, device=p.device, dtype=torch.float32)
    counts = torch.empty(B, bins, bins, device=p.device, dtype=torch.int32)
    with torch.cuda.device(p.device):
        L.check(L.lib().r2dm_bev_histogram(L.ptr(p), L.ptr(edges), L.ptr(counts), L.ptr(hist), B, N, int(bins),
                                           float(min_depth), float(max_depth), L.stream_ptr()), "r2dm_bev_histogram")
    return hist


def point_cloud_to_histogram(point_cloud: torch.Tensor, field_size: float = 160.0, bins: int = 100,
                             min_depth: float = 3.0, max_depth: float = 70.0) -> torch.Tensor:
    """metrics/bev.py:5-24: point_cloud [N,3] -> [bins,bins]."""
    assert point_cloud.ndim == 2, "must be (N, 3)"
    return point_clouds_to_histograms(point_cloud[None], field_size, bins, min_depth, max_depth)[0]


__all__ = ["make_Rt", "render_point_clouds", "bilinear_rasterizer", "estimate_surface_normal", "colorize",
           "point_cloud_to_histogram", "point_clouds_to_histograms"]
