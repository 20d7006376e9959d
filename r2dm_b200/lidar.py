"""LiDARUtility — host-side mirror of utils/lidar.py:9-120 of the reference, plus the fused
post-processing epilogue (`postprocess`) that replaces the five-op chain of
sample_and_save.py:52-57 with one CUDA kernel (`r2dm_lidar_postprocess`)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L

_FORMATS = ("log_depth", "inverse_depth", "depth")


def get_hdl64e_linear_ray_angles(H: int = 64, W: int = 2048, device="cpu") -> torch.Tensor:
    """Velodyne HDL-64E: elevation +3..-25 deg (top row first), azimuth 180..-180 deg; lidar.py:9-20."""
    elevation = (1 - torch.arange(H, device=device) / H) * 28 - 25
    azimuth = (1 - torch.arange(W, device=device) / W) * 360 - 180
    grid = torch.meshgrid([elevation, azimuth], indexing="ij")
    return torch.stack(grid)[None].deg2rad()


class LiDARUtility(nn.Module):
    def __init__(self, resolution, depth_format: str, min_depth: float, max_depth: float,
                 ray_angles: torch.Tensor = None):
        super().__init__()
        assert depth_format in _FORMATS
        self.resolution = tuple(resolution)
        self.depth_format = depth_format
        self.min_depth = min_depth
        self.max_depth = max_depth
        if ray_angles is None:
            ray_angles = get_hdl64e_linear_ray_angles(*self.resolution)
        else:
            assert ray_angles.ndim == 4 and ray_angles.shape[1] == 2
        ray_angles = F.interpolate(ray_angles, size=self.resolution, mode="nearest-exact")
        self.register_buffer("ray_angles", ray_angles.float())

    @staticmethod
    def denormalize(x: torch.Tensor) -> torch.Tensor:
        """[-1, +1] -> [0, 1]"""
        return (x + 1) / 2

    @staticmethod
    def normalize(x: torch.Tensor) -> torch.Tensor:
        """[0, 1] -> [-1, +1]"""
        return x * 2 - 1

    def get_mask(self, metric):
        return ((metric > self.min_depth) & (metric < self.max_depth)).float()

    @torch.no_grad()
    def to_xyz(self, metric: torch.Tensor) -> torch.Tensor:
        assert metric.dim() == 4
        phi, theta = self.ray_angles[:, [0]], self.ray_angles[:, [1]]
        planar = metric * phi.cos()
        xyz = torch.cat((planar * theta.cos(), planar * theta.sin(), metric * phi.sin()), dim=1)
        return xyz * self.get_mask(metric)

    @torch.no_grad()
    def convert_depth(self, metric, mask=None, depth_format: str = None) -> torch.Tensor:
        """metric depth [0, max_depth] -> normalized [0, 1]"""
        depth_format = self.depth_format if depth_format is None else depth_format
        mask = self.get_mask(metric) if mask is None else mask
        if depth_format == "log_depth":
            normalized = torch.log2(metric + 1) / np.log2(self.max_depth + 1)
        elif depth_format == "inverse_depth":
            normalized = self.min_depth / metric.add(1e-8)
        elif depth_format == "depth":
            normalized = metric.div(self.max_depth)
        else:
            raise ValueError
        return normalized.clamp(0, 1) * mask

    @torch.no_grad()
    def revert_depth(self, normalized, image_format: str = None) -> torch.Tensor:
        """normalized [0, 1] -> metric depth [0, max_depth]"""
        image_format = self.depth_format if image_format is None else image_format
        if image_format == "log_depth":
            metric = torch.exp2(normalized * np.log2(self.max_depth + 1)) - 1
        elif image_format == "inverse_depth":
            metric = self.min_depth / normalized.add(1e-8)
        elif image_format == "depth":
            metric = normalized.mul(self.max_depth)
        else:
            raise ValueError
        return metric * self.get_mask(metric)

    @torch.no_grad()
    def postprocess(self, sample: torch.Tensor) -> torch.Tensor:
        """Fused denormalize -> revert_depth -> to_xyz -> cat: [B,2,H,W] in [-1,1] -> [B,5,H,W]
        (depth, x, y, z, reflectance), the per-sample format sample_and_save.py:52-57 stores."""
        if not sample.is_cuda:
            raise L.R2dmError("LiDARUtility.postprocess runs on CUDA only")
        B, Cc, H, W = sample.shape
        assert Cc == 2 and (H, W) == tuple(self.ray_angles.shape[-2:])
        s = L.f32c(sample)
        ang = L.f32c(self.ray_angles.to(s.device)[0])
        out = torch.empty(B, 5, H, W, device=s.device, dtype=torch.float32)
        with torch.cuda.device(s.device):
            L.check(L.lib().r2dm_lidar_postprocess(L.ptr(s), L.ptr(ang), L.ptr(out), B, H, W,
                                                   _FORMATS.index(self.depth_format), float(self.min_depth),
                                                   float(self.max_depth), L.stream_ptr()), "r2dm_lidar_postprocess")
        return out
