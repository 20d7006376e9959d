"""Synthetic (random) weights for benchmarking and smoke tests when no checkpoint is available.

The reference zero-initialises every `conv2`, `out_proj` and `out_conv` (efficient_unet.py:39,84,267), so a
freshly constructed network outputs zeros; for timing / sanity runs the bench fills ALL parameters with
scaled random values instead.  This is product-side plumbing (no oracle / test imports): it writes
through the module's own parameters, so the normal weight-packing path is exercised.
"""
from __future__ import annotations

import math

import torch

from .config import Config
from .inference import build_model
from .lidar import LiDARUtility


@torch.no_grad()
def randomize_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Variance-preserving random fill of every parameter (weights ~ N(0, 1/fan_in), biases ~ 0.02 N(0,1),
    normalisation gains ~ 1 + 0.1 N(0,1)); buffers (FIR windows, residual scales, coordinates) untouched."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if name.endswith("norm1.weight") or name.endswith("norm.weight"):
            v = 1 + 0.1 * torch.randn(p.shape, generator=g)
        elif name.endswith("weight") and p.dim() >= 2:
            fan_in = p[0].numel()
            v = torch.randn(p.shape, generator=g) / math.sqrt(fan_in)
        else:
            v = 0.02 * torch.randn(p.shape, generator=g)
        p.copy_(v.to(p.dtype))
    if hasattr(module, "mark_weights_changed"):
        module.mark_weights_changed()
    for m in module.modules():
        if m is not module and hasattr(m, "mark_weights_changed"):
            m.mark_weights_changed()
    return module


def synthetic_model(cfg: Config = None, device="cuda", precision: str = "bf16", seed: int = 0):
    """(ddpm, lidar_utils, cfg) like `setup_model`, with random weights; `cfg` defaults to config H
    (utils/option.py defaults = r2dm-h-kitti360-300k: 2x64x1024, base 64, multipliers 1-2-4-8, 3 blocks)."""
    cfg = cfg or Config()
    ddpm = build_model(cfg, precision)
    randomize_(ddpm, seed)
    ddpm.eval()
    ddpm.to(device)
    lidar_utils = LiDARUtility(resolution=cfg.data.resolution, depth_format=cfg.data.depth_format,
                               min_depth=cfg.data.min_depth, max_depth=cfg.data.max_depth,
                               ray_angles=ddpm.model.coords)
    lidar_utils.eval()
    lidar_utils.to(device)
    return ddpm, lidar_utils, cfg
