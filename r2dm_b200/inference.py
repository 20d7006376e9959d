"""setup_model / setup_rng — host-side mirror of utils/inference.py:20-114 of the reference."""
from __future__ import annotations

from pathlib import Path

import torch

from .config import Config
from .diffusion import (ContinuousTimeGaussianDiffusion, DiscreteTimeGaussianDiffusion,
                        GaussianDiffusion)
from .lidar import LiDARUtility
from .unet import EfficientUNet


def count_parameters(model: torch.nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def build_model(cfg: Config, precision: str = "fp32") -> GaussianDiffusion:
    """Construct the (randomly initialised) diffusion model a checkpoint's cfg describes."""
    in_channels = int(bool(cfg.data.train_depth)) + int(bool(cfg.data.train_reflectance))
    if cfg.model.architecture != "efficient_unet":
        raise ValueError(f"Unknown / unsupported architecture: {cfg.model.architecture} "
                         "(only the EfficientUNet sampling path is implemented)")
    model = EfficientUNet(
        in_channels=in_channels,
        resolution=cfg.data.resolution,
        base_channels=cfg.model.base_channels,
        temb_channels=cfg.model.temb_channels,
        channel_multiplier=cfg.model.channel_multiplier,
        num_residual_blocks=cfg.model.num_residual_blocks,
        gn_num_groups=cfg.model.gn_num_groups,
        gn_eps=cfg.model.gn_eps,
        attn_num_heads=cfg.model.attn_num_heads,
        coords_encoding=cfg.model.coords_encoding,
        ring=True,
        precision=precision,
    )
    if cfg.diffusion.timestep_type == "discrete":
        return DiscreteTimeGaussianDiffusion(
            model=model, loss_type=cfg.diffusion.loss_type,
            num_training_steps=cfg.diffusion.num_training_steps,
            prediction_type=cfg.diffusion.prediction_type, noise_schedule=cfg.diffusion.noise_schedule)
    if cfg.diffusion.timestep_type == "continuous":
        return ContinuousTimeGaussianDiffusion(
            model=model, loss_type=cfg.diffusion.loss_type,
            prediction_type=cfg.diffusion.prediction_type, noise_schedule=cfg.diffusion.noise_schedule)
    raise ValueError(f"Unknown: {cfg.diffusion.timestep_type}")


def setup_model(ckpt, device="cpu", ema: bool = True, show_info: bool = True, compile: bool = False,
                precision: str = "fp32"):
    """(ddpm, lidar_utils, cfg) from a reference checkpoint (path or dict with keys `cfg`, `weights`,
    `ema_weights`, `global_step`; train.py:294-304).  `compile` is accepted for signature
    compatibility and ignored: the network already runs as hand-written kernels + CUDA graph.
    `precision`: "fp32" (tf32 tensor cores, the reference's GPU default) or "bf16"; with the default,
    a surrounding bf16 `torch.autocast` selects the bf16 engine (fp16 autocast keeps fp32, see
    EfficientUNet._active_precision)."""
    if isinstance(ckpt, (str, Path)):
        ckpt = torch.load(ckpt, map_location="cpu")
    cfg = ckpt["cfg"] if isinstance(ckpt["cfg"], Config) else Config(**ckpt["cfg"])
    ddpm = build_model(cfg, precision)
    if precision != "fp32":
        ddpm.model.set_precision(precision)     # explicit choice: not overridden by autocast
    state_dict = ckpt["ema_weights"] if ema else ckpt["weights"]
    ddpm.load_state_dict(state_dict)
    ddpm.eval()
    ddpm.to(device)
    lidar_utils = LiDARUtility(
        resolution=cfg.data.resolution, depth_format=cfg.data.depth_format,
        min_depth=cfg.data.min_depth, max_depth=cfg.data.max_depth, ray_angles=ddpm.model.coords)
    lidar_utils.eval()
    lidar_utils.to(device)
    if show_info:
        print(*[
            f"resolution: {ddpm.model.resolution}",
            f"model: {ddpm.model.__class__.__name__}",
            f"ddpm: {ddpm.__class__.__name__}",
            f'#steps:  {ckpt.get("global_step", 0):,}',
            f"#params: {count_parameters(ddpm):,}",
        ], sep="\n")
    return ddpm, lidar_utils, cfg


def setup_rng(seeds, device):
    return [torch.Generator(device=device).manual_seed(int(i)) for i in seeds]
