"""r2dm_b200 — B200-native implementation of R2DM's reverse-diffusion sampling hot path.

Keeps the reference's API surface (`setup_model`, `setup_rng`, `ddpm.sample`, `ddpm.repaint`,
`lidar_utils.*`); the U-Net forward and the sampler arithmetic run as hand-written sm_100a CUDA
kernels behind the C ABI in include/r2dm_b200.h.  No CPU fallback exists.
"""
from .config import Config, DataConfig, DiffusionConfig, ModelConfig, TrainingConfig  # noqa: F401
from .diffusion import (ContinuousTimeGaussianDiffusion, DiscreteTimeGaussianDiffusion,  # noqa: F401
                        GaussianDiffusion)
from .inference import build_model, setup_model, setup_rng  # noqa: F401
from .lidar import LiDARUtility, get_hdl64e_linear_ray_angles  # noqa: F401
from .synthetic import randomize_, synthetic_model  # noqa: F401
from .unet import EfficientUNet  # noqa: F401

__all__ = [
    "Config", "DataConfig", "DiffusionConfig", "ModelConfig", "TrainingConfig",
    "GaussianDiffusion", "ContinuousTimeGaussianDiffusion", "DiscreteTimeGaussianDiffusion",
    "EfficientUNet", "LiDARUtility", "get_hdl64e_linear_ray_angles",
    "build_model", "setup_model", "setup_rng", "randomize_", "synthetic_model",
]
