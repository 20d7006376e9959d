"""Build an r2dm_b200 diffusion model for an oracle UNetCfg + state dict (test helper)."""
import torch

import r2dm_b200 as R


def make_cfg(ucfg, timestep_type="continuous", schedule="cosine", objective="eps", num_training_steps=None):
    cfg = R.Config()
    cfg.data.resolution = tuple(ucfg.resolution)
    assert ucfg.in_channels in (1, 2)
    cfg.data.train_reflectance = ucfg.in_channels == 2     # utils/inference.py:31-36: depth [+ reflectance]
    cfg.model.base_channels = ucfg.base_channels
    cfg.model.temb_channels = ucfg.temb_channels
    cfg.model.channel_multiplier = tuple(ucfg.channel_multiplier)
    cfg.model.num_residual_blocks = tuple(ucfg.num_residual_blocks)
    cfg.model.gn_num_groups = ucfg.gn_num_groups
    cfg.model.gn_eps = ucfg.gn_eps
    cfg.model.attn_num_heads = ucfg.attn_num_heads
    cfg.model.coords_encoding = ucfg.coords_encoding
    cfg.diffusion.timestep_type = timestep_type
    cfg.diffusion.noise_schedule = schedule
    cfg.diffusion.prediction_type = objective
    cfg.diffusion.num_training_steps = num_training_steps
    return cfg


def make_ddpm(ucfg, sd, precision="fp32", device="cuda", **kw):
    """Goes through the public checkpoint path: setup_model(ckpt dict)."""
    cfg = make_cfg(ucfg, **kw)
    probe = R.build_model(cfg)
    full = {k: v for k, v in probe.state_dict().items() if not k.startswith("model.")}
    full.update({"model." + k: v for k, v in sd.items()})
    ckpt = {"cfg": cfg.to_dict(), "weights": full, "ema_weights": full, "global_step": 0}
    ddpm, lidar_utils, _ = R.setup_model(ckpt, device=device, show_info=False, precision=precision)
    return ddpm
