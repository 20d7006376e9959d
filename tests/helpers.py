"""Shared test helpers: the two model configs used by the golden fixtures and the teacher-forced
noise draw order of the reference samplers (models/diffusion/base.py:71-94: one
`torch.randn(C, H, W, generator=g_i)` per sample per draw)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.r2dm_oracle import UNetCfg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# config H = what setup_model builds for r2dm-h-kitti360-300k (utils/option.py:9-19,58-69)
H_CFG = UNetCfg(in_channels=2, resolution=(64, 1024), base_channels=64,
                channel_multiplier=(1, 2, 4, 8), num_residual_blocks=(3, 3, 3, 3))
# same widths, quarter height, fewer residual blocks: cheap enough for CPU trajectories
SMALL_CFG = UNetCfg(in_channels=2, resolution=(16, 1024), base_channels=64,
                    channel_multiplier=(1, 2, 4, 8), num_residual_blocks=(2, 1, 1, 2))


def draw_noise(seeds, n_draws, cfg, device="cpu"):
    """n_draws successive batches [B, C, H, W]; sample i always draws from its own generator."""
    gens = [torch.Generator().manual_seed(s) for s in seeds]
    shape = (cfg.in_channels, *cfg.resolution)
    return [torch.stack([torch.randn(*shape, generator=g) for g in gens]).to(device)
            for _ in range(n_draws)]


def repaint_masks(B, cfg):
    """Corruption masks in the spirit of completion_demo.py:81-87 (1 = known)."""
    H, W = cfg.resolution
    g = torch.Generator().manual_seed(3)
    mask = torch.zeros(B, cfg.in_channels, H, W)
    mask[0, :, ::4] = 1
    if B > 1:
        mask[1] = (torch.rand(H, W, generator=g) < 0.1).float()
    if B > 2:
        mask[2] = 1
    if B > 3:
        mask[3, :] = (torch.rand(H, 1, generator=g) < 0.5).float()
    return mask


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()
