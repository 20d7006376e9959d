"""Input-independent coordinate encodings (host-side mirror of models/encoding.py of the reference).

These produce a per-model constant `[1, extra_ch, H, W]` table that is computed once in fp32 when the
weights are uploaded and then lives, pre-packed, inside the network's input staging tensor (the
reference recomputes and re-concatenates it every forward, efficient_unet.py:278-281).  Keeping it
fp32 also avoids the reference's bf16-autocast precision loss on the 512*pi phases.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def generate_polar_coords(H: int, W: int, device="cpu") -> torch.Tensor:
    """phi (polar, top-down) and theta (azimuth, decreasing) grids; models/encoding.py:80-89."""
    rows = torch.arange(H, device=device) / H
    cols = torch.arange(W, device=device) / W
    phi = (0.5 - rows) * torch.pi
    theta = (1 - cols) * 2 * torch.pi - torch.pi
    return torch.stack(torch.meshgrid([phi, theta], indexing="ij"))[None]


_SH_COEF = (
    0.28209479177387814, 0.4886025119029199, 1.0925484305920792, 0.9461746957575601, 0.31539156525251999,
    0.5462742152960396, 0.5900435899266435, 2.890611442640554, 0.4570457994644658, 0.3731763325901154,
    1.445305721320277, 2.5033429417967046, 1.7701307697799304, 0.6690465435572892, 0.10578554691520431,
    0.47308734787878004, 0.6258357354491761,
)


class SphericalHarmonics(nn.Module):
    """Real SH basis up to `levels` (levels^2 channels) of the ray directions; encoding.py:10-77,92-114."""

    def __init__(self, levels: int = 4) -> None:
        super().__init__()
        assert 1 <= levels <= 5
        self.levels = levels
        self.extra_ch = levels ** 2

    def forward(self, coords: torch.Tensor) -> torch.Tensor:
        phi, theta = coords[:, 0], coords[:, 1]
        x, y, z = theta.cos() * phi.cos(), -theta.sin() * phi.cos(), phi.sin()
        xx, yy, zz = x * x, y * y, z * z
        k = _SH_COEF
        bands = [[torch.full_like(x, k[0])]]
        bands.append([k[1] * y, k[1] * z, k[1] * x])
        bands.append([k[2] * x * y, k[2] * y * z, k[3] * zz - k[4], k[2] * x * z, k[5] * (xx - yy)])
        bands.append([k[6] * y * (3 * xx - yy), k[7] * x * y * z, k[8] * y * (5 * zz - 1),
                      k[9] * z * (5 * zz - 3), k[8] * x * (5 * zz - 1), k[10] * z * (xx - yy),
                      k[6] * x * (xx - 3 * yy)])
        bands.append([k[11] * x * y * (xx - yy), k[12] * y * z * (3 * xx - yy), k[3] * x * y * (7 * zz - 1),
                      k[13] * y * z * (7 * zz - 3), k[14] * (35 * zz * zz - 30 * zz + 3),
                      k[13] * x * z * (7 * zz - 3), k[15] * (xx - yy) * (7 * zz - 1),
                      k[12] * x * z * (xx - 3 * yy), k[16] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))])
        comps = [c for band in bands[: self.levels] for c in band]
        return torch.stack(comps, dim=1)

    def extra_repr(self):
        return f"levels={self.levels}"


class FourierFeatures(nn.Module):
    """[sin || cos] of 2^k multiples of the two angles; encoding.py:120-146.  Buffers `freqs`
    ([F, 2, 1, 1]) and `phase` ([F]) keep the reference's state-dict keys."""

    def __init__(self, resolution):
        super().__init__()
        self.resolution = tuple(resolution)
        self.L_h = int(math.ceil(math.log2(self.resolution[0])))
        self.L_w = int(math.ceil(math.log2(self.resolution[1])))
        n = self.L_h + self.L_w
        freqs = torch.zeros(n, 2)
        freqs[: self.L_h, 0] = 2.0 ** torch.arange(self.L_h)
        freqs[self.L_h:, 1] = 2.0 ** torch.arange(self.L_w)
        self.register_buffer("freqs", freqs[..., None, None])
        self.register_buffer("phase", torch.zeros(n))
        self.extra_ch = 2 * n

    def forward(self, coords: torch.Tensor) -> torch.Tensor:
        f = self.freqs[:, :, 0, 0].to(coords)
        ang = f[None, :, 0, None, None] * coords[:, None, 0] + f[None, :, 1, None, None] * coords[:, None, 1]
        ang = ang + self.phase.to(coords)[None, :, None, None]
        return torch.cat([ang.sin(), ang.cos()], dim=1)

    def extra_repr(self):
        return f"shape={self.resolution}, num_freqs={self.extra_ch}, L=({self.L_h}, {self.L_w})"
