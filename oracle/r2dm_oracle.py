"""CPU oracle for the R2DM reverse-diffusion sampling path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (fp32/fp64, CPU) functional restatement of the reference algorithm on the hot path:
EfficientUNet.forward, the continuous/discrete-time samplers and LiDARUtility.  It exists so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg can check / time the CUDA path
on a GPU box where `/root/reference` does not exist.  The product package `r2dm_b200` must never
import this module.

Parity pinning: the reference ships no tests / golden vectors, so this oracle is
pinned against outputs of the reference implementation itself, imported read-only from
/root/reference in the build container by `tests/golden/make_golden.py`; the resulting fixtures
live in `tests/golden/*.pt` and `tests/test_oracle_golden.py` re-checks the oracle against them.

Every function cites the reference file:line (relative to /root/reference) it restates.  The network
is expressed functionally over a flat state-dict (the checkpoint schema of SURVEY.md appendix B),
with the FIR resamplers in closed form rather than pad / zero-insert / depthwise-conv.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------- config
@dataclass
class UNetCfg:
    """Constructor arguments of EfficientUNet (models/efficient_unet.py:194-209) as used by
    utils/inference.py:38-51 (ring=True always)."""

    in_channels: int = 2
    resolution: Tuple[int, int] = (64, 1024)
    base_channels: int = 64
    temb_channels: Optional[int] = None
    channel_multiplier: Tuple[int, int, int, int] = (1, 2, 4, 8)
    num_residual_blocks: Tuple[int, int, int, int] = (3, 3, 3, 3)
    gn_num_groups: int = 8
    gn_eps: float = 1e-6
    attn_num_heads: int = 8
    coords_encoding: Optional[str] = "fourier_features"

    @property
    def temb(self) -> int:
        return self.base_channels * 4 if self.temb_channels is None else self.temb_channels

    @property
    def channels(self) -> List[int]:
        return [self.base_channels] + [self.base_channels * m for m in self.channel_multiplier]

    @property
    def extra_channels(self) -> int:
        if self.coords_encoding == "fourier_features":
            lh = int(math.ceil(math.log2(self.resolution[0])))
            lw = int(math.ceil(math.log2(self.resolution[1])))
            return 2 * (lh + lw)
        if self.coords_encoding == "spherical_harmonics":
            return 25
        if self.coords_encoding == "polar_coordinates":
            return 2
        return 0


def block_table(cfg: UNetCfg):
    """(name, cin, cout, nres, down, up, attn) for the eight Blocks, efficient_unet.py:254-265."""
    C, N = cfg.channels, cfg.num_residual_blocks
    return [
        ("d_block1", C[0], C[1], N[0], 1, 1, False),
        ("d_block2", C[1], C[2], N[1], 2, 1, False),
        ("d_block3", C[2], C[3], N[2], 2, 1, False),
        ("d_block4", C[3], C[4], N[3], 2, 1, True),
        ("u_block4", C[4], C[3], N[3], 1, 2, True),
        ("u_block3", C[3] + C[3], C[2], N[2], 1, 2, False),
        ("u_block2", C[2] + C[2], C[1], N[1], 1, 2, False),
        ("u_block1", C[1] + C[1], C[0], N[0], 1, 1, False),
    ]


def state_dict_schema(cfg: UNetCfg) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of EfficientUNet.state_dict() (checked against the
    reference module in tests/golden/make_golden.py)."""
    H, W = cfg.resolution
    s: Dict[str, Tuple[int, ...]] = {}
    s["coords"] = (1, 2, H, W)
    if cfg.coords_encoding == "fourier_features":
        nf = cfg.extra_channels // 2
        s["coords_encoding.freqs"] = (nf, 2, 1, 1)
        s["coords_encoding.phase"] = (nf,)
    T, C0 = cfg.temb, cfg.base_channels
    s["time_embedding.1.weight"] = (T, C0)
    s["time_embedding.1.bias"] = (T,)
    s["time_embedding.3.weight"] = (T, T)
    s["time_embedding.3.bias"] = (T,)
    cin0 = cfg.in_channels + cfg.extra_channels
    s["in_conv.weight"] = (C0, cin0, 3, 3)
    s["in_conv.bias"] = (C0,)
    for name, cin, cout, nres, down, up, attn in block_table(cfg):
        if down > 1:
            s[f"{name}.downsample.0.weight"] = (cout, cin, 3, 3)
            s[f"{name}.downsample.0.bias"] = (cout,)
            s[f"{name}.downsample.1.kernel"] = (4,)
        for i in range(nres):
            ci = cout if (i != 0 or down > 1) else cin
            p = f"{name}.residual_blocks.{i}"
            s[f"{p}.scale"] = ()
            s[f"{p}.norm1.weight"] = (ci,)
            s[f"{p}.norm1.bias"] = (ci,)
            s[f"{p}.conv1.weight"] = (cout, ci, 3, 3)
            s[f"{p}.conv1.bias"] = (cout,)
            s[f"{p}.norm2.proj.1.weight"] = (2 * cout, T)
            s[f"{p}.norm2.proj.1.bias"] = (2 * cout,)
            s[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
            s[f"{p}.conv2.bias"] = (cout,)
            if ci != cout:
                s[f"{p}.skip.weight"] = (cout, ci, 1, 1)
                s[f"{p}.skip.bias"] = (cout,)
        if attn:
            p = f"{name}.self_attn_block"
            s[f"{p}.scale"] = ()
            s[f"{p}.norm.weight"] = (cout,)
            s[f"{p}.norm.bias"] = (cout,)
            s[f"{p}.attn.in_proj_weight"] = (3 * cout, cout)
            s[f"{p}.attn.in_proj_bias"] = (3 * cout,)
            s[f"{p}.attn.out_proj.weight"] = (cout, cout)
            s[f"{p}.attn.out_proj.bias"] = (cout,)
        if up > 1:
            s[f"{name}.upsample.0.kernel"] = (4,)
            s[f"{name}.upsample.1.weight"] = (cout, cout, 3, 3)
            s[f"{name}.upsample.1.bias"] = (cout,)
    s["out_conv.weight"] = (cfg.in_channels, C0, 3, 3)
    s["out_conv.bias"] = (cfg.in_channels,)
    return s


def hdl64e_linear_ray_angles(H: int = 64, W: int = 2048) -> Tensor:
    """utils/lidar.py:9-20 — elevation in [-25, 3] deg top-down, azimuth in (-180, 180] deg."""
    elev = (1 - torch.arange(H) / H) * (3 - (-25)) + (-25)
    azim = (1 - torch.arange(W) / W) * (180 - (-180)) + (-180)
    e, a = torch.meshgrid([elev, azim], indexing="ij")
    return torch.stack([e, a])[None].deg2rad()


def polar_coords(H: int, W: int) -> Tensor:
    """models/encoding.py:80-89."""
    phi = (0.5 - torch.arange(H) / H) * torch.pi
    theta = (1 - torch.arange(W) / W) * 2 * torch.pi - torch.pi
    p, t = torch.meshgrid([phi, theta], indexing="ij")
    return torch.stack([p, t])[None]


def random_state_dict(cfg: UNetCfg, seed: int, std: float = 0.05) -> Dict[str, Tensor]:
    """Deterministic synthetic weights for parity tests / the benchmark.  Independent of nn.Module
    init order; every tensor that the reference zero-initialises (conv2, out_proj, out_conv:
    efficient_unet.py:39,84,267) is drawn non-zero so that outputs are not identically 0
    (SURVEY.md appendix C.1).  Conv / linear weights ~ N(0, 1/fan_in) so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    H, W = cfg.resolution
    sd: Dict[str, Tensor] = {}
    for name, shape in state_dict_schema(cfg).items():
        if name == "coords":
            sd[name] = hdl64e_linear_ray_angles(H, W).float()
        elif name.endswith("coords_encoding.freqs"):
            sd[name] = fourier_freqs(cfg.resolution)
        elif name.endswith("coords_encoding.phase"):
            sd[name] = torch.zeros(shape)
        elif name.endswith(".scale"):
            sd[name] = torch.tensor(1 / math.sqrt(2)).float()
        elif name.endswith("downsample.1.kernel"):
            sd[name] = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8
        elif name.endswith("upsample.0.kernel"):
            sd[name] = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 4
        elif name.endswith("norm1.weight") or name.endswith("norm.weight"):
            sd[name] = 1 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = std * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return sd


# --------------------------------------------------------------------------------------- ops
def sinusoidal_embedding(t: Tensor, channels: int, max_period: float = 10_000) -> Tensor:
    """models/ops.py:14-26 — [sin(t f_k) || cos(t f_k)], f_k = exp(-ln(P) k / (C/2 - 1))."""
    assert t.dim() == 1
    half = channels // 2
    f = torch.exp(-math.log(max_period) / (half - 1) * torch.arange(half, device=t.device))
    a = t[:, None] * f[None, :]
    return torch.cat([a.sin(), a.cos()], dim=-1).to(t)


def fourier_freqs(resolution: Sequence[int]) -> Tensor:
    """models/encoding.py:127-138 — 2^k on phi for k<L_h, 2^k on theta for k<L_w."""
    lh = int(math.ceil(math.log2(resolution[0])))
    lw = int(math.ceil(math.log2(resolution[1])))
    fh = torch.cat([torch.arange(lh).exp2(), torch.zeros(lw)])
    fw = torch.cat([torch.zeros(lh), torch.arange(lw).exp2()])
    return torch.stack([fh, fw], dim=-1)[..., None, None].float()


def fourier_features(coords: Tensor, freqs: Tensor, phase: Tensor) -> Tensor:
    """models/encoding.py:141-146 — a 1x1 conv of the two angles then [sin || cos]."""
    ang = torch.einsum("bchw,fc->bfhw", coords, freqs[:, :, 0, 0]) + phase[None, :, None, None]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


def spherical_harmonics(coords: Tensor, levels: int = 5) -> Tensor:
    """models/encoding.py:10-77,98-114 (nerfstudio real SH basis, levels^2 channels)."""
    phi, theta = coords[:, 0], coords[:, 1]
    x = torch.cos(theta) * torch.cos(phi)
    y = -torch.sin(theta) * torch.cos(phi)
    z = torch.sin(phi)
    xx, yy, zz = x * x, y * y, z * z
    c = [torch.full_like(x, 0.28209479177387814)]
    if levels > 1:
        c += [0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x]
    if levels > 2:
        c += [1.0925484305920792 * x * y, 1.0925484305920792 * y * z,
              0.9461746957575601 * zz - 0.31539156525251999, 1.0925484305920792 * x * z,
              0.5462742152960396 * (xx - yy)]
    if levels > 3:
        c += [0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * x * y * z,
              0.4570457994644658 * y * (5 * zz - 1), 0.3731763325901154 * z * (5 * zz - 3),
              0.4570457994644658 * x * (5 * zz - 1), 1.445305721320277 * z * (xx - yy),
              0.5900435899266435 * x * (xx - 3 * yy)]
    if levels > 4:
        c += [2.5033429417967046 * x * y * (xx - yy), 1.7701307697799304 * y * z * (3 * xx - yy),
              0.9461746957575601 * x * y * (7 * zz - 1), 0.6690465435572892 * y * z * (7 * zz - 3),
              0.10578554691520431 * (35 * zz * zz - 30 * zz + 3),
              0.6690465435572892 * x * z * (7 * zz - 3),
              0.47308734787878004 * (xx - yy) * (7 * zz - 1),
              1.7701307697799304 * x * z * (xx - 3 * yy),
              0.6258357354491761 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(c, dim=1)


def coords_encoding(cfg: UNetCfg, sd: Dict[str, Tensor]) -> Optional[Tensor]:
    """efficient_unet.py:216-229,278-279 — the input-independent [1, extra, H, W] encoding."""
    if cfg.coords_encoding is None:
        return None
    coords = sd["coords"]
    if cfg.coords_encoding == "fourier_features":
        return fourier_features(coords, sd["coords_encoding.freqs"], sd["coords_encoding.phase"])
    if cfg.coords_encoding == "spherical_harmonics":
        return spherical_harmonics(coords, 5)
    if cfg.coords_encoding == "polar_coordinates":
        return coords
    raise ValueError(cfg.coords_encoding)


def ring_pad(x: Tensor, p: int) -> Tensor:
    """models/ops.py:39-43 — circular along azimuth (W), zeros along elevation (H)."""
    x = torch.cat([x[..., -p:], x, x[..., :p]], dim=-1)
    return F.pad(x, (0, 0, p, p))


def ring_conv3x3(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """models/ops.py:149-173 with kernel 3, padding 1, ring=True."""
    return F.conv2d(ring_pad(x, 1), w, b)


def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """models/ops.py:149-173 with kernel 1, padding 0 (efficient_unet.py:87-91)."""
    return F.conv2d(x, w, b)


def resample_down2(x: Tensor) -> Tensor:
    """models/ops.py:52-146 with down=2: y[i,j] = sum_ab w_a w_b x[2i-1+a, 2j-1+b], w=[1,3,3,1]/8."""
    w = (0.125, 0.375, 0.375, 0.125)     # exact in binary; Python scalars keep the function CUDA-graph capturable
    xp = ring_pad(x, 1)
    H, W = x.shape[-2:]
    y = 0
    for a in range(4):
        for b in range(4):
            y = y + w[a] * w[b] * xp[..., a:a + H:2, b:b + W:2]
    return y


def resample_up2(x: Tensor) -> Tensor:
    """models/ops.py:52-146 with up=2: per axis y[2i] = (x[i-1] + 3 x[i]) / 4,
    y[2i+1] = (3 x[i] + x[i+1]) / 4; W circular, H zero outside."""
    xp = ring_pad(x, 1)
    H, W = x.shape[-2:]
    c, l, r = xp[..., :, 1:W + 1], xp[..., :, 0:W], xp[..., :, 2:W + 2]
    even, odd = (l + 3 * c) / 4, (3 * c + r) / 4
    xw = torch.stack([even, odd], dim=-1).reshape(*xp.shape[:-1], 2 * W)
    c, u, d = xw[..., 1:H + 1, :], xw[..., 0:H, :], xw[..., 2:H + 2, :]
    even, odd = (u + 3 * c) / 4, (3 * c + d) / 4
    return torch.stack([even, odd], dim=-2).reshape(*x.shape[:-2], 2 * H, 2 * W)


def group_norm(x: Tensor, groups: int, eps: float, w: Optional[Tensor], b: Optional[Tensor]) -> Tensor:
    """nn.GroupNorm as used at efficient_unet.py:33,72 — biased variance over (C/G, H, W)."""
    B, C = x.shape[:2]
    xg = x.reshape(B, groups, -1)
    mean = xg.mean(dim=-1, keepdim=True)
    var = xg.var(dim=-1, unbiased=False, keepdim=True)
    y = ((xg - mean) * torch.rsqrt(var + eps)).reshape(x.shape)
    if w is not None:
        y = y * w[None, :, None, None] + b[None, :, None, None]
    return y


def adagn(x: Tensor, emb: Tensor, groups: int, eps: float, pw: Tensor, pb: Tensor) -> Tensor:
    """models/ops.py:176-200 — GN without affine, then h (1 + scale) + shift with
    [scale || shift] = Linear(SiLU(emb))."""
    h = group_norm(x, groups, eps, None, None)
    ss = F.linear(F.silu(emb), pw, pb)
    scale, shift = ss.chunk(2, dim=1)
    return h * (1 + scale[:, :, None, None]) + shift[:, :, None, None]


def residual_block(sd, p: str, x: Tensor, temb: Tensor, cfg: UNetCfg) -> Tensor:
    """efficient_unet.py:95-110."""
    G, eps = cfg.gn_num_groups, cfg.gn_eps
    h = F.silu(group_norm(x, G, eps, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"]))
    h = ring_conv3x3(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"])
    h = F.silu(adagn(h, temb, G, eps, sd[f"{p}.norm2.proj.1.weight"], sd[f"{p}.norm2.proj.1.bias"]))
    h = ring_conv3x3(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"])
    if f"{p}.skip.weight" in sd:
        x = conv1x1(x, sd[f"{p}.skip.weight"], sd[f"{p}.skip.bias"])
    return (x + h) * sd[f"{p}.scale"]


def self_attention_block(sd, p: str, x: Tensor, cfg: UNetCfg) -> Tensor:
    """efficient_unet.py:42-53 + nn.MultiheadAttention(batch_first) semantics: packed
    in_proj = [q; k; v], heads split the embedding contiguously, softmax(q k^T / sqrt(hd)) v."""
    B, C, H, W = x.shape
    nh = cfg.attn_num_heads
    hd = C // nh
    h = group_norm(x, cfg.gn_num_groups, cfg.gn_eps, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"])
    tok = h.flatten(2).transpose(1, 2)  # [B, L, C], L = H*W row-major
    qkv = F.linear(tok, sd[f"{p}.attn.in_proj_weight"], sd[f"{p}.attn.in_proj_bias"])
    q, k, v = qkv.split(C, dim=-1)
    q = q.reshape(B, -1, nh, hd).transpose(1, 2)
    k = k.reshape(B, -1, nh, hd).transpose(1, 2)
    v = v.reshape(B, -1, nh, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, -1, C)
    o = F.linear(o, sd[f"{p}.attn.out_proj.weight"], sd[f"{p}.attn.out_proj.bias"])
    o = o.transpose(1, 2).reshape(B, C, H, W)
    return (x + o) * sd[f"{p}.scale"]


def time_embedding(sd, cond: Tensor, cfg: UNetCfg) -> Tensor:
    """efficient_unet.py:232-237,273-275."""
    e = sinusoidal_embedding(cond, cfg.base_channels)
    e = F.linear(e, sd["time_embedding.1.weight"], sd["time_embedding.1.bias"])
    return F.linear(F.silu(e), sd["time_embedding.3.weight"], sd["time_embedding.3.bias"])


def unet_forward(sd: Dict[str, Tensor], cfg: UNetCfg, x: Tensor, cond: Tensor,
                 taps: Optional[dict] = None) -> Tensor:
    """efficient_unet.py:269-295.  `taps`, if given, receives named intermediate activations."""
    if cond.dim() == 0:
        cond = cond[None].repeat_interleave(x.shape[0], dim=0)
    temb = time_embedding(sd, cond.to(x), cfg)
    h = x
    cenc = coords_encoding(cfg, sd)
    if cenc is not None:
        h = torch.cat([h, cenc.to(h).repeat_interleave(h.shape[0], dim=0)], dim=1)
    h = ring_conv3x3(h, sd["in_conv.weight"], sd["in_conv.bias"])
    if taps is not None:
        taps["in_conv"] = h
    skips = []
    for name, cin, cout, nres, down, up, attn in block_table(cfg):
        if name.startswith("u_") and name != "u_block4":
            h = torch.cat([h, skips.pop()], dim=1)
        if down > 1:
            h = ring_conv3x3(h, sd[f"{name}.downsample.0.weight"], sd[f"{name}.downsample.0.bias"])
            h = resample_down2(h)
        for i in range(nres):
            h = residual_block(sd, f"{name}.residual_blocks.{i}", h, temb, cfg)
            if taps is not None:
                taps[f"{name}.rb{i}"] = h
        if attn:
            h = self_attention_block(sd, f"{name}.self_attn_block", h, cfg)
        if up > 1:
            h = resample_up2(h)
            h = ring_conv3x3(h, sd[f"{name}.upsample.1.weight"], sd[f"{name}.upsample.1.bias"])
        if taps is not None:
            taps[name] = h
        if name in ("d_block1", "d_block2", "d_block3"):
            skips.append(h)
    return ring_conv3x3(h, sd["out_conv.weight"], sd["out_conv.bias"])


# --------------------------------------------------------------------------------------- sampler
def _log(t: Tensor, eps: float = 1e-20) -> Tensor:
    return torch.log(t.clamp(min=eps))


def log_snr(t: Tensor, schedule: str = "cosine", image_d: float = None, noise_d_low: float = None,
            noise_d_high: float = None, lo: float = -15.0, hi: float = 15.0) -> Tensor:
    """models/diffusion/continuous_time.py:14-58."""
    def cos(tt):
        t_min = math.atan(math.exp(-0.5 * hi))
        t_max = math.atan(math.exp(-0.5 * lo))
        return -2 * _log(torch.tan(t_min + tt * (t_max - t_min)))
    if schedule == "linear":
        return -_log(torch.special.expm1(1e-4 + 10 * (t ** 2)))
    if schedule == "cosine":
        return cos(t)
    if schedule == "cosine_shifted":
        return cos(t) + 2 * math.log(noise_d_low / image_d)
    if schedule == "cosine_interpolated":
        a = cos(t) + 2 * math.log(noise_d_low / image_d)
        b = cos(t) + 2 * math.log(noise_d_high / image_d)
        return t * a + (1 - t) * b
    raise ValueError(schedule)


def alpha_sigma(lam: Tensor) -> Tuple[Tensor, Tensor]:
    """continuous_time.py:61-63."""
    return lam.sigmoid().sqrt(), (-lam).sigmoid().sqrt()


def _b(v: Tensor) -> Tensor:
    return v[:, None, None, None]


def x0_from_prediction(x_t, pred, alpha_t, sigma_t, objective: str, clip: Optional[float]):
    """continuous_time.py:208-217."""
    if objective == "eps":
        x0 = (x_t - sigma_t * pred) / alpha_t
    elif objective == "v":
        x0 = alpha_t * x_t - sigma_t * pred
    elif objective == "x_0":
        x0 = pred
    else:
        raise ValueError(objective)
    if clip is not None:
        x0 = x0.clamp(-clip, clip)
    return x0


def p_step_update(x_t, pred, noise, lam_t, lam_s, mode="ddpm", eta=0.0, objective="eps",
                  clip: Optional[float] = 1.0):
    """continuous_time.py:203-232 given the network prediction and the drawn noise."""
    lam_t, lam_s = _b(lam_t), _b(lam_s)
    a_t, s_t = alpha_sigma(lam_t)
    a_s, s_s = alpha_sigma(lam_s)
    x0 = x0_from_prediction(x_t, pred, a_t, s_t, objective, clip)
    if mode == "ddpm":
        c = -torch.special.expm1(lam_t - lam_s)
        mean = a_s * (x_t * (1 - c) / a_t + c * x0)
        return mean + s_s * c.sqrt() * noise
    if mode == "ddim":
        c1 = eta * s_s / s_t * (1 - a_t ** 2 / a_s ** 2).sqrt()
        c2 = (1 - a_s ** 2 - c1 ** 2).sqrt()
        eps = (x_t - a_t * x0) / s_t
        return a_s * x0 + c1 * noise + c2 * eps
    raise ValueError(mode)


def q_step_from_x0(x0, noise, lam_t):
    """continuous_time.py:169-176."""
    a, s = alpha_sigma(_b(lam_t))
    return x0 * a + noise * s


def q_step(x_s, noise, lam_t, lam_s):
    """continuous_time.py:178-190."""
    a_t, s_t = alpha_sigma(_b(lam_t))
    a_s, s_s = alpha_sigma(_b(lam_s))
    a_ts = a_t / a_s
    var = s_t ** 2 - a_ts ** 2 * s_s ** 2
    return x_s * a_ts + var.sqrt() * noise


@dataclass
class OracleDiffusion:
    """Continuous-time sampler driving the oracle U-Net with externally supplied ("teacher
    forced") noise so that CPU and CUDA trajectories can be compared draw for draw."""

    sd: Dict[str, Tensor]
    cfg: UNetCfg
    schedule: str = "cosine"
    objective: str = "eps"
    clip: Optional[float] = 1.0
    sched_kwargs: dict = field(default_factory=dict)

    def lam(self, t: Tensor) -> Tensor:
        return log_snr(t, self.schedule, **self.sched_kwargs)

    def model(self, x, cond):
        return unet_forward(self.sd, self.cfg, x, cond)

    def p_step(self, x_t, t, s, noise, mode="ddpm", eta=0.0):
        lt, ls = self.lam(t), self.lam(s)
        pred = self.model(x_t, lt)
        return p_step_update(x_t, pred, noise, lt, ls, mode, eta, self.objective, self.clip)

    def sample(self, x_T: Tensor, noises: Sequence[Tensor], mode="ddpm", eta=0.0, return_all=False):
        """continuous_time.py:234-258 with x_T and the per-step noises given."""
        n = len(noises)
        B = x_T.shape[0]
        steps = torch.linspace(1.0, 0.0, n + 1)[None].repeat_interleave(B, dim=0)
        x, out = x_T, [x_T]
        for i in range(n):
            x = self.p_step(x, steps[:, i], steps[:, i + 1], noises[i], mode, eta)
            out.append(x)
        return torch.stack(out) if return_all else x

    def repaint(self, known, mask, x_T, draw, num_steps, num_resample_steps=1, jump_length=1,
                return_all=False):
        """continuous_time.py:260-317; `draw()` returns the next noise tensor (the reference's
        draw order: q_step_from_x_0, then p_step, then q_step re-noising)."""
        B = known.shape[0]
        steps = torch.linspace(1, 0, num_steps + 1)[None].repeat_interleave(B, dim=0)
        x_t, out, x_s = x_T, [x_T], None
        for i in range(num_steps):
            for j in range(num_resample_steps):
                t, s = steps[:, [i]], steps[:, [i + 1]]
                interp = torch.linspace(0, 1, jump_length + 1)
                r = t + interp[None] * (s - t)
                x = x_t
                for k in range(jump_length):
                    known_s = q_step_from_x0(known, draw(), self.lam(r[:, k + 1]))
                    unknown_s = self.p_step(x, r[:, k], r[:, k + 1], draw(), "ddpm")
                    x = mask * known_s + (1 - mask) * unknown_s
                x_s = x
                out.append(x_s)
                if i == num_steps - 1 or j == num_resample_steps - 1:
                    x_t = x
                    break
                x = x_s
                for k in range(jump_length, 0, -1):
                    x = q_step(x, draw(), self.lam(r[:, k - 1]), self.lam(r[:, k]))
                x_t = x
        return torch.stack(out) if return_all else x_s


# --------------------------------------------------------------------------------------- discrete
def beta_schedule(name: str, steps: int) -> Tensor:
    """models/diffusion/discrete_time.py:12-48 (float64)."""
    if name == "linear":
        scale = 1000 / steps
        return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float64)
    t = torch.linspace(0, steps, steps + 1, dtype=torch.float64) / steps
    if name == "cosine":
        ab = torch.cos((t + 0.008) / 1.008 * math.pi * 0.5) ** 2
    elif name == "sigmoid":
        start, end, tau = -3.0, 3.0, 1.0
        v0, v1 = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
        ab = (-((t * (end - start) + start) / tau).sigmoid() + v1) / (v1 - v0)
    else:
        raise ValueError(name)
    ab = ab / ab[0]
    return torch.clip(1 - ab[1:] / ab[:-1], 0, 0.999)


def discrete_tables(name: str, steps: int):
    """discrete_time.py:57-78 -> float32 beta, alpha_bar, alpha_bar_prev (each [T])."""
    beta = beta_schedule(name, steps)
    ab = torch.cumprod(1 - beta, dim=0)
    abp = torch.cat([torch.ones(1, dtype=ab.dtype), ab[:-1]])
    return beta.float(), ab.float(), abp.float()


def discrete_p_step_update(x_t, pred, noise, t: Tensor, tables, mode="ddim", eta=0.0,
                           objective="eps", clip: Optional[float] = 1.0):
    """discrete_time.py:126-180 given prediction and noise (noise zeroed where t == 0)."""
    beta, ab, abp = (_b(v[t]) for v in tables)
    alpha = 1 - beta
    if objective == "eps":
        x0 = ab.rsqrt() * x_t - (ab.reciprocal() - 1).sqrt() * pred
    elif objective == "x_0":
        x0 = pred
    elif objective == "v":
        x0 = ab.sqrt() * x_t - (1 - ab).sqrt() * pred
    else:
        raise ValueError(objective)
    if clip is not None:
        x0 = x0.clamp(-clip, clip)
    nz = noise * _b((t != 0).to(noise))
    if mode == "ddpm":
        mean = abp.sqrt() * beta / (1 - ab) * x0 + (1 - abp) * alpha.sqrt() / (1 - ab) * x_t
        var = (beta * (1 - abp) / (1 - ab)).clamp(min=1e-20)
        return mean + (0.5 * var.log()).exp() * nz
    if mode == "ddim":
        var = (1 - abp) / (1 - ab) * (1 - ab / abp)
        std = eta * torch.sqrt(var)
        eps = (x_t - ab.sqrt() * x0) / (1 - ab).sqrt()
        x_s = abp.sqrt() * x0 + (1 - abp - std ** 2).sqrt() * eps
        return x_s + std * nz if eta > 0 else x_s
    raise ValueError(mode)


# --------------------------------------------------------------------------------------- lidar
def lidar_denormalize(x):  # utils/lidar.py:49-52
    return (x + 1) / 2


def lidar_normalize(x):  # utils/lidar.py:54-57
    return x * 2 - 1


def lidar_mask(metric, min_depth, max_depth):  # utils/lidar.py:118-120
    return ((metric > min_depth) & (metric < max_depth)).float()


def lidar_convert_depth(metric, fmt, min_depth, max_depth, mask=None):  # utils/lidar.py:72-97
    if mask is None:
        mask = lidar_mask(metric, min_depth, max_depth)
    if fmt == "log_depth":
        n = torch.log2(metric + 1) / math.log2(max_depth + 1)
    elif fmt == "inverse_depth":
        n = min_depth / (metric + 1e-8)
    elif fmt == "depth":
        n = metric / max_depth
    else:
        raise ValueError(fmt)
    return n.clamp(0, 1) * mask


def lidar_revert_depth(normalized, fmt, min_depth, max_depth):  # utils/lidar.py:99-116
    if fmt == "log_depth":
        m = torch.exp2(normalized * math.log2(max_depth + 1)) - 1
    elif fmt == "inverse_depth":
        m = min_depth / (normalized + 1e-8)
    elif fmt == "depth":
        m = normalized * max_depth
    else:
        raise ValueError(fmt)
    return m * lidar_mask(m, min_depth, max_depth)


def lidar_to_xyz(metric, ray_angles, min_depth, max_depth):  # utils/lidar.py:59-70
    mask = lidar_mask(metric, min_depth, max_depth)
    phi, theta = ray_angles[:, [0]], ray_angles[:, [1]]
    xyz = torch.cat([metric * phi.cos() * theta.cos(), metric * phi.cos() * theta.sin(),
                     metric * phi.sin()], dim=1)
    return xyz * mask


def lidar_postprocess(sample, ray_angles, fmt, min_depth, max_depth):
    """sample_and_save.py:52-57 — [depth, xyz(3), reflectance] from a clamped [-1, 1] sample."""
    s = lidar_denormalize(sample)
    depth = lidar_revert_depth(s[:, [0]], fmt, min_depth, max_depth)
    xyz = lidar_to_xyz(depth, ray_angles, min_depth, max_depth)
    return torch.cat([depth, xyz, s[:, [1]]], dim=1)
