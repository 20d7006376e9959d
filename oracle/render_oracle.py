"""CPU oracle for the caller-side consumers.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch CPU restatement of the reference's `utils/render.py` (render_point_clouds, bilinear_rasterizer,
estimate_surface_normal, make_Rt) and `metrics/bev.py` (point_cloud_to_histogram); imported only by `tests/`.
The product package `r2dm_b200` never imports it.

Parity pinning: `tests/golden/make_golden_render.py` imports the reference's own `utils/render.py` and
`metrics/bev.py` from /root/reference in the build container, asserts that this oracle reproduces them, and
stores input/output fixtures in `tests/golden/render.pt`.  `utils/render.py` imports kornia and matplotlib,
which are not installed here: the generating script supplies the two kornia functions the module calls
(`project_points`, `axis_angle_to_rotation_matrix`) as restatements of kornia 0.7.0's published source (the
version the reference pins, environment.yaml:14), so the splat / rasteriser / normal code that runs is the
reference's own, while `make_Rt` and the projection formula are pinned only against that restatement
("parity unpinned" for those two third-party formulas).
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


def axis_angle_to_rotation_matrix(v: Tensor) -> Tensor:
    """kornia 0.7.0 geometry/conversions.py `axis_angle_to_rotation_matrix` for v [N,3] (published algorithm:
    Rodrigues' formula with axis v / (theta + 1e-6); first-order form for theta^2 <= 1e-6)."""
    out = torch.empty(v.shape[0], 3, 3, dtype=v.dtype)
    for n in range(v.shape[0]):
        th2 = (v[n] * v[n]).sum()
        if th2 > 1e-6:
            th = th2.sqrt()
            k = v[n] / (th + 1e-6)
            c, s = th.cos(), th.sin()
            K = torch.tensor([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]], dtype=v.dtype)
            out[n] = c * torch.eye(3, dtype=v.dtype) + s * K + (1 - c) * torch.outer(k, k)
        else:
            out[n] = torch.tensor([[1.0, -v[n, 2], v[n, 1]], [v[n, 2], 1.0, -v[n, 0]], [-v[n, 1], v[n, 0], 1.0]],
                                  dtype=v.dtype)
    return out


def make_Rt(roll=0.0, pitch=0.0, yaw=0.0, x=0.0, y=0.0, z=0.0):
    """utils/render.py:10-29: R = Rz(yaw) @ Ry(pitch) @ Rx(roll), t = [x, y, z]."""
    def rot(axis, a):
        v = torch.zeros(1, 3)
        v[0, axis] = a
        return axis_angle_to_rotation_matrix(v)
    return rot(2, yaw) @ rot(1, pitch) @ rot(0, roll), torch.tensor([[x, y, z]])


def project_points(p: Tensor, focal: float) -> Tensor:
    """kornia 0.7.0 `project_points` with K = [[f,0,.5],[0,f,.5],[0,0,1]] (utils/render.py:57-66): perspective
    divide by z + 1e-8 (scale 1 where |z| <= 1e-8), then u = x f + 0.5."""
    z = p[..., 2:3]
    scale = torch.where(z.abs() > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
    return (scale * p[..., :2]) * focal + 0.5


def _inside(v: Tensor, n: int) -> Tensor:
    return ((v >= 0) & (v <= n - 1)).to(v.dtype)


def bilinear_rasterizer(coords: Tensor, values: Tensor, out_shape) -> Tensor:
    """utils/render.py:83-142: every point adds values * bilinear weight to its four neighbouring pixels;
    neighbours outside the image and weights below 1e-3 contribute nothing.  coords = (row, column)."""
    B, N, C = values.shape
    H, W = out_shape
    h, w = coords[..., 0], coords[..., 1]
    h0, w0 = h.floor(), w.floor()
    out = torch.zeros(B, H * W, C, dtype=values.dtype)
    for dh in (0, 1):
        for dw in (0, 1):
            hh, ww = h0 + dh, w0 + dw
            wh = ((h0 + 1) - h) if dh == 0 else (h - h0)      # top row: h_b - h, bottom row: h - h_t
            wv = ((w0 + 1) - w) if dw == 0 else (w - w0)
            wt = (wh * _inside(hh, H)) * (wv * _inside(ww, W))
            wt = wt * (wt >= 1e-3).to(wt.dtype)
            idx = (ww.clamp(0, W - 1) + W * hh.clamp(0, H - 1)).long()
            out.scatter_add_(1, idx[..., None].expand(-1, -1, C), values * wt[..., None])
    return out.reshape(B, H, W, C).permute(0, 3, 1, 2)


def render_point_clouds(points: Tensor, colors: Tensor = None, size: int = 800, R: Tensor = None, t: Tensor = None,
                        focal_length: float = 1.0) -> Tensor:
    """utils/render.py:32-80: flip z, apply extrinsics (row vectors: p @ R + t), project, weight every point by
    exp(-3 |p|), splat weighted colours and weights bilinearly, divide."""
    p = points.clone()
    p[..., 2] = -p[..., 2]
    if colors is None:
        colors = torch.ones_like(p)
    if R is not None:
        p = p @ R
    if t is not None:
        p = p + t
    uv = project_points(p, focal_length) * size
    ok = ((uv > 0) & (uv < size - 1)).all(dim=-1, keepdim=True)
    colors = colors * ok
    uv = size - uv
    depth = p.norm(dim=-1, keepdim=True)
    weight = 1.0 / torch.exp(3.0 * depth)
    weight = weight * (depth > 1e-8)
    num = bilinear_rasterizer(uv, weight * colors, (size, size))
    den = bilinear_rasterizer(uv, weight, (size, size))
    return num / (den + 1e-8)


def estimate_surface_normal(points: Tensor, d: int = 2, mode: str = "closest") -> Tensor:
    """utils/render.py:145-234: normal = cross(p_k - a, p_{k+2} - a) over the 8 neighbours at distance d in the
    reference's order (rows replicated at the border, columns circular); "closest" keeps the pair with the
    smallest |p_k - a| + |p_{k+2} - a|, "mean" averages all eight; normalised with + 1e-8."""
    B, _, H, W = points.shape
    P = points.permute(0, 2, 3, 1)
    off = [(-d, 0), (-d, d), (0, d), (d, d), (d, 0), (d, -d), (0, -d), (-d, -d)]
    hh = torch.arange(H)[:, None]
    ww = torch.arange(W)[None, :]
    nb = []
    for dh, dw in off:
        nb.append(P[:, (hh + dh).clamp(0, H - 1), (ww + dw) % W] - P)        # [B,H,W,3]
    nb = torch.stack(nb, dim=1)                                              # [B,8,H,W,3]
    nb2 = torch.roll(nb, shifts=-2, dims=1)
    if mode == "closest":
        cost = nb.norm(dim=-1) + nb2.norm(dim=-1)
        i = cost.argmin(dim=1)[:, None, :, :, None].expand(-1, 1, -1, -1, 3)
        n = torch.cross(nb.gather(1, i)[:, 0], nb2.gather(1, i)[:, 0], dim=-1)
    elif mode == "mean":
        n = torch.cross(nb, nb2, dim=-1).mean(dim=1)
    else:
        raise NotImplementedError(mode)
    n = n / (n.norm(dim=-1, keepdim=True) + 1e-8)
    return n.permute(0, 3, 1, 2)


def point_cloud_to_histogram(pc: Tensor, field_size: float = 160.0, bins: int = 100, min_depth: float = 3.0,
                             max_depth: float = 70.0) -> Tensor:
    """metrics/bev.py:5-24 without torch.histogramdd: counts of (x, y) in a bins x bins grid over
    [-field/2, field/2]^2 for the points with min_depth < |p| < max_depth; bin i = [e_i, e_{i+1}), the last
    one closed; e = torch.linspace(-field/2, field/2, bins + 1) in fp32."""
    depth = pc.norm(p=2, dim=1)
    xy = pc[(depth > min_depth) & (depth < max_depth)][:, :2]
    e = torch.linspace(-field_size / 2, field_size / 2, bins + 1, dtype=pc.dtype)
    inside = ((xy >= e[0]) & (xy <= e[-1])).all(dim=1)
    xy = xy[inside]
    pos = (torch.searchsorted(e, xy.contiguous(), right=True) - 1).clamp(max=bins - 1)
    hist = torch.zeros(bins * bins, dtype=pc.dtype)
    hist.index_add_(0, pos[:, 0] * bins + pos[:, 1], torch.ones(pos.shape[0], dtype=pc.dtype))
    return hist.view(bins, bins)
