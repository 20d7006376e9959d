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


# --------------------------------------------------------------------------------------- PointNet features
def random_pointnet_state_dict(seed: int, k: int = 16):
    """Seeded synthetic weights with the key set of the reference's PointNet1 (metrics/extractor/pointnet.py:67-81;
    the pretrained SpareNet checkpoint is not available offline): default-init-like weights, BatchNorm affine and
    running statistics drawn away from (1, 0, 0, 1) so that the BatchNorm folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def layer(name, cout, cin, conv):
        bound = 1.0 / cin ** 0.5
        shape = (cout, cin, 1) if conv else (cout, cin)
        sd[name + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1)

    for pre in ("feat.stn.", "feat."):
        layer(pre + "conv1", 64, 3, True); layer(pre + "conv2", 128, 64, True); layer(pre + "conv3", 1024, 128, True)
        bn(pre + "bn1", 64); bn(pre + "bn2", 128); bn(pre + "bn3", 1024)
    layer("feat.stn.fc1", 512, 1024, False); layer("feat.stn.fc2", 256, 512, False); layer("feat.stn.fc3", 9, 256, False)
    bn("feat.stn.bn4", 512); bn("feat.stn.bn5", 256)
    layer("fc1", 512, 1024, False); layer("fc2", 256, 512, False); layer("fc3", k, 256, False)
    bn("bn1", 512); bn("bn2", 256)
    return sd


def pointnet_features(sd, x: Tensor) -> Tensor:
    """PointNet1.forward in eval mode (metrics/extractor/pointnet.py:22-33, 47-58, 74-81) as plain matrix algebra:
    x [B,3,N] -> cat(global feature [1024], fc1 [512], fc2 [256], logits [k])."""
    def bn(h, name, eps=1e-5):
        shape = (1, -1, 1) if h.dim() == 3 else (1, -1)
        m, v = sd[name + ".running_mean"].view(shape), sd[name + ".running_var"].view(shape)
        return (h - m) / torch.sqrt(v + eps) * sd[name + ".weight"].view(shape) + sd[name + ".bias"].view(shape)

    def pw(h, name):     # Conv1d with kernel size 1
        return torch.einsum("oc,bcn->bon", sd[name + ".weight"][:, :, 0], h) + sd[name + ".bias"].view(1, -1, 1)

    def fc(h, name):
        return h @ sd[name + ".weight"].t() + sd[name + ".bias"]

    def trunk(h, pre, relu_last):
        h = torch.relu(bn(pw(h, pre + "conv1"), pre + "bn1"))
        h = torch.relu(bn(pw(h, pre + "conv2"), pre + "bn2"))
        h = bn(pw(h, pre + "conv3"), pre + "bn3")
        return (torch.relu(h) if relu_last else h).amax(dim=2)

    g = trunk(x, "feat.stn.", True)
    g = torch.relu(bn(fc(g, "feat.stn.fc1"), "feat.stn.bn4"))
    g = torch.relu(bn(fc(g, "feat.stn.fc2"), "feat.stn.bn5"))
    trans = fc(g, "feat.stn.fc3").view(-1, 3, 3) + torch.eye(3)
    moved = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)
    x1 = trunk(moved, "feat.", False)
    x2 = torch.relu(bn(fc(x1, "fc1"), "bn1"))
    x3 = torch.relu(bn(fc(x2, "fc2"), "bn2"))
    x4 = fc(x3, "fc3")
    return torch.cat((x1, x2, x3, x4), dim=1)
