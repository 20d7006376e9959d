"""Generate tests/golden/render.pt from the REFERENCE's `utils/render.py` and `metrics/bev.py`.

Runs only in the build container (needs /root/reference, imported read-only).  `utils/render.py` imports
kornia and matplotlib, neither of which is installed in this image; this script registers two stand-in modules
BEFORE importing it:
  * `kornia` with the two functions the module calls - `geometry.project_points` and
    `geometry.conversions.axis_angle_to_rotation_matrix` - restated from kornia 0.7.0's published source (the
    version pinned by the reference's environment.yaml:14);
  * `matplotlib.cm` with a synthetic `turbo` table (only `colorize`'s default argument touches it; the fixtures
    pass an explicit table).
Everything else that runs (render_point_clouds, bilinear_rasterizer, estimate_surface_normal, make_Rt's
composition order, colorize, point_cloud_to_histogram incl. torch.histogramdd) is the reference's own code.
The script asserts that oracle/render_oracle.py reproduces the reference and stores inputs + reference outputs.
Usage:  python tests/golden/make_golden_render.py
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


# ---------------------------------------------------------------- kornia 0.7.0 stand-ins (published algorithms)
def _k_axis_angle_to_rotation_matrix(axis_angle):
    def normal(aa, theta2, eps=1e-6):
        theta = torch.sqrt(theta2)
        wxyz = aa / (theta + eps)
        wx, wy, wz = torch.chunk(wxyz, 3, dim=1)
        c, s = torch.cos(theta), torch.sin(theta)
        r00 = c + wx * wx * (1.0 - c)
        r10 = wz * s + wx * wy * (1.0 - c)
        r20 = -wy * s + wx * wz * (1.0 - c)
        r01 = wx * wy * (1.0 - c) - wz * s
        r11 = c + wy * wy * (1.0 - c)
        r21 = wx * s + wy * wz * (1.0 - c)
        r02 = wy * s + wx * wz * (1.0 - c)
        r12 = -wx * s + wy * wz * (1.0 - c)
        r22 = c + wz * wz * (1.0 - c)
        return torch.cat([r00, r01, r02, r10, r11, r12, r20, r21, r22], dim=1).view(-1, 3, 3)

    def taylor(aa):
        rx, ry, rz = torch.chunk(aa, 3, dim=1)
        one = torch.ones_like(rx)
        return torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)

    aa1 = axis_angle.unsqueeze(1)
    theta2 = torch.matmul(aa1, aa1.transpose(1, 2)).squeeze(1)
    mask = (theta2 > 1e-6).view(-1, 1, 1)
    pos, neg = mask.type_as(theta2), (~mask).type_as(theta2)
    return pos * normal(axis_angle, theta2) + neg * taylor(axis_angle)


def _k_project_points(point_3d, camera_matrix):
    z = point_3d[..., -1:]
    mask = torch.abs(z) > 1e-8
    scale = torch.where(mask, 1.0 / (z + 1e-8), torch.ones_like(z))
    xy = scale * point_3d[..., :-1]
    u = xy[..., 0] * camera_matrix[..., 0, 0] + camera_matrix[..., 0, 2]
    v = xy[..., 1] * camera_matrix[..., 1, 1] + camera_matrix[..., 1, 2]
    return torch.stack([u, v], dim=-1)


def _lut(x):
    """synthetic 256-entry RGBA table (stands in for a matplotlib colormap)."""
    x = np.asarray(x, dtype=np.float64)
    return np.stack([x, 1 - x, 0.5 + 0.5 * np.sin(6.0 * x), np.ones_like(x)], axis=-1)


kornia = types.ModuleType("kornia")
kornia.geometry = types.ModuleType("kornia.geometry")
kornia.geometry.conversions = types.ModuleType("kornia.geometry.conversions")
kornia.geometry.project_points = _k_project_points
kornia.geometry.conversions.axis_angle_to_rotation_matrix = _k_axis_angle_to_rotation_matrix
sys.modules.update({"kornia": kornia, "kornia.geometry": kornia.geometry,
                    "kornia.geometry.conversions": kornia.geometry.conversions})
mpl = types.ModuleType("matplotlib")
mpl.cm = types.ModuleType("matplotlib.cm")
mpl.cm.turbo = _lut
sys.modules.update({"matplotlib": mpl, "matplotlib.cm": mpl.cm})

import utils.render as ref_render  # noqa: E402
from metrics import bev as ref_bev  # noqa: E402
from utils.lidar import LiDARUtility  # noqa: E402

from oracle import render_oracle as RO  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()


def synthetic_scene(B, H, W, seed):
    """A LiDAR-like scene: ground plane + a few walls seen from an HDL-64E, metric depth image [B,1,H,W]."""
    g = torch.Generator().manual_seed(seed)
    lu = LiDARUtility((H, W), "log_depth", 1.45, 80.0)
    phi, theta = lu.ray_angles[:, [0]], lu.ray_angles[:, [1]]
    depth = torch.empty(B, 1, H, W)
    for b in range(B):
        ground = 1.73 / (-phi.sin()).clamp(min=1e-3)                       # sensor height / sin(depression)
        r_wall = 6.0 + 30.0 * torch.rand(1, generator=g)
        wall = r_wall / (theta + torch.rand(1, generator=g) * 6.28).cos().abs().clamp(min=0.2) / phi.cos()
        d = torch.minimum(ground, wall) * (1 + 0.01 * torch.randn(1, 1, H, W, generator=g))
        d[torch.rand(1, 1, H, W, generator=g) < 0.05] = 0.0                # dropped rays
        depth[b] = d[0].clamp(max=120.0)
    xyz = lu.to_xyz(depth)
    return depth, xyz, lu


def main():
    out = {}
    # ---------------------------------------------------------------- make_Rt
    cams = [dict(pitch=torch.pi / 3, yaw=torch.pi / 4, z=0.8),       # generate.py:52
            dict(pitch=torch.pi / 4, yaw=torch.pi / 4, z=0.6),       # completion_demo.py:118-120
            dict(roll=0.2, pitch=-0.1, yaw=2.5, x=0.1, y=-0.2, z=0.3), dict()]
    out["make_Rt"] = []
    for kw in cams:
        R, t = ref_render.make_Rt(**kw)
        Ro, to = RO.make_Rt(**kw)
        assert relerr(Ro, R) < 1e-6 and torch.equal(to, t), kw
        out["make_Rt"].append(dict(kw={k: float(v) for k, v in kw.items()}, R=R, t=t))
    print("  make_Rt ok")

    # ---------------------------------------------------------------- render_point_clouds (two callers' settings)
    H, W = 64, 1024
    depth, xyz, lu = synthetic_scene(2, H, W, seed=1)
    pts = (xyz / lu.max_depth).flatten(2).transpose(1, 2).contiguous()        # B (H W) C
    z_min, z_max = -2 / lu.max_depth, 0.5 / lu.max_depth
    zc = ((xyz[:, [2]] / lu.max_depth - z_min) / (z_max - z_min)).clamp(0, 1)
    table = torch.from_numpy(_lut(np.linspace(0, 1, 256))[:, :3]).float()
    col_img = ref_render.colorize(zc, _lut) / 255                             # B 3 H W
    colors = 1 - col_img.flatten(2).transpose(1, 2).contiguous()
    # fixtures keep only depth + xyz; tests rebuild points / colours with scene_inputs() below (same torch CPU ops)
    out["scene"] = dict(depth=depth, xyz=xyz, max_depth=lu.max_depth, min_depth=lu.min_depth, table=table)
    out["render"] = dict(cases=[])
    for ci, size, focal, use_col in [(0, 800, 1.0, True), (1, 256, 1.0, False), (2, 200, 1.7, True), (3, 128, 1.0, True)]:
        R, t = out["make_Rt"][ci]["R"], out["make_Rt"][ci]["t"]
        kw = dict(size=size, focal_length=focal)
        if ci != 3:
            kw.update(R=R, t=t)
        else:
            kw.update(t=torch.tensor([[0.05, -0.02, 0.4]]))       # translation only
        ref = ref_render.render_point_clouds(points=pts, colors=colors if use_col else None, **kw)
        ora = RO.render_point_clouds(pts, colors if use_col else None, **kw)
        e = relerr(ora, ref)
        print(f"  render size={size} focal={focal} colors={use_col}: oracle vs reference l2-rel {e:.2e}, "
              f"coverage {(ref.sum(1) > 0).float().mean().item():.3f}")
        assert e < 1e-5
        st = 2 if size > 400 else 1
        out["render"]["cases"].append(dict(R=kw.get("R"), t=kw.get("t"), size=size, focal=focal, colors=use_col,
                                           stride=st, y=ref[..., ::st, ::st].half()))
    # colorize (explicit table)
    out["colorize"] = dict(y=ref_render.colorize(zc, _lut))

    # ---------------------------------------------------------------- bilinear_rasterizer incl. the border cases
    g = torch.Generator().manual_seed(2)
    B, N, C, Hh, Ww = 2, 4096, 5, 37, 53
    coords = torch.rand(B, N, 2, generator=g) * torch.tensor([Hh + 4.0, Ww + 4.0]) - 2.0     # some fall outside
    coords[0, :8] = torch.tensor([[0.0, 0.0], [Hh - 1.0, Ww - 1.0], [-0.5, 3.0], [3.0, -0.5], [Hh - 0.5, 1.0],
                                  [1.0, Ww - 0.5], [5.0005, 7.9995], [-3.0, -3.0]])
    values = torch.randn(B, N, C, generator=g)
    ref = ref_render.bilinear_rasterizer(coords, values, (Hh, Ww))
    assert relerr(RO.bilinear_rasterizer(coords, values, (Hh, Ww)), ref) < 1e-6
    out["rasterizer"] = dict(coords=coords, values=values, shape=(Hh, Ww), y=ref.contiguous())
    print("  bilinear_rasterizer ok")

    # ---------------------------------------------------------------- estimate_surface_normal
    out["normal"] = dict(cases=[])
    for d, mode in [(2, "closest"), (1, "closest"), (2, "mean"), (3, "mean")]:
        ref = ref_render.estimate_surface_normal(xyz, d=d, mode=mode)
        ora = RO.estimate_surface_normal(xyz, d=d, mode=mode)
        e = relerr(ora, ref)
        print(f"  normal d={d} {mode}: oracle vs reference l2-rel {e:.2e}")
        assert e < 1e-5
        out["normal"]["cases"].append(dict(d=d, mode=mode, y=ref[..., ::2, ::3].half()))

    # ---------------------------------------------------------------- point_cloud_to_histogram
    mask = lu.get_mask(depth)
    clouds = (xyz * mask).flatten(2).transpose(1, 2).contiguous()             # metric, B N 3 (evaluate.py:114-117)
    clouds[0, :4] = torch.tensor([[80.0, 0.0, 0.0], [-80.0, 10.0, 0.0], [10.0, 80.0, 0.0], [3.0, 0.0, 0.0]])
    hists = []
    for b in range(clouds.shape[0]):
        ref = ref_bev.point_cloud_to_histogram(clouds[b])
        assert torch.equal(RO.point_cloud_to_histogram(clouds[b]), ref), "histogram oracle != reference"
        hists.append(ref)
    ref_small = ref_bev.point_cloud_to_histogram(clouds[1], field_size=40.0, bins=16, min_depth=1.0, max_depth=30.0)
    assert torch.equal(RO.point_cloud_to_histogram(clouds[1], 40.0, 16, 1.0, 30.0), ref_small)
    out["bev"] = dict(head=clouds[0, :4].clone(), hists=torch.stack(hists).to(torch.int16), small=ref_small.to(torch.int16),
                      small_kw=dict(field_size=40.0, bins=16, min_depth=1.0, max_depth=30.0))
    print(f"  bev histogram ok ({int(hists[0].sum())} / {clouds.shape[1]} points counted)")

    # ---------------------------------------------------------------- PointNet1 features (random weights)
    from metrics.extractor.pointnet import PointNet1 as RefPointNet1
    sd = RO.random_pointnet_state_dict(seed=5, k=16)
    ref_net = RefPointNet1(k=16)
    ref_net.load_state_dict(sd)
    ref_net.eval()
    g = torch.Generator().manual_seed(6)
    small = torch.randn(3, 3, 2048, generator=g) * 0.5
    clouds_n = (clouds / 80.0).transpose(1, 2).contiguous()                  # evaluate.py:121: point_clouds / max depth
    feats = {}
    for name, pc in (("small", small), ("scene", clouds_n)):
        ref = ref_net(pc)
        e = relerr(RO.pointnet_features(sd, pc), ref)
        print(f"  pointnet {name} {tuple(pc.shape)}: oracle vs reference l2-rel {e:.2e}")
        assert e < 1e-5
        feats[name] = ref
    out["pointnet"] = dict(seed=5, k=16, small=small, feats=feats)

    path = os.path.join(HERE, "render.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
