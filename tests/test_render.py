"""Caller-side consumers: render splat, bilinear rasteriser, surface normals, BEV histogram.

CPU: oracle/render_oracle.py against the reference's outputs in tests/golden/render.pt (generated from
/root/reference by tests/golden/make_golden_render.py).  GPU (-m gpu): the CUDA kernels, called through the C ABI
via r2dm_b200.render, against the oracle on the same inputs and against the reference goldens.

Tolerances: scatter-adds run as fp32 atomics in arbitrary order (the reference's scatter_add_ has its own
order), so images are compared at l2-rel 1e-5 (fp32) / 1e-3 (against the fp16-stored goldens); a floor() or
1e-3-threshold decision that flips under a 1-ulp difference moves a single corner weight, bounded by a max-abs
check on the few affected pixels.  The histogram is integer work: bit-exact.
"""
import os

import pytest
import torch

from oracle import render_oracle as RO
from tests.helpers import GOLDEN, rel_l2


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN, "render.pt"))


def colorize_ref(x, table):
    """utils/render.py:237-246 with an explicit table (plain torch; only used to rebuild the fixture's colours)."""
    ids = (x.squeeze(1) * 256).clamp(0, 255).long()
    return table[ids].permute(0, 3, 1, 2).mul(255).clamp(0, 255).byte()


def scene_inputs(fx):
    sc = fx["scene"]
    xyz, md = sc["xyz"], sc["max_depth"]
    pts = (xyz / md).flatten(2).transpose(1, 2).contiguous()
    z_min, z_max = -2 / md, 0.5 / md
    zc = ((xyz[:, [2]] / md - z_min) / (z_max - z_min)).clamp(0, 1)
    col = colorize_ref(zc, sc["table"])
    colors = 1 - (col / 255).flatten(2).transpose(1, 2).contiguous()
    return pts, colors, zc, col


def bev_clouds(fx):
    sc = fx["scene"]
    mask = ((sc["depth"] > sc["min_depth"]) & (sc["depth"] < sc["max_depth"])).float()
    clouds = (sc["xyz"] * mask).flatten(2).transpose(1, 2).contiguous()
    clouds[0, :4] = fx["bev"]["head"]
    return clouds


# ------------------------------------------------------------------------------------------- CPU: oracle pinning
def test_oracle_make_rt(fx):
    for c in fx["make_Rt"]:
        R, t = RO.make_Rt(**c["kw"])
        assert rel_l2(R, c["R"]) < 1e-6 and torch.equal(t, c["t"])


def test_oracle_render_against_golden(fx):
    pts, colors, zc, col = scene_inputs(fx)
    assert torch.equal(col, fx["colorize"]["y"])
    for c in fx["render"]["cases"]:
        if c["size"] > 400:
            continue   # the 800 x 800 case takes a few seconds on the CPU: covered on the GPU box
        y = RO.render_point_clouds(pts, colors if c["colors"] else None, size=c["size"], R=c["R"], t=c["t"],
                                   focal_length=c["focal"])
        assert rel_l2(y[..., ::c["stride"], ::c["stride"]], c["y"].float()) < 1e-3


def test_oracle_rasterizer_normals_histogram(fx):
    r = fx["rasterizer"]
    assert rel_l2(RO.bilinear_rasterizer(r["coords"], r["values"], r["shape"]), r["y"]) < 1e-6
    xyz = fx["scene"]["xyz"]
    for c in fx["normal"]["cases"]:
        y = RO.estimate_surface_normal(xyz, d=c["d"], mode=c["mode"])
        assert (y[..., ::2, ::3] - c["y"].float()).abs().max() < 2e-3
    clouds = bev_clouds(fx)
    for b in range(clouds.shape[0]):
        assert torch.equal(RO.point_cloud_to_histogram(clouds[b]), fx["bev"]["hists"][b].float())
    assert torch.equal(RO.point_cloud_to_histogram(clouds[1], **fx["bev"]["small_kw"]), fx["bev"]["small"].float())


def test_render_module_fails_loudly_on_cpu(fx):
    from r2dm_b200 import render as R
    from r2dm_b200._lib import R2dmError
    with pytest.raises(R2dmError):
        R.render_point_clouds(torch.zeros(1, 4, 3))
    with pytest.raises(R2dmError):
        R.point_cloud_to_histogram(torch.zeros(4, 3))
    Rm, t = R.make_Rt(pitch=torch.pi / 3, yaw=torch.pi / 4, z=0.8)      # host-side, device-agnostic
    assert rel_l2(Rm, fx["make_Rt"][0]["R"]) < 1e-6 and torch.equal(t, fx["make_Rt"][0]["t"])


# ------------------------------------------------------------------------------------------- GPU: kernels
def close_image(y, ref, tol, flips_abs=None):
    """l2-rel within tol; if a handful of pixels differ by a flipped floor / threshold decision, they must be few."""
    y, ref = y.float().cpu(), ref.float().cpu()
    bad = (y - ref).abs() > 1e-4 + 1e-4 * ref.abs()
    frac = bad.float().mean().item()
    return rel_l2(y, ref) < tol or frac < 2e-5, (rel_l2(y, ref), frac)


@pytest.mark.gpu
def test_render_point_clouds_gpu(fx):
    from r2dm_b200 import render as R
    pts, colors, _, _ = scene_inputs(fx)
    for c in fx["render"]["cases"]:
        kw = dict(size=c["size"], focal_length=c["focal"])
        y = R.render_point_clouds(pts.cuda(), colors.cuda() if c["colors"] else None,
                                  R=None if c["R"] is None else c["R"].cuda(),
                                  t=None if c["t"] is None else c["t"].cuda(), **kw)
        assert y.shape == (pts.shape[0], 3, c["size"], c["size"])
        ref = RO.render_point_clouds(pts, colors if c["colors"] else None, R=c["R"], t=c["t"], **kw)
        ok, info = close_image(y, ref, 1e-5)
        assert ok, (c["size"], info)
        assert rel_l2(y[..., ::c["stride"], ::c["stride"]], c["y"].float()) < 1e-3     # the reference's own output


@pytest.mark.gpu
def test_render_callers_protocol_gpu(fx):
    """generate.py:44-59 / completion_demo.py:117-133 as written there, on CUDA tensors."""
    from r2dm_b200 import render as R
    pts, colors, zc, col = scene_inputs(fx)
    table = fx["scene"]["table"]
    assert torch.equal(R.colorize(zc.cuda(), table).cpu(), col)
    Rm, t = R.make_Rt(pitch=torch.pi / 3, yaw=torch.pi / 4, z=0.8, device="cuda")
    bev = 1 - R.render_point_clouds(points=pts.cuda(), colors=colors.cuda(), R=Rm, t=t)
    # make_Rt ran on the device: its sin / cos differ from the CPU's by an ulp, so the oracle gets the same matrix
    assert rel_l2(Rm, fx["make_Rt"][0]["R"]) < 1e-6 and torch.equal(t.cpu(), fx["make_Rt"][0]["t"])
    ref = 1 - RO.render_point_clouds(pts, colors, R=Rm.cpu(), t=t.cpu())
    ok, info = close_image(bev, ref, 1e-5)
    assert ok, info
    # white splat without colours / extrinsics, single image, non-default size
    y = R.render_point_clouds(pts[:1].cuda() + torch.tensor([0.0, 0.0, -0.3], device="cuda"), size=96)
    ref = RO.render_point_clouds(pts[:1] + torch.tensor([0.0, 0.0, -0.3]), size=96)
    ok, info = close_image(y, ref, 1e-5)
    assert ok, info


@pytest.mark.gpu
def test_bilinear_rasterizer_gpu(fx):
    from r2dm_b200 import render as R
    r = fx["rasterizer"]
    y = R.bilinear_rasterizer(r["coords"].cuda(), r["values"].cuda(), r["shape"])
    assert rel_l2(y, r["y"]) < 1e-5
    # one point exactly on a pixel centre: all weight on one pixel; a point outside: nothing
    co = torch.tensor([[[2.0, 3.0], [-5.0, 1.0]]])
    va = torch.tensor([[[1.5], [7.0]]])
    y = R.bilinear_rasterizer(co.cuda(), va.cuda(), (4, 5)).cpu()
    assert y[0, 0, 2, 3] == 1.5 and y.sum() == 1.5
    assert torch.equal(y, RO.bilinear_rasterizer(co, va, (4, 5)))


@pytest.mark.gpu
def test_surface_normal_gpu(fx):
    from r2dm_b200 import render as R
    xyz = fx["scene"]["xyz"]
    for c in fx["normal"]["cases"]:
        y = R.estimate_surface_normal(xyz.cuda(), d=c["d"], mode=c["mode"]).cpu()
        ref = RO.estimate_surface_normal(xyz, d=c["d"], mode=c["mode"])
        # argmin ties / near-degenerate cross products can pick another neighbour pair: bounded fraction
        bad = ((y - ref).abs().amax(dim=1) > 1e-4).float().mean().item()
        assert bad < 2e-3, (c["d"], c["mode"], bad)
        good = (y - ref).abs().amax(dim=1, keepdim=True) <= 1e-4
        assert ((y - ref) * good).abs().max() <= 1e-4
        assert ((y[..., ::2, ::3] - c["y"].float()).abs().amax(dim=1) > 2e-3).float().mean().item() < 2e-3
    with pytest.raises(NotImplementedError):
        R.estimate_surface_normal(xyz.cuda(), mode="nearest")


@pytest.mark.gpu
def test_bev_histogram_gpu(fx):
    from r2dm_b200 import render as R
    clouds = bev_clouds(fx)
    h = R.point_clouds_to_histograms(clouds.cuda()).cpu()
    assert torch.equal(h, fx["bev"]["hists"].float())                    # bit-exact vs the reference's histogramdd
    kw = fx["bev"]["small_kw"]
    assert torch.equal(R.point_cloud_to_histogram(clouds[1].cuda(), **kw).cpu(), fx["bev"]["small"].float())
    # a larger random cloud against the oracle (size-independent property: counts sum to the points kept)
    g = torch.Generator().manual_seed(7)
    pc = (torch.rand(3, 200_000, 3, generator=g) - 0.5) * torch.tensor([200.0, 200.0, 10.0])
    hb = R.point_clouds_to_histograms(pc.cuda()).cpu()
    for b in range(3):
        assert torch.equal(hb[b], RO.point_cloud_to_histogram(pc[b]))


# ------------------------------------------------------------------------------------------- PointNet features
def test_oracle_pointnet_against_golden(fx):
    pn = fx["pointnet"]
    sd = RO.random_pointnet_state_dict(pn["seed"], pn["k"])
    assert rel_l2(RO.pointnet_features(sd, pn["small"]), pn["feats"]["small"]) < 1e-5
    clouds = (bev_clouds(fx) / 80.0).transpose(1, 2).contiguous()
    assert rel_l2(RO.pointnet_features(sd, clouds), pn["feats"]["scene"]) < 1e-5


def test_pointnet_module_mirrors_reference_state_dict(fx):
    """Same module tree / parameter names as metrics/extractor/pointnet.py: a reference state dict loads strictly;
    no CPU path."""
    from r2dm_b200 import pointnet as P
    from r2dm_b200._lib import R2dmError
    pn = fx["pointnet"]
    net = P.PointNet1(k=pn["k"])
    net.load_state_dict(RO.random_pointnet_state_dict(pn["seed"], pn["k"]), strict=True)
    with pytest.raises(NotImplementedError):
        net.train()(pn["small"])
    with pytest.raises(R2dmError):
        net.eval()(pn["small"])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pointnet_features_gpu(fx, precision):
    """Tolerances: the point-wise layers multiply tf32 / bf16 operands on the tensor cores (inputs and hidden
    activations rounded to the operand type), three layers deep, followed by a max over up to 65 536 points:
    l2-rel 3e-3 (tf32) / 3e-2 (bf16) on every feature block."""
    from r2dm_b200 import pointnet as P
    pn = fx["pointnet"]
    net = P.PointNet1(k=pn["k"], precision=precision)
    net.load_state_dict(RO.random_pointnet_state_dict(pn["seed"], pn["k"]))
    net = net.eval().cuda()
    tol = 3e-3 if precision == "fp32" else 3e-2
    clouds = (bev_clouds(fx) / 80.0).transpose(1, 2).contiguous()
    for name, pc in (("small", pn["small"]), ("scene", clouds)):
        y = net(pc.cuda()).cpu()
        ref = pn["feats"][name]
        assert y.shape == ref.shape
        for lo, hi in ((0, 1024), (1024, 1536), (1536, 1792), (1792, 1792 + pn["k"])):
            e = rel_l2(y[:, lo:hi], ref[:, lo:hi])
            assert e < tol, (name, precision, lo, e)
    # batch composition must not matter (per-cloud max pool, per-cloud transform)
    y2 = net(pn["small"][1:2].cuda()).cpu()
    assert torch.equal(y2[0], net(pn["small"].cuda()).cpu()[1])


@pytest.mark.gpu
def test_consumer_edge_cases_gpu():
    """Degenerate inputs: points on the camera plane (z = 0: kornia's divide guard), at the origin (zero weight),
    behind the camera, the smallest point image PointNet accepts, and shapes the C ABI must reject."""
    from r2dm_b200 import pointnet as P
    from r2dm_b200 import render as R
    from r2dm_b200._lib import R2dmError
    pts = torch.tensor([[[0.1, 0.2, 0.0], [0.0, 0.0, 0.0], [0.2, -0.1, 0.5], [0.3, 0.3, -0.7], [1e-9, 0.0, 1e-9]]])
    y = R.render_point_clouds(pts.cuda(), size=64).cpu()
    ref = RO.render_point_clouds(pts, size=64)
    assert torch.isfinite(y).all() and rel_l2(y, ref) < 1e-5
    # a single bin pair, every point on an edge or outside
    pc = torch.tensor([[-5.0, 0.0, 0.0], [5.0, 5.0, 0.0], [0.0, -5.0, 3.0], [0.0, 0.0, 4.0], [6.0, 0.0, 0.0]])
    kw = dict(field_size=10.0, bins=2, min_depth=1.0, max_depth=8.0)
    assert torch.equal(R.point_cloud_to_histogram(pc.cuda(), **kw).cpu(), RO.point_cloud_to_histogram(pc, **kw))
    # normals of a plane are constant; d larger than the image height clamps at the border rows
    hh, ww = torch.meshgrid(torch.arange(8.0), torch.arange(256.0), indexing="ij")
    plane = torch.stack([ww, hh, 0.5 * ww + 2.0 * hh])[None]
    n = R.estimate_surface_normal(plane.cuda(), d=3, mode="mean").cpu()
    assert rel_l2(n, RO.estimate_surface_normal(plane, d=3, mode="mean")) < 1e-5
    # PointNet: 128 points is the smallest image; 100 is rejected before any launch
    net = P.PointNet1(k=4).eval().cuda()
    net.load_state_dict(RO.random_pointnet_state_dict(9, 4))
    x = torch.randn(2, 3, 128, generator=torch.Generator().manual_seed(3))
    assert rel_l2(net(x.cuda()), RO.pointnet_features(RO.random_pointnet_state_dict(9, 4), x)) < 3e-3
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 100, device="cuda"))
    with pytest.raises(R2dmError):
        R.render_point_clouds(pts.cuda(), size=8192)          # accumulator index would leave fp32's exact range
