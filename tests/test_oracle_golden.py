"""CPU: the oracle restatement must reproduce the reference's golden outputs (tests/golden/*.pt,
generated from /root/reference by tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import r2dm_oracle as O
from tests.helpers import GOLDEN, H_CFG, SMALL_CFG, draw_noise, rel_l2, repaint_masks


def sub(t):
    return t[..., ::2, ::7]


@pytest.fixture(scope="module")
def ops():
    return torch.load(os.path.join(GOLDEN, "ops_small.pt"))


def test_ops_against_golden(ops):
    c = ops["conv"]
    assert rel_l2(O.ring_conv3x3(c["x"], c["w"], c["b"]), c["y"]) < 1e-6
    r = ops["resample"]
    assert rel_l2(O.resample_down2(r["x"]), r["down"]) < 1e-6
    assert rel_l2(O.resample_up2(r["x"]), r["up"]) < 1e-6
    n = ops["norm"]
    assert rel_l2(O.group_norm(n["x"], 8, 1e-6, n["gw"], n["gb"]), n["y"]) < 1e-6
    assert rel_l2(O.adagn(n["x"], n["emb"], 8, 1e-6, n["pw"], n["pb"]), n["ya"]) < 1e-6
    s = ops["sinusoid"]
    assert rel_l2(O.sinusoidal_embedding(s["t"], 64), s["e"]) < 1e-6
    coords = O.hdl64e_linear_ray_angles(64, 1024).float()
    f = O.fourier_features(coords, O.fourier_freqs((64, 1024)), torch.zeros(16))
    assert rel_l2(f[:, :, ::7, ::13], ops["encoding"]["fourier_sub"]) < 1e-5
    assert rel_l2(O.spherical_harmonics(coords, 5)[:, :, ::7, ::13], ops["encoding"]["sh_sub"]) < 1e-5
    rb = ops["resblock"]
    assert rel_l2(O.residual_block(rb["sd"], "rb", rb["x"], rb["temb"], O.UNetCfg(base_channels=16)), rb["y"]) < 1e-6
    ab = ops["attn"]
    assert rel_l2(O.self_attention_block(ab["sd"], "ab", ab["x"], O.UNetCfg()), ab["y"]) < 5e-6


def test_empty_and_edge_shapes():
    # ring seam: a one-hot at column 0 must leak to column W-1 through the circular padding
    x = torch.zeros(1, 1, 4, 8)
    x[0, 0, 1, 0] = 1
    w = torch.ones(1, 1, 3, 3)
    y = O.ring_conv3x3(x, w, None)
    assert y[0, 0, 1, 7] == 1 and y[0, 0, 0, 7] == 1 and y[0, 0, 3, 7] == 0
    # elevation border is zero padded: top row sees only 2 rows of a constant image
    y = O.ring_conv3x3(torch.ones(1, 1, 4, 8), w, None)
    assert y[0, 0, 0, 0] == 6 and y[0, 0, 1, 0] == 9
    # resamplers preserve constants away from the elevation border and are linear
    c = torch.ones(1, 1, 8, 16)
    assert torch.allclose(O.resample_down2(c)[..., 1:-1, :], torch.ones(1, 1, 2, 8))
    assert torch.allclose(O.resample_up2(c)[..., 1:-1, :], torch.ones(1, 1, 14, 32))
    a, b = torch.randn(1, 2, 4, 8), torch.randn(1, 2, 4, 8)
    assert torch.allclose(O.resample_up2(a + 2 * b), O.resample_up2(a) + 2 * O.resample_up2(b), atol=1e-6)


def test_unet_small_against_golden():
    gd = torch.load(os.path.join(GOLDEN, "unet.pt"))["small"]
    sd = O.random_state_dict(SMALL_CFG, gd["seed_weights"])
    g = torch.Generator().manual_seed(gd["seed_x"])
    x = torch.randn(2, 2, *SMALL_CFG.resolution, generator=g)
    y = O.unet_forward(sd, SMALL_CFG, x, gd["cond"])
    assert rel_l2(y, gd["y"]) < 1e-5


def test_schema_matches_config_h():
    s = O.state_dict_schema(H_CFG)
    assert len(s) == 267
    n_params = sum(int(torch.tensor(v).prod()) if v else 1 for k, v in s.items()
                   if not (k == "coords" or "coords_encoding" in k or k.endswith(".scale") or k.endswith(".kernel")))
    assert n_params == 31_099_650   # SURVEY.md §6 probe


def test_sampler_against_golden():
    gd = torch.load(os.path.join(GOLDEN, "sampler.pt"))
    cfg = SMALL_CFG
    sd = O.random_state_dict(cfg, 77)
    orc = O.OracleDiffusion(sd, cfg)
    draws = draw_noise([100, 101], 7, cfg)
    y = orc.sample(draws[0], draws[1:], mode="ddim", eta=0.5, return_all=True)
    assert rel_l2(sub(y[1]), gd["sample_ddim_0.5"]["step1_sub"]) < 2e-4
    assert rel_l2(y[-1], gd["sample_ddim_0.5"]["final"]) < 2e-4
    # q steps
    g = torch.Generator().manual_seed(10)
    x0 = torch.randn(2, 2, *cfg.resolution, generator=g).clamp(-1, 1)
    t, s = torch.tensor([0.7, 0.2]), torch.tensor([0.6, 0.1])
    noise = draw_noise([400, 401], 2, cfg)
    assert rel_l2(sub(O.q_step_from_x0(x0, noise[0], orc.lam(t))), gd["q"]["xt_sub"]) < 1e-6
    assert rel_l2(sub(O.q_step(x0, noise[1], orc.lam(t), orc.lam(s))), gd["q"]["xq_sub"]) < 1e-6
    # discrete-time step (noise must be zeroed where t == 0)
    tb = O.discrete_tables("cosine", 40)
    g = torch.Generator().manual_seed(12)
    x_t = torch.randn(2, 2, *cfg.resolution, generator=g)
    steps = torch.tensor([17, 0])
    pred = O.unet_forward(sd, cfg, x_t, steps)
    nz = draw_noise([600, 601], 1, cfg)[0]
    y = O.discrete_p_step_update(x_t, pred, nz, steps, tb, "ddpm", 0.0)
    assert rel_l2(sub(y), gd["discrete_cosine_ddpm_0.0"]["y_sub"]) < 2e-5


def test_lidar_against_golden():
    gd = torch.load(os.path.join(GOLDEN, "lidar.pt"))
    g = torch.Generator().manual_seed(21)
    ang = O.hdl64e_linear_ray_angles(64, 1024).float()
    for fmt in ("log_depth", "inverse_depth", "depth"):
        metric = torch.rand(2, 1, 64, 1024, generator=g) * 90
        n = O.lidar_convert_depth(metric, fmt, 1.45, 80.0)
        assert rel_l2(n[..., ::5, ::17], gd[fmt]["n_sub"]) < 1e-6
        assert rel_l2(O.lidar_revert_depth(n, fmt, 1.45, 80.0)[..., ::5, ::17], gd[fmt]["r_sub"]) < 1e-6
        assert rel_l2(O.lidar_to_xyz(metric, ang, 1.45, 80.0)[..., ::5, ::17], gd[fmt]["xyz_sub"]) < 1e-6
    sample = torch.rand(2, 2, 64, 1024, generator=g) * 2 - 1
    out = O.lidar_postprocess(sample, ang, "log_depth", 1.45, 80.0)
    assert rel_l2(out[..., ::5, ::17], gd["postprocess_log_sub"]) < 1e-6
