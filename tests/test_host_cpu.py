"""CPU tests of the host logic: C-ABI exports, state-dict schema, schedules / coefficient folding,
config handling, LiDARUtility, error behaviour without a GPU, and the world_size-2 shard/gather
path over gloo."""
import ctypes
import math
import os
import re
import socket

import pytest
import torch
import torch.multiprocessing as mp

import r2dm_b200 as R
from oracle import r2dm_oracle as O
from r2dm_b200 import _lib, parallel
from r2dm_b200.diffusion import continuous_coefficients
from tests.helpers import H_CFG, ROOT, SMALL_CFG, rel_l2


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "r2dm_b200.h")).read()
    declared = set(re.findall(r"\b(r2dm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/r2dm_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().r2dm_version() >= 100


def test_create_rejects_unsupported_configs_without_gpu():
    cfg = _lib.R2dmConfig()
    cfg.in_channels, cfg.height, cfg.width, cfg.base_channels = 2, 64, 1000, 64
    cfg.channel_multiplier = (ctypes.c_int * 4)(1, 2, 4, 8)
    cfg.num_residual_blocks = (ctypes.c_int * 4)(3, 3, 3, 3)
    cfg.gn_num_groups, cfg.gn_eps, cfg.attn_num_heads, cfg.extra_channels, cfg.dtype = 8, 1e-6, 8, 32, 1
    h = ctypes.c_void_p()
    assert _lib.lib().r2dm_create(ctypes.byref(cfg), ctypes.byref(h)) < 0
    assert b"width" in _lib.lib().r2dm_last_error()
    cfg.width = 1024
    assert _lib.lib().r2dm_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    assert _lib.lib().r2dm_film_width(h) == 8832          # sum of 2*C over the 24 AdaGN layers
    assert _lib.lib().r2dm_weight_arena_bytes(h) > 31_000_000 * 2
    ws1 = _lib.lib().r2dm_workspace_bytes(h, 1)
    ws8 = _lib.lib().r2dm_workspace_bytes(h, 8)
    assert 0 < ws1 < ws8 < 8 * ws1 * 1.02 + (8 << 20)      # linear in the batch up to buffer-recycling slack
    # (buffers are recycled by exact size: at B=1 one more pair of sizes coincides than at B>=2, 125.9 vs 127.5 MB per image)
    _lib.lib().r2dm_destroy(h)


def test_state_dict_schema_matches_reference_layout():
    ddpm = R.build_model(R.Config())
    sd = ddpm.state_dict()
    schema = O.state_dict_schema(H_CFG)
    assert set(sd) == {"model." + k for k in schema} | {"_dummy"}
    for k, shape in schema.items():
        assert tuple(sd["model." + k].shape) == shape
    assert R.inference.count_parameters(ddpm) == 31_099_650
    # zero-initialised tensors like the reference (efficient_unet.py:39,84,267)
    assert sd["model.out_conv.weight"].abs().sum() == 0
    assert sd["model.d_block1.residual_blocks.0.conv2.weight"].abs().sum() == 0
    assert float(sd["model.u_block4.self_attn_block.scale"]) == pytest.approx(1 / math.sqrt(2))
    disc = R.DiscreteTimeGaussianDiffusion(model=ddpm.model, num_training_steps=100, noise_schedule="cosine")
    assert {"beta", "alpha_bar", "alpha_bar_prev", "snr"} <= set(disc.state_dict())
    tb = O.discrete_tables("cosine", 100)
    assert torch.equal(disc.beta.flatten(), tb[0]) and torch.equal(disc.alpha_bar_prev.flatten(), tb[2])


def test_no_cpu_fallback():
    ddpm = R.build_model(R.Config())
    with pytest.raises(_lib.R2dmError):
        ddpm.model(torch.zeros(1, 2, 64, 1024), torch.zeros(1))
    with pytest.raises(NotImplementedError):
        ddpm(torch.zeros(1, 2, 64, 1024))
    with pytest.raises(ValueError):
        R.EfficientUNet(2, (64, 1024), base_channels=64, ring=False)
    with pytest.raises(ValueError):
        R.ContinuousTimeGaussianDiffusion(ddpm.model, prediction_type="bogus")
    with pytest.raises(ValueError):
        R.ContinuousTimeGaussianDiffusion(ddpm.model, noise_schedule="bogus")


def test_schedules_and_coefficients_match_oracle():
    ddpm = R.build_model(R.Config())
    t = torch.linspace(0, 1, 33)
    for sched, kw in (("cosine", {}), ("linear", {}), ("cosine_shifted", dict(image_d=64.0, noise_d_low=32.0)),
                      ("cosine_interpolated", dict(image_d=64.0, noise_d_low=32.0, noise_d_high=256.0))):
        d = R.ContinuousTimeGaussianDiffusion(ddpm.model, noise_schedule=sched,
                                              **{k: v for k, v in kw.items()})
        lam = d.log_snr(t)
        assert lam.shape == (33, 1, 1, 1)
        assert torch.allclose(lam[:, 0, 0, 0], O.log_snr(t, sched, **kw), atol=1e-6)
    assert float(ddpm.log_snr(torch.tensor([0.0]))) == pytest.approx(15.0, abs=1e-4)
    assert float(ddpm.log_snr(torch.tensor([1.0]))) == pytest.approx(-15.0, abs=1e-4)
    # folded coefficients reproduce the reference update for every mode / objective
    g = torch.Generator().manual_seed(0)
    x, pred, nz = (torch.randn(3, 2, 4, 8, generator=g).double() for _ in range(3))
    tt, ss = torch.tensor([1.0, 0.6, 0.02]), torch.tensor([0.9, 0.55, 0.0])
    lt, ls = O.log_snr(tt).double(), O.log_snr(ss).double()
    for mode, eta in (("ddpm", 0.0), ("ddim", 0.0), ("ddim", 1.0)):
        for obj in ("eps", "v", "x_0"):
            c = continuous_coefficients(lt, ls, mode, eta, obj)[:, :, None, None, None]
            x0 = (c[:, 0] * x + c[:, 1] * pred).clamp(-1, 1)
            mine = c[:, 2] * x + c[:, 3] * x0 + c[:, 4] * nz
            ref = O.p_step_update(x, pred, nz, lt, ls, mode, eta, obj, 1.0)
            assert rel_l2(mine, ref) < 1e-9, (mode, eta, obj)
    # discrete folding
    disc = R.DiscreteTimeGaussianDiffusion(model=ddpm.model, num_training_steps=40, noise_schedule="sigmoid")
    steps = torch.tensor([39, 17, 0])
    tb = O.discrete_tables("sigmoid", 40)
    for mode, eta in (("ddpm", 0.0), ("ddim", 0.0), ("ddim", 0.7)):
        c = disc._coefficients(steps, mode, eta)[:, :, None, None, None]
        x0 = (c[:, 0] * x + c[:, 1] * pred).clamp(-1, 1)
        mine = c[:, 2] * x + c[:, 3] * x0 + c[:, 4] * nz
        ref = O.discrete_p_step_update(x.float(), pred.float(), nz.float(), steps, tb, mode, eta)
        assert rel_l2(mine, ref) < 2e-6, (mode, eta)


def test_randn_protocol():
    ddpm = R.build_model(R.Config())
    a = ddpm.randn(2, 3, 4, rng=[torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)])
    b = torch.stack([torch.randn(3, 4, generator=torch.Generator().manual_seed(s)) for s in (1, 2)])
    assert torch.equal(a, b)
    c = ddpm.randn(2, 3, rng=torch.Generator().manual_seed(3))
    assert torch.equal(c, torch.randn(2, 3, generator=torch.Generator().manual_seed(3)))
    assert ddpm.randn(2, 3).shape == (2, 3)
    with pytest.raises(AssertionError):
        ddpm.randn(3, 2, rng=[torch.Generator()])
    with pytest.raises(ValueError):
        ddpm.randn(3, 2, rng="nope")
    assert ddpm.sampling_shape == (2, 64, 1024) and ddpm.device.type == "cpu"


def test_config_roundtrip_and_lidar_utility():
    cfg = R.Config(**{"data": {"resolution": [64, 1024], "depth_format": "log_depth"},
                      "model": {"base_channels": 64, "coords_encoding": "fourier_features"},
                      "diffusion": {"timestep_type": "continuous"}, "training": {"mixed_precision": "fp16"}})
    assert cfg.data.resolution == (64, 1024) and cfg.data.max_depth == 80.0
    assert R.Config(**cfg.to_dict()).model.channel_multiplier == (1, 2, 4, 8)
    lu = R.LiDARUtility((64, 1024), "log_depth", 1.45, 80.0)
    assert torch.equal(lu.ray_angles, O.hdl64e_linear_ray_angles(64, 1024).float())
    g = torch.Generator().manual_seed(0)
    metric = torch.rand(1, 1, 64, 1024, generator=g) * 90
    for fmt in ("log_depth", "inverse_depth", "depth"):
        u = R.LiDARUtility((64, 1024), fmt, 1.45, 80.0)
        n = u.convert_depth(metric)
        assert torch.equal(n, O.lidar_convert_depth(metric, fmt, 1.45, 80.0))
        assert torch.equal(u.revert_depth(n), O.lidar_revert_depth(n, fmt, 1.45, 80.0))
        assert torch.equal(u.to_xyz(metric), O.lidar_to_xyz(metric, u.ray_angles, 1.45, 80.0))
        back = u.revert_depth(n)
        m = u.get_mask(metric).bool() & u.get_mask(back).bool()
        assert torch.allclose(back[m], metric[m], rtol=2e-4)          # convert -> revert round trip
    assert torch.allclose(lu.normalize(lu.denormalize(metric)), metric, rtol=1e-6)
    with pytest.raises(_lib.R2dmError):
        lu.postprocess(torch.zeros(1, 2, 64, 1024))
    # encodings equal the oracle's (and hence the reference's, see make_golden.py)
    m = R.EfficientUNet(2, (64, 1024), base_channels=64, coords_encoding="fourier_features")
    coords = O.hdl64e_linear_ray_angles(64, 1024).float()
    assert torch.equal(m.coords_encoding(coords), O.fourier_features(coords, O.fourier_freqs((64, 1024)), torch.zeros(16)))
    sh = R.EfficientUNet(2, (64, 1024), base_channels=64, coords_encoding="spherical_harmonics")
    assert rel_l2(sh.coords_encoding(coords), O.spherical_harmonics(coords, 5)) < 1e-6
    assert torch.equal(m.coords, O.polar_coords(64, 1024))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 64, 10_000):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seeds = list(range(100, 100 + n))

    def fake_sample(local):   # sample i depends on seed i only, like the real sampler
        assert len(local) > 0, "sample_fn must not be called for an empty shard (ddpm.sample(batch_size=0) raises)"
        return torch.stack([torch.randn(2, 4, 8, generator=torch.Generator().manual_seed(s)) for s in local])

    out = parallel.sample_sharded(fake_sample, seeds)
    ref = fake_sample(seeds)
    q.put((rank, bool(torch.equal(out, ref)), out.shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 5, 8])   # n = 1: fewer seeds than ranks (rank 1 has an empty shard)
def test_world_size_2_gloo_shard_and_gather(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok and cnt == n for _, ok, cnt in res)


# ------------------------------------------------------------------------------------ boundary
def test_checkpoint_file_round_trip(tmp_path):
    """setup_model(path) ingests the full checkpoint dict that train.py:294-304 writes (incl. the
    optimizer / lr_scheduler entries it does not need) and restores the weights exactly; the module can be
    deep-copied and pickled (EMA-style) even though live engines hold raw C handles."""
    import copy
    import pickle
    cfg = R.Config()
    cfg.model.num_residual_blocks = (1, 1, 1, 1)
    src = R.randomize_(R.build_model(cfg), seed=3)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    ckpt = {"cfg": cfg.to_dict(), "weights": sd, "ema_weights": sd, "optimizer": {"state": {}, "param_groups": []},
            "lr_scheduler": {"last_epoch": 7}, "global_step": 1234}
    path = tmp_path / "r2dm-test.pth"
    torch.save(ckpt, path)
    ddpm, lidar_utils, cfg2 = R.setup_model(str(path), device="cpu", show_info=False)
    assert cfg2.model.num_residual_blocks == (1, 1, 1, 1)
    got = ddpm.state_dict()
    assert set(got) == set(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    assert isinstance(lidar_utils, R.LiDARUtility)
    # hubconf.pretrained_r2dm(ckpt=...) is the same path (reference hubconf.py:21-37)
    import hubconf
    ddpm2, _, _ = hubconf.pretrained_r2dm(ckpt=str(path), device="cpu", show_info=False)
    assert all(torch.equal(ddpm2.state_dict()[k], sd[k]) for k in sd)
    clone = copy.deepcopy(ddpm)
    assert clone.model._engines == {} and torch.equal(clone.state_dict()["model.in_conv.weight"], sd["model.in_conv.weight"])
    again = pickle.loads(pickle.dumps(ddpm.model))
    assert again._engines == {} and torch.equal(again.in_conv.weight, ddpm.model.in_conv.weight)


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_signatures_match_reference():
    """Drop-in check: every public callable of the sampling path takes the reference's parameters, in the
    reference's order, with the reference's defaults (extra trailing keyword parameters are allowed)."""
    import importlib
    import inspect
    import sys
    sys.path.insert(0, REFERENCE)
    try:
        ref_ct = importlib.import_module("models.diffusion.continuous_time")
        ref_dt = importlib.import_module("models.diffusion.discrete_time")
        ref_unet = importlib.import_module("models.efficient_unet")
        ref_lidar = importlib.import_module("utils.lidar")
        ref_inf_src = open(os.path.join(REFERENCE, "utils", "inference.py")).read()
    finally:
        sys.path.remove(REFERENCE)
    import r2dm_b200.diffusion as D

    def check(ours, ref, skip=()):
        po, pr = inspect.signature(ours).parameters, inspect.signature(ref).parameters
        names_o = [n for n in po if not n.startswith("_")]
        names_r = [n for n in pr if n not in skip]
        assert names_o[:len(names_r)] == names_r, (ours.__qualname__, names_o, names_r)
        for n in names_r:
            if pr[n].default is not inspect.Parameter.empty:
                assert po[n].default == pr[n].default, (ours.__qualname__, n, po[n].default, pr[n].default)

    for name in ("sample", "repaint", "p_step", "q_step", "q_step_from_x_0", "__init__"):
        check(getattr(D.ContinuousTimeGaussianDiffusion, name), getattr(ref_ct.ContinuousTimeGaussianDiffusion, name))
    for name in ("sample", "p_step", "q_step_from_x_0"):
        check(getattr(D.DiscreteTimeGaussianDiffusion, name), getattr(ref_dt.DiscreteTimeGaussianDiffusion, name))
    check(R.EfficientUNet.__init__, ref_unet.EfficientUNet.__init__)
    check(R.EfficientUNet.forward, ref_unet.EfficientUNet.forward)
    for name in ("__init__", "normalize", "denormalize", "to_xyz", "convert_depth", "revert_depth", "get_mask"):
        if hasattr(ref_lidar.LiDARUtility, name):
            check(getattr(R.LiDARUtility, name), getattr(ref_lidar.LiDARUtility, name))
    # utils/inference.py does not import on Python >= 3.11 (pydantic dataclass defaults): compare with its source
    for fn, sig in (("setup_model", ["ckpt", "device", "ema", "show_info", "compile"]), ("setup_rng", ["seeds", "device"])):
        assert f"def {fn}(" in ref_inf_src
        assert list(inspect.signature(getattr(R, fn)).parameters)[:len(sig)] == sig
