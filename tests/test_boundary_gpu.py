"""Drop-in boundary on the GPU: configurations the reference's scripts can produce but the golden fixtures
do not cover (depth-only models, 2048-wide projections, the other coordinate encodings on the bf16
engine), checkpoint files, torch.compile(ddpm.sample) as in sample_and_save.py:45, autocast handling,
deep copies, and several devices / engines in one process."""
import copy
import dataclasses

import pytest
import torch

import r2dm_b200 as R
from oracle import r2dm_oracle as O
from tests.helpers import SMALL_CFG, rel_l2
from tests.util_model import make_cfg, make_ddpm

pytestmark = pytest.mark.gpu
TOL = {"fp32": 5e-3, "bf16": 3e-2}      # single-forward tolerances of DESIGN.md section 2


def _forward_vs_oracle(cfg, precision, B=2, seed=5):
    sd = O.random_state_dict(cfg, 31)
    ddpm = make_ddpm(cfg, sd, precision=precision)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.in_channels, *cfg.resolution, generator=g)
    cond = O.log_snr(torch.tensor([0.25, 0.8][:B]))
    y = ddpm.model(x.cuda(), cond.cuda())
    torch.cuda.synchronize()
    return rel_l2(y, O.unet_forward(sd, cfg, x, cond)), ddpm


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_depth_only_model(precision):
    """in_channels = 1 (train_reflectance = False, utils/inference.py:31-36)."""
    cfg = dataclasses.replace(SMALL_CFG, in_channels=1)
    e, ddpm = _forward_vs_oracle(cfg, precision)
    assert e <= TOL[precision], e
    y = ddpm.sample(batch_size=2, num_steps=3, progress=False, rng=R.setup_rng([1, 2], "cuda"), mode="ddpm")
    assert y.shape == (2, 1, *cfg.resolution) and torch.isfinite(y).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_2048_wide_projection(precision):
    """W = 2048 (`spherical-2048` projections, utils/option.py:60-65): 16 column tiles per row at full
    resolution, 2 per row at the bottleneck."""
    cfg = dataclasses.replace(SMALL_CFG, resolution=(16, 2048), num_residual_blocks=(1, 1, 1, 1))
    e, _ = _forward_vs_oracle(cfg, precision, B=1)
    assert e <= TOL[precision], e


@pytest.mark.parametrize("enc", ["spherical_harmonics", "polar_coordinates"])
def test_other_encodings_on_the_bf16_engine(enc):
    cfg = dataclasses.replace(SMALL_CFG, coords_encoding=enc)
    e, _ = _forward_vs_oracle(cfg, "bf16")
    assert e <= TOL["bf16"], (enc, e)


def test_checkpoint_file_to_gpu(tmp_path):
    """The full train.py:294-304 dict from disk -> setup_model(path, device='cuda') -> same prediction as the
    in-memory dict; hubconf.pretrained_r2dm(ckpt=path) is the same path (reference hubconf.py:21-37)."""
    cfg = make_cfg(SMALL_CFG)
    sd = O.random_state_dict(SMALL_CFG, 9)
    probe = R.build_model(cfg)
    full = {k: v for k, v in probe.state_dict().items() if not k.startswith("model.")}
    full.update({"model." + k: v for k, v in sd.items()})
    ckpt = {"cfg": cfg.to_dict(), "weights": full, "ema_weights": full, "optimizer": {"state": {}},
            "lr_scheduler": {"last_epoch": 3}, "global_step": 300_000}
    path = tmp_path / "ckpt.pth"
    torch.save(ckpt, path)
    import hubconf
    a, lu, _ = R.setup_model(str(path), device="cuda", show_info=False)
    b, _, _ = hubconf.pretrained_r2dm(ckpt=str(path), device="cuda", show_info=False, precision="bf16")
    x = torch.randn(1, 2, *SMALL_CFG.resolution).cuda()
    t = torch.tensor([0.3]).cuda()
    ya, yb = a.model(x, t), b.model(x, t)
    ref = O.unet_forward(sd, SMALL_CFG, x.cpu(), t.cpu())
    assert rel_l2(ya, ref) <= TOL["fp32"] and rel_l2(yb, ref) <= TOL["bf16"]
    assert lu.ray_angles.device.type == "cuda"


def test_torch_compile_of_sample_is_a_pass_through():
    """sample_and_save.py:45 does `ddpm.sample = torch.compile(ddpm.sample)`."""
    ddpm = make_ddpm(SMALL_CFG, O.random_state_dict(SMALL_CFG, 77), precision="fp32")
    kw = dict(batch_size=2, num_steps=4, progress=False, mode="ddpm")
    ref = ddpm.sample(rng=R.setup_rng([3, 4], "cuda"), **kw)
    compiled = torch.compile(ddpm.sample)
    out = compiled(rng=R.setup_rng([3, 4], "cuda"), **kw)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_autocast_and_explicit_precision():
    """bf16 autocast selects the bf16 engine only when no precision was requested explicitly; fp16 autocast
    (the reference's default mixed_precision) keeps the tf32 engine."""
    sd = O.random_state_dict(SMALL_CFG, 3)
    x = torch.randn(1, 2, *SMALL_CFG.resolution).cuda()
    t = torch.tensor([1.5]).cuda()
    default = make_ddpm(SMALL_CFG, sd, precision="fp32")
    y32 = default.model(x, t)
    with torch.autocast("cuda", dtype=torch.float16):
        assert torch.equal(default.model(x, t).float(), y32)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_auto = default.model(x, t).float()
    pinned = make_ddpm(SMALL_CFG, sd, precision="fp32")
    pinned.model.set_precision("fp32")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(pinned.model(x, t).float(), y32)
    bf = make_ddpm(SMALL_CFG, sd, precision="bf16")
    assert torch.equal(bf.model(x, t), y_auto) and not torch.equal(y_auto, y32)


def test_deepcopy_and_in_place_weight_edits():
    """EMA-style use: deep copies get their own engines; in-place parameter edits are picked up."""
    ddpm = make_ddpm(SMALL_CFG, O.random_state_dict(SMALL_CFG, 3), precision="fp32")
    x = torch.randn(1, 2, *SMALL_CFG.resolution).cuda()
    t = torch.tensor([0.5]).cuda()
    y0 = ddpm.model(x, t)
    clone = copy.deepcopy(ddpm)
    assert torch.equal(clone.model(x, t), y0)
    with torch.no_grad():
        clone.model.out_conv.weight.mul_(2.0)
        clone.model.out_conv.bias.mul_(2.0)
    y2 = clone.model(x, t)
    torch.cuda.synchronize()
    assert rel_l2(y2, 2 * y0) < 1e-5          # the edit reached the packed weights ...
    assert torch.equal(ddpm.model(x, t), y0)  # ... of the clone only


def test_invalid_configs_are_rejected_up_front():
    """Configurations the launch program cannot run must fail in r2dm_create, not in the middle of a forward
    or a CUDA-graph capture."""
    def build(**kw):
        cfg = dataclasses.replace(SMALL_CFG, **kw)
        return make_ddpm(cfg, O.random_state_dict(cfg, 1), precision="bf16")
    x = torch.zeros(1, 2, *SMALL_CFG.resolution).cuda()
    t = torch.zeros(1).cuda()
    with pytest.raises(R._lib.R2dmError, match="exceeds the supported maximum"):
        build(base_channels=192, attn_num_heads=24).model(x, t)    # 8 * 192 = 1536 input channels
    with pytest.raises(R._lib.R2dmError, match="identity skip"):
        build(channel_multiplier=(2, 1, 2, 4), attn_num_heads=4).model(x, t)   # u_block2: concat of 2 x 64 -> 128
    with pytest.raises(R._lib.R2dmError, match="gn_num_groups"):
        build(gn_num_groups=4).model(x, t)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in) and the SM count are per device."""
    sd = O.random_state_dict(SMALL_CFG, 3)
    x = torch.randn(1, 2, *SMALL_CFG.resolution)
    t = torch.tensor([0.5])
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        ddpm = make_ddpm(SMALL_CFG, sd, precision="bf16", device=dev)
        outs.append(ddpm.model(x.to(dev), t.to(dev)).cpu())
    assert torch.equal(outs[0], outs[1])
    moved = make_ddpm(SMALL_CFG, sd, precision="bf16", device="cuda:0")
    moved.model(x.cuda(0), t.cuda(0))
    moved.to("cuda:1")
    assert torch.equal(moved.model(x.to("cuda:1"), t.to("cuda:1")).cpu(), outs[0])
    assert all(k[0] == "cuda:1" for k in moved.model._engines)
