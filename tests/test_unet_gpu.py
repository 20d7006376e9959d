"""Whole-network parity: r2dm_b200.EfficientUNet (CUDA, through setup_model + the C ABI) vs the
golden outputs of the reference (tests/golden/unet.pt) and the CPU oracle.

Tolerances (l2-relative to the fp32 reference output, random weights):
  fp32 mode (tf32 tensor cores, like the reference's own GPU default): 5e-3
  bf16 mode (bf16 operands/storage, fp32 accumulate/statistics):                           3e-2
    (the reference's own bf16-autocast forward sits at 2.4e-2 vs fp64)
"""
import os

import pytest
import torch

from oracle import r2dm_oracle as O
from tests.helpers import GOLDEN, H_CFG, SMALL_CFG, rel_l2
from tests.util_model import make_ddpm

pytestmark = pytest.mark.gpu
TOL = {"fp32": 5e-3, "bf16": 3e-2}


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "unet.pt"))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["small", "H"])
def test_forward_matches_reference_golden(golden, tag, precision):
    cfg, B = (SMALL_CFG, 2) if tag == "small" else (H_CFG, 1)
    gd = golden[tag]
    sd = O.random_state_dict(cfg, gd["seed_weights"])
    ddpm = make_ddpm(cfg, sd, precision=precision)
    g = torch.Generator().manual_seed(gd["seed_x"])
    x = torch.randn(B, cfg.in_channels, *cfg.resolution, generator=g)
    y = ddpm.model(x.cuda(), gd["cond"].cuda())
    torch.cuda.synchronize()
    assert y.shape == gd["y"].shape and y.dtype == torch.float32
    e = rel_l2(y, gd["y"])
    assert e <= TOL[precision], f"{tag}/{precision}: l2-rel {e:.3e}"


@pytest.mark.parametrize("enc", ["spherical_harmonics", "polar_coordinates", None])
def test_forward_other_coordinate_encodings(enc):
    """efficient_unet.py:221-229: the spherical-harmonics (25 ch), polar (2 ch) and no encoding variants of
    the input stage against the CPU oracle (the goldens cover Fourier features only)."""
    import dataclasses
    cfg = dataclasses.replace(SMALL_CFG, coords_encoding=enc)
    sd = O.random_state_dict(cfg, 21)
    ddpm = make_ddpm(cfg, sd, precision="fp32")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, cfg.in_channels, *cfg.resolution, generator=g)
    cond = torch.tensor([1.5])
    y = ddpm.model(x.cuda(), cond.cuda())
    torch.cuda.synchronize()
    ref = O.unet_forward(sd, cfg, x, cond)
    e = rel_l2(y, ref)
    assert e <= TOL["fp32"], f"{enc}: l2-rel {e:.3e}"


def test_batch_composition_invariance():
    """Sample i's prediction must not depend on its batch neighbours."""
    cfg = SMALL_CFG
    sd = O.random_state_dict(cfg, 3)
    ddpm = make_ddpm(cfg, sd, precision="bf16")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 2, *cfg.resolution, generator=g).cuda()
    cond = torch.tensor([-2.0, 0.5, 7.0]).cuda()
    y3 = ddpm.model(x, cond)
    y1 = ddpm.model(x[1:2].contiguous(), cond[1:2])
    torch.cuda.synchronize()
    assert torch.equal(y3[1:2], y1), "batch split changed a sample's result"


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_batch_composition_invariance_config_h_b8(precision):
    """The benchmarked shape: config H, batch 8.  Every sample of the batch-8 forward equals its own batch-1
    forward bit for bit (tile shapes and summation orders never depend on the batch), so the B=1 golden
    parity of config H carries over to B=8."""
    from tests.helpers import H_CFG
    sd = O.random_state_dict(H_CFG, 1234)
    ddpm = make_ddpm(H_CFG, sd, precision=precision)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 2, *H_CFG.resolution, generator=g).cuda()
    cond = torch.linspace(-12.0, 12.0, 8).cuda()
    y8 = ddpm.model(x, cond).clone()
    for i in (0, 3, 7):
        y1 = ddpm.model(x[i:i + 1].contiguous(), cond[i:i + 1])
        assert torch.equal(y8[i:i + 1], y1), f"sample {i} differs between B=8 and B=1"
    ref = O.unet_forward(sd, H_CFG, x[5:6].cpu(), cond[5:6].cpu())
    assert rel_l2(y8[5:6], ref) <= TOL[precision]


def test_scalar_timestep_broadcast_and_autocast():
    cfg = SMALL_CFG
    sd = O.random_state_dict(cfg, 3)
    ddpm = make_ddpm(cfg, sd, precision="fp32")
    x = torch.randn(2, 2, *cfg.resolution).cuda()
    ya = ddpm.model(x, torch.tensor(1.5).cuda())
    yb = ddpm.model(x, torch.tensor([1.5, 1.5]).cuda())
    assert torch.equal(ya, yb)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yc = ddpm.model(x, torch.tensor([1.5, 1.5]).cuda())   # selects the bf16 engine
    assert rel_l2(yc, yb) < 3e-2 and not torch.equal(yc, yb)


def test_errors():
    import r2dm_b200 as R
    cfg = SMALL_CFG
    sd = O.random_state_dict(cfg, 3)
    ddpm = make_ddpm(cfg, sd)
    with pytest.raises(ValueError):
        ddpm.model(torch.randn(1, 2, 8, 1024).cuda(), torch.zeros(1).cuda())
    cpu = R.build_model(R.Config())
    with pytest.raises(R._lib.R2dmError):
        cpu.model(torch.randn(1, 2, 64, 1024), torch.zeros(1))
    with pytest.raises(NotImplementedError):
        ddpm(torch.randn(1, 2, 16, 1024).cuda())


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_conv_chain_option_is_bit_identical(precision):
    """Option `chain` (off by default): the ResidualBlock convolutions of a resolution level run as ONE persistent
    launch with per-(layer, image) dependency counters (csrc/conv_chain.cu).  Same tiles, same MMA order, same
    epilogue arithmetic: the outputs must be bit-identical to the one-launch-per-layer path, for an odd batch
    (unequal image groups), a single image (one group) and the benchmarked shape."""
    from r2dm_b200 import _lib as L
    from tests.helpers import H_CFG
    cases = [(SMALL_CFG, 3), (SMALL_CFG, 1), (H_CFG, 8)] if precision == "bf16" else [(SMALL_CFG, 2)]
    for cfg, B in cases:
        sd = O.random_state_dict(cfg, 21)
        g = torch.Generator().manual_seed(4)
        x = torch.randn(B, 2, *cfg.resolution, generator=g).cuda()
        cond = torch.linspace(-5.0, 5.0, B).cuda()
        outs, launches = [], []
        for chain in (0, 1):
            L.check(L.lib().r2dm_set_option(b"chain", chain))
            try:
                ddpm = make_ddpm(cfg, sd, precision=precision)      # the option is read when the workspace is planned
                y = ddpm.model(x, cond)
                y2 = ddpm.model(x, cond)                               # counters are re-armed on every launch
                torch.cuda.synchronize()
                assert torch.equal(y, y2)
                outs.append(y.clone())
                launches.append(ddpm.model.engine(precision).launches_per_forward)
            finally:
                L.check(L.lib().r2dm_set_option(b"chain", 0))
        assert launches[1] < launches[0], launches
        assert torch.equal(outs[0], outs[1]), f"chain launch changed the result ({cfg.resolution}, B={B})"
