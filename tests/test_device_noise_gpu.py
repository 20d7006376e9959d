"""Noise drawn inside the sampler kernels (Philox4x32-10 + Box-Muller, csrc/elementwise.cu) must be
BIT-identical to what the reference draws on the same device: `torch.randn(C, H, W, generator=g_i)`
per sample per draw (models/diffusion/base.py:71-94 with utils/inference.py:113-114 generators).
Integer / bit-exact work, so every comparison here is torch.equal."""
import ctypes

import pytest
import torch

from oracle import r2dm_oracle as O
from tests.helpers import SMALL_CFG, repaint_masks
from tests.util_model import make_ddpm

pytestmark = pytest.mark.gpu


def _philox_draws(seeds, offsets, shape, n_draws):
    """n_draws successive draws of every generator through r2dm_philox_normal."""
    from r2dm_b200 import _lib as L
    from r2dm_b200.diffusion import _torch_randn_geometry, _u64_tensor
    dev = torch.device("cuda", torch.cuda.current_device())
    per = 1
    for s in shape:
        per *= s
    threads, inc = _torch_randn_geometry(per, dev)
    sd, od = _u64_tensor(seeds, dev), _u64_tensor(offsets, dev)
    ph = L.R2dmPhilox(L.ptr(sd), L.ptr(od), None, None, 0, 0, inc, threads)
    outs = []
    for k in range(n_draws):
        out = torch.empty(len(seeds), *shape, device=dev)
        L.check(L.lib().r2dm_philox_normal(L.ptr(out), ctypes.byref(ph), k, len(seeds), per, L.stream_ptr()),
                "r2dm_philox_normal")
        outs.append(out)
    torch.cuda.synchronize()
    return outs, inc


@pytest.mark.parametrize("shape", [(2, 64, 1024), (1, 64, 1024), (2, 64, 2048), (2, 16, 1024), (5, 64, 2048), (3, 7, 12)])
def test_philox_normal_equals_torch_randn(shape):
    """Shapes cover: one element per ATen thread (the 2x64x1024 sampling shape, in_channels = 1,
    W = 2048), several elements per thread with the unroll-by-4 interleave (5x64x2048 > 1184 * 256
    threads), and a ragged size that is not a multiple of the block."""
    from r2dm_b200.diffusion import _cuda_gen_seed_offset
    seeds = [0, 7, 2 ** 40 + 3, 2 ** 63 + 11]
    gens = [torch.Generator("cuda").manual_seed(s) for s in seeds]
    for g in gens[1:]:                       # non-zero starting offsets
        torch.randn(33, generator=g, device="cuda")
    so = [_cuda_gen_seed_offset(g) for g in gens]
    assert [a for a, _ in so] == seeds
    per = shape[0] * shape[1] * shape[2]
    if per % 4:
        pytest.skip("kernel API draws whole tensors of a multiple of 4 elements")
    outs, inc = _philox_draws(seeds, [b for _, b in so], shape, 3)
    for k in range(3):
        ref = torch.stack([torch.randn(*shape, generator=g, device="cuda") for g in gens])
        assert torch.equal(outs[k], ref), f"draw {k} differs from torch.randn"
    # and the offsets torch consumed are what the host bookkeeping assumes
    for g, (_, off0) in zip(gens, so):
        assert _cuda_gen_seed_offset(g)[1] == off0 + 3 * inc


@pytest.fixture(scope="module")
def ddpm():
    return make_ddpm(SMALL_CFG, O.random_state_dict(SMALL_CFG, 77), precision="fp32")


@pytest.mark.parametrize("mode,eta,steps", [("ddpm", 0.0, 19), ("ddim", 0.0, 5), ("ddim", 0.6, 5)])
def test_sample_device_noise_equals_host_noise(ddpm, mode, eta, steps):
    """Same seeds -> the in-kernel draws and the host-side torch.randn draws give the same trajectory
    bit for bit, and leave the generators in the same state (19 steps = 1 eager + a 16-step graph +
    a 2-step remainder graph)."""
    import r2dm_b200 as R
    from r2dm_b200.diffusion import _cuda_gen_seed_offset
    seeds = [11, 12, 13]
    ddpm.device_noise = False
    rng_h = R.setup_rng(seeds, "cuda")
    host = ddpm.sample(batch_size=3, num_steps=steps, progress=False, rng=rng_h, mode=mode, ddim_eta=eta)
    ddpm.device_noise = True
    rng_d = R.setup_rng(seeds, "cuda")
    devn = ddpm.sample(batch_size=3, num_steps=steps, progress=False, rng=rng_d, mode=mode, ddim_eta=eta)
    again = ddpm.sample(batch_size=3, num_steps=steps, progress=False, rng=R.setup_rng(seeds, "cuda"), mode=mode,
                        ddim_eta=eta)            # warm graph cache, fresh generators
    torch.cuda.synchronize()
    assert torch.equal(devn, host)
    assert torch.equal(again, host)
    assert [_cuda_gen_seed_offset(g) for g in rng_d] == [_cuda_gen_seed_offset(g) for g in rng_h]
    # the generators keep working as a stream: a second call continues where the first stopped
    h2 = ddpm.sample(batch_size=3, num_steps=3, progress=False, rng=rng_d, mode="ddpm")
    ddpm.device_noise = False
    d2 = ddpm.sample(batch_size=3, num_steps=3, progress=False, rng=rng_h, mode="ddpm")
    ddpm.device_noise = True
    assert torch.equal(h2, d2)


def test_repaint_device_noise_equals_host_noise(ddpm):
    import r2dm_b200 as R
    from r2dm_b200.diffusion import _cuda_gen_seed_offset
    g = torch.Generator().manual_seed(10)
    known = torch.randn(2, 2, *SMALL_CFG.resolution, generator=g).clamp(-1, 1).cuda()
    mask = repaint_masks(2, SMALL_CFG).cuda()
    outs, states = [], []
    for flag in (False, True):
        ddpm.device_noise = flag
        rng = R.setup_rng([21, 22], "cuda")
        outs.append(ddpm.repaint(known, mask, num_steps=4, num_resample_steps=3, jump_length=2, progress=False,
                                 rng=rng))
        states.append([_cuda_gen_seed_offset(r) for r in rng])
    ddpm.device_noise = True
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert states[0] == states[1]


def test_discrete_sample_device_noise(ddpm):
    import r2dm_b200 as R
    d = make_ddpm(SMALL_CFG, O.random_state_dict(SMALL_CFG, 77), precision="fp32", timestep_type="discrete",
                  schedule="cosine", num_training_steps=40)
    outs = []
    for flag in (False, True):
        d.device_noise = flag
        outs.append(d.sample(batch_size=2, num_steps=6, progress=False, rng=R.setup_rng([5, 6], "cuda"), mode="ddpm"))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
