"""Sampler parity: ddpm.sample / p_step / q_step / repaint (CUDA) vs the reference's golden
trajectories (tests/golden/sampler.pt), with the reference's own CPU generators supplying the noise
("teacher forcing": CPU generators draw on the CPU and are copied to the GPU, diffusion.py randn).

Tolerances, l2-relative to the fp32 CPU reference, fp32(tf32) engine:
  single p_step                     : 5e-3
  6-step trajectories / repaint     : 2e-2  (errors of the stochastic steps compound)
The sampler arithmetic itself (given identical predictions) is checked to 2e-6.
"""
import os

import pytest
import torch

from oracle import r2dm_oracle as O
from tests.helpers import GOLDEN, SMALL_CFG, draw_noise, rel_l2, repaint_masks
from tests.util_model import make_ddpm

pytestmark = pytest.mark.gpu
B = 2


def sub(t):
    return t[..., ::2, ::7]


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "sampler.pt"))


@pytest.fixture(scope="module")
def sd():
    return O.random_state_dict(SMALL_CFG, 77)


@pytest.fixture(scope="module")
def ddpm(sd):
    return make_ddpm(SMALL_CFG, sd, precision="fp32")


def cpu_rng(base):
    return [torch.Generator().manual_seed(base + i) for i in range(B)]


@pytest.mark.parametrize("mode,eta", [("ddpm", 0.0), ("ddim", 0.0), ("ddim", 0.5)])
@pytest.mark.parametrize("graph", [True, False])
def test_sample_matches_reference(ddpm, golden, mode, eta, graph):
    gd = golden[f"sample_{mode}_{eta}"]
    ddpm.use_cuda_graph = graph
    ys = ddpm.sample(batch_size=B, num_steps=gd["steps"], progress=False, rng=cpu_rng(100),
                     return_all=True, mode=mode, ddim_eta=eta)
    torch.cuda.synchronize()
    ddpm.use_cuda_graph = True
    assert ys.shape[0] == gd["steps"] + 1
    # step 1 divides by alpha(t=1) = 5.5e-4 before clipping: prediction error is
    # amplified wherever x0 is not saturated, so it gets the trajectory tolerance, not the p_step one
    assert rel_l2(sub(ys[1]), gd["step1_sub"]) <= 2e-2
    e = rel_l2(ys[-1], gd["final"])
    assert e <= 3e-2, f"{mode}/{eta}: final l2-rel {e:.3e}"


def test_sample_graph_equals_eager(ddpm):
    a = ddpm.sample(batch_size=B, num_steps=5, progress=False, rng=cpu_rng(5), mode="ddpm")
    ddpm.use_cuda_graph = False
    b = ddpm.sample(batch_size=B, num_steps=5, progress=False, rng=cpu_rng(5), mode="ddpm")
    ddpm.use_cuda_graph = True
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_sample_multi_step_graph_cache(ddpm):
    """sample() replays cached CUDA graphs of `graph_steps` steps: the result must not depend on the chunk
    size, on whether the cache is cold or warm, on calls with other step counts / batch sizes in between,
    and `return_all` must still deliver every intermediate state (N + 1 entries, x_T first)."""
    N = 11
    ddpm.graph_steps = 1
    ddpm._loop_state = None
    ref = ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=cpu_rng(3), mode="ddim", ddim_eta=0.3)
    for k in (4, 8):                                    # 11 steps = 1 eager + 4+4+2 / 8+2 (remainder graphs)
        ddpm.graph_steps = k
        ddpm._loop_state = None
        cold = ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=cpu_rng(3), mode="ddim", ddim_eta=0.3)
        ddpm.sample(batch_size=B, num_steps=5, progress=False, rng=cpu_rng(4), mode="ddpm")       # other N, other mode
        ddpm.sample(batch_size=1, num_steps=3, progress=False, rng=cpu_rng(4)[:1], mode="ddpm")   # other batch: rebind
        warm = ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=cpu_rng(3), mode="ddim", ddim_eta=0.3)
        torch.cuda.synchronize()
        assert torch.equal(cold, ref), f"graph_steps={k}: cold cache differs from single-step graphs"
        assert torch.equal(warm, ref), f"graph_steps={k}: warm cache differs from single-step graphs"
    traj = ddpm.sample(batch_size=B, num_steps=N, progress=False, rng=cpu_rng(3), mode="ddim", ddim_eta=0.3,
                       return_all=True)
    assert traj.shape[0] == N + 1 and torch.equal(traj[-1], ref)
    ddpm.graph_steps = 8


def test_sample_cuda_generators_batch_split(ddpm):
    """rng = per-sample CUDA generators (utils/inference.py:113-114): sample i depends on seed i only."""
    import r2dm_b200 as R
    full = ddpm.sample(batch_size=3, num_steps=4, progress=False, rng=R.setup_rng([7, 8, 9], "cuda"), mode="ddim")
    one = ddpm.sample(batch_size=1, num_steps=4, progress=False, rng=R.setup_rng([8], "cuda"), mode="ddim")
    torch.cuda.synchronize()
    assert torch.equal(full[1:2], one)
    single = ddpm.sample(batch_size=2, num_steps=3, progress=False, rng=torch.Generator("cuda").manual_seed(1))
    none = ddpm.sample(batch_size=2, num_steps=3, progress=False)
    assert single.shape == none.shape == (2, 2, *SMALL_CFG.resolution)
    assert torch.isfinite(single).all() and torch.isfinite(none).all()


@pytest.mark.parametrize("obj", ["eps", "v", "x_0"])
@pytest.mark.parametrize("sched", ["cosine", "linear"])
@pytest.mark.parametrize("mode", ["ddpm", "ddim"])
def test_p_step_matches_reference(sd, golden, obj, sched, mode):
    d = make_ddpm(SMALL_CFG, sd, precision="fp32", schedule=sched, objective=obj)
    g = torch.Generator().manual_seed(9)
    x_t = torch.randn(B, 2, *SMALL_CFG.resolution, generator=g)
    t, s = torch.tensor([0.9, 0.4]), torch.tensor([0.8, 0.35])
    y = d.p_step(x_t.cuda(), t, s, rng=cpu_rng(300), mode=mode, ddim_eta=0.3)
    torch.cuda.synchronize()
    e = rel_l2(sub(y), golden[f"p_step_{obj}_{sched}_{mode}"]["y_sub"])
    assert e <= 5e-3, f"{obj}/{sched}/{mode}: {e:.3e}"


def test_sampler_update_arithmetic_exact(ddpm, sd):
    """Given the SAME prediction, the fused update must match the oracle formula to fp32 rounding."""
    from r2dm_b200.diffusion import continuous_coefficients
    g = torch.Generator().manual_seed(4)
    shape = (B, 2, *SMALL_CFG.resolution)
    x_t, pred, noise = (torch.randn(*shape, generator=g) for _ in range(3))
    t, s = torch.tensor([0.95, 0.3]), torch.tensor([0.9, 0.25])
    lt, ls = O.log_snr(t), O.log_snr(s)
    for mode, eta in (("ddpm", 0.0), ("ddim", 0.0), ("ddim", 0.7)):
        for obj in ("eps", "v", "x_0"):
            ref = O.p_step_update(x_t.double(), pred.double(), noise.double(), lt.double(), ls.double(),
                                  mode, eta, obj, 1.0)
            coef = continuous_coefficients(lt, ls, mode, eta, obj).float().cuda()
            out = torch.empty_like(x_t).cuda()
            ddpm.objective_backup = None
            ddpm._update(out, x_t.cuda(), pred.cuda(), noise.cuda(), coef, None, 0, 1)
            torch.cuda.synchronize()
            assert rel_l2(out, ref) <= 2e-6, (mode, eta, obj, rel_l2(out, ref))


def test_q_steps(ddpm, golden):
    g = torch.Generator().manual_seed(10)
    x0 = torch.randn(B, 2, *SMALL_CFG.resolution, generator=g).clamp(-1, 1)
    t, s = torch.tensor([0.7, 0.2]), torch.tensor([0.6, 0.1])
    rng = cpu_rng(400)
    xt, nz = ddpm.q_step_from_x_0(x0.cuda(), t, rng=rng)
    xq = ddpm.q_step(x0.cuda(), t, s, rng=rng)
    torch.cuda.synchronize()
    assert rel_l2(sub(xt), golden["q"]["xt_sub"]) <= 2e-6
    assert rel_l2(sub(xq), golden["q"]["xq_sub"]) <= 2e-6


@pytest.mark.parametrize("n,r,j", [(4, 2, 1), (3, 2, 2)])
def test_repaint_matches_reference(ddpm, golden, n, r, j):
    g = torch.Generator().manual_seed(10)
    known = torch.randn(B, 2, *SMALL_CFG.resolution, generator=g).clamp(-1, 1)
    mask = repaint_masks(B, SMALL_CFG)
    gd = golden[f"repaint_{n}_{r}_{j}"]
    ys = ddpm.repaint(known.cuda(), mask.cuda(), num_steps=n, num_resample_steps=r, jump_length=j,
                      progress=False, rng=cpu_rng(500), return_all=True)
    torch.cuda.synchronize()
    assert ys.shape[0] == gd["n_states"]
    e = rel_l2(ys[-1], gd["final"])
    assert e <= 2e-2, f"repaint {n},{r},{j}: {e:.3e}"
    # the known region of the result is the (re-noised at s=0, i.e. ~clean) known image
    y = ys[-1].cpu()
    assert (y - known)[mask.bool()].abs().max() < 5e-3


@pytest.mark.parametrize("sched", ["linear", "cosine", "sigmoid"])
@pytest.mark.parametrize("mode,eta", [("ddpm", 0.0), ("ddim", 0.0), ("ddim", 0.7)])
def test_discrete_p_step(sd, golden, sched, mode, eta):
    d = make_ddpm(SMALL_CFG, sd, precision="fp32", timestep_type="discrete", schedule=sched,
                  num_training_steps=40)
    g = torch.Generator().manual_seed(12)
    x_t = torch.randn(B, 2, *SMALL_CFG.resolution, generator=g)
    y = d.p_step(x_t.cuda(), torch.tensor([17, 0]), rng=cpu_rng(600), mode=mode, eta=eta)
    torch.cuda.synchronize()
    e = rel_l2(sub(y), golden[f"discrete_{sched}_{mode}_{eta}"]["y_sub"])
    assert e <= 5e-3, f"{sched}/{mode}/{eta}: {e:.3e}"


def test_discrete_sample(sd, golden):
    d = make_ddpm(SMALL_CFG, sd, precision="fp32", timestep_type="discrete", schedule="cosine",
                  num_training_steps=40)
    y = d.sample(batch_size=B, num_steps=4, progress=False, rng=cpu_rng(700), mode="ddpm")
    torch.cuda.synchronize()
    e = rel_l2(y, golden["discrete_sample_ddpm"]["y"])
    assert e <= 2e-2, f"{e:.3e}"


def test_bf16_trajectory_tracks_fp32(sd):
    """bf16 engine vs fp32(tf32) engine over a DDIM trajectory with shared noise: stated bound 6e-2."""
    a = make_ddpm(SMALL_CFG, sd, precision="fp32")
    b = make_ddpm(SMALL_CFG, sd, precision="bf16")
    ya = a.sample(batch_size=B, num_steps=6, progress=False, rng=cpu_rng(1), mode="ddim")
    yb = b.sample(batch_size=B, num_steps=6, progress=False, rng=cpu_rng(1), mode="ddim")
    torch.cuda.synchronize()
    assert rel_l2(yb, ya) <= 6e-2


def test_lidar_postprocess(sd):
    import r2dm_b200 as R
    gd = torch.load(os.path.join(GOLDEN, "lidar.pt"))
    g = torch.Generator().manual_seed(21)
    for fmt in ("log_depth", "inverse_depth", "depth"):
        torch.rand(2, 1, 64, 1024, generator=g)   # keep the generator in step with make_golden.py
    sample = torch.rand(2, 2, 64, 1024, generator=g) * 2 - 1
    for fmt in ("log_depth", "depth", "inverse_depth"):
        lu = R.LiDARUtility((64, 1024), fmt, 1.45, 80.0).cuda()
        out = lu.postprocess(sample.cuda())
        ref = O.lidar_postprocess(sample, lu.ray_angles.cpu(), fmt, 1.45, 80.0)
        torch.cuda.synchronize()
        assert rel_l2(out, ref) <= 1e-5, fmt
        if fmt == "log_depth":
            assert rel_l2(out[..., ::5, ::17], gd["postprocess_log_sub"]) <= 1e-5
