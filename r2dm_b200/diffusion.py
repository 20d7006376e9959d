"""Gaussian diffusion samplers — host-side mirror of models/diffusion/{base,continuous_time,
discrete_time}.py of the reference, driving the CUDA U-Net and the fused sampler-update kernels.

Public protocol kept from the reference: `sample`, `repaint`, `p_step`, `q_step`,
`q_step_from_x_0`, `randn`, `randn_like`, `log_snr`, `device`, `sampling_shape`, `model(x, cond)`,
`load_state_dict` of reference checkpoints (keys `model.*`, `_dummy`, discrete-time tables).
Training-only members (`p_loss`, `forward`, `get_target`, `get_loss_weight`, `sample_timesteps`) are
out of scope and raise NotImplementedError.

How the hot loop runs (no host sync between steps):
  * the network condition of step i depends only on i, so the time-embedding MLP and all 24 AdaGN
    projections are evaluated for all N steps up front into a FiLM table [N, F] (one kernel pair);
  * the per-step scalar algebra of p_step (alpha/sigma, x0 coefficients, DDPM/DDIM update) is folded
    on the host, in fp64, into a coefficient table [N, 5(+2)];
  * one step = U-Net forward + fused update + counter increment, captured once in a CUDA graph and
    replayed N times; the kernels pick their table row through a device-side step counter;
  * noise: with `rng` = a list of per-sample CUDA generators on the model's device (what
    utils/inference.py:113-114 builds) the draws happen INSIDE the update kernel - Philox4x32-10 +
    Box-Muller laid out like ATen's normal_ kernel, bit-identical to the reference's
    `torch.randn(C,H,W, generator=g_i)` per sample per step (base.py:71-94) - and the generators'
    offsets are advanced on the host afterwards, so the whole loop is graph replays with nothing in
    between.  Any other `rng` (None, one generator, CPU generators) is drawn on the host exactly like
    the reference does, into the graph's static noise buffer.
"""
from __future__ import annotations

import math
from functools import partial
from typing import List, Optional

import torch
from torch import nn

from . import _lib as L

# `sample_and_save.py:45` wraps `ddpm.sample` in torch.compile: the loop here already IS hand-written kernels
# inside CUDA graphs and talks to them through ctypes, which a tracing compiler cannot see - so the public
# entry points opt out of tracing and `torch.compile(ddpm.sample)` simply calls them.
_no_compile = getattr(getattr(torch, "compiler", None), "disable", None) or (lambda fn: fn)

try:  # progress bars are optional plumbing
    from tqdm.auto import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kwargs):
        return it


# ------------------------------------------------------------------------------------- schedules
def _log(t: torch.Tensor, eps: float = 1e-20) -> torch.Tensor:
    return torch.log(t.clamp(min=eps))


def _log_snr_schedule_linear(t: torch.Tensor) -> torch.Tensor:
    """continuous_time.py:18-19."""
    return -_log(torch.special.expm1(1e-4 + 10 * (t ** 2)))[:, None, None, None]


def _log_snr_schedule_cosine(t: torch.Tensor, logsnr_min: float = -15, logsnr_max: float = 15) -> torch.Tensor:
    """continuous_time.py:22-29: lambda(t) = -2 log tan(t_min + t (t_max - t_min))."""
    t_min = math.atan(math.exp(-0.5 * logsnr_max))
    t_max = math.atan(math.exp(-0.5 * logsnr_min))
    return -2 * _log(torch.tan(t_min + t * (t_max - t_min)))[:, None, None, None]


def _log_snr_schedule_cosine_shifted(t, image_d, noise_d, logsnr_min=-15, logsnr_max=15):
    """continuous_time.py:32-41."""
    return _log_snr_schedule_cosine(t, logsnr_min, logsnr_max) + 2 * math.log(noise_d / image_d)


def _log_snr_schedule_cosine_interpolated(t, image_d, noise_d_low, noise_d_high, logsnr_min=-15, logsnr_max=15):
    """continuous_time.py:44-58."""
    low = _log_snr_schedule_cosine_shifted(t, image_d, noise_d_low, logsnr_min, logsnr_max)
    high = _log_snr_schedule_cosine_shifted(t, image_d, noise_d_high, logsnr_min, logsnr_max)
    tt = t[:, None, None, None]
    return tt * low + (1 - tt) * high


def _log_snr_to_alpha_sigma(log_snr: torch.Tensor):
    """continuous_time.py:61-63."""
    return log_snr.sigmoid().sqrt(), (-log_snr).sigmoid().sqrt()


def continuous_coefficients(lam_t: torch.Tensor, lam_s: torch.Tensor, mode: str, eta: float,
                            objective: str) -> torch.Tensor:
    """Fold continuous_time.py:203-229 into per-row scalars [n, 5] = (ux, up, kx, k0, kn):
        x0 = clamp(ux x_t + up pred);  x_s = kx x_t + k0 x0 + kn noise.   Evaluated in fp64."""
    lt, ls = lam_t.double().flatten(), lam_s.double().flatten()
    a_t, s_t = _log_snr_to_alpha_sigma(lt)
    a_s, s_s = _log_snr_to_alpha_sigma(ls)
    if objective == "eps":
        ux, up = 1 / a_t, -s_t / a_t
    elif objective == "v":
        ux, up = a_t, -s_t
    elif objective == "x_0":
        ux, up = torch.zeros_like(a_t), torch.ones_like(a_t)
    else:
        raise ValueError(f"invalid objective {objective}")
    if mode == "ddpm":
        c = -torch.special.expm1(lt - ls)
        kx, k0, kn = a_s * (1 - c) / a_t, a_s * c, s_s * c.sqrt()
    elif mode == "ddim":
        c1 = eta * s_s / s_t * (1 - a_t ** 2 / a_s ** 2).clamp(min=0).sqrt()
        c2 = (1 - a_s ** 2 - c1 ** 2).clamp(min=0).sqrt()
        kx, k0, kn = c2 / s_t, a_s - c2 * a_t / s_t, c1
    else:
        raise ValueError(f"invalid mode {mode}")
    return torch.stack([ux, up, kx, k0, kn], dim=1)


# ------------------------------------------------------------------------------------- device noise
def _torch_randn_geometry(numel: int, device) -> tuple:
    """(threads, offset_per_draw) of ATen's CUDA normal_ kernel for a tensor of `numel` elements
    (ATen/native/cuda/DistributionTemplates.h, distribution_nullary_kernel: block 256, unroll 4, grid
    capped at SMs * maxThreadsPerSM / 256; every call advances the Philox offset by
    ceil(numel / (threads * 4)) * 4)."""
    props = torch.cuda.get_device_properties(device)
    block, unroll = 256, 4
    grid = min((numel + block - 1) // block,
               props.multi_processor_count * (props.max_threads_per_multi_processor // block))
    threads = grid * block
    return threads, ((numel - 1) // (threads * unroll) + 1) * 4


def _cuda_gen_seed_offset(g: torch.Generator) -> tuple:
    """(seed, philox offset) of a CUDA generator (state = two little-endian uint64)."""
    st = g.get_state()
    assert st.numel() == 16, "unexpected CUDA generator state layout"
    v = st.view(torch.int64)
    return int(v[0].item()) & (2 ** 64 - 1), int(v[1].item()) & (2 ** 64 - 1)


def _cuda_gen_advance(g: torch.Generator, delta: int) -> None:
    st = g.get_state().clone()
    v = st.view(torch.int64)
    off = ((int(v[1].item()) & (2 ** 64 - 1)) + delta) & (2 ** 64 - 1)
    v[1] = off - 2 ** 64 if off >= 2 ** 63 else off
    g.set_state(st)


def _u64_tensor(vals, device) -> torch.Tensor:
    return torch.tensor([v - 2 ** 64 if v >= 2 ** 63 else v for v in vals], dtype=torch.int64, device=device)


# ------------------------------------------------------------------------------------- base
class GaussianDiffusion(nn.Module):
    """Mirror of models/diffusion/base.py:9-163 (sampling members only)."""

    def __init__(
        self,
        model: nn.Module,
        sampling: str = "ddpm",
        prediction_type: str = "eps",
        loss_type="l2",
        num_training_steps: Optional[int] = 1000,
        noise_schedule: str = "linear",
        min_snr_loss_weight: bool = True,
        min_snr_gamma: float = 5.0,
        sampling_resolution=None,
        clip_sample: bool = True,
        clip_sample_range: float = 1,
    ):
        super().__init__()
        self.model = model
        self.sampling = sampling
        self.num_training_steps = num_training_steps
        self.objective = prediction_type
        self.noise_schedule = noise_schedule
        self.min_snr_loss_weight = min_snr_loss_weight
        self.min_snr_gamma = min_snr_gamma
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.loss_type = loss_type
        if prediction_type not in ("eps", "v", "x_0"):
            raise ValueError(f"invalid objective {prediction_type}")
        if sampling_resolution is None:
            assert hasattr(self.model, "resolution")
            assert hasattr(self.model, "in_channels")
            self.sampling_shape = (self.model.in_channels, *self.model.resolution)
        else:
            assert len(sampling_resolution) == 2
            assert hasattr(self.model, "in_channels")
            self.sampling_shape = (self.model.in_channels, *sampling_resolution)
        self.use_cuda_graph = True
        self.graph_steps = 8   # denoising steps captured per CUDA graph in sample() (host-drawn noise)
        self.graph_steps_device_noise = 16   # ... when the noise is drawn inside the update kernel
        self.device_noise = True   # False: always draw on the host (debugging / A-B tests)
        self._graphs = {}
        self.setup_parameters()
        self.register_buffer("_dummy", torch.tensor([]))

    @property
    def device(self):
        return self._dummy.device

    # -- random draws: base.py:71-94 ------------------------------------------------------------
    def randn(self, *shape, rng=None, out: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        """None / one Generator / list of per-sample Generators (len == batch).  Generators living on
        another device (e.g. CPU generators with a CUDA model) draw on their own device and are
        copied over, which makes CPU-reference trajectories reproducible on the GPU."""
        dev = kwargs.pop("device", None)
        dtype = kwargs.pop("dtype", torch.float32)

        def draw(shp, gen):
            gdev = gen.device if gen is not None else dev
            t = torch.randn(*shp, generator=gen, device=gdev, dtype=dtype)
            return t

        if rng is None:
            res = draw(shape, None)
        elif isinstance(rng, torch.Generator):
            res = draw(shape, rng)
        elif isinstance(rng, list):
            assert len(rng) == shape[0]
            if out is not None and all(r.device == out.device for r in rng):
                for i, r in enumerate(rng):
                    torch.randn(*shape[1:], generator=r, out=out[i])
                return out
            res = torch.stack([draw(shape[1:], r) for r in rng])
        else:
            raise ValueError(f"invalid rng: {rng}")
        if out is not None:
            out.copy_(res, non_blocking=True)
            return out
        return res.to(dev) if dev is not None else res

    def randn_like(self, x: torch.Tensor, rng=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.randn(*x.shape, rng=rng, out=out, device=x.device, dtype=x.dtype)

    # -- training-only API: out of scope ----------------------------------------------------------
    def setup_parameters(self) -> None:
        raise NotImplementedError

    def _training_only(self, *a, **k):
        raise NotImplementedError("training (p_loss / forward) is outside the scope of r2dm_b200; "
                                  "this package accelerates the sampling path only")

    p_loss = forward = get_target = get_loss_weight = sample_timesteps = _training_only

    # -- engine helpers ---------------------------------------------------------------------------
    def _engine(self):
        return self.model.engine()

    def _clip(self) -> float:
        return float(self.clip_sample_range) if self.clip_sample else 0.0

    def _device_noise_ok(self, rng, dev) -> bool:
        return (bool(getattr(self, "device_noise", True)) and isinstance(rng, list) and len(rng) > 0
                and dev.type == "cuda" and all(isinstance(r, torch.Generator) and r.device == dev for r in rng))

    def _philox(self, st, rng, per, dev, ctr0=None, mul0=0, ctr1=None, mul1=0):
        """Fill the loop state's static seed / offset tensors from the generators' CURRENT state and
        return (r2dm_philox struct, offset_per_draw)."""
        so = [_cuda_gen_seed_offset(r) for r in rng]
        st["seeds"].copy_(_u64_tensor([a for a, _ in so], dev))
        st["offsets"].copy_(_u64_tensor([b for _, b in so], dev))
        threads, inc = _torch_randn_geometry(per, dev)
        ph = L.R2dmPhilox(L.ptr(st["seeds"]), L.ptr(st["offsets"]), L.ptr(ctr0) if ctr0 is not None else None,
                          L.ptr(ctr1) if ctr1 is not None else None, mul0, mul1, inc, threads)
        return ph, inc

    def _update_philox(self, x_out, x, pred, coef, step_ptr, rows_per_step, row_batch_stride, ph,
                       draw_noise=0, draw_noise2=0, known=None, mask=None):
        import ctypes
        L.check(L.lib().r2dm_sampler_update_philox(
            L.ptr(x_out), L.ptr(x), L.ptr(pred), L.ptr(coef), coef.shape[1],
            L.ptr(step_ptr) if step_ptr is not None else None, rows_per_step, row_batch_stride, self._clip(),
            L.ptr(known) if known is not None else None, L.ptr(mask) if mask is not None else None,
            ctypes.byref(ph), draw_noise, draw_noise2, x.shape[0], x[0].numel(), L.stream_ptr()),
            "r2dm_sampler_update_philox")
        return x_out

    def _update(self, x_out, x, pred, noise, coef, step_ptr, rows_per_step, row_batch_stride,
                known=None, mask=None, noise2=None):
        B = x.shape[0]
        per = x[0].numel()
        L.check(L.lib().r2dm_sampler_update(
            L.ptr(x_out), L.ptr(x), L.ptr(pred), L.ptr(noise), L.ptr(coef), coef.shape[1],
            L.ptr(step_ptr) if step_ptr is not None else None, rows_per_step, row_batch_stride,
            self._clip(), L.ptr(known) if known is not None else None,
            L.ptr(mask) if mask is not None else None, L.ptr(noise2) if noise2 is not None else None,
            B, per, L.stream_ptr()), "r2dm_sampler_update")
        return x_out

    def _axpby(self, x, noise, ac):
        y = torch.empty_like(x)
        L.check(L.lib().r2dm_axpby(L.ptr(y), L.ptr(x), L.ptr(noise), L.ptr(ac), x.shape[0], x[0].numel(),
                                   L.stream_ptr()), "r2dm_axpby")
        return y

    def _run_table(self, x, conds, coefs, rng, return_all, progress, desc, draw_noise):
        """Shared hot loop: N steps over precomputed (cond, coefficient) tables with a device-side
        step counter.  The loop runs as CUDA graphs of `graph_steps` denoising steps; graphs and their
        static buffers (state, prediction, noise, tables) are cached on this object and reused by later
        calls with the same batch size, so a call costs N step replays and nothing else."""
        eng = self._engine()
        dev = x.device
        N = conds.numel()
        B = x.shape[0]
        with torch.cuda.device(dev):
            lib = L.lib()
            philox = draw_noise and self._device_noise_ok(rng, dev)
            K = 1 if return_all else max(1, int(getattr(self, "graph_steps_device_noise", 16) if philox
                                                else getattr(self, "graph_steps", 8)))
            ncoef = coefs.shape[1]
            key = (B, K, ncoef, self._clip(), tuple(x.shape[1:]), philox)
            eng.bind(B)
            st = getattr(self, "_loop_state", None)
            if (st is None or st["key"] != key or st["eng"] is not eng or st["epoch"] != eng.bind_epoch
                    or st["film"].shape[0] < N):
                st = {
                    "key": key, "eng": eng, "epoch": eng.bind_epoch,
                    "x": torch.empty((B,) + tuple(x.shape[1:]), device=dev, dtype=torch.float32),
                    "pred": torch.empty((B,) + tuple(x.shape[1:]), device=dev, dtype=torch.float32),
                    "noise": torch.zeros((1 if philox else K, B) + tuple(x.shape[1:]), device=dev,
                                         dtype=torch.float32),
                    "seeds": torch.zeros(B, dtype=torch.int64, device=dev),
                    "offsets": torch.zeros(B, dtype=torch.int64, device=dev),
                    "film": torch.zeros(max(N, 256), eng.film_width, device=dev, dtype=torch.float32),
                    "coef": torch.zeros(max(N, 256), ncoef, device=dev, dtype=torch.float32),
                    "step": torch.zeros(1, dtype=torch.int32, device=dev),
                    "graphs": {}, "warm": False,
                }
                self._loop_state = st
            xs, pred, noise, step = st["x"], st["pred"], st["noise"], st["step"]
            st["film"][:N].copy_(eng.cond_embed(conds.to(dev)))
            st["coef"][:N].copy_(coefs.to(device=dev, dtype=torch.float32))
            xs.copy_(L.f32c(x))
            step.zero_()
            if not draw_noise:
                noise.zero_()
            ph = None
            if philox:
                # draw number of step i is i (one draw per step; the x_T draw has already been consumed)
                ph, inc = self._philox(st, rng, xs[0].numel(), dev, ctr0=step, mul0=1)
                st["ph"] = ph   # keeps the struct alive; graphs hold only the device pointers inside it
            out = [xs.clone()] if return_all else None

            def one_step(k=0):
                eng.forward_film(xs, st["film"], pred, step_ptr=step, rows_per_step=1, row_batch_stride=0)
                if philox:
                    self._update_philox(xs, xs, pred, st["coef"], step, 1, 0, ph)
                else:
                    self._update(xs, xs, pred, noise[k], st["coef"], step, 1, 0)
                L.check(lib.r2dm_advance_step(L.ptr(step), 1, L.stream_ptr()))

            def draw(n):
                if draw_noise and not philox:
                    for k in range(n):
                        self.randn_like(xs, rng=rng, out=noise[k])

            def capture(n):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for k in range(n):
                        one_step(k)
                return g   # capture does not execute: the device counter is unchanged

            use_graph = self.use_cuda_graph and N > 2
            done = 0
            bar = tqdm(total=N, desc=desc, leave=False, disable=not progress)
            while done < N:
                warmup = use_graph and not st["warm"]
                n = 1 if warmup else min(K, N - done)
                draw(n)
                if warmup:
                    # very first step of this cache entry runs eagerly (lazy kernel attribute setup)
                    one_step(0)
                    torch.cuda.current_stream().synchronize()
                    st["warm"] = True
                elif use_graph:
                    if n not in st["graphs"]:
                        st["graphs"][n] = capture(n)
                    st["graphs"][n].replay()
                else:
                    for k in range(n):
                        one_step(k)
                done += n
                bar.update(n)
                if return_all:
                    out.append(xs.clone())
            bar.close()
            if philox:   # leave the generators where N host-side draws would have left them
                for r in rng:
                    _cuda_gen_advance(r, N * inc)
        return torch.stack(out) if return_all else xs.clone()


# ------------------------------------------------------------------------------------- continuous
class ContinuousTimeGaussianDiffusion(GaussianDiffusion):
    """Mirror of models/diffusion/continuous_time.py:66-317 (https://arxiv.org/abs/2107.00630)."""

    def __init__(
        self,
        model: nn.Module,
        prediction_type: str = "eps",
        loss_type="l2",
        noise_schedule: str = "cosine",
        min_snr_loss_weight: bool = True,
        min_snr_gamma: float = 5.0,
        sampling_resolution=None,
        clip_sample: bool = True,
        clip_sample_range: float = 1,
        image_d: float = None,
        noise_d_low: float = None,
        noise_d_high: float = None,
    ):
        # (the reference assigns these after super().__init__, which makes its shifted /
        #  interpolated schedules unconstructible, continuous_time.py:89-123; here they work)
        self._sched_args = dict(image_d=image_d, noise_d_low=noise_d_low, noise_d_high=noise_d_high)
        super().__init__(
            model=model, sampling="ddpm", prediction_type=prediction_type, loss_type=loss_type,
            num_training_steps=None, noise_schedule=noise_schedule, min_snr_loss_weight=min_snr_loss_weight,
            min_snr_gamma=min_snr_gamma, sampling_resolution=sampling_resolution, clip_sample=clip_sample,
            clip_sample_range=clip_sample_range)
        self.image_d, self.noise_d_low, self.noise_d_high = image_d, noise_d_low, noise_d_high

    def setup_parameters(self) -> None:
        a = self._sched_args
        if self.noise_schedule == "linear":
            self.log_snr = _log_snr_schedule_linear
        elif self.noise_schedule == "cosine":
            self.log_snr = _log_snr_schedule_cosine
        elif self.noise_schedule == "cosine_shifted":
            assert a["image_d"] is not None and a["noise_d_low"] is not None
            self.log_snr = partial(_log_snr_schedule_cosine_shifted, image_d=a["image_d"], noise_d=a["noise_d_low"])
        elif self.noise_schedule == "cosine_interpolated":
            assert a["image_d"] is not None and a["noise_d_low"] is not None and a["noise_d_high"] is not None
            self.log_snr = partial(_log_snr_schedule_cosine_interpolated, image_d=a["image_d"],
                                   noise_d_low=a["noise_d_low"], noise_d_high=a["noise_d_high"])
        else:
            raise ValueError(f"invalid beta schedule: {self.noise_schedule}")

    def get_network_condition(self, steps):
        return self.log_snr(steps)[:, 0, 0, 0]

    def _lam(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 log-SNR of times t ([n]) evaluated on the host like the reference evaluates it."""
        return self.log_snr(t.detach().float().cpu())[:, 0, 0, 0]

    # -- forward process ---------------------------------------------------------------------------
    @_no_compile
    def q_step_from_x_0(self, x_0, step_t, rng=None):
        """continuous_time.py:169-176: x_t = alpha x_0 + sigma eps; returns (x_t, eps)."""
        x_0 = L.f32c(x_0)
        noise = self.randn_like(x_0, rng=rng)
        alpha, sigma = _log_snr_to_alpha_sigma(self._lam(step_t).double())
        ac = torch.stack([alpha, sigma], dim=1).float().to(x_0.device)
        with torch.cuda.device(x_0.device):
            return self._axpby(x_0, noise, ac), noise

    @_no_compile
    def q_step(self, x_s, step_t, step_s, rng=None):
        """continuous_time.py:178-190: q(z_t | z_s), 0 < s < t < 1."""
        x_s = L.f32c(x_s)
        a_t, s_t = _log_snr_to_alpha_sigma(self._lam(step_t).double())
        a_s, s_s = _log_snr_to_alpha_sigma(self._lam(step_s).double())
        a_ts = a_t / a_s
        var = s_t ** 2 - a_ts ** 2 * s_s ** 2
        noise = self.randn_like(x_s, rng=rng)
        ac = torch.stack([a_ts, var.clamp(min=0).sqrt()], dim=1).float().to(x_s.device)
        with torch.cuda.device(x_s.device):
            return self._axpby(x_s, noise, ac)

    # -- reverse process ---------------------------------------------------------------------------
    @_no_compile
    @torch.inference_mode()
    def p_step(self, x_t, step_t, step_s, rng=None, mode="ddpm", ddim_eta: float = 0.0,
               _known=None, _mask=None):
        """continuous_time.py:192-232: one reverse step p(z_s | z_t) with per-sample times."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        x_t = L.f32c(x_t)
        lam_t, lam_s = self._lam(step_t), self._lam(step_s)
        coef = continuous_coefficients(lam_t, lam_s, mode, ddim_eta, self.objective)
        eng = self._engine()
        with torch.cuda.device(x_t.device):
            noise2 = None
            if _known is not None:  # RePaint: the known-region draw comes first (continuous_time.py:296)
                noise2 = self.randn_like(x_t, rng=rng)
                a_s, s_s = _log_snr_to_alpha_sigma(lam_s.double())
                coef = torch.cat([coef, a_s[:, None], s_s[:, None]], dim=1)
            film = eng.cond_embed(lam_t.to(x_t.device))
            pred = torch.empty_like(x_t)
            eng.forward_film(x_t, film, pred, step_ptr=None, rows_per_step=0, row_batch_stride=1)
            noise = self.randn_like(x_t, rng=rng)
            coef = coef.float().contiguous().to(x_t.device)
            x_s = torch.empty_like(x_t)
            self._update(x_s, x_t, pred, noise, coef, None, 0, 1, _known, _mask, noise2)
        return x_s

    @_no_compile
    @torch.inference_mode()
    def sample(self, batch_size: int, num_steps: int, progress: bool = True, rng=None,
               return_all: bool = False, mode: str = "ddpm", ddim_eta: float = 0.0):
        """continuous_time.py:234-258."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        x = self.randn(batch_size, *self.sampling_shape, rng=rng, device=self.device)
        steps = torch.linspace(1.0, 0.0, num_steps + 1)
        lam = self._lam(steps)
        coefs = continuous_coefficients(lam[:-1], lam[1:], mode, ddim_eta, self.objective)
        return self._run_table(x, lam[:-1], coefs, rng, return_all, progress, "sampling", draw_noise=True)

    @_no_compile
    @torch.inference_mode()
    def repaint(self, known, mask, num_steps: int, num_resample_steps: int = 1, jump_length: int = 1,
                progress: bool = True, rng=None, return_all: bool = False):
        """continuous_time.py:260-317 (RePaint, https://arxiv.org/abs/2201.09865); mask == 1 is known.

        The (t, s) sequence of all reverse / re-noising steps is static, so like `sample` the loop runs
        from precomputed tables: reverse steps (U-Net + update + known-region blend) and forward steps
        (q_step re-noising) are two CUDA graphs driven by device-side counters.  Draw order per reverse
        step is the reference's: known-region noise first, then the p_step noise."""
        assert num_resample_steps > 0
        assert jump_length > 0
        batch_size = known.shape[0]
        dev = self.device
        known = L.f32c(known.to(dev))
        mask = L.f32c(mask.to(dev).expand_as(known))
        x = L.f32c(self.randn(batch_size, *self.sampling_shape, rng=rng, device=dev)).clone()
        steps = torch.linspace(1, 0, num_steps + 1)
        interp = torch.linspace(0, 1, jump_length + 1)
        # ---- static schedule: 'p' = reverse step r[k] -> r[k+1]; 'q' = re-noise r[k] -> r[k-1]
        prog, p_t, p_s, q_t, q_s = [], [], [], [], []
        for i in range(num_steps):
            for j in range(num_resample_steps):
                r = steps[i] + interp * (steps[i + 1] - steps[i])
                for k in range(jump_length):
                    prog.append("p"); p_t.append(r[k]); p_s.append(r[k + 1])
                prog.append("out")
                if (i == num_steps - 1) or (j == num_resample_steps - 1):
                    break
                for k in range(jump_length, 0, -1):
                    prog.append("q"); q_t.append(r[k - 1]); q_s.append(r[k])
        lam_pt, lam_ps = self._lam(torch.stack(p_t)), self._lam(torch.stack(p_s))
        coef = continuous_coefficients(lam_pt, lam_ps, "ddpm", 0.0, self.objective)
        a_s, s_s = _log_snr_to_alpha_sigma(lam_ps.double())
        coef = torch.cat([coef, a_s[:, None], s_s[:, None]], dim=1)           # known_s = a_s known + s_s noise2
        eng = self._engine()
        lib = L.lib()
        with torch.cuda.device(dev):
            film = eng.cond_embed(lam_pt.to(dev))
            coef = coef.to(device=dev, dtype=torch.float32).contiguous()
            if q_t:
                a_t, s_t = _log_snr_to_alpha_sigma(self._lam(torch.stack(q_t)).double())
                a_q, s_q = _log_snr_to_alpha_sigma(self._lam(torch.stack(q_s)).double())
                a_ts = a_t / a_q
                qtab = torch.stack([a_ts, (s_t ** 2 - a_ts ** 2 * s_q ** 2).clamp(min=0).sqrt()], dim=1)
                qtab = qtab.to(device=dev, dtype=torch.float32).contiguous()
            pstep = torch.zeros(1, dtype=torch.int32, device=dev)
            qstep = torch.zeros(1, dtype=torch.int32, device=dev)
            pred, noise, noise2 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
            per = x[0].numel()
            philox = self._device_noise_ok(rng, dev)
            if philox:
                # draws so far = 2 per reverse step (known-region noise first, then the p_step noise,
                # continuous_time.py:296-299) + 1 per re-noising step
                pst = {"seeds": torch.zeros(batch_size, dtype=torch.int64, device=dev),
                       "offsets": torch.zeros(batch_size, dtype=torch.int64, device=dev)}
                ph, inc = self._philox(pst, rng, per, dev, ctr0=pstep, mul0=2, ctr1=qstep, mul1=1)

            def reverse_step():
                eng.forward_film(x, film, pred, step_ptr=pstep, rows_per_step=1, row_batch_stride=0)
                if philox:
                    self._update_philox(x, x, pred, coef, pstep, 1, 0, ph, draw_noise=1, draw_noise2=0,
                                        known=known, mask=mask)
                else:
                    self._update(x, x, pred, noise, coef, pstep, 1, 0, known, mask, noise2)
                L.check(lib.r2dm_advance_step(L.ptr(pstep), 1, L.stream_ptr()))

            def renoise_step():
                if philox:
                    import ctypes
                    L.check(lib.r2dm_axpby_table_philox(L.ptr(x), L.ptr(x), L.ptr(qtab), L.ptr(qstep), 1, 0,
                                                        ctypes.byref(ph), 0, batch_size, per, L.stream_ptr()),
                            "r2dm_axpby_table_philox")
                else:
                    L.check(lib.r2dm_axpby_table(L.ptr(x), L.ptr(x), L.ptr(noise), L.ptr(qtab), L.ptr(qstep), 1, 0,
                                                 batch_size, per, L.stream_ptr()), "r2dm_axpby_table")
                L.check(lib.r2dm_advance_step(L.ptr(qstep), 1, L.stream_ptr()))

            graphs = {"p": None, "q": None}
            fns = {"p": reverse_step, "q": renoise_step}
            seen = {"p": 0, "q": 0}
            out = [x.clone()] if return_all else None
            n_p = sum(1 for o in prog if o == "p")
            for op in tqdm(prog, desc="RePaint", leave=False, disable=not progress):
                if op == "out":
                    if return_all:
                        out.append(x.clone())
                    continue
                if not philox:
                    if op == "p":
                        self.randn_like(x, rng=rng, out=noise2)     # q_step_from_x_0(known) draw
                    self.randn_like(x, rng=rng, out=noise)
                if graphs[op] is not None:
                    graphs[op].replay()
                    continue
                fns[op]()                                       # first occurrence runs eagerly
                seen[op] += 1
                if self.use_cuda_graph and n_p > 2 and seen[op] == 1:
                    torch.cuda.current_stream().synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fns[op]()
                    graphs[op] = g
            if philox:
                n_q = sum(1 for o in prog if o == "q")
                for r in rng:
                    _cuda_gen_advance(r, (2 * n_p + n_q) * inc)
        return torch.stack(out) if return_all else x


# ------------------------------------------------------------------------------------- discrete
def _linear_beta_schedule(steps):
    scale = 1000 / steps
    return torch.linspace(scale * 0.0001, scale * 0.02, steps, dtype=torch.float64)


def _alpha_bar_to_beta(alphas_bar):
    alphas_bar = alphas_bar / alphas_bar[0]
    return torch.clip(1 - (alphas_bar[1:] / alphas_bar[:-1]), 0, 0.999)


def _cosine_beta_schedule(steps, s=0.008):
    t = torch.linspace(0, steps, steps + 1, dtype=torch.float64) / steps
    return _alpha_bar_to_beta(torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2)


def _sigmoid_beta_schedule(steps, start=-3, end=3, tau=1):
    t = torch.linspace(0, steps, steps + 1, dtype=torch.float64) / steps
    v0, v1 = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
    return _alpha_bar_to_beta((-((t * (end - start) + start) / tau).sigmoid() + v1) / (v1 - v0))


class DiscreteTimeGaussianDiffusion(GaussianDiffusion):
    """Mirror of models/diffusion/discrete_time.py:51-201 (https://arxiv.org/abs/2006.11239)."""

    def setup_parameters(self) -> None:
        assert self.num_training_steps is not None
        if self.noise_schedule == "linear":
            beta = _linear_beta_schedule(self.num_training_steps)
        elif self.noise_schedule == "cosine":
            beta = _cosine_beta_schedule(self.num_training_steps)
        elif self.noise_schedule == "sigmoid":
            beta = _sigmoid_beta_schedule(self.num_training_steps)
        else:
            raise ValueError(f"invalid beta schedule {self.noise_schedule}")
        beta = beta[:, None, None, None]
        alpha_bar = torch.cumprod(1 - beta, dim=0)
        alpha_bar_prev = torch.cat([torch.ones_like(alpha_bar[:1]), alpha_bar[:-1]])
        self.register_buffer("beta", beta.float())
        self.register_buffer("alpha_bar", alpha_bar.float())
        self.register_buffer("alpha_bar_prev", alpha_bar_prev.float())
        self.register_buffer("snr", (alpha_bar / (1 - alpha_bar)).float())

    def get_network_condition(self, steps: torch.Tensor) -> torch.Tensor:
        return steps

    def _coefficients(self, steps: torch.Tensor, mode: str, eta: float) -> torch.Tensor:
        """discrete_time.py:135-179 folded into (ux, up, kx, k0, kn) rows, fp64."""
        idx = steps.detach().long().cpu()
        beta = self.beta.flatten().cpu().double()[idx]
        ab = self.alpha_bar.flatten().cpu().double()[idx]
        abp = self.alpha_bar_prev.flatten().cpu().double()[idx]
        alpha = 1 - beta
        nz = (idx != 0).double()
        if self.objective == "eps":
            ux, up = ab.rsqrt(), -(ab.reciprocal() - 1).sqrt()
        elif self.objective == "x_0":
            ux, up = torch.zeros_like(ab), torch.ones_like(ab)
        elif self.objective == "v":
            ux, up = ab.sqrt(), -(1 - ab).sqrt()
        else:
            raise ValueError(f"invalid objective {self.objective}")
        if mode == "ddpm":
            k0 = abp.sqrt() * beta / (1 - ab)
            kx = (1 - abp) * alpha.sqrt() / (1 - ab)
            var = (beta * (1 - abp) / (1 - ab)).clamp(min=1e-20)
            kn = (0.5 * var.log()).exp() * nz
        elif mode == "ddim":
            var = (1 - abp) / (1 - ab) * (1 - ab / abp)
            std = eta * var.clamp(min=0).sqrt()
            c2 = (1 - abp - std ** 2).clamp(min=0).sqrt()
            kx = c2 / (1 - ab).sqrt()
            k0 = abp.sqrt() - c2 * ab.sqrt() / (1 - ab).sqrt()
            kn = std * nz if eta > 0 else torch.zeros_like(std)
        else:
            raise ValueError(f"invalid mode {mode}")
        return torch.stack([ux, up, kx, k0, kn], dim=1)

    @_no_compile
    def q_step_from_x_0(self, x_0, steps, rng=None):
        """discrete_time.py:119-124."""
        x_0 = L.f32c(x_0)
        noise = self.randn_like(x_0, rng=rng)
        ab = self.alpha_bar.flatten().cpu().double()[steps.detach().long().cpu()]
        ac = torch.stack([ab.sqrt(), (1 - ab).sqrt()], dim=1).float().to(x_0.device)
        with torch.cuda.device(x_0.device):
            return self._axpby(x_0, noise, ac), noise

    @_no_compile
    @torch.inference_mode()
    def p_step(self, x_t, steps, rng=None, mode="ddim", eta: float = 0.0):
        """discrete_time.py:126-180."""
        x_t = L.f32c(x_t)
        coef = self._coefficients(steps, mode, eta).float().contiguous().to(x_t.device)
        eng = self._engine()
        with torch.cuda.device(x_t.device):
            film = eng.cond_embed(steps.to(x_t.device).float())
            pred = torch.empty_like(x_t)
            eng.forward_film(x_t, film, pred, step_ptr=None, rows_per_step=0, row_batch_stride=1)
            draws = mode == "ddpm" or eta > 0   # the reference draws no noise for deterministic DDIM
            noise = self.randn_like(x_t, rng=rng) if draws else torch.zeros_like(x_t)
            x_s = torch.empty_like(x_t)
            self._update(x_s, x_t, pred, noise, coef, None, 0, 1)
        return x_s

    @_no_compile
    @torch.inference_mode()
    def sample(self, batch_size: int, num_steps: int, progress: bool = True, rng=None,
               return_all: bool = False, mode: str = "ddpm"):
        """discrete_time.py:182-201 (visits t = num_steps-1 ... 0 without re-spacing, like the reference)."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        x = self.randn(batch_size, *self.sampling_shape, rng=rng, device=self.device)
        ts = torch.arange(num_steps - 1, -1, -1)
        coefs = self._coefficients(ts, mode, 0.0)
        return self._run_table(x, ts.float(), coefs, rng, return_all, progress, "sampling",
                               draw_noise=(mode == "ddpm"))
