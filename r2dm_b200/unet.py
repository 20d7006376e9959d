"""EfficientUNet — host-side mirror of models/efficient_unet.py:188-295 of the reference.

Same constructor arguments, attributes (`resolution`, `in_channels`, `out_channels`, `coords`,
`coords_encoding`), state-dict keys / shapes (so reference checkpoints load with
`load_state_dict`) and call signature `forward(images[B,C,H,W], timesteps[B]) -> [B,C,H,W]`.
There is no PyTorch implementation of the network here: `forward` runs the hand-written sm_100a
kernels through the C ABI (`r2dm_unet_forward`) and raises if the CUDA extension is missing or the
module is not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import encoding


# ctypes calls into hand-written kernels: nothing for a tracing compiler to see (see diffusion.py)
_no_compile = getattr(getattr(torch, "compiler", None), "disable", None) or (lambda fn: fn)


def _n_tuple(x, N: int) -> tuple:
    if isinstance(x, Iterable):
        x = tuple(x)
        assert len(x) == N
        return x
    return (x,) * N


def _register(root: nn.Module, name: str, tensor: torch.Tensor, buffer: bool) -> None:
    """Create nested containers so that `root.state_dict()` yields exactly `name`."""
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, nn.Module())
        mod = getattr(mod, p)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class EfficientUNet(nn.Module):
    """B200-native Efficient U-Net for LiDAR range images (drop-in for the reference class)."""

    def __init__(
        self,
        in_channels: int,
        resolution,
        out_channels: Optional[int] = None,
        base_channels: int = 128,
        temb_channels: Optional[int] = None,
        channel_multiplier=(1, 2, 4, 8),
        num_residual_blocks=(3, 3, 3, 3),
        gn_num_groups: int = 32 // 4,
        gn_eps: float = 1e-6,
        attn_num_heads: int = 8,
        coords_encoding: Optional[str] = "spherical_harmonics",
        ring: bool = True,
        precision: str = "fp32",
    ):
        super().__init__()
        if not ring:
            raise ValueError("only ring=True (circular azimuth padding) is implemented; "
                             "utils/inference.py:50 always builds the network with ring=True")
        if out_channels is not None and out_channels != in_channels:
            raise ValueError("out_channels must equal in_channels on the sampling path")
        self.resolution = _n_tuple(resolution, 2)
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.base_channels = base_channels
        self.temb_channels = base_channels * 4 if temb_channels is None else temb_channels
        self.channel_multiplier = _n_tuple(channel_multiplier, 4)
        self.num_residual_blocks = _n_tuple(num_residual_blocks, 4)
        self.gn_num_groups = gn_num_groups
        self.gn_eps = gn_eps
        self.attn_num_heads = attn_num_heads
        self.coords_encoding_type = coords_encoding
        self.precision = precision or "fp32"  # "fp32" (tf32 tensor cores, the reference's GPU default) | "bf16"
        self._precision_explicit = False      # set_precision() pins the engine regardless of autocast

        H, W = self.resolution
        self.register_buffer("coords", encoding.generate_polar_coords(H, W))
        self.extra_ch = 0
        self.coords_encoding = None
        if coords_encoding == "spherical_harmonics":
            self.coords_encoding = encoding.SphericalHarmonics(levels=5)
            self.extra_ch = self.coords_encoding.extra_ch
        elif coords_encoding == "polar_coordinates":
            self.coords_encoding = nn.Identity()
            self.extra_ch = 2
        elif coords_encoding == "fourier_features":
            self.coords_encoding = encoding.FourierFeatures(self.resolution)
            self.extra_ch = self.coords_encoding.extra_ch
        elif coords_encoding is not None:
            raise ValueError(f"invalid coords_encoding: {coords_encoding}")

        # parameters / buffers with the reference's names, shapes and default initialisation
        g = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 31))
        for name, shape, kind in self._schema():
            if kind == "scale":
                _register(self, name, torch.tensor(1 / np.sqrt(2)).float(), True)
            elif kind == "kernel_down":
                _register(self, name, torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8, True)
            elif kind == "kernel_up":
                _register(self, name, torch.tensor([1.0, 3.0, 3.0, 1.0]) / 4, True)
            elif kind == "zero":       # zero_out() of conv2 / out_proj / out_conv
                _register(self, name, torch.zeros(shape), False)
            elif kind == "one":
                _register(self, name, torch.ones(shape), False)
            else:                      # kaiming-uniform like nn.Conv2d / nn.Linear defaults
                fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(kind)
                bound = 1 / math.sqrt(fan_in)
                _register(self, name, (torch.rand(shape, generator=g) * 2 - 1) * bound, False)
        self._engines: Dict[Tuple[str, str], "UNetEngine"] = {}
        self._weights_version = 0

    # ------------------------------------------------------------------------------- schema
    def _blocks(self):
        Cs = [self.base_channels] + [self.base_channels * m for m in self.channel_multiplier]
        N = self.num_residual_blocks
        return [
            ("d_block1", Cs[0], Cs[1], N[0], 1, 1, False), ("d_block2", Cs[1], Cs[2], N[1], 2, 1, False),
            ("d_block3", Cs[2], Cs[3], N[2], 2, 1, False), ("d_block4", Cs[3], Cs[4], N[3], 2, 1, True),
            ("u_block4", Cs[4], Cs[3], N[3], 1, 2, True), ("u_block3", 2 * Cs[3], Cs[2], N[2], 1, 2, False),
            ("u_block2", 2 * Cs[2], Cs[1], N[1], 1, 2, False), ("u_block1", 2 * Cs[1], Cs[0], N[0], 1, 1, False),
        ]

    def _schema(self):
        """(name, shape, kind) in the order of the reference module tree (efficient_unet.py:212-267)."""
        T, C0 = self.temb_channels, self.base_channels

        def conv(name, co, ci, k, zero=False):
            yield (f"{name}.weight", (co, ci, k, k), "zero" if zero else "w")
            yield (f"{name}.bias", (co,), "zero" if zero else str(ci * k * k))

        def linear(name, co, ci, zero=False):
            yield (f"{name}.weight", (co, ci), "zero" if zero else "w")
            yield (f"{name}.bias", (co,), "zero" if zero else str(ci))

        yield from linear("time_embedding.1", T, C0)
        yield from linear("time_embedding.3", T, T)
        yield from conv("in_conv", C0, self.in_channels + self.extra_ch, 3)
        for name, cin, cout, nres, down, up, attn in self._blocks():
            if down > 1:
                yield from conv(f"{name}.downsample.0", cout, cin, 3)
                yield (f"{name}.downsample.1.kernel", (4,), "kernel_down")
            for i in range(nres):
                ci = cout if (i != 0 or down > 1) else cin
                p = f"{name}.residual_blocks.{i}"
                yield (f"{p}.scale", (), "scale")
                yield (f"{p}.norm1.weight", (ci,), "one")
                yield (f"{p}.norm1.bias", (ci,), "zero")
                yield from conv(f"{p}.conv1", cout, ci, 3)
                yield from linear(f"{p}.norm2.proj.1", 2 * cout, T)
                yield from conv(f"{p}.conv2", cout, cout, 3, zero=True)
                if ci != cout:
                    yield from conv(f"{p}.skip", cout, ci, 1)
            if attn:
                p = f"{name}.self_attn_block"
                yield (f"{p}.scale", (), "scale")
                yield (f"{p}.norm.weight", (cout,), "one")
                yield (f"{p}.norm.bias", (cout,), "zero")
                yield (f"{p}.attn.in_proj_weight", (3 * cout, cout), "w")
                yield (f"{p}.attn.in_proj_bias", (3 * cout,), "zero")
                yield from linear(f"{p}.attn.out_proj", cout, cout, zero=True)
            if up > 1:
                yield (f"{name}.upsample.0.kernel", (4,), "kernel_up")
                yield from conv(f"{name}.upsample.1", cout, cout, 3)
        yield from conv("out_conv", self.in_channels, C0, 3, zero=True)

    # ------------------------------------------------------------------------------- engine
    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._weights_version += 1

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_version += 1
        # the module may have moved: drop engines (C handle + weight arena + workspace) of other devices
        dev = str(self.coords.device)
        for key in [k for k in self._engines if k[0] != dev]:
            self._engines.pop(key).close()
        return out

    def set_precision(self, precision: str) -> "EfficientUNet":
        """Pin the engine ("fp32" | "bf16"); an explicit choice is not overridden by autocast."""
        if precision not in ("fp32", "tf32", "bf16"):
            raise ValueError(f"invalid precision: {precision}")
        self.precision = "fp32" if precision == "tf32" else precision
        self._precision_explicit = True
        return self

    def _param_versions(self) -> int:
        """Sum of the parameters' in-place modification counters: detects `p.data.copy_()` / optimizer-style
        edits that bypass load_state_dict / .to()."""
        return sum(p._version for p in self.parameters())

    def mark_weights_changed(self) -> None:
        """Call after editing parameters in place so the packed device copy is rebuilt."""
        self._weights_version += 1

    def _active_precision(self) -> str:
        """An explicit `precision=` always wins.  With the default ("fp32") a surrounding
        `torch.autocast("cuda", dtype=torch.bfloat16)` selects the bf16 engine; fp16 autocast
        (sample_and_save.py:70 with the reference's default mixed_precision="fp16") keeps the fp32 (tf32
        tensor core) engine, whose error vs the fp32 reference (1e-3) is below the reference's own fp16
        autocast error (3e-3) - there is no fp16 engine."""
        if self._precision_explicit or self.precision != "fp32":
            return self.precision
        get = getattr(torch, "get_autocast_dtype", None)
        ac_dtype = get("cuda") if get is not None else torch.get_autocast_gpu_dtype()
        if torch.is_autocast_enabled() and ac_dtype == torch.bfloat16:
            return "bf16"
        return self.precision

    # engines hold raw C handles and device arenas: never copy / pickle them, rebuild lazily instead
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
        return new

    def engine(self, precision: Optional[str] = None) -> "UNetEngine":
        precision = precision or self._active_precision()
        dev = self.coords.device
        if dev.type != "cuda":
            raise L.R2dmError("r2dm_b200.EfficientUNet runs on CUDA only (no CPU fallback): "
                              "move the module to a B200 with .to('cuda')")
        key = (str(dev), precision)
        eng = self._engines.get(key)
        pv = self._param_versions()
        if pv != getattr(self, "_seen_param_versions", pv):
            self._weights_version += 1       # parameters were edited in place since the last forward
        self._seen_param_versions = pv
        if eng is None or eng.weights_version != self._weights_version:
            if eng is not None:
                eng.close()
            eng = UNetEngine(self, precision)
            self._engines[key] = eng
        return eng

    def coords_table(self) -> Optional[torch.Tensor]:
        """[extra_ch, H, W] fp32 constant coordinate encoding (efficient_unet.py:278-279)."""
        if self.coords_encoding is None:
            return None
        with torch.no_grad():
            return self.coords_encoding(self.coords.float())[0].contiguous()

    @_no_compile
    def forward(self, images: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        if timesteps.dim() == 0:
            timesteps = timesteps[None].repeat_interleave(images.shape[0], dim=0)
        eng = self.engine()
        return eng.forward(images, timesteps).to(images.dtype)


class UNetEngine:
    """Owns the C handle, the packed-weight arena and per-batch workspaces for one (device, precision)."""

    def __init__(self, model: EfficientUNet, precision: str):
        if precision not in ("fp32", "tf32", "bf16"):
            raise ValueError(f"invalid precision: {precision}")
        self.lib = L.lib()
        self.device = model.coords.device
        self.precision = precision
        self.weights_version = model._weights_version
        H, W = model.resolution
        cfg = L.R2dmConfig()
        cfg.in_channels = model.in_channels
        cfg.height, cfg.width = H, W
        cfg.base_channels = model.base_channels
        cfg.temb_channels = model.temb_channels
        cfg.channel_multiplier = (C.c_int * 4)(*model.channel_multiplier)
        cfg.num_residual_blocks = (C.c_int * 4)(*model.num_residual_blocks)
        cfg.gn_num_groups = model.gn_num_groups
        cfg.gn_eps = model.gn_eps
        cfg.attn_num_heads = model.attn_num_heads
        cfg.extra_channels = model.extra_ch
        sd = model.state_dict()
        scales = {float(v) for k, v in sd.items() if k.endswith(".scale")}
        if len(scales) != 1:
            raise L.R2dmError(f"residual scale buffers differ across blocks: {scales}")
        cfg.residual_scale = scales.pop()
        for k, v in sd.items():
            if k.endswith("downsample.1.kernel"):
                assert torch.allclose(v.cpu(), torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8), "unsupported FIR window"
            if k.endswith("upsample.0.kernel"):
                assert torch.allclose(v.cpu(), torch.tensor([1.0, 3.0, 3.0, 1.0]) / 4), "unsupported FIR window"
        cfg.dtype = L.BF16 if precision == "bf16" else L.F32
        self.cfg = cfg
        self.in_channels, self.H, self.W = model.in_channels, H, W
        self.temb_channels = model.temb_channels
        h = C.c_void_p()
        L.check(self.lib.r2dm_create(C.byref(cfg), C.byref(h)), "r2dm_create")
        self.h = h
        with torch.cuda.device(self.device):
            nbytes = self.lib.r2dm_weight_arena_bytes(self.h)
            self.arena = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            L.check(self.lib.r2dm_bind_weight_arena(self.h, self._aligned(self.arena, 256), nbytes), "bind arena")
            stream = L.stream_ptr()
            tensors = dict(sd)
            table = model.coords_table()
            if table is not None:
                tensors["coords_encoding.table"] = table
            keep = []
            for name, t in tensors.items():
                t = L.f32c(t.to(self.device))
                keep.append(t)
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(1)
                L.check(self.lib.r2dm_load_tensor(self.h, name.encode(), L.ptr(t), shape, t.dim(), stream),
                        f"load {name}")
            buf = C.create_string_buffer(1 << 16)
            missing = self.lib.r2dm_missing_tensors(self.h, buf, len(buf))
            if missing:
                raise L.R2dmError(f"{missing} tensors missing from the state dict: {buf.value.decode()[:400]}")
            torch.cuda.current_stream().synchronize()
        self.film_width = self.lib.r2dm_film_width(self.h)
        self._ws: Dict[int, torch.Tensor] = {}
        self._bound_batch = None
        self.bind_epoch = 0   # bumped whenever the workspace binding (and with it every captured pointer) changes
        self._zero_step = torch.zeros(1, dtype=torch.int32, device=self.device)

    @staticmethod
    def _aligned(t: torch.Tensor, a: int) -> int:
        p = t.data_ptr()
        return (p + a - 1) // a * a

    def close(self):
        if getattr(self, "h", None):
            self.lib.r2dm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind(self, batch: int) -> None:
        if self._bound_batch == batch:
            return
        with torch.cuda.device(self.device):
            ws = self._ws.get(batch)
            nbytes = self.lib.r2dm_workspace_bytes(self.h, batch)
            if ws is None:
                ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
                self._ws = {batch: ws}  # keep one workspace alive (the binding is exclusive anyway)
            L.check(self.lib.r2dm_bind_workspace(self.h, self._aligned(ws, 1024), nbytes, batch, L.stream_ptr()),
                    "bind workspace")
        self._bound_batch = batch
        self.bind_epoch += 1

    @property
    def launches_per_forward(self) -> int:
        return self.lib.r2dm_num_launches(self.h)

    def cond_embed(self, cond: torch.Tensor) -> torch.Tensor:
        """[rows] conditions -> FiLM table [rows, film_width] (time MLP + all AdaGN projections)."""
        cond = L.f32c(cond.to(self.device))
        rows = cond.numel()
        scratch = torch.empty(rows, self.temb_channels, device=self.device, dtype=torch.float32)
        film = torch.empty(rows, self.film_width, device=self.device, dtype=torch.float32)
        L.check(self.lib.r2dm_cond_embed(self.h, L.ptr(cond), rows, L.ptr(scratch), L.ptr(film), L.stream_ptr()),
                "r2dm_cond_embed")
        return film

    def forward_film(self, x: torch.Tensor, film: torch.Tensor, pred: torch.Tensor,
                     step_ptr: Optional[torch.Tensor] = None, rows_per_step: int = 0,
                     row_batch_stride: int = 1) -> torch.Tensor:
        """Enqueue one U-Net forward; x / pred are fp32 [B, C, H, W] on this device."""
        assert x.dtype == torch.float32 and x.is_contiguous() and pred.is_contiguous()
        self.bind(x.shape[0])
        L.check(self.lib.r2dm_unet_forward(self.h, L.ptr(x), L.ptr(film),
                                           L.ptr(step_ptr) if step_ptr is not None else None,
                                           rows_per_step, row_batch_stride, L.ptr(pred), L.stream_ptr()),
                "r2dm_unet_forward")
        return pred

    def forward(self, images: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        B, Cc, H, W = images.shape
        if (Cc, H, W) != (self.in_channels, self.H, self.W):
            raise ValueError(f"expected images of shape [B, {self.in_channels}, {self.H}, {self.W}], got {tuple(images.shape)}")
        if cond.shape != (B,):
            raise ValueError("timesteps must have shape [B]")
        x = L.f32c(images)
        with torch.cuda.device(self.device):
            film = self.cond_embed(cond.to(torch.float32))
            pred = torch.empty_like(x)
            self.forward_film(x, film, pred)
        return pred

    KINDS = ("pack_input", "conv3x3", "conv1x1", "gn_apply", "down2", "up2", "attention")

    def profile_forward(self, x: torch.Tensor, cond: torch.Tensor):
        """Measurement aid: per-launch (kind, ms, algorithmic flops, algorithmic bytes) of one eager
        forward, timed with CUDA events on the launching stream.  Synchronises."""
        x = L.f32c(x)
        with torch.cuda.device(self.device):
            film = self.cond_embed(cond)
            pred = torch.empty_like(x)
            self.bind(x.shape[0])
            cap = 1024     # program entries (>= launches: a chain of convolutions is one launch, several entries)
            kind, ms = (C.c_int * cap)(), (C.c_float * cap)()
            fl, by = (C.c_double * cap)(), (C.c_double * cap)()
            n = L.check(self.lib.r2dm_profile_forward(self.h, L.ptr(x), L.ptr(film), L.ptr(pred), L.stream_ptr(),
                                                      cap, kind, ms, fl, by), "r2dm_profile_forward")
        return [(self.KINDS[kind[i]], ms[i], fl[i], by[i]) for i in range(n)]

    def forward_kinds(self, x: torch.Tensor, film: torch.Tensor, pred: torch.Tensor, kinds) -> None:
        """Measurement aid: enqueue only the launches of the given kinds (names from KINDS) of one forward."""
        mask = 0
        for k in kinds:
            mask |= 1 << self.KINDS.index(k)
        self.bind(x.shape[0])
        L.check(self.lib.r2dm_debug_forward_kinds(self.h, L.ptr(x), L.ptr(film), L.ptr(pred), mask, L.stream_ptr()),
                "r2dm_debug_forward_kinds")

    def debug_tensor(self, name: str) -> torch.Tensor:
        """fp32 NCHW copy of a named intermediate of the last forward (needs R2DM_KEEP_ACTIVATIONS=1)."""
        c, hh, ww = C.c_int(), C.c_int(), C.c_int()
        L.check(self.lib.r2dm_debug_tensor(self.h, name.encode(), None, C.byref(c), C.byref(hh), C.byref(ww), None))
        out = torch.empty(self._bound_batch, c.value, hh.value, ww.value, device=self.device, dtype=torch.float32)
        L.check(self.lib.r2dm_debug_tensor(self.h, name.encode(), L.ptr(out), None, None, None, L.stream_ptr()))
        return out
