"""Single-operator entry points backed by the same CUDA kernels the network uses.

Host-side mirror of models/ops.py of the reference (Conv2d ring padding :149-173, Resample :52-146,
AdaGN :176-200) plus nn.GroupNorm / the attention core, operating on fp32 NCHW CUDA tensors through
the C ABI (`r2dm_op_*`).  These exist for parity testing of the individual kernels; the network
itself runs fused inside `r2dm_unet_forward`.
"""
from __future__ import annotations

import torch

from . import _lib as L

_DT = {"fp32": L.F32, "tf32": L.F32, "bf16": L.BF16, torch.float32: L.F32, torch.bfloat16: L.BF16}


def _dt(dtype) -> int:
    if dtype not in _DT:
        raise ValueError(f"invalid dtype: {dtype}")
    return _DT[dtype]


def _scratch(B, C, H, W, device):
    n = L.lib().r2dm_op_scratch_bytes(B, C, H, W)
    return torch.empty(n, dtype=torch.uint8, device=device), n


def conv2d(x, weight, bias=None, residual=None, scale=1.0, dtype="bf16"):
    """ops.Conv2d(k=3, padding=1, ring=True) or (k=1, padding=0); y = (conv + bias [+ residual]) * scale."""
    x = L.f32c(x)
    w = L.f32c(weight)
    B, Cin, H, W = x.shape
    Cout, Cin2, kh, kw = w.shape
    assert Cin2 == Cin and kh == kw and kh in (1, 3)
    b = L.f32c(bias) if bias is not None else None
    r = L.f32c(residual) if residual is not None else None
    y = torch.empty(B, Cout, H, W, device=x.device, dtype=torch.float32)
    sc, n = _scratch(B, max(Cin, Cout), H, W, x.device)
    L.check(L.lib().r2dm_op_conv(_dt(dtype), kh * kw, L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(r), float(scale),
                                 L.ptr(y), B, Cin, Cout, H, W, L.ptr(sc), n, L.stream_ptr()), "r2dm_op_conv")
    return y


def gn_conv2d(x, weight, bias=None, gamma=None, beta=None, film=None, eps=1e-6, silu=True, dtype="bf16"):
    """conv(act(GroupNorm(x))) with the normalisation fused into the conv kernel's operand path, as the
    network runs efficient_unet.py:99-106 (film [B, 2*Cin] = [scale || shift] selects AdaGN)."""
    x = L.f32c(x)
    w = L.f32c(weight)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    b = L.f32c(bias) if bias is not None else None
    g = L.f32c(gamma) if gamma is not None else None
    bt = L.f32c(beta) if beta is not None else None
    f = L.f32c(film) if film is not None else None
    y = torch.empty(B, Cout, H, W, device=x.device, dtype=torch.float32)
    sc, n = _scratch(B, max(Cin, Cout), H, W, x.device)
    L.check(L.lib().r2dm_op_gn_conv(_dt(dtype), kh * kw, L.ptr(x), L.ptr(g), L.ptr(bt), L.ptr(f), float(eps),
                                    int(silu), L.ptr(w), L.ptr(b), L.ptr(y), B, Cin, Cout, H, W, L.ptr(sc), n,
                                    L.stream_ptr()), "r2dm_op_gn_conv")
    return y


def gn_conv2d_skip(h, film, weight, bias, xs, skip_weight, skip_bias, scale, eps=1e-6, dtype="bf16"):
    """ResidualBlock tail with the skip projection folded into conv2 (efficient_unet.py:99-110 on blocks with a
    skip): (conv3x3(silu(adagn(h, film))) + bias + conv1x1(xs, skip_weight) + skip_bias) * scale."""
    h, xs = L.f32c(h), L.f32c(xs)
    w, w2 = L.f32c(weight), L.f32c(skip_weight)
    B, Cc, H, W = h.shape
    Cs = xs.shape[1]
    assert w.shape == (Cc, Cc, 3, 3) and w2.shape[:2] == (Cc, Cs)
    b = L.f32c(bias) if bias is not None else None
    b2 = L.f32c(skip_bias) if skip_bias is not None else None
    f = L.f32c(film)
    y = torch.empty(B, Cc, H, W, device=h.device, dtype=torch.float32)
    n = 2 * L.lib().r2dm_op_scratch_bytes(B, max(Cc, Cs), H, W)
    sc = torch.empty(n, dtype=torch.uint8, device=h.device)
    L.check(L.lib().r2dm_op_gn_conv_skip(_dt(dtype), L.ptr(h), L.ptr(f), float(eps), L.ptr(w), L.ptr(b), L.ptr(xs),
                                         L.ptr(w2), L.ptr(b2), float(scale), L.ptr(y), B, Cc, Cs, H, W, L.ptr(sc), n,
                                         L.stream_ptr()), "r2dm_op_gn_conv_skip")
    return y


def group_norm(x, gamma=None, beta=None, film=None, eps=1e-6, silu=False, dtype="bf16"):
    """nn.GroupNorm(8, C, eps) (+SiLU); with `film` [B, 2C] = [scale || shift]: AdaGN (ops.py:196-199)."""
    x = L.f32c(x)
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    sc, n = _scratch(B, Cc, H, W, x.device)
    g = L.f32c(gamma) if gamma is not None else None
    b = L.f32c(beta) if beta is not None else None
    f = L.f32c(film) if film is not None else None
    L.check(L.lib().r2dm_op_groupnorm(_dt(dtype), L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(f), float(eps), int(silu),
                                      L.ptr(y), B, Cc, H, W, L.ptr(sc), n, L.stream_ptr()), "r2dm_op_groupnorm")
    return y


def resample(x, up=1, down=1, dtype="bf16"):
    """ops.Resample(up=2) / ops.Resample(down=2) with the [1,3,3,1] window, ring=True."""
    assert (up, down) in ((2, 1), (1, 2))
    x = L.f32c(x)
    B, Cc, H, W = x.shape
    Ho, Wo = (H * 2, W * 2) if up == 2 else (H // 2, W // 2)
    y = torch.empty(B, Cc, Ho, Wo, device=x.device, dtype=torch.float32)
    sc, n = _scratch(B, Cc, max(H, Ho), max(W, Wo), x.device)
    L.check(L.lib().r2dm_op_resample(_dt(dtype), 2 if up == 2 else -2, L.ptr(x), L.ptr(y), B, Cc, H, W,
                                     L.ptr(sc), n, L.stream_ptr()), "r2dm_op_resample")
    return y


def attention_core(qkv, heads, dtype="bf16"):
    """softmax(q k^T / sqrt(hd)) v over the H*W tokens of packed qkv [B, 3E, H, W] -> [B, E, H, W]."""
    qkv = L.f32c(qkv)
    B, C3, H, W = qkv.shape
    E = C3 // 3
    y = torch.empty(B, E, H, W, device=qkv.device, dtype=torch.float32)
    sc, n = _scratch(B, C3, H, W, qkv.device)
    L.check(L.lib().r2dm_op_attention(_dt(dtype), L.ptr(qkv), L.ptr(y), B, E, heads, H, W, L.ptr(sc), n,
                                      L.stream_ptr()), "r2dm_op_attention")
    return y
