"""Per-kernel parity: CUDA kernels (through the C ABI op hooks) vs the CPU oracle.

Tolerances.  The tensor-core path multiplies bf16 (or tf32) operands exactly and accumulates in
fp32, so when the inputs are pre-rounded to the operand type the only error left is the fp32
accumulation order (~1e-6) plus the rounding of the stored output (bf16: 2^-9 relative).
  fp32/tf32 mode : l2-rel <= 2e-5 (conv, down2: fp32 outputs)
                   l2-rel <= 5e-4 (GroupNorm/AdaGN, up2, attention: their outputs only ever feed a
                   tensor-core operand and are therefore stored rounded-to-nearest tf32, 2^-11 rel.)
  bf16 mode      : l2-rel <= 4e-3, max-abs <= 2^-7 * max|y|
"""
import math

import pytest
import torch

from oracle import r2dm_oracle as O
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _round_to(x, dtype):
    if dtype == "bf16":
        return x.bfloat16().float()
    # tf32: keep 10 mantissa bits (exactly representable inputs make the tensor-core read exact)
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _check(y, ref, dtype, name, tf32_out=False):
    e = rel_l2(y, ref)
    tol = 4e-3 if dtype == "bf16" else (5e-4 if tf32_out else 2e-5)
    assert e <= tol, f"{name}[{dtype}]: l2-rel {e:.3e} > {tol}"
    if dtype == "bf16":
        m = (y.cpu() - ref).abs().max().item()
        assert m <= 2 ** -7 * ref.abs().max().item() + 1e-6, f"{name}[{dtype}]: max-abs {m:.3e}"


CONV_CASES = [
    # B, Cin, Cout, H, W, k, residual
    (2, 64, 64, 8, 256, 3, True),
    (1, 34, 64, 4, 128, 3, False),
    (2, 128, 128, 4, 128, 3, True),
    (1, 64, 2, 4, 128, 3, False),
    (1, 256, 64, 2, 256, 3, False),
    (1, 64, 128, 6, 128, 3, False),
    (2, 128, 64, 4, 128, 1, False),
    (1, 512, 1536, 2, 128, 1, False),
    (1, 256, 256, 2, 128, 1, True),
    (2, 128, 256, 4, 256, 3, True),
    (1, 64, 512, 2, 128, 3, False),
    # full-resolution shapes of config H (four-row tiles across the whole 64 x 1024 image, 7 tiles per CTA;
    # two-row 128-channel tiles at 32 x 512; single-row tiles at the 8 x 128 bottleneck)
    (2, 64, 64, 64, 1024, 3, True),
    (2, 128, 128, 32, 512, 3, True),
    (4, 512, 512, 8, 128, 3, False),
]


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv(case, dtype):
    from r2dm_b200 import ops
    B, Cin, Cout, H, W, k, use_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = _round_to(torch.randn(B, Cin, H, W, generator=g), dtype)
    w = _round_to(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k), dtype)
    b = torch.randn(Cout, generator=g) * 0.1
    res = _round_to(torch.randn(B, Cout, H, W, generator=g), dtype) if use_res else None
    scale = 1 / math.sqrt(2) if use_res else 1.0
    ref = O.ring_conv3x3(x, w, b) if k == 3 else O.conv1x1(x, w, b)
    if use_res:
        ref = (ref + res) * scale
    y = ops.conv2d(x.cuda(), w.cuda(), b.cuda(), res.cuda() if use_res else None, scale, dtype=dtype)
    torch.cuda.synchronize()
    _check(y, ref, dtype, f"conv{k}x{k} {case}")


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("film", [False, True])
@pytest.mark.parametrize("shape", [(2, 64, 4, 128), (1, 128, 8, 256), (2, 512, 2, 128)])
def test_groupnorm_silu(shape, film, dtype):
    from r2dm_b200 import ops
    B, Cc, H, W = shape
    g = torch.Generator().manual_seed(7)
    x = _round_to(torch.randn(B, Cc, H, W, generator=g) * 1.7 + 0.4, dtype)
    if film:
        ss = torch.randn(B, 2 * Cc, generator=g) * 0.3
        h = O.group_norm(x, 8, 1e-6, None, None)
        ref = torch.nn.functional.silu(h * (1 + ss[:, :Cc, None, None]) + ss[:, Cc:, None, None])
        y = ops.group_norm(x.cuda(), film=ss.cuda(), eps=1e-6, silu=True, dtype=dtype)
    else:
        gm, bt = 1 + 0.1 * torch.randn(Cc, generator=g), 0.1 * torch.randn(Cc, generator=g)
        ref = torch.nn.functional.silu(O.group_norm(x, 8, 1e-6, gm, bt))
        y = ops.group_norm(x.cuda(), gm.cuda(), bt.cuda(), eps=1e-6, silu=True, dtype=dtype)
    torch.cuda.synchronize()
    _check(y, ref, dtype, f"groupnorm {shape} film={film}", tf32_out=True)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("shape", [(2, 64, 4, 256), (1, 128, 8, 512)])
def test_resample(shape, dtype):
    from r2dm_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = _round_to(torch.randn(*shape, generator=g), dtype)
    yd = ops.resample(x.cuda(), down=2, dtype=dtype)
    yu = ops.resample(x.cuda(), up=2, dtype=dtype)
    torch.cuda.synchronize()
    _check(yd, O.resample_down2(x), dtype, "down2")
    _check(yu, O.resample_up2(x), dtype, "up2", tf32_out=True)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("case", [(2, 256, 8, 2, 128), (1, 512, 8, 4, 128)])
def test_attention_core(case, dtype):
    from r2dm_b200 import ops
    B, E, heads, H, W = case
    hd = E // heads
    g = torch.Generator().manual_seed(9)
    qkv = _round_to(torch.randn(B, 3 * E, H, W, generator=g), dtype)
    y = ops.attention_core(qkv.cuda(), heads, dtype=dtype)
    torch.cuda.synchronize()
    tok = qkv.flatten(2).transpose(1, 2)
    q, k, v = tok.split(E, dim=-1)
    q = q.reshape(B, -1, heads, hd).transpose(1, 2)
    k = k.reshape(B, -1, heads, hd).transpose(1, 2)
    v = v.reshape(B, -1, heads, hd).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(B, -1, E).transpose(1, 2).reshape(B, E, H, W)
    _check(y, ref, dtype, f"attention {case}", tf32_out=True)


def test_attention_exact_option():
    """fp32 engine: `attn_exact` swaps the kind::tf32 tensor-core attention for the fp32 FMA kernel (no operand
    rounding at all: unrounded inputs, 2e-5 before the tf32-rounded store)."""
    from r2dm_b200 import _lib as L
    from r2dm_b200 import ops
    B, E, heads, H, W = 1, 256, 8, 2, 128
    hd = E // heads
    g = torch.Generator().manual_seed(10)
    qkv = torch.randn(B, 3 * E, H, W, generator=g)
    tok = qkv.flatten(2).transpose(1, 2)
    q, k, v = [t.reshape(B, -1, heads, hd).transpose(1, 2) for t in tok.split(E, dim=-1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(B, -1, E).transpose(1, 2).reshape(B, E, H, W)
    L.check(L.lib().r2dm_set_option(b"attn_exact", 1))
    try:
        y_exact = ops.attention_core(qkv.cuda(), heads, dtype="fp32")
    finally:
        L.check(L.lib().r2dm_set_option(b"attn_exact", 0))
    y_tf32 = ops.attention_core(qkv.cuda(), heads, dtype="fp32")
    torch.cuda.synchronize()
    e_exact, e_tf32 = rel_l2(y_exact, ref), rel_l2(y_tf32, ref)
    assert e_exact <= 5e-4, e_exact
    assert e_tf32 <= 2e-3, e_tf32          # unrounded q / k / v: the tensor core truncates them to tf32
    assert not torch.equal(y_exact, y_tf32)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("case", [(2, 64, 64, 8, 256, 3, False), (1, 128, 64, 4, 128, 3, True),
                                  (2, 256, 128, 2, 128, 3, True), (1, 512, 1536, 4, 128, 1, False),
                                  (2, 256, 256, 4, 128, 3, True), (1, 128, 512, 6, 256, 3, False)])
def test_fused_groupnorm_conv(case, dtype):
    """The network's fused form: GroupNorm/AdaGN (+SiLU) applied to the operand tile inside the conv.
    The normalised activations are rounded to the operand type (bf16 / tf32) before the MMA, so the
    tolerance is that of an operand rounding: bf16 6e-3 l2-rel, tf32 1e-3."""
    from r2dm_b200 import ops
    B, Cin, Cout, H, W, k, film = case
    g = torch.Generator().manual_seed(3)
    x = _round_to(torch.randn(B, Cin, H, W, generator=g) * 1.3 + 0.2, dtype)
    w = _round_to(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k), dtype)
    b = torch.randn(Cout, generator=g) * 0.1
    silu = k == 3
    if film:
        ss = torch.randn(B, 2 * Cin, generator=g) * 0.3
        h = O.group_norm(x, 8, 1e-6, None, None) * (1 + ss[:, :Cin, None, None]) + ss[:, Cin:, None, None]
        kw = dict(film=ss.cuda())
    else:
        gm, bt = 1 + 0.1 * torch.randn(Cin, generator=g), 0.1 * torch.randn(Cin, generator=g)
        h = O.group_norm(x, 8, 1e-6, gm, bt)
        kw = dict(gamma=gm.cuda(), beta=bt.cuda())
    if silu:
        h = torch.nn.functional.silu(h)
    ref = O.ring_conv3x3(h, w, b) if k == 3 else O.conv1x1(h, w, b)
    y = ops.gn_conv2d(x.cuda(), w.cuda(), b.cuda(), eps=1e-6, silu=silu, dtype=dtype, **kw)
    torch.cuda.synchronize()
    e = rel_l2(y, ref)
    tol = 6e-3 if dtype == "bf16" else 1e-3
    assert e <= tol, f"fused gn+conv {case}[{dtype}]: l2-rel {e:.3e} > {tol}"


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("case", [(2, 64, 64, 64, 1024, 3, True), (2, 128, 128, 32, 512, 3, False)])
def test_fused_groupnorm_conv_full_resolution(case, dtype):
    """The fused form at config H's real tile counts (several tiles per CTA, image changes inside a CTA's
    range, statistics folded from 128 slots)."""
    test_fused_groupnorm_conv(case, dtype)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("case", [
    # B, C, Cs, H, W  (the four up-path skip shapes of config H, at reduced and at full size)
    (2, 64, 128, 8, 256), (1, 64, 256, 4, 128), (2, 128, 512, 4, 256), (1, 256, 512, 2, 128), (3, 256, 512, 8, 128),
    (2, 64, 128, 64, 1024), (2, 128, 512, 16, 256),
])
def test_folded_skip_projection(case, dtype):
    """ResidualBlock tail of the up-path blocks (efficient_unet.py:99-110): the 1x1 skip projection runs as extra
    K stages of conv2 - (conv3x3(silu(adagn(h))) + b + conv1x1(x) + b_skip) / sqrt 2 - against the oracle ops."""
    from r2dm_b200 import ops
    B, Cc, Cs, H, W = case
    g = torch.Generator().manual_seed(17)
    h = _round_to(torch.randn(B, Cc, H, W, generator=g) * 1.3 + 0.2, dtype)
    xs = _round_to(torch.randn(B, Cs, H, W, generator=g), dtype)
    w = _round_to(torch.randn(Cc, Cc, 3, 3, generator=g) / math.sqrt(Cc * 9), dtype)
    w2 = _round_to(torch.randn(Cc, Cs, 1, 1, generator=g) / math.sqrt(Cs), dtype)
    b, b2 = torch.randn(Cc, generator=g) * 0.1, torch.randn(Cc, generator=g) * 0.1
    ss = torch.randn(B, 2 * Cc, generator=g) * 0.3
    scale = 1 / math.sqrt(2)
    hn = O.group_norm(h, 8, 1e-6, None, None) * (1 + ss[:, :Cc, None, None]) + ss[:, Cc:, None, None]
    ref = (O.ring_conv3x3(torch.nn.functional.silu(hn), w, b) + O.conv1x1(xs, w2, b2)) * scale
    y = ops.gn_conv2d_skip(h.cuda(), ss.cuda(), w.cuda(), b.cuda(), xs.cuda(), w2.cuda(), b2.cuda(), scale, dtype=dtype)
    torch.cuda.synchronize()
    e = rel_l2(y, ref)
    tol = 6e-3 if dtype == "bf16" else 1e-3
    assert e <= tol, f"folded skip {case}[{dtype}]: l2-rel {e:.3e} > {tol}"
