"""Developer probe: fp32 (kind::tf32) attention core against torch for structured inputs."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from r2dm_b200 import ops

def ref_attn(qkv, heads):
    B, C3, H, W = qkv.shape
    E = C3 // 3; hd = E // heads
    tok = qkv.flatten(2).transpose(1, 2)
    q, k, v = [t.reshape(B, -1, heads, hd).transpose(1, 2) for t in tok.split(E, dim=-1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    return (att @ v).transpose(1, 2).reshape(B, -1, E).transpose(1, 2).reshape(B, E, H, W)

def rel(a, b):
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()

B, E, heads, H, W = 1, 256, 8, 2, 128
g = torch.Generator().manual_seed(1)
base = torch.randn(B, 3 * E, H, W, generator=g)
base = (base.view(torch.int32) & ~0x1FFF).view(torch.float32)
for name in ["random", "q0", "v1", "vramp"]:
    x = base.clone()
    if name == "q0": x[:, :E] = 0
    if name == "v1": x[:, 2 * E:] = 1
    if name == "vramp": x[:, :E] = 0; x[:, 2 * E:] = torch.arange(E).float()[None, :, None, None] + 0.001 * torch.arange(H * W).float().view(1, 1, H, W)
    for dt in ["fp32", "bf16"]:
        y = ops.attention_core(x.cuda(), heads, dtype=dt).cpu()
        r = ref_attn(x, heads)
        print(f"{name:8s} {dt}: rel {rel(y, r):.3e} | y mean {y.mean():.4f} absmax {y.abs().max():.4f} nan {torch.isnan(y).sum().item()} | ref mean {r.mean():.4f} absmax {r.abs().max():.4f}")
        if name == "vramp" and dt == "fp32":
            print("   y[0,:12,0,0]  ", [round(v, 3) for v in y[0, :12, 0, 0].tolist()])
            print("   ref[0,:12,0,0]", [round(v, 3) for v in r[0, :12, 0, 0].tolist()])
            print("   y[0,0,0,:6]   ", [round(v, 4) for v in y[0, 0, 0, :6].tolist()])
