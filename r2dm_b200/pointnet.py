"""PointNet feature extractor — host mirror of the reference's `metrics/extractor/pointnet.py` (STN3d :8-33,
PointNetfeat :36-64, PointNet1 :67-81, pretrained_pointnet :84-98), the network `evaluate.py:83,121,160` uses
for the Frechet point-cloud distance.  Same module tree and parameter names (the SpareNet checkpoint the
reference downloads loads with `load_state_dict`), inference only.

The arithmetic runs behind the C ABI (`r2dm_pointnet_features`): the point-wise Conv1d + BatchNorm1d + ReLU layers
are 1x1 tensor-core convolutions (`conv_umma_kernel`) over the point image with the global max pool fused into
the epilogue of the 128 -> 1024 layer (the [B,1024,N] tensor never exists), BatchNorm (eval) is folded into the
weights on the host.  CUDA only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L


def _fold(conv_or_fc: nn.Module, bn: nn.BatchNorm1d | None):
    """(weight [out, in], bias [out]) of `bn(layer(x))` in eval mode."""
    w = conv_or_fc.weight.detach().float().flatten(1)
    b = conv_or_fc.bias.detach().float()
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w.contiguous(), b.contiguous()


class STN3d(nn.Module):
    """Input transform net (pointnet.py:8-33); parameters only - evaluated inside PointNet1.forward."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv1d(3, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, 9)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)

    def folded(self):
        w3, b3 = _fold(self.fc3, None)
        b3 = b3 + torch.eye(3, device=b3.device).flatten()      # x + eye(3), pointnet.py:32
        return [_fold(self.conv1, self.bn1), _fold(self.conv2, self.bn2), _fold(self.conv3, self.bn3),
                _fold(self.fc1, self.bn4), _fold(self.fc2, self.bn5), (w3, b3)]


class PointNetfeat(nn.Module):
    """Point feature trunk (pointnet.py:36-64), global_feat=True only (the only mode evaluate.py uses)."""

    def __init__(self, global_feat: bool = True):
        super().__init__()
        if not global_feat:
            raise NotImplementedError("global_feat=False is not on the evaluation path")
        self.stn = STN3d()
        self.conv1 = nn.Conv1d(3, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.global_feat = global_feat

    def folded(self):
        return [_fold(self.conv1, self.bn1), _fold(self.conv2, self.bn2), _fold(self.conv3, self.bn3)]


class PointNet1(nn.Module):
    """pointnet.py:67-81: forward(x [B,3,N]) -> cat(x1 [1024], x2 [512], x3 [256], x4 [k])."""

    def __init__(self, k: int = 2, precision: str = "fp32"):
        super().__init__()
        self.feat = PointNetfeat(global_feat=True)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.k = k
        self.precision = precision      # "fp32" = tf32 tensor cores, "bf16"
        self._cache = None              # (parameter versions, device) -> folded weights, scratch

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_cache"] = None
        return d

    def _folded(self):
        return (self.feat.stn.folded() + self.feat.folded()
                + [_fold(self.fc1, self.bn1), _fold(self.fc2, self.bn2), _fold(self.fc3, None)])

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("PointNet1 is an inference-only feature extractor here: call .eval()")
        assert x.dim() == 3 and x.shape[1] == 3, f"expected (B,3,N), but got {tuple(x.shape)}"
        if not x.is_cuda:
            raise L.R2dmError("PointNet1: expected a CUDA tensor (r2dm_b200 has no CPU path)")
        B, _, N = x.shape
        if N % 128:
            raise ValueError("number of points must be a multiple of 128 (the flattened range image)")
        pts = L.f32c(x)
        # BatchNorm folding is redone only when a parameter / buffer changed (in-place edits bump ._version)
        key = (tuple((id(t), t._version) for t in list(self.parameters()) + list(self.buffers())), pts.device)
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, [(w.to(pts.device), b.to(pts.device)) for w, b in self._folded()], {})
        params, scratches = self._cache[1], self._cache[2]
        st = L.R2dmPointNetWeights()
        for i, (w, b) in enumerate(params):
            st.weight[i] = w.data_ptr()
            st.bias[i] = b.data_ptr()
        st.num_classes = self.k
        dt = L.F32 if self.precision == "fp32" else L.BF16
        out = torch.empty(B, 1024 + 512 + 256 + self.k, device=pts.device, dtype=torch.float32)
        with torch.cuda.device(pts.device):
            n = L.lib().r2dm_pointnet_scratch_bytes(dt, B, N)
            scratch = scratches.get((dt, B, N))
            if scratch is None:
                scratches.clear()
                scratch = scratches[(dt, B, N)] = torch.empty(n, dtype=torch.uint8, device=pts.device)
            L.check(L.lib().r2dm_pointnet_features(dt, L.ptr(pts), C.byref(st), L.ptr(out), B, N, L.ptr(scratch), n,
                                                   L.stream_ptr()), "r2dm_pointnet_features")
        return out


def pretrained_pointnet(dataset: str = "shapenet", device="cuda", compile: bool = True, ckpt: str | None = None):
    """pointnet.py:84-98.  `ckpt`: local path of the state dict (otherwise the reference's URL is fetched through
    torch.hub, which needs network access); `compile` is accepted and ignored (the forward is hand-written kernels)."""
    if dataset != "shapenet":
        raise ValueError(f"Unknown dataset: {dataset}")
    model = PointNet1(k=16)
    if ckpt is not None:
        state_dict = torch.load(ckpt, map_location="cpu")
    else:
        from torch.hub import load_state_dict_from_url
        state_dict = load_state_dict_from_url(
            url="https://github.com/microsoft/SpareNet/raw/main/Frechet/cls_model_39.pth", progress=True)
    model.load_state_dict(state_dict)
    model.eval().requires_grad_(False)
    return model.to(device)


__all__ = ["STN3d", "PointNetfeat", "PointNet1", "pretrained_pointnet"]
