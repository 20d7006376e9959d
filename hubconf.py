"""torch.hub entry point - mirror of the reference's hubconf.py:21-37 for the sampling path.

    ddpm, lidar_utils, cfg = torch.hub.load("<this repo>", "pretrained_r2dm", source="local",
                                            ckpt="r2dm-h-kitti360-300k.pth", device="cuda")

The RangeNet++ entry points of the reference's hubconf (:45-115) belong to evaluation, which is out of
scope here.
"""
from torch.hub import load_state_dict_from_url

from r2dm_b200.inference import setup_model as _setup_model

dependencies = ["torch", "numpy"]


def _get_r2dm_url(key: str) -> str:
    return f"https://github.com/kazuto1011/r2dm/releases/download/weights/{key}.pth"


def pretrained_r2dm(config: str = "r2dm-h-kitti360-300k", ckpt: str = None, **kwargs):
    """R2DM models of "LiDAR Data Synthesis with Denoising Diffusion Probabilistic Models"
    (https://arxiv.org/abs/2309.09256), running on the B200-native kernels.

    Args:
        config: name of a released checkpoint (downloaded when `ckpt` is None; needs network access).
        ckpt: path to a checkpoint file or an already loaded checkpoint dict; `config` is then ignored.
        **kwargs: forwarded to `setup_model` (device, ema, show_info, compile, precision).

    Returns:
        (ddpm, lidar_utils, cfg) exactly like the reference.
    """
    if ckpt is None:
        ckpt = load_state_dict_from_url(_get_r2dm_url(config), map_location="cpu")
    ddpm, lidar_utils, cfg = _setup_model(ckpt, **kwargs)
    return ddpm, lidar_utils, cfg
