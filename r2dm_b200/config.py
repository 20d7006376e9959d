"""Configuration dataclasses with the field names of the reference's utils/option.py:6-77.

The reference's pydantic dataclasses do not import on Python >= 3.11 (mutable defaults,
utils/option.py:74-77), so this is an independent plain-dataclass equivalent that accepts the
nested dict stored in checkpoints (`Config(**ckpt["cfg"])`, utils/inference.py:29)."""
from __future__ import annotations

from dataclasses import asdict, dataclass, field, fields
from typing import Optional, Tuple


def _coerce(cls, value):
    if isinstance(value, cls):
        return value
    if value is None:
        return cls()
    if isinstance(value, dict):
        if cls is TrainingConfig:
            return cls(**value)
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in value.items() if k in known})
    raise TypeError(f"cannot build {cls.__name__} from {type(value).__name__}")


@dataclass
class ModelConfig:
    architecture: str = "efficient_unet"
    base_channels: int = 64
    temb_channels: Optional[int] = None
    channel_multiplier: Tuple[int, int, int, int] = (1, 2, 4, 8)
    num_residual_blocks: Tuple[int, int, int, int] = (3, 3, 3, 3)
    gn_num_groups: int = 32 // 4
    gn_eps: float = 1e-6
    attn_num_heads: int = 8
    coords_encoding: Optional[str] = "fourier_features"
    dropout: float = 0.0

    def __post_init__(self):
        self.channel_multiplier = tuple(self.channel_multiplier)
        self.num_residual_blocks = tuple(self.num_residual_blocks)


@dataclass
class DiffusionConfig:
    num_training_steps: Optional[int] = None
    num_sampling_steps: int = 1024
    prediction_type: str = "eps"
    loss_type: str = "l2"
    noise_schedule: str = "cosine"
    timestep_type: str = "continuous"


class TrainingConfig:
    """Training settings of a checkpoint (optimizer, EMA, mixed precision, ...; utils/option.py:31-50).
    Training is outside the scope of this package, so the section is carried through opaquely: whatever keys
    the checkpoint holds are kept, readable as attributes, and written back unchanged by `Config.to_dict()`."""

    # the two settings the sampling scripts look at (sample_and_save.py:25-27), with the reference's defaults
    _DEFAULTS = {"mixed_precision": "fp16", "dynamo_backend": "inductor"}

    def __init__(self, **settings):
        self.settings = {**self._DEFAULTS, **settings}

    def __getattr__(self, name):
        try:
            return self.__dict__["settings"][name]
        except KeyError:
            raise AttributeError(name) from None

    def __eq__(self, other):
        return isinstance(other, TrainingConfig) and self.settings == other.settings

    def __repr__(self):
        return f"TrainingConfig({self.settings})"


@dataclass
class DataConfig:
    dataset: str = "kitti_360"
    depth_format: str = "log_depth"
    projection: str = "spherical-1024"
    train_depth: bool = True
    train_reflectance: bool = True
    resolution: Tuple[int, int] = (64, 1024)
    # class attributes (not fields) in the reference, utils/option.py:68-69
    min_depth: float = 1.45
    max_depth: float = 80.0

    def __post_init__(self):
        self.resolution = tuple(self.resolution)


@dataclass
class Config:
    data: DataConfig = field(default_factory=DataConfig)
    model: ModelConfig = field(default_factory=ModelConfig)
    diffusion: DiffusionConfig = field(default_factory=DiffusionConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)

    def __post_init__(self):
        self.data = _coerce(DataConfig, self.data)
        self.model = _coerce(ModelConfig, self.model)
        self.diffusion = _coerce(DiffusionConfig, self.diffusion)
        self.training = _coerce(TrainingConfig, self.training)

    def to_dict(self) -> dict:
        return {"data": asdict(self.data), "model": asdict(self.model), "diffusion": asdict(self.diffusion),
                "training": dict(self.training.settings)}
