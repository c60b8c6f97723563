"""Frame containers of python/rcs/camera/interface.py:13-49 (same field names; `data` is a device tensor with a leading
environment axis)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any


@dataclass(kw_only=True)
class DataFrame:
    data: Any
    timestamp: Any = None      # simulation time of every environment
    intrinsics: Any = None     # 3 x 4
    extrinsics: Any = None     # [num_envs, 4, 4] (or 4 x 4 for one environment)


@dataclass(kw_only=True)
class CameraFrame:
    color: DataFrame | None
    ir: DataFrame | None = None
    depth: DataFrame | None = None
    temperature: float | None = None


@dataclass(kw_only=True)
class Frame:
    camera: CameraFrame
    imu: Any = None
    avg_timestamp: Any = None


@dataclass(kw_only=True)
class FrameSet:
    frames: dict
    avg_timestamp: Any


class BaseCameraSet:
    DEPTH_SCALE: int = 1000  # camera/interface.py:53
