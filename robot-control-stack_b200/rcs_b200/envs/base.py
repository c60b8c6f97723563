"""Enums and space helpers mirroring /root/reference/python/rcs/envs/base.py:171-174,357-361 and the
Box spaces of base.py:28-168 (gymnasium is not a dependency: `Box`/`Dict` below carry only what the
vector env needs: bounds and `sample()`)."""
from __future__ import annotations

from enum import Enum, auto

import numpy as np
import torch


class ControlMode(Enum):
    JOINTS = auto()
    CARTESIAN_TRPY = auto()
    CARTESIAN_TQuat = auto()


class RelativeTo(Enum):
    LAST_STEP = auto()
    CONFIGURED_ORIGIN = auto()


class Box:
    def __init__(self, low, high, device="cpu", generator: torch.Generator | None = None):
        self.low = np.asarray(low, dtype=np.float64)
        self.high = np.asarray(high, dtype=np.float64)
        self.shape = self.low.shape
        self.device, self.generator = device, generator

    def sample(self, n: int | None = None):
        lo = torch.as_tensor(np.where(np.isfinite(self.low), self.low, -1.0), device=self.device)
        hi = torch.as_tensor(np.where(np.isfinite(self.high), self.high, 1.0), device=self.device)
        shape = (n, *self.shape) if n is not None else self.shape
        u = torch.rand(shape, dtype=torch.float64, device=self.device, generator=self.generator)
        return lo + (hi - lo) * u


class Dict:
    def __init__(self, spaces: dict, num_envs: int):
        self.spaces, self.num_envs = spaces, num_envs

    def sample(self):
        return {k: s.sample(self.num_envs) for k, s in self.spaces.items()}
