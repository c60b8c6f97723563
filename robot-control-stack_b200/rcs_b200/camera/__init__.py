"""Mirror of `rcs.camera` (python/rcs/camera/{interface,sim}.py) for the batched backend: depth frames only."""
from .interface import BaseCameraSet, CameraFrame, DataFrame, Frame, FrameSet  # noqa: F401
from .sim import CameraType, SimCameraConfig, SimCameraSet  # noqa: F401
