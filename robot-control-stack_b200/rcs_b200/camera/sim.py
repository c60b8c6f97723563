"""`SimCameraSet` on the batched backend: depth observations for every environment at once.

The reference renders every camera with OpenGL (src/sim/camera.cpp:100-140: mjv_updateScene, mjr_render,
mjr_readPixels) and python/rcs/camera/sim.py:45-115 turns the z-buffer into metres (`near / (1 - d (1 - near / far))`
with near / far = vis.map.znear / zfar x stat.extent), scales by DEPTH_SCALE = 1000 into uint16, flips the rows and
attaches intrinsics (from cam_fovy) and extrinsics (camera pose with a pi rotation about x, inverted). Here one kernel
(rcsb_camera_depth) casts a ray per pixel against the COLLISION geoms of the scene (convex hulls for meshes): the same
conventions and output format, a coarser scene than the rasterised visual meshes. Colour frames are not produced
(`Frame.camera.color` is None): there is no rasteriser in this backend.
"""
from __future__ import annotations

import enum

import numpy as np
import torch

from rcs_b200 import common, devmodel
from rcs_b200.camera.interface import BaseCameraSet, CameraFrame, DataFrame, Frame, FrameSet

ZNEAR, ZFAR = 0.01, 50.0  # mjVisual.map defaults; no shipped scene overrides them


class CameraType(enum.IntEnum):  # src/sim/camera.h:19-24
    free = 0
    tracking = 1
    fixed = 2
    default_free = 3


class SimCameraConfig:  # src/sim/camera.h:26-34, rcs.cpp BaseCameraConfig
    def __init__(self, identifier: str, frame_rate: int, resolution_width: int, resolution_height: int,
                 type: CameraType = CameraType.fixed):
        self.identifier, self.frame_rate = identifier, frame_rate
        self.resolution_width, self.resolution_height, self.type = resolution_width, resolution_height, type


class SimCameraSet(BaseCameraSet):
    def __init__(self, simulation, cameras: dict[str, SimCameraConfig], physical_units: bool = False, render_on_demand: bool = True):
        self._sim, self.cameras, self.physical_units, self.render_on_demand = simulation, cameras, physical_units, render_on_demand
        M = simulation._M
        self._cams = {}
        for name, cfg in cameras.items():
            if cfg.type != CameraType.fixed:
                raise NotImplementedError("only CameraType.fixed cameras (the MJCF <camera> elements) are supported")
            if cfg.identifier not in list(M["cam_names"]):
                raise RuntimeError(f"No camera named {cfg.identifier}")  # as mj_name2id failing, camera.cpp:33-38
            cid = list(M["cam_names"]).index(cfg.identifier)
            body, pos, rot = devmodel.fold_frame(M, int(M["cam_bodyid"][cid]), M["cam_pos"][cid], M["cam_quat"][cid])
            self._cams[name] = dict(id=cid, body=body, pos=pos, rot=rot, fovy=float(M["cam_fovy"][cid]))
        extent = float(M.get("stat_extent", 1.0))
        self._near, self._far = ZNEAR * extent, ZFAR * extent
        self._latest: FrameSet | None = None
        self._flip = torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64, device=simulation.batch.dev)

    # ---- reference surface
    def buffer_size(self) -> int:
        return 0 if self._latest is None else 1

    def clear_buffer(self):
        self._latest = None

    @property
    def camera_names(self) -> list[str]:
        return list(self.cameras.keys())

    @property
    def name_to_identifier(self) -> dict[str, str]:
        return {name: cfg.identifier for name, cfg in self.cameras.items()}

    def config(self, camera_name: str) -> SimCameraConfig:
        return self.cameras[camera_name]

    def calibrate(self) -> bool:
        return True

    def close(self):
        pass

    def _intrinsics(self, camera_name) -> np.ndarray:  # camera/sim.py:97-107
        cfg, fovy = self.cameras[camera_name], self._cams[camera_name]["fovy"]
        fx = fy = 0.5 * cfg.resolution_height / np.tan(fovy * np.pi / 360)
        return np.array([[fx, 0, (cfg.resolution_width - 1) / 2, 0], [0, fy, (cfg.resolution_height - 1) / 2, 0], [0, 0, 1, 0]])

    def _extrinsics(self, cam_frames: torch.Tensor) -> torch.Tensor:  # camera/sim.py:109-119: (cam * Rx(pi))^-1 as 4 x 4
        """cam_frames [n, 12]: the camera's world frame of every environment, as the depth kernel wrote it."""
        p, R = cam_frames[:, :3], cam_frames[:, 3:].reshape(-1, 3, 3)
        Rc = R * self._flip                      # columns y and z negated: the pi rotation about x
        E = torch.zeros((p.shape[0], 4, 4), dtype=torch.float64, device=p.device)
        E[:, :3, :3] = Rc.transpose(1, 2)
        E[:, :3, 3] = -torch.einsum("nji,nj->ni", Rc, p)
        E[:, 3, 3] = 1
        return E[0] if self._sim.num_envs == 1 else E

    def render(self) -> FrameSet:
        b = self._sim.batch
        frames = {}
        ts = b.time.clone()
        for name, cfg in self.cameras.items():
            c = self._cams[name]
            cf = torch.empty((b.n, 12), dtype=torch.float64, device=b.dev)
            d = b.camera_depth(c["body"], c["pos"], c["rot"], c["fovy"], cfg.resolution_width, cfg.resolution_height, self._near,
                               self._far, self.physical_units, cam_frames=cf)
            depth = DataFrame(data=d.unsqueeze(-1), timestamp=ts, intrinsics=self._intrinsics(name), extrinsics=self._extrinsics(cf))
            frames[name] = Frame(camera=CameraFrame(color=None, depth=depth), avg_timestamp=ts)
        self._latest = FrameSet(frames=frames, avg_timestamp=ts)
        return self._latest

    def get_latest_frames(self) -> FrameSet | None:
        """render_on_demand (the reference's default): render now, from the current state."""
        if self.render_on_demand or self._latest is None:
            return self.render()
        return self._latest

    def get_timestamp_frames(self, ts) -> FrameSet | None:
        return self.get_latest_frames()


class CameraSetWrapper:
    """python/rcs/envs/base.py:585-677 for the vector env: adds obs["frames"][camera]["depth"] (a dict with data /
    intrinsics / extrinsics) after every reset() / step(); rgb is absent (depth-only backend)."""
    RGB_KEY, DEPTH_KEY, CAMERA_KEY = "rgb", "depth", "frames"

    def __init__(self, env, camera_set: SimCameraSet, include_depth: bool = True):
        self.env, self.camera_set, self.include_depth = env, camera_set, include_depth

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _observe(self, obs: dict, info: dict):
        fs = self.camera_set.get_latest_frames()
        obs = dict(obs)
        obs[self.CAMERA_KEY] = {
            name: {self.DEPTH_KEY: dict(data=f.camera.depth.data, intrinsics=f.camera.depth.intrinsics, extrinsics=f.camera.depth.extrinsics)}
            for name, f in fs.frames.items()}
        info = dict(info)
        info["camera_available"] = True
        info["frame_timestamp"] = fs.avg_timestamp
        return obs, info

    def reset(self, seed=None, options=None):
        self.camera_set.clear_buffer()
        obs, info = self.env.reset(seed=seed, options=options)
        return self._observe(obs, info)

    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        obs, info = self._observe(obs, info)
        return obs, rew, term, trunc, info

    def close(self):
        self.camera_set.close()
