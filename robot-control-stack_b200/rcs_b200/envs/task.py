"""Task layer of the pick-up environment on the batched backend, mirroring the reference wrappers
(/root/reference/python/rcs/envs/sim.py:290-431): `RandomCubePos` / `RandomObjectPos` re-place the free object after every
reset, `PickCubeSuccessWrapper` adds the success flag and the shaped reward. Everything is per-environment elementwise
arithmetic on the state rows the kernels own (column views of `Batch.sr`), so it stays on the device: no host round trip
inside reset() / step().
"""
from __future__ import annotations

import numpy as np
import torch

from rcs_b200 import common


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):  # gym.Wrapper-style attribute forwarding
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    def reset(self, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)


def _joint_qpos_slice(sim, joint_name: str) -> slice:
    M = sim._M
    names = list(M["jnt_names"])
    if joint_name not in names:
        raise RuntimeError(f"No joint named {joint_name}")
    j = names.index(joint_name)
    adr = int(M["jnt_qposadr"][j])
    return slice(adr, adr + 7)


class RandomObjectPos(_Wrapper):
    """envs/sim.py:290-356: x, y of the object's free joint are re-drawn in +-0.1 m around the initial pose after every
    reset (z fixed); with include_rotation the quaternion's w is re-drawn as in the reference."""

    def __init__(self, env, simulation, joint_name: str, init_object_pose: common.Pose, include_position: bool = True,
                 include_rotation: bool = False, generator: torch.Generator | None = None):
        super().__init__(env)
        self.sim = simulation
        self.joint_name = joint_name
        self.init_object_pose = init_object_pose
        self.include_position, self.include_rotation = include_position, include_rotation
        self.generator = generator
        self._qs = _joint_qpos_slice(simulation, joint_name)

    def _rand(self, n):
        b = self.sim.batch
        return torch.rand((n,), dtype=torch.float64, device=b.dev, generator=self.generator)

    def _place(self, x0, y0, z, quat_xyzw):
        b = self.sim.batch
        n = b.n
        q = b.qpos[:, self._qs]
        q[:, 0] = x0 + (self._rand(n) * 0.2 - 0.1 if self.include_position else 0.0)
        q[:, 1] = y0 + (self._rand(n) * 0.2 - 0.1 if self.include_position else 0.0)
        q[:, 2] = z
        q[:, 3] = (2 * self._rand(n) - quat_xyzw[3]) if self.include_rotation else quat_xyzw[3]
        q[:, 4], q[:, 5], q[:, 6] = quat_xyzw[0], quat_xyzw[1], quat_xyzw[2]

    def reset(self, seed=None, options=None):
        if options is not None and "RandomObjectPos.init_object_pose" in options:
            assert isinstance(options["RandomObjectPos.init_object_pose"], common.Pose), \
                "RandomObjectPos.init_object_pose must be a rcs.common.Pose"
            self.init_object_pose = options.pop("RandomObjectPos.init_object_pose")
        obs, info = self.env.reset(seed=seed, options=options)
        self.sim.step(1)
        t = self.init_object_pose.translation()
        self._place(t[0], t[1], t[2], self.init_object_pose.rotation_q())
        return obs, info


class RandomCubePos(RandomObjectPos):
    """envs/sim.py:359-384: the cube of fr3_simple_pick_up is re-placed around (0.498, 0, 0.226) in the robot frame."""

    def __init__(self, env, simulation, include_rotation: bool = True, generator: torch.Generator | None = None):
        super().__init__(env, simulation, "box_joint", common.Pose(), include_position=True, include_rotation=include_rotation,
                         generator=generator)

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self.sim.step(1)
        iso = common.Pose(translation=np.array([0.498, 0.0, 0.226]), rpy_vector=np.zeros(3))
        iso_w = self.unwrapped.robot.to_pose_in_world_coordinates(iso).translation()
        b = self.sim.batch
        n = b.n
        q = b.qpos[:, self._qs]
        q[:, 0] = iso_w[0] + self._rand(n) * 0.2 - 0.1
        q[:, 1] = iso_w[1] + self._rand(n) * 0.2 - 0.1
        q[:, 2] = 0.0288 / 2
        q[:, 3] = (2 * self._rand(n) - 1) if self.include_rotation else 0.0
        q[:, 4], q[:, 5], q[:, 6] = 0.0, 0.0, 1.0
        return obs, info


class PickCubeSuccessWrapper(_Wrapper):
    """envs/sim.py:387-431: success = cube lifted above 0.15 + 0.852 m with the gripper closed; otherwise the ManiSkill
    style shaped reward (reach + grasp + place), everything divided by 5. `terminated` = success."""

    EE_HOME = np.array([0.34169773, 0.00047028, 0.4309004])
    BINARY_GRIPPER_CLOSED = 0

    def __init__(self, env):
        super().__init__(env)
        self.sim = env.get_wrapper_attr("sim")
        self._qs = _joint_qpos_slice(self.sim, "box_joint")
        self._home = None

    def step(self, action):
        obs, reward, _, truncated, info = self.env.step(action)
        b = self.sim.batch
        box = b.qpos[:, self._qs][:, :3]
        if self._home is None:
            self._home = torch.as_tensor(self.EE_HOME, dtype=torch.float64, device=b.dev)
        success = (box[:, 2] > 0.15 + 0.852) & (obs["gripper"] == self.BINARY_GRIPPER_CLOSED)
        info["success"] = success
        tcp_to_obj = (box - obs["tquat"][:, :3]).norm(dim=1)
        obj_to_goal = (box - self._home).norm(dim=1)
        grasped = info["is_grasped"].to(torch.float64)
        shaped = (1 - torch.tanh(5 * tcp_to_obj)) + grasped + (1 - torch.tanh(5 * obj_to_goal)) * grasped
        reward = torch.where(success, torch.full_like(shaped, 5.0), shaped) / 5
        return obs, reward, success, truncated, info
