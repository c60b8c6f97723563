"""`MultiRobotWrapper` and `CollisionGuard` for vector envs.

Reference: python/rcs/envs/base.py:310-355 (a dict of envs stepped as one) and python/rcs/envs/sim.py:156-287 (a shadow
simulation that executes every action first and vetoes the ones that collide). `MultiSimRobotWrapper` (several arms in ONE
mjData, envs/sim.py:79-117) is broken in the reference and needs several SimRobots per
environment, which the per-warp device layer does not hold: not offered.

Batched semantics: everything is per environment. The guard's shadow env is a second batch of the same size; a colliding
action of environment e is vetoed for e alone (the others step normally): e is masked out of the launch, so not a bit of
its state changes -- the reference returns before stepping -- and it reports its last observation with terminated =
truncated = True; without truncate_on_collision e holds its current joint position instead, as in the reference.
"""
from __future__ import annotations

import torch

from rcs_b200.envs.base import ControlMode


class MultiRobotWrapper:
    """base.py:310-355 over vector envs: reward is the sum, terminated / truncated the OR over the robots (per environment
    index when the envs have equal size)."""

    def __init__(self, envs: dict):
        self.envs = envs
        self.unwrapped_multi = {k: e.unwrapped for k, e in envs.items()}

    def step(self, action: dict):
        obs, info = {}, {}
        reward = terminated = truncated = None
        for key, env in self.envs.items():
            obs[key], r, t, tr, info[key] = env.step(action[key])
            reward = r if reward is None else reward + r
            terminated = t if terminated is None else terminated | t
            truncated = tr if truncated is None else truncated | tr
            info[key]["terminated"], info[key]["truncated"] = t, tr
        return obs, reward, terminated, truncated, info

    def reset(self, seed: dict | None = None, options: dict | None = None):
        obs, info = {}, {}
        seed_ = seed if seed is not None else {k: None for k in self.envs}
        options_ = options if options is not None else {k: None for k in self.envs}
        for key, env in self.envs.items():
            obs[key], info[key] = env.reset(seed=seed_[key], options=options_[key])
        return obs, info

    def get_wrapper_attr(self, name: str):
        if name in self.__dir__():
            return getattr(self, name)
        return {k: e.get_wrapper_attr(name) for k, e in self.envs.items()}

    def close(self):
        for e in self.envs.values():
            e.close()


class CollisionGuard:
    """envs/sim.py:156-287. `env` and `collision_env` are SimVectorEnv instances of equal size in JOINTS control with
    ABSOLUTE actions (the reference: "RelativeActionSpace has to be added after this ... the input expects absolute
    actions"); with to_joint_control the collision env runs a Cartesian control mode and hands its IK solution on as the
    joint action of the guarded env."""

    def __init__(self, env, simulation, collision_env, check_home_collision: bool = True, to_joint_control: bool = False,
                 sim_gui: bool = False, truncate_on_collision: bool = True):
        if sim_gui:
            raise NotImplementedError("the GUI bridge is out of scope (SURVEY.md 8f-3)")
        self.env, self.sim, self.collision_env = env, simulation, collision_env
        self.check_home_collision, self.to_joint_control = check_home_collision, to_joint_control
        self.truncate_on_collision = truncate_on_collision
        assert env.unwrapped.control_mode == ControlMode.JOINTS and not env.unwrapped.relative, \
            "the guarded env takes absolute joint actions"
        assert collision_env.num_envs == env.num_envs
        self.last_obs = None
        self.action_space = collision_env.action_space if to_joint_control else env.action_space

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action: dict):
        env, cenv = self.env.unwrapped, self.collision_env.unwrapped
        cur = env.robot.get_joint_position_tensor()
        cenv.robot.set_joints_hard(cur)                       # the shadow robot starts where the real one is
        _, _, _, _, cinfo = self.collision_env.step(action)
        coll = cinfo["collision"]
        act = dict(action)
        if self.to_joint_control:
            act = {k: v for k, v in act.items() if k == "gripper"}
            act["joints"] = cenv.robot.get_joint_position_tensor()
        joints = act["joints"].to(device=cur.device, dtype=torch.float64).clone()
        joints[coll] = cur[coll]                              # a vetoed action becomes "stay where you are"
        act["joints"] = joints
        if self.truncate_on_collision:
            if self.last_obs is None and bool(coll.any()):
                raise RuntimeError("Collision detected in the first step!")
            rows = env.step_packed(act, mask=(~coll).to(torch.uint8).contiguous())
            if self.last_obs is not None:
                rows = torch.where(coll.unsqueeze(1), self.last_obs, rows)
            obs, info, truncated = env.unpack(rows)
            self.last_obs = rows
            zeros = torch.zeros(env.num_envs, dtype=torch.float64, device=rows.device)
            info["guard_collision"] = coll
            return obs, zeros, coll.clone(), truncated | coll, info
        rows = env.step_packed(act)
        obs, info, truncated = env.unpack(rows)
        self.last_obs = rows
        info["guard_collision"] = coll
        zeros = torch.zeros(env.num_envs, dtype=torch.float64, device=rows.device)
        return obs, zeros, torch.zeros_like(coll), truncated, info

    def reset(self, seed=None, options=None):
        cenv = self.collision_env.unwrapped
        if self.check_home_collision:  # is the way home free?
            cenv.robot.move_home()
            cenv.sim.step_until_convergence()
            b = cenv.sim.batch
            if bool((b.si[:, 1] != 0).any()) or bool((b.si[:, 0] == 0).any()):
                raise RuntimeError("Collision detected while moving to home position!")
        else:
            cenv.robot.reset()
        obs, info = self.env.reset(seed=seed, options=options)
        self.last_obs = self.env.unwrapped.sim.batch.obs
        return obs, info

    def close(self):
        self.env.close()
        self.collision_env.close()
