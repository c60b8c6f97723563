"""Vector Gym-style env over the fused per-env programs of the CUDA backend.

One `step()` is ONE kernel launch that restates, per environment, the reference wrapper stack
RelativeActionSpace -> GripperWrapperSim -> GripperWrapper -> RobotSimWrapper -> RobotEnv
(/root/reference/python/rcs/envs/base.py:246-288,469-488,710-735; envs/sim.py:49-76,125-131): action
transform, gripper command, dedupe against the previous action, set_joint_position, Sim.step(k) or
step_until_convergence, then the observation / info pack. Keys and conventions follow the reference
(`tquat` = xyz + quat(xyzw), `joints`, `xyzrpy`, `gripper`; info `collision`, `ik_success`,
`is_sim_converged`, `gripper_width`, `is_grasped`), with a leading env axis.
"""
from __future__ import annotations

import numpy as np
import torch

from rcs_b200 import _lib, common
from rcs_b200.envs.base import Box, ControlMode, Dict, RelativeTo


class SimVectorEnv:
    DEFAULT_MAX_JOINT_MOV = np.deg2rad(5)
    DEFAULT_MAX_CART_MOV = 0.5
    DEFAULT_MAX_CART_ROT = np.deg2rad(90)

    def __init__(self, simulation, robot, gripper, control_mode: ControlMode, max_relative_movement=None,
                 relative_to: RelativeTo = RelativeTo.LAST_STEP, binary_gripper: bool = True):
        self.sim, self.robot, self.gripper = simulation, robot, gripper
        self.control_mode, self.relative_to = control_mode, relative_to
        self.num_envs = simulation.num_envs
        self.relative = max_relative_movement is not None
        self.max_mov = max_relative_movement
        b = simulation.batch
        self.dev = b.dev
        self.obs_dim = b.model.obs_dim
        meta = common.robots_meta_config(robot.get_config().robot_type)
        self.jlow, self.jhigh = meta.joint_limits[0].copy(), meta.joint_limits[1].copy()
        self.dof = meta.dof
        if control_mode != ControlMode.JOINTS and self.relative:  # base.py:377-390
            mm = self.max_mov
            if isinstance(mm, (int, float)):
                mm = (float(mm), self.DEFAULT_MAX_CART_ROT)
            assert isinstance(mm, tuple) and len(mm) == 2, \
                "in cartesian control max_mov must be a tuple of maximum translation (in m) and maximum rotation in (rad)"
            self.max_mov = (float(mm[0]), float(mm[1]))
        self.binary_gripper = binary_gripper
        self._origin = None        # RelativeActionSpace._origin / _last_action for CONFIGURED_ORIGIN (base.py:443-467)
        self._last_action = None
        spaces = {}
        if control_mode == ControlMode.JOINTS:
            if self.relative:
                assert isinstance(self.max_mov, float), "in joint control max_mov must be a float (rad)"
                spaces["joints"] = Box([-self.max_mov] * self.dof, [self.max_mov] * self.dof, self.dev)
            else:
                spaces["joints"] = Box(self.jlow, self.jhigh, self.dev)
        elif control_mode == ControlMode.CARTESIAN_TRPY:
            if self.relative:  # LimitedTRPYRelDictType, base.py:40-50
                mt, mr = self.max_mov
                spaces["xyzrpy"] = Box([-mt] * 3 + [-mr] * 3, [mt] * 3 + [mr] * 3, self.dev)
            else:
                spaces["xyzrpy"] = Box([-0.855, -0.855, 0, -np.pi, -np.pi, -np.pi], [0.855, 0.855, 1.188, np.pi, np.pi, np.pi], self.dev)
        else:
            if self.relative:  # LimitedTQuatRelDictType, base.py:63-73
                mt = self.max_mov[0]
                spaces["tquat"] = Box([-mt] * 3 + [-1] * 4, [mt] * 3 + [1] * 4, self.dev)
            else:
                spaces["tquat"] = Box([-0.855, -0.855, 0, -1, -1, -1, -1], [0.855, 0.855, 1.188, 1, 1, 1, 1], self.dev)
        if gripper is not None:
            spaces["gripper"] = Box(np.zeros(()), np.ones(()), self.dev)
        self.action_space = Dict(spaces, self.num_envs)
        self._zeros = torch.zeros(self.num_envs, dtype=torch.float64, device=self.dev)
        self._false = torch.zeros(self.num_envs, dtype=torch.bool, device=self.dev)
        # pinned host staging for the host-buffer path
        self._h_obs = None

    # ------------------------------------------------------------------ helpers
    def _substeps(self):
        cfg = self.sim._cfg  # read-only use: no copy on the per-step path
        return round(1 / cfg.frequency / self.sim.model.opt.timestep)  # envs/sim.py:53

    def _step_ops(self):
        cfg = self.sim._cfg
        ops = _lib.OBS | (_lib.STEP_K if cfg.async_control else _lib.STEP_CONV)
        if self.control_mode == ControlMode.JOINTS:
            ops |= _lib.ACT_JOINTS_REL if self.relative else _lib.ACT_JOINTS_ABS
        if self.gripper is not None:
            ops |= _lib.ACT_GRIPPER_BIN if self.binary_gripper else _lib.ACT_GRIPPER_CONT
        return ops, cfg

    def unpack(self, o: torch.Tensor):
        """Packed observation rows [n, obs_dim] (any n: the local block or the gathered block of all ranks) -> the
        reference's obs / info dicts and the truncated flag. Columns: tquat 0:7, joints 7:14, xyzrpy 14:20, gripper
        command 20, gripper width 21, then collision, ik_success, is_sim_converged, is_grasped, truncated, robot /
        gripper collision, convergence steps as reals."""
        obs = {"tquat": o[:, 0:7], "joints": o[:, 7:7 + self.dof], "xyzrpy": o[:, 14:20]}
        info = {"collision": o[:, 22] != 0, "ik_success": o[:, 23] != 0, "is_sim_converged": o[:, 24] != 0}
        if self.gripper is not None:
            obs["gripper"] = o[:, 20] if self.binary_gripper else o[:, 21]  # last command | normalised width (base.py:710-718)
            info["gripper_width"] = o[:, 21]
            info["is_grasped"] = o[:, 25] != 0
        return obs, info, o[:, 26] != 0

    def _pack(self):
        return self.unpack(self.sim.batch.obs)

    # ------------------------------------------------------------------ gym API
    def reset_packed(self, obs_out: torch.Tensor | None = None) -> torch.Tensor:
        """reset() that returns the packed observation rows (written into obs_out when given)."""
        b = self.sim.batch
        ops = _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
        if self.gripper is not None:
            ops |= _lib.GRIPPER_RESET
        b.run(ops, k=1, want_obs=True, fresh_obs=True, obs_out=obs_out)
        return b.obs

    def reset(self, seed=None, options=None):
        """envs/sim.py:68-76 under base.py:703-708 and base.py:462-467: gripper.reset(); sim.reset();
        robot.reset(); sim.step(1); get_obs()."""
        b = self.sim.batch
        obs, info, _ = self.unpack(self.reset_packed())
        if self.relative and self.relative_to == RelativeTo.CONFIGURED_ORIGIN:  # base.py:462-467: set_origin_to_current()
            if self.control_mode == ControlMode.JOINTS:
                self._origin = obs["joints"].clone()
                self._last_action = None
            else:  # origin pose, last clipped offset and its flag live on the device, next to the state rows
                self._origin = obs["tquat"].clone().contiguous()
                self._last_action = torch.zeros((self.num_envs, 7), dtype=torch.float64, device=self.dev)
                self._have_last = torch.zeros((self.num_envs,), dtype=torch.int32, device=self.dev)
        return obs, {}

    def step(self, action: dict):
        obs, info, truncated = self.unpack(self.step_packed(action))  # tensors allocated for this step: never overwritten later
        return obs, torch.zeros_like(self._zeros), torch.zeros_like(self._false), truncated, info

    def step_packed(self, action: dict, obs_out: torch.Tensor | None = None, mask: torch.Tensor | None = None) -> torch.Tensor:
        """env.step() that returns the packed observation rows [n, obs_dim] (see unpack); with obs_out the kernel writes
        them straight into that tensor (this rank's slice of the multi-GPU gather buffer). mask ([n] uint8, JOINTS control):
        environments with mask == 0 are left untouched -- not a bit of their state changes and their rows of the result
        are unspecified (CollisionGuard overwrites them with the last observation)."""
        b = self.sim.batch
        ops, cfg = self._step_ops()
        aj = ag = None
        if self.control_mode == ControlMode.JOINTS:
            if "joints" not in action:
                raise RuntimeError("Given type is not matching control mode!")  # base.py:257-266
            aj = action["joints"].to(device=self.dev, dtype=torch.float64).contiguous()
            if self.relative and self.relative_to == RelativeTo.CONFIGURED_ORIGIN:
                # base.py:479-488: the offset may move by at most max_mov per step; the origin stays where reset() left it
                if self._origin is None:
                    self._origin = b.qpos[:, :self.dof].clone()
                mm = float(self.max_mov)
                lim = aj.clamp(-mm, mm) if self._last_action is None else (aj - self._last_action).clamp(-mm, mm) + self._last_action
                self._last_action = lim
                lo = torch.as_tensor(self.jlow, device=self.dev); hi = torch.as_tensor(self.jhigh, device=self.dev)
                aj = torch.minimum(torch.maximum(self._origin + lim, lo), hi).contiguous()
                ops = (ops & ~_lib.ACT_JOINTS_REL) | _lib.ACT_JOINTS_ABS
        else:
            key = "xyzrpy" if self.control_mode == ControlMode.CARTESIAN_TRPY else "tquat"
            if key not in action:
                raise RuntimeError("Given type is not matching control mode!")
            # relative offset / clip / dedupe / IK / set_joint_position for every env in one launch
            a = action[key].to(device=self.dev, dtype=torch.float64).contiguous()
            kind = 0 if self.control_mode == ControlMode.CARTESIAN_TRPY else 1
            assert a.shape == (self.num_envs, 6 if kind == 0 else 7)
            mt, mr = self.max_mov if self.relative else (0.0, 0.0)
            if self.relative and self.relative_to == RelativeTo.CONFIGURED_ORIGIN:
                if self._origin is None:  # step() before reset(): the origin is where the robot is now
                    self._origin = self.robot.get_cartesian_position_tensor().contiguous()
                    self._last_action = torch.zeros((self.num_envs, 7), dtype=torch.float64, device=self.dev)
                    self._have_last = torch.zeros((self.num_envs,), dtype=torch.int32, device=self.dev)
                _lib.check(_lib.lib().rcsb_env_cartesian_action_origin(
                    b.ptr, a.data_ptr(), kind, 2, float(mt), float(mr), self._origin.data_ptr(), self._last_action.data_ptr(),
                    self._have_last.data_ptr()))
            else:
                _lib.check(_lib.lib().rcsb_env_cartesian_action(b.ptr, a.data_ptr(), kind, int(self.relative), float(mt), float(mr)))
        if self.gripper is not None:
            assert "gripper" in action, "Gripper action not found."  # base.py:724
            ag = action["gripper"].to(device=self.dev, dtype=torch.float64).reshape(-1).contiguous()
        b.run(ops, k=self._substeps(), max_convergence_steps=cfg.max_convergence_steps, act_joints=aj, act_gripper=ag,
              max_mov=float(self.max_mov) if (self.relative and self.control_mode == ControlMode.JOINTS) else 0.0,
              jlow=self.jlow, jhigh=self.jhigh, want_obs=True, fresh_obs=True, obs_out=obs_out, mask=mask)
        return b.obs

    def _to_pose7(self, a: torch.Tensor) -> torch.Tensor:
        a = a.to(device=self.dev, dtype=torch.float64)
        if a.shape[-1] == 7:
            q = a[:, 3:7]
            return torch.cat([a[:, :3], q / q.norm(dim=1, keepdim=True)], dim=1).contiguous()
        r, p, y = a[:, 3] / 2, a[:, 4] / 2, a[:, 5] / 2  # Rz(yaw) Ry(pitch) Rx(roll), Pose.h:37-43
        cr, sr, cp, sp, cy, sy = r.cos(), r.sin(), p.cos(), p.sin(), y.cos(), y.sin()
        q = torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                         cr * cp * cy + sr * sp * sy], dim=1)
        return torch.cat([a[:, :3], q], dim=1).contiguous()

    # ------------------------------------------------------------------ host-buffer path (what a CPU-side policy sees)
    def step_host(self, act_host: torch.Tensor) -> torch.Tensor:
        """The same env.step() for a policy that lives on the HOST: act_host is a pinned [n, njoints + 1] float64 block
        (joint action, then the gripper action). One H2D copy, the fused launch, one D2H copy of the packed observation
        rows and a stream synchronise happen inside the call (C ABI rcsb_env_step_host). Returns the pinned [n, obs_dim]
        staging tensor, which the NEXT call overwrites: unpack() / copy what must outlive a step."""
        b = self.sim.batch
        if self.control_mode != ControlMode.JOINTS or (self.relative and self.relative_to != RelativeTo.LAST_STEP) or self.gripper is None:
            raise NotImplementedError("step_host covers joint control with a gripper (absolute or relative to the last step)")
        if self._h_obs is None:
            self._h_obs = torch.zeros((self.num_envs, b.model.obs_dim), dtype=torch.float64).pin_memory()
        assert act_host.shape == (self.num_envs, self.dof + 1) and act_host.dtype == torch.float64 and act_host.is_contiguous()
        ops, cfg = self._step_ops()
        b.step_host(ops, self._substeps(), cfg.max_convergence_steps, act_host,
                    float(self.max_mov) if self.relative else 0.0, self.jlow, self.jhigh, self._h_obs)
        return self._h_obs

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    def get_wrapper_attr(self, name):
        return getattr(self, name)
