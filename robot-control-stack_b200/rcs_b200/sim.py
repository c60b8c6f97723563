"""Host-side mirror of `rcs.sim` / `rcs._core.sim` (/root/reference/python/rcs/sim/sim.py:44-62,
/root/reference/src/pybind/rcs.cpp:420-527) on top of the batched CUDA backend.

Same class names, method names, argument meaning and error behaviour as the reference; the additive
extension is `num_envs`: with num_envs == 1 (default) getters return numpy arrays / Pose / bool exactly
like the reference; with num_envs > 1 they return torch CUDA tensors with a leading env axis.
Model loading replaces `mujoco.MjModel.from_xml_path` by the in-tree MJCF compiler (rcs_b200.mjcf) or a
precompiled `.npz` scene (the analogue of the reference's build-time `.mjb`).
"""
from __future__ import annotations

import copy
import math
import sys
from dataclasses import dataclass, field
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib, batch as _batch, common, mjcf


@dataclass
class SimConfig:
    """/root/reference/src/sim/sim.h:29-34"""
    async_control: bool = False
    realtime: bool = False
    frequency: int = 30
    max_convergence_steps: int = 500


@dataclass
class SimRobotConfig(common.RobotConfig):
    """/root/reference/src/sim/SimRobot.h:14-47"""
    joint_rotational_tolerance: float = 0.05 * (math.pi / 180.0)
    seconds_between_callbacks: float = 0.1
    trajectory_trace: bool = False
    arm_collision_geoms: list = field(default_factory=lambda: [f"fr3_link{i}_collision" for i in range(8)])
    joints: list = field(default_factory=lambda: [f"fr3_joint{i}" for i in range(1, 8)])
    actuators: list = field(default_factory=lambda: [f"fr3_joint{i}" for i in range(1, 8)])
    base: str = "base"
    mjcf_scene_path: str = "assets/scenes/fr3_empty_world/scene.xml"

    def add_id(self, id: str):
        self.arm_collision_geoms = [s + "_" + id for s in self.arm_collision_geoms]
        self.joints = [s + "_" + id for s in self.joints]
        self.actuators = [s + "_" + id for s in self.actuators]
        self.attachment_site = self.attachment_site + "_" + id
        self.base = self.base + "_" + id


@dataclass
class SimRobotState:
    """/root/reference/src/sim/SimRobot.h:49-57"""
    previous_angles: object = None
    target_angles: object = None
    inverse_tcp_offset: common.Pose = field(default_factory=common.Pose)
    ik_success: object = True
    collision: object = False
    is_moving: object = False
    is_arrived: object = False


@dataclass
class SimGripperConfig:
    """/root/reference/src/sim/SimGripper.h:15-45"""
    epsilon_inner: float = 0.005
    epsilon_outer: float = 0.005
    seconds_between_callbacks: float = 0.05
    max_actuator_width: float = 255
    min_actuator_width: float = 0
    max_joint_width: float = 0.04
    min_joint_width: float = 0.0
    ignored_collision_geoms: list = field(default_factory=list)
    collision_geoms: list = field(default_factory=lambda: ["hand_c", "d435i_collision", "finger_0_left", "finger_0_right"])
    collision_geoms_fingers: list = field(default_factory=lambda: ["finger_0_left", "finger_0_right"])
    joint: str = "finger_joint1"
    actuator: str = "actuator8"

    def add_id(self, id: str):
        self.collision_geoms = [s + "_" + id for s in self.collision_geoms]
        self.collision_geoms_fingers = [s + "_" + id for s in self.collision_geoms_fingers]
        self.ignored_collision_geoms = [s + "_" + id for s in self.ignored_collision_geoms]
        self.joint = self.joint + "_" + id
        self.actuator = self.actuator + "_" + id


@dataclass
class SimGripperState:
    """/root/reference/src/sim/SimGripper.h:47-52"""
    last_commanded_width: object = 0
    is_moving: object = False
    last_width: object = 0
    collision: object = False


def load_compiled_scene(path) -> dict:
    """`.xml` -> compile with rcs_b200.mjcf; `.npz` -> precompiled scene (the `.mjb` analogue)."""
    path = Path(path)
    if path.suffix == ".xml":
        return mjcf.compile_mjcf(str(path))
    if path.suffix == ".npz":
        return mjcf.load_model(str(path))
    if path.suffix == ".mjb":
        alt = path.with_suffix(".npz")
        if alt.exists():
            return mjcf.load_model(str(alt))
    msg = f"Filetype {path.suffix} is unknown"
    raise ValueError(msg)


class _Opt:
    def __init__(self, M):
        self.timestep = M["opt_timestep"]


class _ModelShim:
    """The attributes Python callers of the reference reach on `sim.model` (envs/sim.py:53)."""

    def __init__(self, M):
        self.opt = _Opt(M)
        self.nq, self.nv, self.nu = M["nq"], M["nv"], M["nu"]
        self._M = M


class Sim:
    """`rcs.sim.Sim(mjmdl, cfg)`; additive: num_envs, device. The device objects are created lazily when
    the first SimRobot / SimGripper is attached (they contribute the callback configuration)."""

    def __init__(self, mjmdl, cfg: SimConfig | None = None, num_envs: int = 1, device: int = 0, maxcon: int | None = None):
        self._M = mjmdl if isinstance(mjmdl, dict) else load_compiled_scene(mjmdl)
        self.model = _ModelShim(self._M)
        self.num_envs = int(num_envs)
        self.device = device
        self._maxcon = maxcon
        self._cfg = SimConfig()
        self._robot_cfg = None
        self._gripper_cfg = None
        self._register_convergence = True
        self._dm = None
        self._b = None
        self._converged = True
        if cfg is not None:
            self.set_config(cfg)

    # ---- device objects
    def _attach(self, robot_cfg=None, gripper_cfg=None, register_convergence=True):
        if self._b is not None and (robot_cfg is not None or gripper_cfg is not None):
            # configuration changed after the first step: rebuild, keeping the dynamic state
            old = self._b.sr.clone(), self._b.sd.clone(), self._b.si.clone()
        else:
            old = None
        if robot_cfg is not None:
            self._robot_cfg, self._register_convergence = robot_cfg, register_convergence
        if gripper_cfg is not None:
            self._gripper_cfg = gripper_cfg
        self._dm, self._b = None, None
        self._build()
        if old is not None and old[0].shape == self._b.sr.shape:
            self._b.sr.copy_(old[0]); self._b.sd.copy_(old[1]); self._b.si.copy_(old[2])

    def _build(self):
        rc = self._robot_cfg
        ns = None
        if rc is not None:
            meta = common.robots_meta_config(rc.robot_type)
            ns = SimpleNamespace(joints=rc.joints, actuators=rc.actuators, arm_collision_geoms=rc.arm_collision_geoms,
                                 attachment_site=rc.attachment_site, base=rc.base, tcp_offset=rc.tcp_offset.as7(),
                                 q_home=meta.q_home, joint_rotational_tolerance=rc.joint_rotational_tolerance,
                                 seconds_between_callbacks=rc.seconds_between_callbacks,
                                 register_convergence_callback=self._register_convergence,
                                 ik_nq=min(self._M["nq"], 9))
        self._dm = _batch.DeviceModel(self._M, ns, self._gripper_cfg, self._maxcon, self.device)
        self._b = _batch.Batch(self._dm, self.num_envs)

    @property
    def batch(self) -> _batch.Batch:
        if self._b is None:
            self._build()
        return self._b

    # ---- reference API (rcs.cpp:493-506)
    def set_config(self, cfg: SimConfig) -> bool:
        self._cfg = copy.copy(cfg)
        return True

    def get_config(self) -> SimConfig:
        return copy.copy(self._cfg)

    def step(self, k: int):
        self.batch.run(_lib.STEP_K, k=int(k))

    def step_until_convergence(self):
        self.batch.run(_lib.STEP_CONV, max_convergence_steps=self._cfg.max_convergence_steps)
        conv = self.batch.si[:, 6]
        if self.num_envs == 1:
            self._converged = bool(conv[0].item())
            if int(self.batch.si[0, 7].item()) == self._cfg.max_convergence_steps:
                print("WARNING: Max convergence steps reached!", file=sys.stderr)  # sim.cpp:103-105
        else:
            self._converged = conv.bool()

    def is_converged(self):
        return self._converged

    def reset(self):
        self.batch.run(_lib.SIM_RESET)

    def _start_gui_server(self, id: str):
        raise NotImplementedError("GUI bridge is out of scope of the batched backend (SURVEY.md 8f-3)")

    def _stop_gui_server(self):
        pass

    def open_gui(self):
        raise NotImplementedError("GUI bridge is out of scope of the batched backend (SURVEY.md 8f-3)")

    # ---- state export / import: the layout of MuJoCo's mjSTATE_FULLPHYSICS vector that the
    #      reference's GUI bridge ships between processes (src/sim/gui.h:20: time | qpos | qvel | act; act is empty here)
    def get_state(self) -> torch.Tensor:
        b = self.batch
        return torch.cat([b.time[:, None], b.qpos, b.qvel], dim=1)

    def set_state(self, state: torch.Tensor):
        """Inverse of get_state() for every environment ([num_envs, 1 + nq + nv]); derived quantities follow on the
        next step, as after mj_setState."""
        b = self.batch
        nq, nv = self.model.nq, self.model.nv
        st = torch.as_tensor(state, dtype=torch.float64, device=b.dev).reshape(self.num_envs, 1 + nq + nv)
        b.sd[:, 0] = st[:, 0]
        b.qpos.copy_(st[:, 1:1 + nq])
        b.qvel.copy_(st[:, 1 + nq:])

    def save_checkpoint(self) -> dict:
        """Everything the kernels persist per environment (dynamic state, warm start, RCS device-layer state, callback
        clocks, flags): restoring it resumes bit-identically."""
        b = self.batch
        torch.cuda.synchronize(b.dev)
        return {"sr": b.sr.clone(), "sd": b.sd.clone(), "si": b.si.clone(), "nsr": b.model.nsr}

    def load_checkpoint(self, ckpt: dict):
        b = self.batch
        if ckpt["nsr"] != b.model.nsr or ckpt["sr"].shape != b.sr.shape:
            raise ValueError("checkpoint does not match this scene / number of environments")
        b.sr.copy_(ckpt["sr"]); b.sd.copy_(ckpt["sd"]); b.si.copy_(ckpt["si"])

    # ---- mjData-like views
    @property
    def data(self):
        b = self.batch
        return SimpleNamespace(qpos=b.qpos, qvel=b.qvel, ctrl=b.ctrl, time=b.time, ncon=b.si[:, 14])


def _scalarize(t: torch.Tensor, n: int):
    if n == 1:
        return t[0].detach().cpu().numpy()
    return t


class SimRobot(common.Robot):
    """`rcs.sim.SimRobot(sim, ik, cfg, register_convergence_callback=True)` (rcs.cpp:516-527)."""

    def __init__(self, sim: Sim, ik, cfg: SimRobotConfig, register_convergence_callback: bool = True):
        self.sim, self._ik, self._cfg = sim, ik, copy.deepcopy(cfg)
        sim._attach(robot_cfg=self._cfg, register_convergence=register_convergence_callback)  # raises "No ... named"
        if ik is not None and hasattr(ik, "_bind"):
            ik._bind(sim)
        self._n = sim.num_envs
        self._meta = common.robots_meta_config(cfg.robot_type)

    def get_config(self) -> SimRobotConfig:
        return copy.deepcopy(self._cfg)

    def set_config(self, cfg: SimRobotConfig) -> bool:
        self._cfg = copy.deepcopy(cfg)
        self.sim._attach(robot_cfg=self._cfg, register_convergence=self.sim._register_convergence)
        return True

    def get_state(self) -> SimRobotState:
        b, o = self.sim.batch, self.sim.batch.model.o_tail
        nj = b.model.njoints
        si = b.si
        f = (lambda t: bool(t[0].item())) if self._n == 1 else (lambda t: t.bool())
        return SimRobotState(previous_angles=_scalarize(b.sr[:, o:o + nj], self._n),
                             target_angles=_scalarize(b.sr[:, o + 8:o + 8 + nj], self._n),
                             inverse_tcp_offset=self._cfg.tcp_offset.inverse(), ik_success=f(si[:, 0]),
                             collision=f(si[:, 1]), is_moving=f(si[:, 2]), is_arrived=f(si[:, 3]))

    def _as_dev(self, q, width):
        b = self.sim.batch
        if isinstance(q, torch.Tensor):
            t = q.to(device=b.dev, dtype=torch.float64)
        else:
            t = torch.as_tensor(np.asarray(q, dtype=np.float64), device=b.dev)
        if t.dim() == 1:
            t = t.unsqueeze(0).expand(self._n, -1)
        return t[:, :width].contiguous()

    def get_cartesian_position_tensor(self) -> torch.Tensor:
        """[num_envs, 7] xyz + quat (xyzw) of every environment (a fresh tensor)."""
        b = self.sim.batch
        b.run(_lib.OBS, want_obs=True, fresh_obs=True)
        return b.obs[:, :7].clone()

    def get_cartesian_position(self):
        o = self.get_cartesian_position_tensor()
        if self._n == 1:
            o = o[0].cpu().numpy()
            return common.Pose(translation=o[:3], quaternion=o[3:7])
        return o

    def set_joint_position(self, q):
        b = self.sim.batch
        b.run(_lib.SET_JOINTS, act_joints=self._as_dev(q, b.model.njoints))

    def get_joint_position_tensor(self) -> torch.Tensor:
        """[num_envs, njoints] joint positions of every environment (a fresh device tensor)."""
        b = self.sim.batch
        idx = torch.as_tensor(np.asarray(b.model.fields["rb_qadr"][0]), device=b.dev, dtype=torch.long)
        return b.qpos.index_select(1, idx)

    def get_joint_position(self):
        return _scalarize(self.get_joint_position_tensor(), self._n)

    def move_home(self):
        self.set_joint_position(self._meta.q_home)

    def reset(self):
        self.sim.batch.run(_lib.ROBOT_RESET)

    def close(self):
        pass

    def set_joints_hard(self, q):
        b = self.sim.batch
        b.run(_lib.SET_JOINTS_HARD, act_joints=self._as_dev(q, b.model.njoints))

    def set_cartesian_position(self, pose):
        b = self.sim.batch
        if isinstance(pose, common.Pose):
            p = torch.as_tensor(pose.as7(), device=b.dev).unsqueeze(0).expand(self._n, -1).contiguous()
        else:
            p = self._as_dev(pose, 7)
        b.set_cartesian_position(p)

    def get_ik(self):
        return self._ik

    def get_base_pose_in_world_coordinates(self) -> common.Pose:
        f = self.sim.batch.model.fields
        q = f["rb_base_quat"][0]
        return common.Pose(translation=f["rb_base_pos"][0], quaternion=np.array([q[1], q[2], q[3], q[0]]))


class SimGripper(common.Gripper):
    """`rcs.sim.SimGripper(sim, cfg)` (rcs.cpp:508-515)."""

    def __init__(self, sim: Sim, cfg: SimGripperConfig):
        self.sim, self._cfg = sim, copy.deepcopy(cfg)
        sim._attach(gripper_cfg=self._cfg)
        self._n = sim.num_envs
        self.sim.batch.run(_lib.GRIPPER_RESET)

    def get_config(self):
        return copy.deepcopy(self._cfg)

    def set_config(self, cfg: SimGripperConfig) -> bool:
        self._cfg = copy.deepcopy(cfg)
        self.sim._attach(gripper_cfg=self._cfg)
        return True

    def get_state(self) -> SimGripperState:
        b, o = self.sim.batch, self.sim.batch.model.o_tail
        f = (lambda t: bool(t[0].item())) if self._n == 1 else (lambda t: t.bool())
        g = (lambda t: float(t[0].item())) if self._n == 1 else (lambda t: t.clone())
        return SimGripperState(last_commanded_width=g(b.sr[:, o + 16]), is_moving=f(b.si[:, 4]),
                               last_width=g(b.sr[:, o + 17]), collision=f(b.si[:, 5]))

    def set_normalized_width(self, width, force=0):
        b = self.sim.batch
        if isinstance(width, torch.Tensor):
            w = width.to(device=b.dev, dtype=torch.float64).reshape(-1)
            bad = bool(((w < 0) | (w > 1)).any().item())
        else:
            bad = width < 0 or width > 1
            w = torch.full((self._n,), float(width), dtype=torch.float64, device=b.dev)
        if bad or force < 0:
            raise ValueError("width must be between 0 and 1, force must be positive")  # SimGripper.cpp:80-83
        b.run(_lib.SET_GRIPPER, act_gripper=w.contiguous())

    def get_normalized_width(self):
        b, c = self.sim.batch, self._cfg
        qadr = int(b.model.fields["gr_qadr"][0][0])
        w = ((b.qpos[:, qadr] - c.min_joint_width) / (c.max_joint_width - c.min_joint_width)).clamp(0, 1)
        return float(w[0].item()) if self._n == 1 else w

    def is_grasped(self):
        w = self.get_normalized_width()
        lcw = self.get_state().last_commanded_width
        r = (lcw - self._cfg.epsilon_inner < w) & (w < lcw + self._cfg.epsilon_outer) if self._n > 1 else \
            (lcw - self._cfg.epsilon_inner < w < lcw + self._cfg.epsilon_outer)
        return r

    def grasp(self):
        self.shut()

    def open(self):
        self.set_normalized_width(1)

    def shut(self):
        self.set_normalized_width(0)

    def reset(self):
        self.sim.batch.run(_lib.GRIPPER_RESET)

    def close(self):
        pass


def _pose7_mul_const(p: torch.Tensor, c7) -> torch.Tensor:
    """[N, 7] poses (xyz + quat xyzw) times one constant pose, Pose.cpp:173-178 semantics (result quaternion normalised)."""
    c = torch.as_tensor(np.asarray(c7, dtype=np.float64), device=p.device)
    x, y, z, w = p[:, 3], p[:, 4], p[:, 5], p[:, 6]
    qv = p[:, 3:6]
    t = c[:3].expand_as(qv)
    uv = 2 * torch.linalg.cross(qv, t)
    rot = t + w[:, None] * uv + torch.linalg.cross(qv, uv)
    cx, cy, cz, cw = c[3], c[4], c[5], c[6]
    q = torch.stack([w * cx + x * cw + y * cz - z * cy, w * cy + y * cw + z * cx - x * cz,
                     w * cz + z * cw + x * cy - y * cx, w * cw - x * cx - y * cy - z * cz], dim=1)
    q = q / q.norm(dim=1, keepdim=True)
    return torch.cat([p[:, :3] + rot, q], dim=1)


class Pin(common.Kinematics):
    """`rcs.common.Pin(path, frame_id, urdf)` (/root/reference/src/rcs/Kinematics.cpp:13-81) on the batched
    DLS-CLIK kernel. The kinematic model is the robot of the scene the Sim was built from (the reference's
    default `kinematic_model_path` is that same robot.xml, envs/utils.py:22); the solver runs on the GPU of
    the Sim it is bound to. `rcs.common.RL` is an alias (README.md:41 of the reference)."""

    def __init__(self, path: str = "", frame_id: str = "fr3_link8", urdf: bool = True):
        self.path, self.frame_id, self.urdf = path, frame_id, urdf
        self._sim = None

    def _bind(self, sim: Sim):
        self._sim = sim

    def _need(self):
        if self._sim is None:
            raise RuntimeError("Pin is not bound to a Sim yet (construct SimRobot(sim, ik, cfg) first)")
        return self._sim.batch

    def inverse(self, pose, q0, tcp_offset: common.Pose = None):
        b = self._need()
        n = b.n
        # Kinematics::inverse(pose, q0, tcp_offset = Identity) drives the frame to pose * tcp_offset^-1
        # (Kinematics.cpp:28-40); the kernel applies the robot config's baked tcp_offset instead, so the goal is
        # re-expressed as pose * tcp_offset^-1 * cfg_tcp whenever the two differ
        cfg_tcp = self._sim._robot_cfg.tcp_offset
        call_tcp = tcp_offset if tcp_offset is not None else common.Pose()
        fix = None if call_tcp.is_close(cfg_tcp, 1e-12, 1e-12) else call_tcp.inverse() * cfg_tcp
        if isinstance(pose, common.Pose):
            if fix is not None:
                pose = pose * fix
            p = torch.as_tensor(pose.as7(), device=b.dev).unsqueeze(0).expand(n, -1).contiguous()
        else:
            p = pose.to(device=b.dev, dtype=torch.float64)
            if fix is not None:
                p = _pose7_mul_const(p, fix.as7())
            p = p.contiguous()
        nj = b.model.njoints
        q0t = torch.as_tensor(np.asarray(q0, dtype=np.float64), device=b.dev) if not isinstance(q0, torch.Tensor) else q0
        if q0t.dim() == 1:
            q0t = q0t.unsqueeze(0).expand(n, -1)
        q0t = q0t[:, :nj].to(torch.float64).contiguous()
        q, ok, _ = b.ik_inverse(p, q0t)
        if n == 1 and isinstance(pose, common.Pose):
            return q[0].cpu().numpy() if bool(ok[0].item()) else None
        return q, ok.bool()

    def forward(self, q0, tcp_offset: common.Pose = None) -> common.Pose:
        # forward kinematics through one kinematics-only device step would perturb the sim; use the host
        # restatement on the compiled scene instead (not a hot path: the reference calls it from planners only)
        from .hostkin import site_pose
        M = self._sim._M
        R, p = site_pose(M, M["site_names"].index(self._sim._robot_cfg.attachment_site), np.asarray(q0, dtype=np.float64))
        f = common.Pose(rotation=R, translation=p)
        return f * (tcp_offset if tcp_offset is not None else common.Pose()).inverse()  # sic, Kinematics.cpp:80


RL = Pin
