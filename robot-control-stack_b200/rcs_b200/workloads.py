"""Benchmark / test workloads of the hot path: the compiled scenes, the joint limits the action spaces clip
to, and the seeded synthetic action streams. Shared by bench.py, tools/ and tests/ (nothing here touches the oracle)."""
from __future__ import annotations

import os

import numpy as np

from . import common, mjcf, scenes

MODELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")
# robots_meta_config(FR3).joint_limits (include/rcs/Robot.h:35-40): tighter than the MJCF ranges
FR3_JLOW = np.array([-2.3093, -1.5133, -2.4937, -2.7478, -2.4800, 0.8521, -2.6895])
FR3_JHIGH = np.array([2.3093, 1.5133, 2.4937, -0.4461, 2.4800, 4.2094, 2.6895])
FR3_Q_HOME = np.array([0, -np.pi / 4, 0, -3 * np.pi / 4, 0, np.pi / 2, np.pi / 4])
_cache: dict = {}


def has_scene(name: str) -> bool:
    return os.path.exists(os.path.join(MODELS, name + ".npz"))


def scene(name: str = "fr3_empty_world") -> dict:
    """The compiled scene (rcs_b200.mjcf model dict) shipped as models/<name>.npz."""
    if name not in _cache:
        _cache[name] = mjcf.load_model(os.path.join(MODELS, name + ".npz"))
    return _cache[name]


def workload_actions(nenv: int, nsteps: int, seed: int = 0, dof: int = 7, gripper: bool = True) -> np.ndarray:
    """BASELINE.md 3 / SURVEY.md 8d C2: joints ~ U(-5 deg, 5 deg)^dof, gripper ~ Bernoulli(0.5), env-major [nenv, nsteps, dof + 1]."""
    rng = np.random.default_rng(seed)
    a = np.zeros((nenv, nsteps, dof + 1))
    a[:, :, :dof] = rng.uniform(-np.deg2rad(5), np.deg2rad(5), (nenv, nsteps, dof))
    if gripper:
        a[:, :, dof] = rng.integers(0, 2, (nenv, nsteps))
    return a


def xarm7_robot_cfg(scene_name: str = "xarm7_empty_world"):
    """The reference's xArm7 wiring (examples/xarm7/xarm7_env_joint_control.py:44-66)."""
    from . import sim
    cfg = sim.SimRobotConfig()
    cfg.actuators = [f"act{i}" for i in range(1, 8)]
    cfg.joints = [f"joint{i}" for i in range(1, 8)]
    cfg.base = "base"
    cfg.robot_type = common.RobotType.XArm7
    cfg.attachment_site = "attachment_site"
    cfg.arm_collision_geoms = []
    cfg.mjcf_scene_path = scenes[scene_name].mjb
    cfg.kinematic_model_path = scenes[scene_name].mjcf_robot
    return cfg


def xarm7_tabletop_robot_cfg():
    """Config C4 (synthetic): xArm7 on a table with duplo-sized bricks, joint control."""
    return xarm7_robot_cfg("xarm7_tabletop")
