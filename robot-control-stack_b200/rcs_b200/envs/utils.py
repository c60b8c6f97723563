"""Default configs, mirroring /root/reference/python/rcs/envs/utils.py:17-38."""
from __future__ import annotations

import rcs_b200
from rcs_b200 import sim


def default_sim_robot_cfg(scene: str = "fr3_empty_world", idx: str = "0") -> sim.SimRobotConfig:
    robot_cfg = sim.SimRobotConfig()
    robot_cfg.robot_type = rcs_b200.scenes[scene].robot_type
    robot_cfg.add_id(idx)
    robot_cfg.mjcf_scene_path = rcs_b200.scenes[scene].mjb
    robot_cfg.kinematic_model_path = rcs_b200.scenes[scene].mjcf_robot
    return robot_cfg


def default_sim_gripper_cfg(idx: str = "0") -> sim.SimGripperConfig:
    cfg = sim.SimGripperConfig()
    cfg.add_id(idx)
    return cfg
