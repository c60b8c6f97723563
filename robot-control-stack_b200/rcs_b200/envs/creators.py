"""`SimEnvCreator` with the reference's call signature (/root/reference/python/rcs/envs/creators.py:43-128)
plus the additive `num_envs` / `device`. It wires the same objects in the same order (Sim -> Pin -> SimRobot ->
[SimGripper]) and returns a vector env whose step/reset run as single fused launches."""
from __future__ import annotations

from rcs_b200 import sim
from rcs_b200.envs.base import ControlMode, RelativeTo
from rcs_b200.envs.vector import SimVectorEnv


class SimEnvCreator:
    def __call__(self, control_mode: ControlMode, robot_cfg: sim.SimRobotConfig, collision_guard: bool = False,
                 gripper_cfg: sim.SimGripperConfig | None = None, sim_cfg: sim.SimConfig | None = None, hand_cfg=None,
                 cameras=None, max_relative_movement: float | tuple[float, float] | None = None,
                 relative_to: RelativeTo = RelativeTo.LAST_STEP, sim_wrapper=None, num_envs: int = 1, device: int = 0):
        if hand_cfg is not None:
            raise NotImplementedError("SimTilburgHand is out of scope (SURVEY.md 2 row 15)")
        if cameras is not None:
            raise NotImplementedError("SimCameraSet is a 'next' row (SURVEY.md 8f-2)")
        if sim_wrapper is not None:
            raise NotImplementedError("SimWrapper task layers are a 'next' row (SURVEY.md 8f-1)")
        simulation = sim.Sim(robot_cfg.mjcf_scene_path, sim_cfg, num_envs=num_envs, device=device)
        ik = sim.Pin(robot_cfg.kinematic_model_path, robot_cfg.attachment_site,
                     urdf=robot_cfg.kinematic_model_path.endswith(".urdf"))
        robot = sim.SimRobot(simulation, ik, robot_cfg)
        gripper = sim.SimGripper(simulation, gripper_cfg) if gripper_cfg is not None else None
        return SimVectorEnv(simulation, robot, gripper, control_mode, max_relative_movement, relative_to)
