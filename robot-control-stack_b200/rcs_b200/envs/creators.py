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
                 relative_to: RelativeTo = RelativeTo.LAST_STEP, sim_wrapper=None, num_envs: int = 1, device: int | None = None,
                 shard: bool = False):
        """num_envs / device / shard are the additive batched extension. With shard=True under an
        initialised torch.distributed process group (one process per GPU), num_envs is the TOTAL number of environments:
        this rank simulates its contiguous block on GPU `device` (default: LOCAL_RANK) and step()/reset() return the
        all-gathered [num_envs, ...] observation on every rank (rcs_b200.envs.sharded.ShardedVectorEnv)."""
        n_total = num_envs
        if shard:
            import os
            import torch.distributed as dist
            from rcs_b200.shard import shard_range
            if not dist.is_initialized():
                raise RuntimeError("shard=True needs an initialised torch.distributed process group (one process per GPU)")
            b0, e0 = shard_range(n_total, dist.get_rank(), dist.get_world_size())
            num_envs = e0 - b0
            if device is None:
                device = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = 0
        if hand_cfg is not None:
            raise NotImplementedError("SimTilburgHand is out of scope (SURVEY.md 2 row 15)")
        simulation = sim.Sim(robot_cfg.mjcf_scene_path, sim_cfg, num_envs=num_envs, device=device)
        ik = sim.Pin(robot_cfg.kinematic_model_path, robot_cfg.attachment_site,
                     urdf=robot_cfg.kinematic_model_path.endswith(".urdf"))
        robot = sim.SimRobot(simulation, ik, robot_cfg)
        gripper = sim.SimGripper(simulation, gripper_cfg) if gripper_cfg is not None else None
        env = SimVectorEnv(simulation, robot, gripper, control_mode, max_relative_movement, relative_to)
        if sim_wrapper is not None:  # creators.py:101-103: the task layer wraps the sim env
            env = sim_wrapper(env, simulation)
        if cameras is not None:  # creators.py:105-110: SimCameraSet + CameraSetWrapper (depth frames of every environment)
            from rcs_b200.camera.sim import CameraSetWrapper, SimCameraSet
            camera_set = SimCameraSet(simulation, cameras, physical_units=True, render_on_demand=True)
            env = CameraSetWrapper(env, camera_set, include_depth=True)
        if shard:
            from rcs_b200.envs.sharded import ShardedVectorEnv
            env = ShardedVectorEnv(env, n_total)
        return env


class SimTaskEnvCreator:
    """creators.py:131-189: relative Cartesian control + a re-placement wrapper + the pick-up success / reward wrapper."""

    def __call__(self, robot_cfg: sim.SimRobotConfig, render_mode: str = "none", control_mode: ControlMode = ControlMode.CARTESIAN_TRPY,
                 delta_actions: bool = True, cameras=None, hand_cfg=None, gripper_cfg: sim.SimGripperConfig | None = None,
                 sim_cfg: sim.SimConfig | None = None, random_pos_args: dict | None = None, num_envs: int = 1, device: int = 0):
        import numpy as np
        from functools import partial
        from rcs_b200.envs.task import PickCubeSuccessWrapper, RandomCubePos, RandomObjectPos
        from rcs_b200.envs.utils import default_sim_gripper_cfg
        if hand_cfg is not None:
            raise NotImplementedError("SimTilburgHand is out of scope (SURVEY.md 2 row 15)")
        if render_mode == "human":
            raise NotImplementedError("the GUI bridge is out of scope (SURVEY.md 8f-3)")
        random_env = RandomCubePos
        if random_pos_args is not None and all(k in random_pos_args for k in ("joint_name", "init_object_pose")):
            random_env = partial(RandomObjectPos, **random_pos_args)
        env = SimEnvCreator()(control_mode=control_mode, robot_cfg=robot_cfg, collision_guard=False,
                              gripper_cfg=gripper_cfg if gripper_cfg is not None else default_sim_gripper_cfg(), sim_cfg=sim_cfg,
                              cameras=cameras, max_relative_movement=(0.2, float(np.deg2rad(45))) if delta_actions else None,
                              relative_to=RelativeTo.LAST_STEP, sim_wrapper=random_env, num_envs=num_envs, device=device)
        return PickCubeSuccessWrapper(env)


class FR3SimplePickUpSimEnvCreator:
    """creators.py:192-224 (gym id rcs/FR3SimplePickUpSim-v0): fr3_simple_pick_up, async 30 Hz, relative TRPY control."""

    def __call__(self, render_mode: str = "none", control_mode: ControlMode = ControlMode.CARTESIAN_TRPY, resolution=None,
                 frame_rate: int = 0, delta_actions: bool = True, cam_list=None, num_envs: int = 1, device: int = 0):
        import numpy as np
        from rcs_b200 import common
        from rcs_b200.envs.utils import default_sim_robot_cfg
        if cam_list:
            raise NotImplementedError("SimCameraSet is a 'next' row (SURVEY.md 8f-2)")
        robot_cfg = default_sim_robot_cfg(scene="fr3_simple_pick_up")
        robot_cfg.tcp_offset = common.Pose(translation=np.array([0.0, 0.0, 0.1034]),
                                           rotation=np.array([[0.707, 0.707, 0], [-0.707, 0.707, 0], [0, 0, 1]]))
        sim_cfg = sim.SimConfig()
        sim_cfg.realtime = False
        sim_cfg.async_control = True
        sim_cfg.frequency = 30
        return SimTaskEnvCreator()(robot_cfg, render_mode, control_mode, delta_actions, None, sim_cfg=sim_cfg, num_envs=num_envs,
                                   device=device)
