#!/usr/bin/env python
"""Config C3 probe (BASELINE.json configs[2]): N x FR3 + Franka hand on fr3_simple_pick_up (free cube resting on the
floor: contacts, elliptic cones, noslip every step), ControlMode.CARTESIAN_TQuat with the on-GPU damped-least-squares IK,
async 30 Hz (17 substeps). Prints env-steps/s; not the headline bench (bench.py)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from rcs_b200 import sim
from rcs_b200.envs.base import ControlMode
from rcs_b200.envs.creators import SimEnvCreator
from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
scene = sys.argv[3] if len(sys.argv) > 3 else "fr3_simple_pick_up"
env = SimEnvCreator()(ControlMode.CARTESIAN_TQuat, default_sim_robot_cfg(scene), gripper_cfg=default_sim_gripper_cfg(),
                      sim_cfg=sim.SimConfig(async_control=True, frequency=30), num_envs=N)
obs, _ = env.reset()
gen = torch.Generator(device=obs["tquat"].device).manual_seed(1)
def act(obs):
    t = obs["tquat"].clone()
    t[:, :3] += (torch.rand((N, 3), dtype=torch.float64, device=t.device, generator=gen) * 2 - 1) * 0.01
    g = torch.randint(0, 2, (N,), device=t.device, generator=gen).to(torch.float64)
    return {"tquat": t, "gripper": g}
for _ in range(3):
    obs, _, _, trunc, info = env.step(act(obs))
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    obs, _, _, trunc, info = env.step(act(obs))
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
# the IK launch alone (SimRobot::set_cartesian_position for every env)
a = act(obs); pose = env._to_pose7(a["tquat"])
torch.cuda.synchronize(); ev0.record()
for _ in range(5):
    env.robot.set_cartesian_position(pose)
ev1.record(); torch.cuda.synchronize()
ik_ms = ev0.elapsed_time(ev1) / 5
b = env.sim.batch
print(json.dumps({"scene": scene, "envs": N, "ms_per_env_step": ms, "ik_ms": ik_ms, "env_steps_per_s": N / (ms * 1e-3), "occupancy": b.occupancy(),
                  "ik_success_frac": float(info["ik_success"].double().mean()), "collision_frac": float(info["collision"].double().mean()),
                  "ncon_mean": float(b.si[:, 14].double().mean()), "warn_max": int(b.si[:, 17].max())}))
