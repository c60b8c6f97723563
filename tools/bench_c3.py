#!/usr/bin/env python
"""Config C3 (BASELINE.json configs[2]): N x FR3 + Franka hand on fr3_simple_pick_up (free cube resting on the floor:
contacts, elliptic cones, noslip every step), relative CARTESIAN_TRPY control (0.2 m, 45 deg) with the on-GPU
damped-least-squares IK, async 30 Hz (17 substeps), RandomCubePos resets every 10 steps, PickCubeSuccess reward: the
reference's FR3SimplePickUpSimEnvCreator. Prints one JSON line; the headline bench is bench.py (config C2)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from rcs_b200.envs.creators import FR3SimplePickUpSimEnvCreator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = FR3SimplePickUpSimEnvCreator()(num_envs=N)
obs, _ = env.reset()
b = env.sim.batch
gen = torch.Generator(device=b.dev).manual_seed(1)
scale = torch.tensor([0.01] * 3 + [0.05] * 3, dtype=torch.float64, device=b.dev)
def act():
    return {"xyzrpy": (torch.rand((N, 6), dtype=torch.float64, device=b.dev, generator=gen) * 2 - 1) * scale,
            "gripper": torch.randint(0, 2, (N,), device=b.dev, generator=gen).to(torch.float64)}
for _ in range(3):
    env.step(act())
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(steps):
    if i % 10 == 0:
        env.reset()
    obs, reward, term, trunc, info = env.step(act())
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
print(json.dumps({"config": "C3 fr3_simple_pick_up, relative CARTESIAN_TRPY + IK, async 17 substeps, reset every 10 steps", "envs": N,
                  "ms_per_env_step": ms, "env_steps_per_s": N / (ms * 1e-3), "occupancy": b.occupancy(),
                  "ik_success_frac": float(info["ik_success"].double().mean()), "mean_reward": float(reward.mean()),
                  "ncon_mean": float(b.si[:, 14].double().mean()), "warn_max": int(b.si[:, 17].max())}))
