#!/usr/bin/env python
"""Secondary mode: SimConfig(async_control=False): env.step() = set target + step_until_convergence.
Reports env-steps/s and physics-steps/s for random relative joint actions."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers as H
from rcs_b200 import _lib, batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dm = batch.DeviceModel(H.scene(), H.robot_ns(), H.gripper_ns())
b = batch.Batch(dm, N)
acts = H.workload_actions(N, T + 2, seed=0)
reset = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
step = _lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_CONV | _lib.OBS
b.run(reset, k=1, want_obs=True)
def one(t):
    b.run(step, max_convergence_steps=500, act_joints=torch.as_tensor(acts[:, t, :7].copy(), device=b.dev),
          act_gripper=torch.as_tensor(acts[:, t, 7].copy(), device=b.dev), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
for t in range(2): one(t)
torch.cuda.synchronize()
s0 = int(b.si[:, 18].sum())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(2, T + 2): one(t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
steps = int(b.si[:, 18].sum()) - s0
print(json.dumps({"mode": "sync (step_until_convergence)", "envs": N, "env_steps_per_s": N * T / (ms * 1e-3), "physics_steps_per_s": steps / (ms * 1e-3),
                  "substeps_per_env_step_mean": steps / (N * T), "conv_steps_min_max": [int(b.si[:, 7].min()), int(b.si[:, 7].max())],
                  "converged_frac": float(b.si[:, 6].double().mean())}))
