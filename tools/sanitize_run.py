#!/usr/bin/env python
"""Small workload for compute-sanitizer (racecheck / memcheck): resets, relative joint steps, floor collision with
contacts + noslip, step_until_convergence, IK."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers as H
from rcs_b200 import _lib, batch
N = 24
dm = batch.DeviceModel(H.scene(), H.robot_ns(), H.gripper_ns())
b = batch.Batch(dm, N)
acts = H.workload_actions(N, 3, seed=0)
reset = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
b.run(reset, k=1, want_obs=True)
for t in range(2):
    b.run(_lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS, k=5, act_joints=torch.as_tensor(acts[:, t, :7].copy(), device=b.dev),
          act_gripper=torch.as_tensor(acts[:, t, 7].copy(), device=b.dev), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
tgt = np.tile(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]), (N, 1))
b.run(_lib.SET_JOINTS | _lib.STEP_K, k=260, act_joints=torch.as_tensor(tgt, device=b.dev))
print("ncon max", int(b.si[:, 14].max()))
b.run(_lib.STEP_CONV, max_convergence_steps=60)
pose = b.obs[:, :7].clone(); pose[:, 0] += 0.05
b.run(_lib.OBS, want_obs=True)
q, ok, it = b.ik_inverse(b.obs[:, :7].clone().contiguous(), b.qpos[:, :7].clone().contiguous())
torch.cuda.synchronize()
print("sanitize workload done", int(ok.sum()))
