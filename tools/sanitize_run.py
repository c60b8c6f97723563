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
# host-buffer step through mapped pinned memory, then a depth frame (two blocks per image)
h_a = torch.zeros((N, 8), dtype=torch.float64).pin_memory(); h_o = torch.zeros((N, dm.obs_dim), dtype=torch.float64).pin_memory()
b.step_host(_lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS, 3, 0, h_a, np.deg2rad(5), H.JLOW, H.JHIGH, h_o)
img = b.camera_depth(-1, np.array([1.5, 0.0, 1.2]), np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0.0]]), 45.0, 72, 64, 0.01, 50.0, True)
torch.cuda.synchronize()
print("host step + depth done", float(h_o[:, 7].abs().max()) > 0, int(img.cpu().numpy().min()), int(img.cpu().numpy().max()))
# several rounds per warp with two alignment groups (RCSB_WARPS / RCSB_BAR_GROUPS set by the caller)
if os.environ.get("SANITIZE_ROUNDS"):
    N2 = int(os.environ["SANITIZE_ROUNDS"])
    b2 = batch.Batch(dm, N2)
    b2.run(reset, k=1, want_obs=True)
    a2 = H.workload_actions(N2, 1, seed=1)
    b2.run(_lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS, k=3, act_joints=torch.as_tensor(a2[:, 0, :7].copy(), device=b2.dev),
           act_gripper=torch.as_tensor(a2[:, 0, 7].copy(), device=b2.dev), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
    torch.cuda.synchronize()
    print("multi-round done", b2.occupancy())
