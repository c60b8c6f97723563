#!/usr/bin/env python
"""Per-stage clock64() breakdown of one physics step (profiling build librcsb_prof.so, warp 0 of CTA 0)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("RCSB_LOCKSTEP", "1")  # a barrier before every stage: the per-stage clocks are comparable
os.environ["RCSB_LIB_PATH"] = os.path.join(ROOT, "robot-control-stack_b200", "csrc", "librcsb_prof.so")
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers as H
from rcs_b200 import _lib, batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dm = batch.DeviceModel(H.scene(), H.robot_ns(), H.gripper_ns())
b = batch.Batch(dm, N)
acts = H.workload_actions(N, 12, seed=0)
reset = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
step = _lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS
b.run(reset, k=1, want_obs=True)
out = (C.c_ulonglong * 16)()
L = _lib.lib(); L.rcsb_debug_stage_cycles.argtypes = [C.POINTER(C.c_ulonglong)]
for t in range(2):
    b.run(step, k=17, act_joints=torch.as_tensor(acts[:, t, :7].copy(), device=b.dev), act_gripper=torch.as_tensor(acts[:, t, 7].copy(), device=b.dev), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
L.rcsb_debug_stage_cycles(out)
nsteps = 0
for t in range(2, 10):
    b.run(step, k=17, act_joints=torch.as_tensor(acts[:, t, :7].copy(), device=b.dev), act_gripper=torch.as_tensor(acts[:, t, 7].copy(), device=b.dev), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
    nsteps += 17 * ((N + 148 * b.occupancy()["warps_per_cta"] - 1) // (148 * b.occupancy()["warps_per_cta"]))
L.rcsb_debug_stage_cycles(out)
names = ["kinematics", "com", "crb+chol", "collision", "velocity", "make_constraint", "actuation+solveM", "constraint_solve", "integrate"]
tot = sum(out[i] for i in range(9))
print(f"envs {N} warps/cta {b.occupancy()['warps_per_cta']} cycles/step (warp 0, CTA 0): {tot / nsteps:.0f}")
for i, n in enumerate(names):
    print(f"  {n:18s} {out[i] / nsteps:9.0f} cycles  {100 * out[i] / tot:5.1f}%")
