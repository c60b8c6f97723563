#!/usr/bin/env python
"""Time the batched Pin::inverse launch (rcsb_ik_inverse): targets 1 cm / 0.05 rad away from the current pose."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers as H
from rcs_b200 import _lib, batch
for N in [int(a) for a in sys.argv[1:]] or [4096, 65536]:
    dm = batch.DeviceModel(H.scene(), H.robot_ns(), H.gripper_ns())
    b = batch.Batch(dm, N)
    b.run(_lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS, k=1, want_obs=True)
    pose = b.obs[:, :7].clone()
    gen = torch.Generator(device=b.dev).manual_seed(0)
    pose[:, :3] += (torch.rand((N, 3), dtype=torch.float64, device=b.dev, generator=gen) * 2 - 1) * 0.01
    q0 = b.qpos[:, :7].clone().contiguous()
    q, ok, it = b.ik_inverse(pose.contiguous(), q0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        q, ok, it = b.ik_inverse(pose.contiguous(), q0)
    e1.record(); torch.cuda.synchronize()
    print(f"N {N}: ik_inverse {e0.elapsed_time(e1) / 5:.3f} ms, success {float(ok.double().mean()):.3f}, iterations mean {float(it.double().mean()):.1f} max {int(it.max())}")
