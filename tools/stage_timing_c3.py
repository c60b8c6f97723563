#!/usr/bin/env python
"""Per-stage clock64() breakdown for the pick-up scene (profiling build)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RCSB_LIB_PATH"] = os.path.join(ROOT, "robot-control-stack_b200", "csrc", "librcsb_prof.so")
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from rcs_b200 import _lib
from rcs_b200.envs.creators import FR3SimplePickUpSimEnvCreator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
env = FR3SimplePickUpSimEnvCreator()(num_envs=N)
env.reset()
b = env.sim.batch
gen = torch.Generator(device=b.dev).manual_seed(1)
scale = torch.tensor([0.01] * 3 + [0.05] * 3, dtype=torch.float64, device=b.dev)
def act():
    return {"xyzrpy": (torch.rand((N, 6), dtype=torch.float64, device=b.dev, generator=gen) * 2 - 1) * scale,
            "gripper": torch.randint(0, 2, (N,), device=b.dev, generator=gen).to(torch.float64)}
out = (C.c_ulonglong * 16)()
L = _lib.lib(); L.rcsb_debug_stage_cycles.argtypes = [C.POINTER(C.c_ulonglong)]
for _ in range(2): env.step(act())
L.rcsb_debug_stage_cycles(out)
T = 6
for _ in range(T): env.step(act())
L.rcsb_debug_stage_cycles(out)
names = ["kinematics", "com", "crb", "collision", "velocity", "make_constraint", "actuation", "constraint_solve", "integrate"]
tot = sum(out[i] for i in range(9)); nsteps = 17 * T
print(f"envs {N} warps/cta {b.occupancy()['warps_per_cta']} cycles/step (warp 0, CTA 0): {tot / nsteps:.0f}; solver iters mean {float(b.si[:,16].double().mean()):.2f} nefc mean {float(b.si[:,15].double().mean()):.1f}")
for i, n in enumerate(names):
    print(f"  {n:18s} {out[i] / nsteps:9.0f} cycles  {100 * out[i] / tot:5.1f}%")
