#!/usr/bin/env python
"""Which stage makes a warp the slowest of its CTA? Per-warp, per-step stage cycles of CTA 0 (profiling build
librcsb_prof.so: make -C robot-control-stack_b200/csrc librcsb_prof.so). Usage: stage_trace.py c2|c3|c4 [envs]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RCSB_LIB_PATH"] = os.path.join(ROOT, "robot-control-stack_b200", "csrc", "librcsb_prof.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from rcs_b200 import _lib
what = sys.argv[1] if len(sys.argv) > 1 else "c3"
h = bench.Harness()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = bench.make_env(h, what, N)
local = env.unwrapped
b = local.sim.batch
gen = torch.Generator(device=h.dev).manual_seed(1)
def act():
    if what == "c3":
        scale = torch.tensor([0.01] * 3 + [0.05] * 3, dtype=torch.float64, device=h.dev)
        return {"xyzrpy": (torch.rand((N, 6), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * scale,
                "gripper": torch.randint(0, 2, (N,), device=h.dev, generator=gen).to(torch.float64)}
    a = {"joints": (torch.rand((N, local.dof), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * bench.MAX_MOV}
    if local.gripper is not None:
        a["gripper"] = torch.randint(0, 2, (N,), device=h.dev, generator=gen).to(torch.float64)
    return a
L = _lib.lib()
L.rcsb_debug_stage_trace.argtypes = [C.c_void_p, C.c_int]
env.reset()
for _ in range(4):
    env.step(act())
L.rcsb_debug_stage_trace(None, 0)  # clear
for _ in range(3):
    env.step(act())
S = 3 * 17
raw = np.zeros(S * 14 * 32, dtype=np.uint32)
L.rcsb_debug_stage_trace(raw.ctypes.data, S)
buf = raw[:S * 10 * 32].reshape(S, 10, 32)
aux = raw[S * 10 * 32:S * 11 * 32].reshape(S, 32)
W = b.occupancy()["warps_per_cta"]
colp = raw[S * 11 * 32:].reshape(S, 3, 32)[:, :, :W].astype(np.float64)
t = buf[:, :9, :W].astype(np.float64)          # [step, stage, warp]
tot = t.sum(axis=1)                            # [step, warp]
names = ["kinematics", "com", "crb", "collision", "velocity", "make_constraint", "actuation", "constraint_solve", "integrate"]
slow = tot.argmax(axis=1)
print(f"{what}: {N} envs, {W} warps/CTA, {S} steps of CTA 0 (only the first round of every launch if there are several)")
print(f"mean step {tot.mean():.0f} cycles, slowest warp of the step {tot.max(axis=1).mean():.0f} (+{100 * (tot.max(axis=1).mean() / tot.mean() - 1):.1f} %)")
print("stage               mean    slowest warp's   excess   std over warps")
for i, n in enumerate(names):
    m = t[:, i, :].mean()
    sl = np.mean([t[s, i, slow[s]] for s in range(S)])
    print(f"  {n:16s} {m:7.0f}   {sl:10.0f}   {sl - m:+8.0f}   {t[:, i, :].std(axis=1).mean():8.0f}")

# the collision stage of the warps that had something due
nar = buf[:, 9, :W].astype(np.float64); col = t[:, 3, :]
due = (aux[:, :W] & 0xff); broad = (aux[:, :W] >> 8) & 0xff; mid = (aux[:, :W] >> 16) & 0xff
ev = due > 0
print(f"collision events: {100 * ev.mean():.1f} % of the warp-steps have a due group; P(some warp of a {W}-warp CTA) = {100 * ev.any(axis=1).mean():.0f} % of the steps")
if ev.any():
    print(f"  per event: due groups {due[ev].mean():.1f}, broad-phase survivors {broad[ev].mean():.1f}, mid-phase survivors {mid[ev].mean():.2f}")
    print(f"  collision stage {col[ev].mean():.0f} cycles (no event: {col[~ev].mean():.0f}); of it the narrow phase {nar[ev].mean():.0f}")
    print(f"  inside an event (cycles from the stage's start): geom centres {colp[:, 0, :][ev].mean():.0f}, broad phase done {colp[:, 1, :][ev].mean():.0f}, mid phase done {colp[:, 2, :][ev].mean():.0f}")
    for k in range(0, int(mid.max()) + 1):
        sel = ev & (mid == k)
        if sel.any():
            print(f"    {k} narrow-phase pairs: {100 * sel.sum() / ev.sum():5.1f} % of the events, stage {col[sel].mean():7.0f} cycles, narrow {nar[sel].mean():7.0f}")
