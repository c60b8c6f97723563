"""Solver attempts / contact and row counts substep by substep right after a C2 reset (why the first env.step of an
episode is slower than the steady state)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200")): sys.path.insert(0, p)
import torch, numpy as np
from collections import Counter
import bench
from rcs_b200 import _lib
h = bench.Harness()
N = 4096
env = bench.make_env(h, "c2", N)
local = env.unwrapped
b = local.sim.batch
gen = torch.Generator(device=h.dev).manual_seed(5)
aj = (torch.rand((30, N, 7), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * bench.MAX_MOV
ag = torch.randint(0, 2, (30, N), device=h.dev, generator=gen).to(torch.float64)
for i in range(12):
    local.step_packed({"joints": aj[i], "gripper": ag[i]})
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(h.stream); fn(); e1.record(h.stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def substeps(tag, i):
    ops, cfg = local._step_ops()
    rows = []
    for s in range(17):
        if s == 0:
            ms = timed(lambda: b.run(ops, k=1, act_joints=aj[i].contiguous(), act_gripper=ag[i].contiguous(), max_mov=float(local.max_mov), jlow=local.jlow, jhigh=local.jhigh, want_obs=True, fresh_obs=True))
        else:
            ms = timed(lambda: b.run(_lib.STEP_K, k=1))
        si = b.si.cpu().numpy()
        rows.append((ms, sorted(Counter(si[:, 16]).items())[:5], sorted(Counter(si[:, 14]).items())[:4], sorted(Counter(si[:, 15]).items())[:5]))
    for s, r in enumerate(rows):
        print(tag, s, f"{r[0]*1000:.0f}us", "iter", r[1], "ncon", r[2], "nefc", r[3])
substeps("steady", 12)
local.reset_packed()
substeps("after-reset", 13)
substeps("second", 14)
