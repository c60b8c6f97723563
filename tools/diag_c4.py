"""Solver attempts / Newton iterations, contacts and rows per environment on the C4 tabletop workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from collections import Counter
h = bench.Harness()
N = 4096
env = bench.make_env(h, "c4", N)
local = env.unwrapped
b = local.sim.batch
gen = torch.Generator(device=h.dev).manual_seed(2)
env.reset()
for i in range(14):
    if i % 10 == 0:
        env.reset()
    a = {"joints": (torch.rand((N, local.dof), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * bench.MAX_MOV}
    if local.gripper is not None:
        a["gripper"] = torch.randint(0, 2, (N,), device=h.dev, generator=gen).to(torch.float64)
    env.step_packed(a)
    si = b.si.cpu().numpy()
    print(i, "solver_iter", sorted(Counter(si[:, 16]).items())[:10], "ncon", sorted(Counter(si[:, 14]).items())[:6], "nefc", sorted(Counter(si[:, 15]).items())[:8])
