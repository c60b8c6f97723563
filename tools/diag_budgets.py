"""Which collision groups run out of separation budget, and how often? Reads the persistent budgets (o_cbud) from the
state rows after every env.step of the C2 workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from rcs_b200 import _lib
h = bench.Harness()
N = 4096
env = bench.make_env(h, "c2", N)
local = env.unwrapped
b = local.sim.batch
F = b.model.fields
pair = np.array(F["pair"][0]).reshape(-1, 2); gbody = np.array(F["g_body"][0]).reshape(-1)
nq, nv, nu = int(F["nq"][0][0]), int(F["nv"][0][0]), int(F["nu"][0][0])
groups = []
for p in pair:
    a, c = sorted((int(gbody[p[0]]), int(gbody[p[1]])))
    if (a, c) not in groups:
        groups.append((a, c))
ngrp = len(groups)
nsr = b.sr.shape[1]
off = nsr - 16 - nq - (ngrp * 4 + 7) // 8   # ... | cbud | cbq | sepcache(16)
gen = torch.Generator(device=h.dev).manual_seed(1)
env.reset()
due = np.zeros(ngrp); val = np.zeros(ngrp); T = 0
for i in range(30):
    if i % 10 == 0:
        env.reset()
    a = {"joints": (torch.rand((N, 7), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * bench.MAX_MOV,
         "gripper": torch.randint(0, 2, (N,), device=h.dev, generator=gen).to(torch.float64)}
    # single physics steps so that every substep's budgets are seen
    ops, cfg = local._step_ops()
    b.run(ops, k=1, act_joints=a["joints"].contiguous(), act_gripper=a["gripper"].contiguous(), max_mov=float(local.max_mov), jlow=local.jlow, jhigh=local.jhigh, want_obs=True, fresh_obs=True)
    for s in range(16):
        b.run(_lib.STEP_K, k=1)
        bud = b.sr[:, off:off + (ngrp * 4 + 7) // 8].contiguous().view(torch.float32)[:, :ngrp].cpu().numpy()
        due += (bud <= 0).mean(axis=0); val += np.median(bud, axis=0); T += 1
print(f"{ngrp} groups; a group is due in {100 * due.sum() / T / 1:.1f} % group-steps summed = {due.sum() / T:.2f} due groups per env per step")
order = np.argsort(-due)
for g in order[:12]:
    np_g = sum(1 for p in pair if tuple(sorted((int(gbody[p[0]]), int(gbody[p[1]])))) == groups[g])
    print(f"  group {g:2d} bodies {groups[g]} ({np_g} geom pairs): due in {100 * due[g] / T:5.1f} % of the steps, median budget {1000 * val[g] / T:7.2f} mm")
