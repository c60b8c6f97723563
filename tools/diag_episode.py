"""Per-step kernel time inside a C2 episode (reset, then 10 steps): where the gap between the mean step time and the
steady-state kernel time comes from."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robot-control-stack_b200"))
sys.path.insert(0, ROOT)
import bench

h = bench.Harness()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = bench.make_env(h, "c2", n)
local = env.unwrapped
gen = torch.Generator(device=h.dev).manual_seed(5)
T = 44
aj = (torch.rand((T, n, 7), dtype=torch.float64, device=h.dev, generator=gen) * 2 - 1) * bench.MAX_MOV
ag = torch.randint(0, 2, (T, n), device=h.dev, generator=gen).to(torch.float64)
rows = []
for i in range(T):
    if i % 11 == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.flush.zero_(); e0.record(h.stream); env.reset(); e1.record(h.stream); torch.cuda.synchronize()
        rows.append(("reset", e0.elapsed_time(e1)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.flush.zero_(); e0.record(h.stream); local.reset_packed(); e1.record(h.stream); torch.cuda.synchronize()
        rows.append(("reset_packed", e0.elapsed_time(e1)))
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.flush.zero_(); e0.record(h.stream); local.step_packed({"joints": aj[i], "gripper": ag[i]}); e1.record(h.stream); torch.cuda.synchronize()
    rows.append((f"step{i % 11}", e0.elapsed_time(e1)))
for name, ms in rows[24:]:
    print(f"{name:14s} {ms:.3f} ms")
