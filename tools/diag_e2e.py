"""Fixed cost of SimVectorEnv.step_host: full step vs a launch that only packs the observation (k = 0), mapped vs staged."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from rcs_b200 import _lib
h = bench.Harness()
N = 4096
env = bench.make_env(h, "c2", N)
local = env.unwrapped
b = local.sim.batch
env.reset()
a = torch.zeros((N, 8), dtype=torch.float64).pin_memory(); a[:, 7] = 1  # zero relative move, gripper open
o = torch.zeros((N, 30), dtype=torch.float64).pin_memory()
ops, cfg = local._step_ops()
def run(label, k, ops_):
    for _ in range(5):
        b.step_host(ops_, k, cfg.max_convergence_steps, a, float(local.max_mov), local.jlow, local.jhigh, o)
    t0 = time.perf_counter()
    for _ in range(200):
        b.step_host(ops_, k, cfg.max_convergence_steps, a, float(local.max_mov), local.jlow, local.jhigh, o)
    dt = (time.perf_counter() - t0) / 200
    print(f"{label:40s} {dt * 1e6:8.1f} us per call")
run("full env.step (k=17)", 17, ops)
run("k=1", 1, ops)
run("obs only (no physics step)", 0, _lib.OBS)
t0 = time.perf_counter()
for _ in range(200):
    local.step_host(a)
print(f"{'SimVectorEnv.step_host (python layer)':40s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us per call")
t0 = time.perf_counter()
for _ in range(200):
    torch.cuda.synchronize()
print(f"{'torch.cuda.synchronize alone':40s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us per call")
