#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
e = d.get("e2e") or {}
print(f"headline n_gpus={d['n_gpus']} value {d['value']:.0f} env-steps/s  ms/step {d['ms_per_step']:.3f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  "
      f"e2e {e.get('value', 0):.0f}  warps {d['config']['warps_per_cta']}  frac {d['roofline']['frac']:.4f}  clocks {d.get('clocks')}")
for r in d.get("sweep", []):
    if "error" in r:
        print("  sweep", r["workload"], r["envs_per_gpu"], "ERROR", r["error"][:200])
    elif r["workload"] == "depth":
        print(f"  sweep depth    {r['envs_per_gpu']:6d}/gpu  {r['value']:.0f} frames/s  {r['ms_per_step']:.3f} ms  {r['roofline']['achieved']:.1f} GB/s  frac {r['roofline']['frac']:.4f}")
    else:
        print(f"  sweep {r['workload']:8s} {r['envs_per_gpu']:6d}/gpu  {r['value']:.0f} env-steps/s  {r['ms_per_step']:.3f} ms  phys {r['physics_steps_per_s']:.3e}  "
              f"e2e {(r.get('e2e') or {}).get('value', 0):.0f}  warps {r.get('warps_per_cta')} {r.get('variant')}")
if "cpu_baseline" in d:
    print("  cpu", {k: d["cpu_baseline"].get(k) for k in ("value", "cores")})
