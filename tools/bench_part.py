#!/usr/bin/env python
"""One record of bench.py's sweep on its own (for ncu captures): tools/bench_part.py depth | c4 | c3 | fleet [envs]."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
what = sys.argv[1] if len(sys.argv) > 1 else "depth"
h = bench.Harness()
if what == "depth":
    rec = bench.measure_depth(h, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, K=4, W=2)
elif what == "fleet":
    rec = bench.measure_fleet(h, int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
else:
    rec = bench.measure(h, what, int(sys.argv[2]) if len(sys.argv) > 2 else 8192, 4, 3, want_e2e=False)
print(json.dumps(rec))
