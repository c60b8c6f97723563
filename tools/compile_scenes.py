#!/usr/bin/env python
"""Regenerate the precompiled scenes (robot-control-stack_b200/rcs_b200/models/*.npz) from the reference's MJCF
assets (the analogue of /root/reference/cmake/compile_scenes.cmake). Needs /root/reference; run in the build container."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robot-control-stack_b200"))
from rcs_b200 import mjcf  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/scenes"
for name in ("fr3_empty_world", "fr3_simple_pick_up", "xarm7_empty_world"):
    M = mjcf.compile_mjcf(os.path.join(SRC, name, "scene.xml"))
    out = os.path.join(ROOT, "robot-control-stack_b200", "rcs_b200", "models", name + ".npz")
    mjcf.save_model(M, out)
    print(name, "->", out, f"nq={M['nq']} nv={M['nv']} ngeom={M['ngeom']} pairs={len(M['pair_geom'])}")
