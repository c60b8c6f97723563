#!/usr/bin/env python
"""Regenerate the precompiled scenes (robot-control-stack_b200/rcs_b200/models/*.npz) from the reference's MJCF
assets (the analogue of /root/reference/cmake/compile_scenes.cmake). Needs /root/reference; run in the build container."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robot-control-stack_b200"))
from rcs_b200 import mjcf  # noqa: E402

import shutil  # noqa: E402
import tempfile  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/scenes"
# synthetic scenes of this repository (tools/scenes/*.xml): compiled against the reference's robot models, laid out the way
# the reference lays out its own scene directories (scene.xml + robot.xml + assets/ links)
SYNTHETIC = {"xarm7_tabletop": ("xarm7_empty_world", os.path.join(ROOT, "tools", "scenes", "xarm7_tabletop.xml"))}


def scene_xml(name):
    if name not in SYNTHETIC:
        return os.path.join(SRC, name, "scene.xml")
    like, xml = SYNTHETIC[name]
    d = tempfile.mkdtemp(prefix="rcsb_scene_")
    for f in ("robot.xml", "assets"):
        os.symlink(os.path.realpath(os.path.join(SRC, like, f)), os.path.join(d, f))
    shutil.copy(xml, os.path.join(d, "scene.xml"))
    return os.path.join(d, "scene.xml")


for name in ("fr3_empty_world", "fr3_simple_pick_up", "xarm7_empty_world", "xarm7_tabletop"):
    M = mjcf.compile_mjcf(scene_xml(name))
    out = os.path.join(ROOT, "robot-control-stack_b200", "rcs_b200", "models", name + ".npz")
    mjcf.save_model(M, out)
    print(name, "->", out, f"nq={M['nq']} nv={M['nv']} ngeom={M['ngeom']} pairs={len(M['pair_geom'])}")
