"""rcs_b200 -- B200-native batched backend behind Robot Control Stack's sim.Sim / SimRobot / SimGripper /
SimEnvCreator surface. Mirrors `rcs/__init__.py` of the reference (/root/reference/python/rcs/__init__.py:17-61):
a `scenes` registry; scene paths point at the precompiled scenes shipped in `rcs_b200/models/*.npz`
(the analogue of the reference's build-time `.mjb`, cmake/compile_scenes.cmake:19)."""
from __future__ import annotations

import os
from dataclasses import dataclass

from . import common  # noqa: F401

__version__ = "0.1.0"
_MODELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


@dataclass(kw_only=True)
class Scene:
    mjb: str
    mjcf_scene: str
    mjcf_robot: str
    urdf: str | None = None
    robot_type: common.RobotType


scenes: dict[str, Scene] = {
    name: Scene(mjb=os.path.join(_MODELS, name + ".npz"), mjcf_scene=os.path.join(_MODELS, name + ".npz"),
                mjcf_robot=os.path.join(_MODELS, name + ".npz"), urdf=None, robot_type=common.RobotType.FR3)
    for name in ("fr3_empty_world", "fr3_simple_pick_up")
}
for _name in ("xarm7_empty_world", "xarm7_tabletop"):  # xarm7_tabletop: synthetic config C4 (tools/scenes/xarm7_tabletop.xml)
    scenes[_name] = Scene(mjb=os.path.join(_MODELS, _name + ".npz"), mjcf_scene=os.path.join(_MODELS, _name + ".npz"),
                          mjcf_robot=os.path.join(_MODELS, _name + ".npz"), urdf=None, robot_type=common.RobotType.XArm7)
