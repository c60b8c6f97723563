"""Shared fixtures for the parity tests: compiled scene, oracle objects, config objects."""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from rcs_b200 import mjcf  # noqa: E402
from oracle import oracle as O  # noqa: E402

from rcs_b200 import workloads as WL  # noqa: E402

MODELS = WL.MODELS
Q_HOME, JLOW, JHIGH = WL.FR3_Q_HOME, WL.FR3_JLOW, WL.FR3_JHIGH
scene = WL.scene


def robot_ns(tcp=(0, 0, 0, 0, 0, 0, 1.0)):
    return NS(joints=[f"fr3_joint{i}_0" for i in range(1, 8)], actuators=[f"fr3_joint{i}_0" for i in range(1, 8)],
              arm_collision_geoms=[f"fr3_link{i}_collision_0" for i in range(8)], attachment_site="attachment_site_0",
              base="base_0", tcp_offset=list(tcp), q_home=Q_HOME, joint_rotational_tolerance=0.05 * np.pi / 180,
              seconds_between_callbacks=0.1, register_convergence_callback=True, ik_nq=9)


def gripper_ns():
    return NS(actuator="actuator8_0", joint="finger_joint1_0",
              collision_geoms=["hand_c_0", "d435i_collision_0", "finger_0_left_0", "finger_0_right_0"],
              collision_geoms_fingers=["finger_0_left_0", "finger_0_right_0"], ignored_collision_geoms=[],
              epsilon_inner=0.005, epsilon_outer=0.005, seconds_between_callbacks=0.05, max_actuator_width=255.0,
              min_actuator_width=0.0, max_joint_width=0.04, min_joint_width=0.0)


def oracle_sim(M, tcp=None):
    m = O.Model(M)
    return m, O.Sim(m, O.robot_cfg(M, tcp_offset=tcp), O.gripper_cfg(M))


workload_actions = WL.workload_actions


XARM_Q_HOME = np.array([0, -45.0, 0, 15.0, 0, -25.0, 0]) * np.pi / 180
XARM_JLOW = np.array([-2 * np.pi, -2.094395, -2 * np.pi, -3.92699, -2 * np.pi, -np.pi, -2 * np.pi])
XARM_JHIGH = np.array([2 * np.pi, 2.059488, 2 * np.pi, 0.191986, 2 * np.pi, 1.692969, 2 * np.pi])


def xarm_robot_ns():
    """examples/xarm7/xarm7_env_joint_control.py:44-66"""
    return NS(joints=[f"joint{i}" for i in range(1, 8)], actuators=[f"act{i}" for i in range(1, 8)], arm_collision_geoms=[],
              attachment_site="attachment_site", base="base", tcp_offset=[0, 0, 0, 0, 0, 0, 1.0], q_home=XARM_Q_HOME,
              joint_rotational_tolerance=0.05 * np.pi / 180, seconds_between_callbacks=0.1, register_convergence_callback=True,
              ik_nq=7)


FRANKA_HAND_TCP = [0, 0, 0.1034, 0, 0, -0.3826834323650898, 0.9238795325112867]  # Pose.cpp:11-15, normalised


def grasp_and_lift_script(M):
    """Grasp-and-lift on fr3_simple_pick_up as a list of (joint target [7], normalised gripper width, physics steps):
    open, move above the cube, descend in stages, close on the cube, raise 12 cm. Joint targets come from the oracle's
    Pin::inverse so that every implementation under test receives identical set_joint_position commands."""
    m = O.Model(M)
    site = O.robot_cfg(M).attachment_site
    q = Q_HOME.copy()
    script = [(q.copy(), 1.0, 200)]
    for z, n in ((0.20, 500), (0.15, 250), (0.10, 250), (0.06, 250), (0.03, 250)):
        sol, _ = O.ik_inverse(m, site, 9, [0.44, 0.1, z, 1, 0, 0, 0], q, FRANKA_HAND_TCP)
        assert sol is not None
        q = sol[:7].copy()
        script.append((q.copy(), 1.0, n))
    script.append((q.copy(), 0.0, 300))
    sol, _ = O.ik_inverse(m, site, 9, [0.44, 0.1, 0.15, 1, 0, 0, 0], q, FRANKA_HAND_TCP)
    assert sol is not None
    script.append((sol[:7].copy(), 0.0, 500))
    return script
