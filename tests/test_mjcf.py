"""The MJCF scene compiler against the reference's own scene files (skipped where /root/reference is absent, e.g.
on the GPU box) and the committed precompiled scenes (always)."""
import os

import numpy as np
import pytest

import helpers as H
from rcs_b200 import devmodel, mjcf

REF = "/root/reference/assets/scenes"


def test_committed_scene_structure():
    M = H.scene()
    assert (M["nq"], M["nv"], M["nu"], M["nbody"], M["njnt"], M["ngeom"]) == (9, 9, 8, 14, 9, 71)
    assert M["opt_integrator"] == "implicitfast" and M["opt_cone"] == "elliptic" and M["opt_impratio"] == 20
    assert M["opt_noslip_iterations"] == 5 and M["opt_timestep"] == 0.002
    col = [g for g in range(M["ngeom"]) if M["geom_contype"][g] or M["geom_conaffinity"][g]]
    assert len(col) == 24
    # position actuators inherit the joint range; the gripper actuator is affine on the tendon
    assert np.allclose(M["actuator_ctrlrange"][0], M["jnt_range"][0]) and M["actuator_ctrllimited"][0] == 1
    assert np.allclose(M["actuator_biasprm"][7], [0, -100, -10]) and np.allclose(M["actuator_ctrlrange"][7], [0, 255])
    assert np.allclose(M["tendon_coef"][0][7:9], [0.5, 0.5])
    assert M["geom_vertnum"].max() <= 256
    P = H.scene("fr3_simple_pick_up")
    assert (P["nq"], P["nv"], P["nbody"]) == (16, 15, 15) and np.allclose(P["qpos0"][9:], [0.44, 0.1, 0.03, 0, 0, 0, 1])


def test_device_model_fusion_preserves_mass_and_tree():
    M = H.scene()
    F, verts = devmodel.build_device_fields(M, H.robot_ns(), H.gripper_ns())
    assert int(F["nb"][0][0]) == 9 and list(F["b_parent"][0]) == [-1, 0, 1, 2, 3, 4, 5, 6, 6]
    assert np.isclose(F["b_mass"][0].sum(), M["body_mass"].sum())
    # link7 carries hand + camera: 0.627143 + 0.73 + 0.072
    assert np.isclose(F["b_mass"][0][6], 0.627143 + 0.73 + 0.072)
    with pytest.raises(RuntimeError, match="No joint named"):
        bad = H.robot_ns(); bad.joints = ["nope"] * 7
        devmodel.build_device_fields(M, bad, H.gripper_ns())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference assets not present on this machine")
def test_recompile_matches_committed_scene():
    M = mjcf.compile_mjcf(os.path.join(REF, "fr3_empty_world", "scene.xml"))
    C = H.scene()
    for k in ("body_mass", "body_pos", "body_quat", "body_inertia", "jnt_range", "dof_invweight0", "body_invweight0",
              "geom_pos", "geom_quat", "geom_size", "mesh_vert", "pair_geom", "actuator_gainprm", "qM0"):
        assert np.allclose(np.asarray(M[k], dtype=float), np.asarray(C[k], dtype=float), atol=1e-12), k
