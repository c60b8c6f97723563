"""Self-checks of the CPU oracle: physics invariants and the behavioural envelopes the reference's
own tests pin (/root/reference/python/tests/test_sim_envs.py, /root/reference/src/sim/test.cpp)."""
import numpy as np
import pytest

import helpers as H
from helpers import O


@pytest.fixture(scope="module")
def fr3():
    M = H.scene()
    return M, O.Model(M)


def test_mass_matrix_spd_symmetric_and_matches_jacobian_form(fr3):
    M, m = fr3
    d = O.Data(m)
    d.forward()
    qM = d.qM.reshape(9, 9)
    assert np.abs(qM - qM.T).max() == 0
    assert np.linalg.eigvalsh(qM).min() > 0.09  # armature 0.1 floor
    assert np.abs(qM - M["qM0"]).max() < 1e-13   # independent sum_b J^T I J computed by the scene compiler
    rng = np.random.default_rng(0)
    for _ in range(10):
        d.qpos[:7] = H.Q_HOME + rng.uniform(-1, 1, 7)
        d.qvel[:] = rng.uniform(-1, 1, 9)
        d.forward()
        qM = d.qM.reshape(9, 9)
        assert np.abs(qM @ d.qacc_smooth - d.qfrc_smooth).max() < 1e-10


def test_gravity_compensation_static_hold_at_home(fr3):
    M, m = fr3
    d = O.Data(m)
    d.qpos[:7] = H.Q_HOME
    d.ctrl[:7] = H.Q_HOME
    d.step(500)
    assert np.abs(d.qpos[:7] - H.Q_HOME).max() < 1e-9
    assert np.abs(d.qvel).max() < 1e-9
    assert np.allclose(d.qfrc_gravcomp[:7], d.qfrc_bias[:7], atol=1e-9)


def test_constraint_kkt_residual_with_contacts(fr3):
    """At the Newton solution, M*qacc - qfrc_smooth - J^T f = 0 (floor collision pose of test_sim_envs.py:347-360)."""
    M, m = fr3
    mm, s = H.oracle_sim(M)
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    s.set_joint_position(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]))
    seen = False
    for _ in range(60):
        s.step(5)
        d = s.data
        if d.ncon[0] > 0:
            seen = True
            d.forward()
            nefc = int(d.nefc[0])
            J = d.efc_J[:nefc * 9].reshape(nefc, 9)
            res = d.qM.reshape(9, 9) @ d.qacc - d.qfrc_smooth - J.T @ d.efc_force[:nefc]
            # noslip modifies forces after the solve: compare against qfrc_constraint instead
            res2 = d.qM.reshape(9, 9) @ d.qacc - d.qfrc_smooth - d.qfrc_constraint
            assert np.abs(res2).max() < 1e-6 * max(1.0, np.abs(d.qfrc_smooth).max())
    assert seen


def test_joint_target_reached_and_converged(fr3):
    """test_sim_envs.py:318-345 (atol 0.01 rad) and src/sim/test.cpp:143-151 (!is_moving && is_arrived)."""
    M, m = fr3
    mm, s = H.oracle_sim(M)
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    tgt = H.Q_HOME + np.array([0.1, 0.1, 0.1, 0.1, -0.1, -0.1, 0.1])  # test_sim_envs.py:337
    s.set_joint_position(tgt)
    s.step_until_convergence()  # the reference does not assert is_sim_converged either (test_sim_envs.py:343)
    st = s.robot_state()
    assert st["ik_success"] and not st["collision"]
    assert np.allclose(s.get_joint_position(), tgt, atol=0.01)
    assert 50 <= s.convergence_steps() <= 500
    # a small move converges properly: !is_moving && is_arrived (src/sim/test.cpp:143-151)
    tgt2 = tgt + 0.01
    s.set_joint_position(tgt2)
    s.step_until_convergence()
    st = s.robot_state()
    assert s.is_converged() and st["is_arrived"] and not st["is_moving"]
    assert np.abs(s.get_joint_position() - tgt2).max() < 0.05 * np.pi / 180


def test_callback_cadence_is_50_and_25_substeps(fr3):
    """sim.cpp:14-23 with float64 time accumulation: 0.1 s clocks fire at substeps 50, 100, ...; the gripper's
    0.05 s clocks at 25, 50, ...."""
    t, fired50, fired25, last50, last25 = 0.0, [], [], 0.0, 0.0
    for k in range(1, 201):
        t += 0.002
        if t - last50 > 0.1:
            fired50.append(k); last50 = t
        if t - last25 > 0.05:
            fired25.append(k); last25 = t
    assert fired50 == [50, 100, 150, 200]
    assert fired25[:4] == [25, 50, 75, 100]


def test_floor_collision_sets_flags(fr3):
    """test_sim_envs.py:347-360: q = [0, 1.78, 0, -1.45, 0, 0, 0] drives the arm into the floor: collision flag
    (robot or gripper, as GripperWrapperSim merges them) with ik_success."""
    M, m = fr3
    mm, s = H.oracle_sim(M)
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    s.set_joint_position(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]))
    s.step_until_convergence()
    assert s.robot_state()["collision"] or s.gripper_state()["collision"]
    assert s.robot_state()["ik_success"]
    assert s.data.ncon[0] > 0


def test_gripper_quirks(fr3):
    """Appendix B 11-13: convergence is true at the first 0.05 s sample; env.reset() leaves the fingers closed."""
    M, m = fr3
    mm, s = H.oracle_sim(M)
    s.gripper_reset()
    assert s.data.qpos[7] == 0.04 and s.data.ctrl[7] == 255
    s.reset()  # mj_resetData wipes it
    assert s.data.qpos[7] == 0.0 and s.data.ctrl[7] == 0.0
    with pytest.raises(ValueError):
        s.gripper_set_normalized_width(1.5)
    s.robot_reset(); s.step(1)
    s.set_joint_position(H.Q_HOME)  # the robot's all-callback needs a target before it can report convergence
    s.gripper_set_normalized_width(1.0)
    s.step_until_convergence()
    assert s.is_converged() and s.convergence_steps() <= 100 and s.gripper_get_normalized_width() < 0.99


def test_ik_roundtrip_and_failure(fr3):
    M, m = fr3
    site = O.robot_cfg(M).attachment_site
    rng = np.random.default_rng(2)
    for _ in range(20):
        qt = H.Q_HOME + rng.uniform(-0.4, 0.4, 7)
        pose = O.ik_forward(m, site, 9, qt)
        q, it = O.ik_inverse(m, site, 9, pose, H.Q_HOME)
        assert q is not None and it < 1000
        back = O.ik_forward(m, site, 9, q[:7])
        assert np.abs(back[:3] - pose[:3]).max() < 1e-4
        assert O.pose_is_close(back, pose, 1e-3, 1e-3)
    far = np.array([2.0, 0, 0.5, 0, 0, 0, 1.0])  # outside the workspace
    q, it = O.ik_inverse(m, site, 9, far, H.Q_HOME)
    assert q is None and it == 1000


def test_free_body_rests_on_floor():
    """fr3_simple_pick_up: the cube (free joint, density 50) settles on the plane with <= 4 box-plane contacts."""
    M = H.scene("fr3_simple_pick_up")
    m = O.Model(M)
    d = O.Data(m)
    d.qpos[:7] = H.Q_HOME; d.ctrl[:7] = H.Q_HOME
    d.step(400)
    z = d.qpos[9 + 2]
    assert abs(z - 0.0288) < 2e-3 and np.abs(d.qvel[9:]).max() < 1e-3
    assert 1 <= d.ncon[0] <= 4
