"""rcs_b200._core -- the compiled pybind11 interface shaped like the reference's rcs._core (src/pybind/rcs.cpp:186-527).
CPU tier: the names the reference's Python layer imports exist, the compiled Pose / RPY agree with the host Pose class
(itself pinned by the reference's own goldens in tests/test_pose.py) and with the oracle, the C++ interfaces
(Kinematics / Robot / Gripper, Robot.h:127-197, Kinematics.h:19-26) can be subclassed from Python. GPU tier: Sim over raw
model / batch addresses, SimRobot / SimGripper / Pin against the Python mirror on the same batch."""
import pickle

import numpy as np
import pytest

import helpers as H
from helpers import O
from rcs_b200 import _core, common


def test_exports_the_names_the_reference_imports():
    for name in ("Pose", "RPY", "Kinematics", "Pin", "Robot", "Gripper", "RobotConfig", "RobotState", "GripperConfig", "GripperState",
                 "RobotType", "RobotPlatform", "RobotMetaConfig", "robots_meta_config", "FrankaHandTCPOffset", "IdentityTranslation",
                 "IdentityRotMatrix", "IdentityRotQuatVec"):
        assert hasattr(_core.common, name), name
    for name in ("Sim", "SimConfig", "SimRobot", "SimRobotConfig", "SimRobotState", "SimGripper", "SimGripperConfig", "SimGripperState"):
        assert hasattr(_core.sim, name), name
    cfg = _core.sim.SimConfig()
    assert (cfg.async_control, cfg.realtime, cfg.frequency, cfg.max_convergence_steps) == (False, False, 30, 500)  # sim.h:29-34
    rc = _core.sim.SimRobotConfig(); rc.add_id("0")
    assert rc.joints[0] == "fr3_joint1_0" and rc.attachment_site == "attachment_site_0" and rc.base == "base_0"
    assert isinstance(rc, _core.common.RobotConfig) and rc.robot_type == _core.common.RobotType.FR3
    meta = _core.common.robots_meta_config(_core.common.RobotType.FR3)
    assert meta.dof == 7 and np.allclose(meta.q_home, H.Q_HOME) and np.allclose(meta.joint_limits, [H.JLOW, H.JHIGH])
    ur = _core.common.robots_meta_config(_core.common.RobotType.UR5e)
    assert ur.dof == 6 and np.isclose(ur.joint_limits[1, 2], np.pi)                                             # Robot.h:43-59


def test_compiled_pose_matches_host_pose_and_oracle():
    P, C = _core.common.Pose, common.Pose
    rng = np.random.default_rng(0)
    for _ in range(200):
        t1, q1, t2, r2 = rng.normal(size=3), rng.normal(size=4), rng.normal(size=3), rng.uniform(-3, 3, 3)
        a, b = P(quaternion=q1, translation=t1), P(rpy_vector=r2, translation=t2)
        ha, hb = C(translation=t1, quaternion=q1), C(translation=t2, rpy_vector=r2)
        as7 = lambda p: np.concatenate([p.translation(), p.rotation_q()])  # noqa: E731
        assert np.allclose(as7(a * b), (ha * hb).as7(), atol=1e-14) and np.allclose(as7(a * b), O.pose_mul(ha.as7(), hb.as7()), atol=1e-14)
        assert np.allclose(as7(a.inverse()), ha.inverse().as7(), atol=1e-14)
        assert np.allclose(a.xyzrpy(), ha.xyzrpy(), atol=1e-12) and np.isclose(a.total_angle(), ha.total_angle(), atol=1e-13)
        assert np.allclose(as7(a.limit_rotation_angle(0.3)), ha.limit_rotation_angle(0.3).as7(), atol=1e-13)
        assert np.allclose(as7(a.limit_translation_length(0.2)), ha.limit_translation_length(0.2).as7(), atol=1e-14)
        assert np.allclose(as7(a.interpolate(b, 0.37)), ha.interpolate(hb, 0.37).as7(), atol=1e-13)
        assert np.allclose(a.pose_matrix(), ha.pose_matrix(), atol=1e-14) and a.is_close(P(pose_matrix=a.pose_matrix()), 1e-12, 1e-12)
        assert pickle.loads(pickle.dumps(a)).is_close(a, 1e-15, 1e-15)
    # constructor overloads of rcs.cpp:224-237
    rpy = _core.common.RPY(0.1, -0.4, 1.2)
    assert P(rpy=rpy, translation=[1, 2, 3.0]).is_close(P(rpy_vector=[0.1, -0.4, 1.2], translation=[1, 2, 3.0]))
    assert np.allclose(P(translation=[1, 2, 3.0]).rotation_q(), [0, 0, 0, 1]) and np.allclose(P().translation(), 0)
    tcp = P(pose_matrix=_core.common.FrankaHandTCPOffset())  # polar rotation of the rounded matrix: exactly Rz(-45 deg)
    assert np.allclose(np.concatenate([tcp.translation(), tcp.rotation_q()]), H.FRANKA_HAND_TCP, atol=1e-15)
    r2 = rpy + _core.common.RPY(0.1, 0.1, 0.1)
    assert np.allclose(r2.as_vector(), [0.2, -0.3, 1.3]) and pickle.loads(pickle.dumps(rpy)).is_close(rpy)


def test_interfaces_can_be_subclassed_from_python():
    class MyIK(_core.common.Kinematics):
        def inverse(self, pose, q0, tcp_offset=_core.common.Pose()):
            return np.asarray(q0) + 1.0

        def forward(self, q0, tcp_offset):
            return _core.common.Pose(translation=[float(np.sum(q0)), 0, 0])

    class MyRobot(_core.common.Robot):
        def __init__(self):
            super().__init__()
            self.q = np.zeros(3)

        def set_joint_position(self, q):
            self.q = np.asarray(q, dtype=float)

        def get_joint_position(self):
            return self.q

        def get_base_pose_in_world_coordinates(self):
            return _core.common.Pose(translation=[1.0, 0, 0])

        def get_ik(self):
            return MyIK()

    r = MyRobot()
    r.set_joint_position([1.0, 2, 3])
    assert np.allclose(r.get_joint_position(), [1, 2, 3])
    p = r.to_pose_in_robot_coordinates(_core.common.Pose(translation=[3.0, 0, 0]))   # C++ Robot.cpp:5-8 calling back into Python
    assert np.allclose(p.translation(), [2, 0, 0])
    ik = r.get_ik()
    assert np.allclose(ik.inverse(_core.common.Pose(), np.zeros(3)), 1.0) and np.isclose(ik.forward(np.ones(3), _core.common.Pose()).translation()[0], 3)
    with pytest.raises(RuntimeError):
        _core.common.Robot().move_home()   # pure virtual


@pytest.mark.gpu
def test_core_sim_over_raw_handles_matches_the_python_mirror():
    import torch
    import rcs_b200
    from rcs_b200 import sim as psim
    s = psim.Sim(rcs_b200.scenes["fr3_empty_world"].mjb, psim.SimConfig(), num_envs=3)
    cfg = psim.SimRobotConfig(); cfg.add_id("0")
    probot = psim.SimRobot(s, psim.Pin(), cfg)
    gcfg = psim.SimGripperConfig(); gcfg.add_id("0")
    pgrip = psim.SimGripper(s, gcfg)
    b = s.batch
    # the reference: Sim(mjmdl: int, mjdata: int) -- raw addresses of objects Python owns (rcs.cpp:493-506)
    cs = _core.sim.Sim(b.model.ptr, b.ptr)
    ccfg = _core.sim.SimRobotConfig(); ccfg.add_id("0")
    cik = _core.common.Pin("", "attachment_site_0", False)
    crobot = _core.sim.SimRobot(cs, cik, ccfg)
    cg = _core.sim.SimGripperConfig(); cg.add_id("0")
    cgrip = _core.sim.SimGripper(cs, cg)
    cs.reset(); crobot.reset(); cs.step(1)
    assert np.allclose(crobot.get_joint_position(), H.Q_HOME, atol=1e-3)
    assert np.allclose(crobot.get_joint_position(), probot.get_joint_position()[0].cpu().numpy())
    cp, pp = crobot.get_cartesian_position(), probot.get_cartesian_position()[0].cpu().numpy()
    assert np.allclose(np.concatenate([cp.translation(), cp.rotation_q()]), pp, atol=1e-12)
    target = H.Q_HOME + 0.05
    crobot.set_joint_position(target)
    cs.step_until_convergence()
    assert cs.is_converged() and np.abs(crobot.get_joint_position() - target).max() < 0.01
    st = crobot.get_state()
    assert st.ik_success and st.is_arrived and not st.is_moving and not st.collision and np.allclose(st.target_angles, target)
    # Pin::inverse through the compiled interface against the oracle
    M = H.scene(); m = O.Model(M); site = O.robot_cfg(M).attachment_site
    goal7 = O.ik_forward(m, site, 9, H.Q_HOME + 0.1)
    q = cik.inverse(_core.common.Pose(quaternion=goal7[3:], translation=goal7[:3]), H.Q_HOME)
    qr, _ = O.ik_inverse(m, site, 9, goal7, H.Q_HOME)
    assert q is not None and np.abs(np.asarray(q) - qr).max() < 1e-9
    assert cik.inverse(_core.common.Pose(translation=[3.0, 0, 0]), H.Q_HOME) is None        # unreachable -> nullopt -> None
    crobot.set_cartesian_position(cp)
    assert crobot.get_state().ik_success
    cgrip.open(); cs.step(300)
    assert cgrip.get_normalized_width() > 0.9 and abs(cgrip.get_normalized_width() - float(pgrip.get_normalized_width()[0])) < 1e-12
    with pytest.raises(ValueError):
        cgrip.set_normalized_width(1.5)                                                      # SimGripper.cpp:80-83 -> ValueError
    assert isinstance(crobot, _core.common.Robot) and isinstance(cgrip, _core.common.Gripper)
    torch.cuda.synchronize()
