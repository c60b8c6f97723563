"""Pose arithmetic pinned by the reference's own golden identities (/root/reference/python/tests/test_common.py),
checked on (a) the CPU oracle's Pose functions and (b) the host `rcs_b200.common.Pose` class."""
import math

import numpy as np
import pytest

import helpers as H
from helpers import O
from rcs_b200 import common

HOME_M = np.array([[9.99999352e-01, 2.51302265e-05, -1.13823380e-03, 3.06764031e-01],
                   [2.65946429e-05, -9.99999172e-01, 1.28657214e-03, 1.39119827e-04],
                   [-1.13820053e-03, -1.28660158e-03, -9.99998525e-01, 4.86190811e-01],
                   [0, 0, 0, 1.0]])  # test_common.py:198-205


def test_identity_quaternion():  # test_common.py:19-26
    assert np.array_equal(common.Pose().rotation_q(), np.array([0, 0, 0, 1]))


def test_interpolate_full_progress():  # test_common.py:28-67
    a = common.Pose(rotation=np.eye(3), translation=np.zeros((3, 1)))
    b = common.Pose(rotation=np.eye(3), translation=np.ones((3, 1)))
    r = a.interpolate(b, progress=1.0)
    assert np.array_equal(r.rotation_m(), np.eye(3)) and np.array_equal(r.translation(), np.ones(3))
    ro = O.pose_interpolate(a.as7(), b.as7(), 1.0)
    assert np.array_equal(ro, r.as7())


@pytest.mark.parametrize("dx,rot,expected", [(0.1, 0, False), (0.09, 0, True), (0, 1e-8, True), (0, 1.0, False)])
def test_is_close(dx, rot, expected):  # test_common.py:69-109
    m1 = np.eye(4); m1[:3, 3] = [1, 2, 3] if dx else 0; m1[0, 1] = rot
    m2 = np.eye(4); m2[:3, 3] = [1 + dx, 2, 3] if dx else 0
    p1, p2 = common.Pose(pose_matrix=m1), common.Pose(pose_matrix=m2)
    assert p1.is_close(p2, eps_t=0.1) == expected
    assert O.pose_is_close(p1.as7(), p2.as7(), 1e-8, 0.1) == expected


def test_multiply_inverse_matrix():  # test_common.py:111-178
    m1 = np.eye(4); m1[:3, 3] = [1, 2, 3]
    m2 = np.eye(4); m2[:3, 3] = [4, 5, 6]
    e = np.eye(4); e[:3, 3] = [5, 7, 9]
    assert np.array_equal((common.Pose(pose_matrix=m1) * common.Pose(pose_matrix=m2)).pose_matrix(), e)
    assert np.array_equal(O.pose_mul(common.Pose(pose_matrix=m1).as7(), common.Pose(pose_matrix=m2).as7())[:3], [5, 7, 9])
    inv = np.eye(4); inv[:3, 3] = [-1, -2, -3]
    assert np.array_equal(common.Pose(pose_matrix=m1).inverse().pose_matrix(), inv)
    assert np.array_equal(O.pose_inverse(common.Pose(pose_matrix=m1).as7())[:3], [-1, -2, -3])
    p = common.Pose(quaternion=np.array([0, 0, 0, 1.0]), translation=np.array([1.0, 1.0, 1.0]))
    e = np.eye(4); e[:3, 3] = 1
    assert np.array_equal(p.pose_matrix(), e)


def test_rpy_identity_and_home_roundtrip():  # test_common.py:180-218
    rpy = common.Pose(pose_matrix=np.eye(4)).rotation_rpy()
    assert all(math.isclose(v, 0, abs_tol=1e-8) for v in (rpy.roll, rpy.pitch, rpy.yaw))
    assert common.Pose(translation=np.zeros(3), rpy_vector=np.zeros(3)).is_close(common.Pose())
    home = common.Pose(pose_matrix=HOME_M)
    assert np.allclose(home.pose_matrix(), HOME_M)
    trpy = home.xyzrpy()
    assert np.allclose(trpy[:3], home.translation())
    home2 = common.Pose(translation=trpy[:3], rpy_vector=trpy[3:])
    assert home.is_close(home2) and np.allclose(HOME_M, home2.pose_matrix())
    # oracle agrees with the host class on the same golden matrix
    o7 = O.pose_from_matrix(HOME_M[:3, :3], HOME_M[:3, 3])
    assert np.allclose(o7, home.as7(), atol=1e-15)
    assert np.allclose(O.pose_xyzrpy(o7), trpy, atol=1e-12)
    assert np.allclose(O.pose_from_rpy(trpy[:3], trpy[3:]), home2.as7(), atol=1e-12)


def test_host_pose_matches_oracle_on_random_poses():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = common.Pose(translation=rng.normal(size=3), quaternion=rng.normal(size=4))
        b = common.Pose(translation=rng.normal(size=3), rpy_vector=rng.uniform(-3, 3, 3))
        assert np.allclose((a * b).as7(), O.pose_mul(a.as7(), b.as7()), atol=1e-14)
        assert np.allclose(a.inverse().as7(), O.pose_inverse(a.as7()), atol=1e-14)
        assert np.allclose(a.xyzrpy(), O.pose_xyzrpy(a.as7()), atol=1e-12)
        assert math.isclose(a.total_angle(), O.pose_total_angle(a.as7()), abs_tol=1e-13)
        assert np.allclose(a.limit_rotation_angle(0.3).as7(), O.pose_limit_rotation_angle(a.as7(), 0.3), atol=1e-13)
        assert np.allclose(a.limit_translation_length(0.2).as7(), O.pose_limit_translation_length(a.as7(), 0.2), atol=1e-14)
        assert np.allclose(a.interpolate(b, 0.37).as7(), O.pose_interpolate(a.as7(), b.as7(), 0.37), atol=1e-13)
        assert np.allclose(a.rotation_m(), O.pose_rotation_m(a.as7()), atol=1e-15)


def test_fk_home_matches_reference_home_matrix():
    """FK(q_home) * FrankaHandTCPOffset lands on the reference's recorded home TCP pose (a rounded measurement)."""
    M = H.scene()
    m = O.Model(M)
    site = O.robot_cfg(M).attachment_site
    flange = O.ik_forward(m, site, 9, H.Q_HOME)
    tcp = common.Pose(pose_matrix=common.FrankaHandTCPOffset())
    p = common.Pose(translation=flange[:3], quaternion=flange[3:]) * tcp
    assert np.abs(p.translation() - HOME_M[:3, 3]).max() < 2e-3
    assert p.is_close(common.Pose(pose_matrix=HOME_M), eps_r=5e-3, eps_t=5e-3)


def test_positional_constructor_overloads_follow_the_pybind_signatures():
    """rcs.cpp:224-237 / Pose.cpp:24-100: Pose(vec3, vec3) is (rpy_vector, translation); RPY instances are accepted."""
    rpy, t = np.array([0.1, -0.4, 1.2]), np.array([0.3, 0.2, 0.1])
    ref = common.Pose(translation=t, rpy_vector=rpy)
    assert common.Pose(rpy, t).is_close(ref, 1e-14, 1e-14)
    assert common.Pose(common.RPY(*rpy), t).is_close(ref, 1e-14, 1e-14)
    assert common.Pose(common.RPY(*rpy)).is_close(common.Pose(rpy_vector=rpy), 1e-14, 1e-14)
    assert np.allclose(common.Pose(t).translation(), t) and np.allclose(common.Pose(t).rotation_q(), [0, 0, 0, 1])
    q = ref.rotation_q()
    assert common.Pose(q, t).is_close(ref, 1e-14, 1e-14) and common.Pose(ref.rotation_m(), t).is_close(ref, 1e-12, 1e-14)


def test_pose_matrix_uses_the_polar_rotation_like_eigen_affine():
    """FrankaHandTCPOffset carries 0.707 entries (Pose.cpp:11-15): Eigen's Affine3d::rotation() returns the polar factor,
    i.e. exactly Rz(-45 deg); the oracle's tcp constant is that rotation."""
    p = common.Pose(pose_matrix=common.FrankaHandTCPOffset())
    assert np.allclose(p.rotation_q(), [0, 0, -np.sin(np.pi / 8), np.cos(np.pi / 8)], atol=1e-15)
    assert np.allclose(p.as7(), H.FRANKA_HAND_TCP, atol=1e-15)


def test_log3_closed_form_including_near_pi():
    """pinocchio::log3 as restated for the IK: w = theta * axis for rotations of any angle, near pi included."""
    def rod(ax, th):
        ax = np.asarray(ax, float) / np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K, ax * th
    for th in (1e-6, 1e-3, 0.5, 2.0, np.pi - 0.02, np.pi - 5e-3, np.pi - 5e-5):
        for ax in ((0.6, 0.8, 0.0), (0.3, -0.5, 0.8), (-1, 2, 3)):
            R, w = rod(ax, th)
            assert np.abs(O.log3(R) - w).max() < 1e-10, (th, ax)
