"""Host-side mirror of `rcs._core.common` (/root/reference/src/pybind/rcs.cpp:204-417): Pose, RPY,
RobotType, RobotPlatform, RobotMetaConfig, robots_meta_config, FrankaHandTCPOffset, Kinematics, Pin,
and the abstract Robot / Gripper device API (/root/reference/include/rcs/Robot.h:127-197).

Scalar `Pose` is plain numpy float64 with Eigen's conventions restated (quaternion order x,y,z,w;
every constructor except Pose(rotation=Matrix3) normalises; eulerAngles(2,1,0) ranges). The batched
equivalents used on the hot path are CUDA device functions (csrc/rcsb_env.cuh).
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field

import numpy as np


# ------------------------------------------------------------------ quaternion helpers (x, y, z, w)
def _q_normalized(q):
    n = math.sqrt(float(np.dot(q, q)))
    return q / n if n > 0 else q


def _q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def _q_conj(a):
    return np.array([-a[0], -a[1], -a[2], a[3]])


def _q_rot(q, v):
    qv = q[:3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[3] * uv + np.cross(qv, uv)


def _q_to_mat(q):
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def _q_from_mat(m):
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t
        q[1] = (m[0, 2] - m[2, 0]) * t
        q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return q


def _q_angular_distance(a, b):
    d = _q_mul(a, _q_conj(b))
    return 2.0 * math.atan2(math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), abs(d[3]))


def _q_slerp(a, t, b):
    one = 1.0 - np.finfo(np.float64).eps
    d = float(np.dot(a, b))
    ad = abs(d)
    if ad >= one:
        s0, s1 = 1.0 - t, t
    else:
        th = math.acos(ad)
        st = math.sin(th)
        s0, s1 = math.sin((1.0 - t) * th) / st, math.sin(t * th) / st
    if d < 0:
        s1 = -s1
    return s0 * a + s1 * b


def _q_from_rpy(roll, pitch, yaw):
    qz = np.array([0, 0, math.sin(yaw / 2), math.cos(yaw / 2)])
    qy = np.array([0, math.sin(pitch / 2), 0, math.cos(pitch / 2)])
    qx = np.array([math.sin(roll / 2), 0, 0, math.cos(roll / 2)])
    return _q_mul(_q_mul(qz, qy), qx)


def IdentityTranslation():
    return np.zeros(3)


def IdentityRotMatrix():
    return np.eye(3)


def IdentityRotQuatVec():
    return np.array([0.0, 0.0, 0.0, 1.0])


def FrankaHandTCPOffset():
    """/root/reference/src/rcs/Pose.cpp:11-15"""
    return np.array([[0.707, 0.707, 0, 0], [-0.707, 0.707, 0, 0], [0, 0, 1, 0.1034], [0, 0, 0, 1]])


class RPY:
    """/root/reference/include/rcs/Pose.h:26-65"""

    def __init__(self, roll=0.0, pitch=0.0, yaw=0.0, rpy=None):
        if rpy is not None:
            roll, pitch, yaw = [float(x) for x in rpy]
        elif not np.isscalar(roll):
            roll, pitch, yaw = [float(x) for x in roll]
        self.roll, self.pitch, self.yaw = float(roll), float(pitch), float(yaw)

    def rotation_matrix(self):
        return _q_to_mat(_q_from_rpy(self.roll, self.pitch, self.yaw))

    def as_quaternion_vector(self):
        return _q_from_rpy(self.roll, self.pitch, self.yaw)

    def as_vector(self):
        return np.array([self.roll, self.pitch, self.yaw])

    def is_close(self, other, eps=1e-8):
        return float(np.abs(self.as_vector() - other.as_vector()).sum()) < eps

    def __add__(self, o):
        return RPY(self.roll + o.roll, self.pitch + o.pitch, self.yaw + o.yaw)

    def __str__(self):
        return f"RPY({self.roll:.6f}, {self.pitch:.6f}, {self.yaw:.6f})"

    __repr__ = __str__

    def __getstate__(self):
        return (self.roll, self.pitch, self.yaw)

    def __setstate__(self, s):
        self.roll, self.pitch, self.yaw = s


class Pose:
    """Immutable SE(3) (/root/reference/src/rcs/Pose.cpp, constructor overloads of rcs.cpp:224-237)."""

    __slots__ = ("_t", "_q")

    def __init__(self, *args, pose_matrix=None, rotation=None, translation=None, quaternion=None, rpy=None,
                 rpy_vector=None, pose=None):
        t, q, normalize = np.zeros(3), IdentityRotQuatVec(), True
        if args:  # positional overloads of rcs.cpp:224-237 / Pose.cpp:24-100
            a = args[0]
            second = np.asarray(args[1], dtype=np.float64) if len(args) > 1 else None
            if isinstance(a, Pose):
                pose = a
            elif isinstance(a, RPY):
                rpy, translation = a, second
            else:
                a = np.asarray(a, dtype=np.float64)
                if a.shape == (4, 4):
                    pose_matrix = a
                elif a.shape == (3, 3):
                    rotation, translation = a, second
                elif a.shape == (4,):
                    quaternion, translation = a, second
                elif a.shape == (3,) and second is not None:
                    rpy_vector, translation = a, second  # Pose(Vector3d rotation, Vector3d translation): rpy first
                elif a.shape == (3,):
                    translation = a
                else:
                    raise TypeError("unsupported positional argument for Pose")
        if pose is not None:
            t, q, normalize = pose._t.copy(), pose._q.copy(), False
        elif pose_matrix is not None:
            m = np.asarray(pose_matrix, dtype=np.float64)
            # Eigen's Affine3d::rotation() is the polar factor of the linear part (SVD), Pose.cpp:24-38: matters for
            # rounded matrices such as FrankaHandTCPOffset (0.707 entries)
            u, _, vt = np.linalg.svd(m[:3, :3])
            r = u @ vt
            if np.linalg.det(r) < 0:
                u[:, -1] = -u[:, -1]
                r = u @ vt
            t, q = m[:3, 3].copy(), _q_from_mat(r)
        else:
            if translation is not None:
                t = np.asarray(translation, dtype=np.float64).reshape(3).copy()
            if rotation is not None:
                q = _q_from_mat(np.asarray(rotation, dtype=np.float64))
                normalize = translation is not None  # Pose(Matrix3d) alone does not normalise (Pose.cpp:94-97)
            elif quaternion is not None:
                q = np.asarray(quaternion, dtype=np.float64).reshape(4).copy()
            elif rpy is not None:
                q = rpy.as_quaternion_vector()
            elif rpy_vector is not None:
                q = RPY(rpy=np.asarray(rpy_vector, dtype=np.float64)).as_quaternion_vector()
            elif translation is not None:
                normalize = False
        self._t = t
        self._q = _q_normalized(q) if normalize else q

    @staticmethod
    def Identity():
        return Pose()

    @classmethod
    def _from_tq(cls, t, q):
        p = cls.__new__(cls)
        p._t = np.asarray(t, dtype=np.float64).copy()
        p._q = _q_normalized(np.asarray(q, dtype=np.float64))
        return p

    def translation(self):
        return self._t.copy()

    def rotation_m(self):
        return _q_to_mat(self._q)

    def rotation_q(self):
        return self._q.copy()

    def pose_matrix(self):
        m = np.eye(4)
        m[:3, :3] = _q_to_mat(self._q)
        m[:3, 3] = self._t
        return m

    def rotation_rpy(self):
        m = _q_to_mat(self._q)
        r0 = math.atan2(m[1, 0], m[0, 0])
        c2 = math.hypot(m[2, 2], m[2, 1])
        if r0 < 0:
            r0 += math.pi
            r1 = math.atan2(-m[2, 0], -c2)
        else:
            r1 = math.atan2(-m[2, 0], c2)
        s1, c1 = math.sin(r0), math.cos(r0)
        r2 = math.atan2(s1 * m[0, 2] - c1 * m[1, 2], c1 * m[1, 1] - s1 * m[0, 1])
        return RPY(r2, r1, r0)

    def xyzrpy(self):
        return np.concatenate([self._t, self.rotation_rpy().as_vector()])

    def interpolate(self, dest_pose, progress):
        progress = min(progress, 1.0)
        return Pose._from_tq(self._t + (dest_pose._t - self._t) * progress, _q_slerp(self._q, progress, dest_pose._q))

    def inverse(self):
        c = _q_conj(self._q)
        return Pose._from_tq(-_q_rot(c, self._t), c)

    def total_angle(self):
        return _q_angular_distance(self._q, IdentityRotQuatVec())

    def limit_rotation_angle(self, max_angle):
        cur = self.total_angle()
        if cur > max_angle and max_angle >= 0:
            return Pose._from_tq(self._t, _q_slerp(IdentityRotQuatVec(), max_angle / cur, self._q))
        return Pose(pose=self)

    def limit_translation_length(self, max_length):
        n = float(np.linalg.norm(self._t))
        if n > max_length and max_length >= 0:
            return Pose._from_tq(self._t / n * max_length, self._q)
        return Pose(pose=self)

    def is_close(self, other, eps_r=1e-8, eps_t=1e-8):
        return float(np.abs(self._t - other._t).sum()) < eps_t and _q_angular_distance(self._q, other._q) < eps_r

    def __mul__(self, other):
        return Pose._from_tq(_q_rot(self._q, other._t) + self._t, _q_mul(self._q, other._q))

    def as7(self):
        """xyz + quat(xyzw): the packed form used by the C ABI."""
        return np.concatenate([self._t, self._q])

    def __str__(self):
        r = self.rotation_rpy()
        return f"{self.pose_matrix()}\nroll: {r.roll}\tpitch: {r.pitch}\tyaw: {r.yaw}"

    __repr__ = __str__

    def __getstate__(self):
        return (self._t, self._q)

    def __setstate__(self, s):
        self._t, self._q = s


class RobotType(enum.IntEnum):
    FR3 = 0
    UR5e = 1
    SO101 = 2
    XArm7 = 3


class RobotPlatform(enum.IntEnum):
    SIMULATION = 0
    HARDWARE = 1


@dataclass
class RobotMetaConfig:
    q_home: np.ndarray
    dof: int
    joint_limits: np.ndarray  # 2 x dof


_META = {  # /root/reference/include/rcs/Robot.h:24-95
    RobotType.FR3: RobotMetaConfig(
        np.array([0.0, -math.pi / 4, 0.0, -3.0 * math.pi / 4, 0.0, math.pi / 2, math.pi / 4]), 7,
        np.array([[-2.3093, -1.5133, -2.4937, -2.7478, -2.4800, 0.8521, -2.6895],
                  [2.3093, 1.5133, 2.4937, -0.4461, 2.4800, 4.2094, 2.6895]])),
    RobotType.UR5e: RobotMetaConfig(
        np.array([-0.4488354, -2.02711196, 1.64630026, -1.18999615, -1.57079762, -2.01963249]), 6,
        np.array([[-2 * math.pi, -2 * math.pi, -math.pi, -2 * math.pi, -2 * math.pi, -2 * math.pi],
                  [2 * math.pi, 2 * math.pi, math.pi, 2 * math.pi, 2 * math.pi, 2 * math.pi]])),
    RobotType.XArm7: RobotMetaConfig(
        np.array([0, -45.0 / 180 * math.pi, 0, 15.0 / 180 * math.pi, 0, -25.0 / 180 * math.pi, 0]), 7,
        np.array([[-2 * math.pi, -2.094395, -2 * math.pi, -3.92699, -2 * math.pi, -math.pi, -2 * math.pi],
                  [2 * math.pi, 2.059488, 2 * math.pi, 0.191986, 2 * math.pi, 1.692969, 2 * math.pi]])),
    RobotType.SO101: RobotMetaConfig(
        np.array([-9.40612320177057, -99.66130397967824, 99.9124726477024, 69.96996996996998, -9.095744680851055]), 5,
        np.array([[-100.0] * 5, [100.0] * 5])),
}


def robots_meta_config(robot_type: RobotType) -> RobotMetaConfig:
    return _META[RobotType(robot_type)]


@dataclass
class RobotConfig:
    """/root/reference/include/rcs/Robot.h:97-104"""
    robot_type: RobotType = RobotType.FR3
    robot_platform: RobotPlatform = RobotPlatform.SIMULATION
    tcp_offset: Pose = field(default_factory=Pose)
    attachment_site: str = "attachment_site"
    kinematic_model_path: str = "assets/scenes/fr3_empty_world/robot.xml"


class Kinematics:
    """/root/reference/include/rcs/Kinematics.h:19-26"""

    def inverse(self, pose: Pose, q0: np.ndarray, tcp_offset: Pose = None):
        raise NotImplementedError

    def forward(self, q0: np.ndarray, tcp_offset: Pose = None) -> Pose:
        raise NotImplementedError


class Robot:
    """Abstract device API (/root/reference/include/rcs/Robot.h:127-169)."""

    def get_config(self): raise NotImplementedError
    def get_state(self): raise NotImplementedError
    def get_cartesian_position(self): raise NotImplementedError
    def set_joint_position(self, q): raise NotImplementedError
    def get_joint_position(self): raise NotImplementedError
    def move_home(self): raise NotImplementedError
    def reset(self): raise NotImplementedError
    def close(self): raise NotImplementedError
    def set_cartesian_position(self, pose): raise NotImplementedError
    def get_ik(self): raise NotImplementedError
    def get_base_pose_in_world_coordinates(self): raise NotImplementedError

    def to_pose_in_world_coordinates(self, pose_in_robot_coordinates: Pose) -> Pose:  # Robot.cpp:10-13
        return self.get_base_pose_in_world_coordinates() * pose_in_robot_coordinates

    def to_pose_in_robot_coordinates(self, pose_in_world_coordinates: Pose) -> Pose:  # Robot.cpp:5-8
        return self.get_base_pose_in_world_coordinates().inverse() * pose_in_world_coordinates


class Gripper:
    """Abstract device API (/root/reference/include/rcs/Robot.h:171-197)."""

    def get_config(self): raise NotImplementedError
    def get_state(self): raise NotImplementedError
    def set_normalized_width(self, width, force=0): raise NotImplementedError
    def get_normalized_width(self): raise NotImplementedError
    def is_grasped(self): raise NotImplementedError
    def grasp(self): raise NotImplementedError
    def open(self): raise NotImplementedError
    def shut(self): raise NotImplementedError
    def reset(self): raise NotImplementedError
    def close(self): raise NotImplementedError
