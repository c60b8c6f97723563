"""Host (numpy) forward kinematics of a site on the compiled scene; used by Pin.forward only
(/root/reference/src/rcs/Kinematics.cpp:70-81). Not on the per-step path."""
from __future__ import annotations

import numpy as np

from .mjcf import JNT_FREE, JNT_SLIDE, axisangle_quat, quat_mul, quat_to_mat


def site_pose(M: dict, site: int, q: np.ndarray):
    chain = []
    b = int(M["site_bodyid"][site])
    while b > 0:
        chain.append(b)
        b = int(M["body_parentid"][b])
    pos, quat = np.zeros(3), np.array([1.0, 0, 0, 0])
    for b in reversed(chain):
        pos = pos + quat_to_mat(quat) @ M["body_pos"][b]
        quat = quat_mul(quat, M["body_quat"][b])
        for jj in range(int(M["body_jntnum"][b])):
            j = int(M["body_jntadr"][b]) + jj
            if M["jnt_type"][j] == JNT_FREE:
                continue
            qa = int(M["jnt_qposadr"][j])
            qj = (q[qa] if qa < len(q) else 0.0) - M["qpos0"][qa]
            R = quat_to_mat(quat)
            axis, anchor = R @ M["jnt_axis"][j], pos + R @ M["jnt_pos"][j]
            if M["jnt_type"][j] == JNT_SLIDE:
                pos = pos + axis * qj
            else:
                quat = quat_mul(quat, axisangle_quat(M["jnt_axis"][j], qj))
                pos = anchor - quat_to_mat(quat) @ M["jnt_pos"][j]
        quat = quat / np.linalg.norm(quat)
    R = quat_to_mat(quat)
    p = pos + R @ M["site_pos"][site]
    Rs = quat_to_mat(quat_mul(quat, M["site_quat"][site]))
    return Rs, p
