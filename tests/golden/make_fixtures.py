#!/usr/bin/env python
"""Generates tests/golden/step_vectors.npz: (qpos, qvel, ctrl) -> state after k steps, contact count and contact geom
pairs, for the three compiled scenes.

Source of truth: a real MuJoCo (`import mujoco`, pinned 3.2.6 in the reference's pyproject.toml:23) if one is importable
in the generating container together with the scene XMLs under /root/reference; otherwise the CPU oracle (oracle/), in
which case the vectors are REGRESSION vectors of the restatement (they pin its behaviour across rounds and are what the
host emulation / CUDA kernels are compared with), not reference goldens -- `source` in the file says which. No golden
vector for mj_step exists anywhere in /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "robot-control-stack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers as H  # noqa: E402
from helpers import O  # noqa: E402

CASES = [  # scene, number of states, steps, seed, spread around home
    ("fr3_empty_world", 12, 1, 1, 0.4),
    ("fr3_empty_world", 6, 60, 2, 0.3),
    ("xarm7_empty_world", 6, 40, 3, 0.3),
    ("fr3_simple_pick_up", 3, 30, 4, 0.1),
    ("xarm7_tabletop", 4, 40, 5, 0.2),
]


def initial_states(scene, n, seed, spread):
    M = H.scene(scene)
    rng = np.random.default_rng(seed)
    nq, nv, nu = M["nq"], M["nv"], M["nu"]
    q = np.tile(M["qpos0"], (n, 1)).astype(float)
    home = H.XARM_Q_HOME if scene.startswith("xarm7") else H.Q_HOME
    if scene == "xarm7_tabletop":
        q[:, 7:9] += rng.uniform(-0.03, 0.03, (n, 2))  # the brick somewhere on the table
    q[:, :7] = home + rng.uniform(-spread, spread, (n, 7))
    v = np.zeros((n, nv)); v[:, :7] = rng.uniform(-0.5, 0.5, (n, 7))
    ctrl = np.zeros((n, nu)); ctrl[:, :7] = q[:, :7] + rng.uniform(-0.15, 0.15, (n, 7))
    if not scene.startswith("xarm7"):
        q[:, 7] = q[:, 8] = rng.uniform(0.002, 0.038, n)
        ctrl[:, 7] = rng.uniform(0, 255, n)
    return M, q, v, ctrl


def floor_case():
    """the arm driven into the floor: contact counts and geom pairs along the way"""
    M = H.scene("fr3_empty_world")
    m, s = H.oracle_sim(M)
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    s.set_joint_position(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]))
    ncon, pairs, q = [], [], []
    for it in range(90):
        s.step(5)
        n = int(s.data.ncon[0])
        ncon.append(n)
        g = s.data.int("contact_geom").reshape(-1, 2)[:n]
        pairs.append(np.pad(g, ((0, 6 - n), (0, 0)), constant_values=-1))
        q.append(s.data.qpos.copy())
    return np.array(ncon), np.array(pairs), np.array(q)


def grasp_case():
    """grasp-and-lift on fr3_simple_pick_up (helpers.grasp_and_lift_script): contact count, geom pairs and qpos every 50 steps"""
    M = H.scene("fr3_simple_pick_up")
    m, s = H.oracle_sim(M, tcp=H.FRANKA_HAND_TCP)
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    ncon, pairs, q = [], [], []
    for qt, w, n in H.grasp_and_lift_script(M):
        s.set_joint_position(qt); s.gripper_set_normalized_width(w)
        for _ in range(n // 50):
            s.step(50)
            k = int(s.data.ncon[0])
            ncon.append(k)
            g = s.data.int("contact_geom").reshape(-1, 2)[:k]
            pairs.append(np.pad(g, ((0, 40 - k), (0, 0)), constant_values=-1))
            q.append(s.data.qpos.copy())
    return np.array(ncon), np.array(pairs), np.array(q)


def main():
    out = {"source": np.array("oracle (CPU restatement; libmujoco 3.2.6 not importable here)")}
    try:
        import mujoco  # noqa: F401
        raise SystemExit("a real MuJoCo is importable: extend this script to step it on the reference's scene XMLs and record "
                         "source='mujoco==%s'" % mujoco.__version__)
    except ImportError:
        pass
    for ci, (scene, n, k, seed, spread) in enumerate(CASES):
        M, q, v, ctrl = initial_states(scene, n, seed, spread)
        m = O.Model(M)
        q1, v1, nc = np.zeros_like(q), np.zeros_like(v), np.zeros(n, dtype=np.int32)
        for i in range(n):
            d = O.Data(m)
            d.qpos[:] = q[i]; d.qvel[:] = v[i]; d.ctrl[:] = ctrl[i]
            d.step(k)
            q1[i], v1[i], nc[i] = d.qpos, d.qvel, int(d.ncon[0])
        for name, arr in (("scene", np.array(scene)), ("k", np.array(k)), ("qpos", q), ("qvel", v), ("ctrl", ctrl), ("qpos_out", q1),
                          ("qvel_out", v1), ("ncon_out", nc)):
            out[f"c{ci}_{name}"] = arr
    out["floor_ncon"], out["floor_pairs"], out["floor_qpos"] = floor_case()
    out["grasp_ncon"], out["grasp_pairs"], out["grasp_qpos"] = grasp_case()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
