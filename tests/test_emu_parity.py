"""CPU-tier parity of the DEVICE CODE's logic: csrc/*.cuh compiled for one host lane (tests/emu, test
infrastructure only) against the CPU oracle. The GPU tier (test_gpu_parity.py) repeats these through the real
kernels; this tier catches arithmetic / indexing errors without a GPU, and runs every parallel-for in reverse
order to catch lane-order dependences (results must be bit-identical)."""
import numpy as np
import pytest

import helpers as H
from helpers import O
from rcs_b200 import devmodel
from emu.emu import Emu

def _emu(M, F, verts, N, **kw):
    return Emu(F, verts, N, graph=devmodel.build_mesh_graph(M), **kw)


RESET = ["GRIPPER_RESET", "SIM_RESET", "ROBOT_RESET", "ENV_RESET_FLAGS", "STEP_K", "OBS"]


@pytest.fixture(scope="module")
def fr3():
    M = H.scene()
    F, verts = devmodel.build_device_fields(M, H.robot_ns(), H.gripper_ns())
    return M, F, verts


def test_single_step_random_states(fr3):
    M, F, verts = fr3
    N = 64
    e = _emu(M, F, verts, N)
    rng = np.random.default_rng(1)
    q = np.zeros((N, 9)); v = np.zeros((N, 9)); ctrl = np.zeros((N, 8))
    q[:, :7] = H.Q_HOME + rng.uniform(-0.4, 0.4, (N, 7)); q[:, 7] = q[:, 8] = rng.uniform(0.001, 0.039, N)
    v[:, :7] = rng.uniform(-1, 1, (N, 7)); v[:, 7] = v[:, 8] = rng.uniform(-0.05, 0.05, N)
    ctrl[:, :7] = q[:, :7] + rng.uniform(-0.2, 0.2, (N, 7)); ctrl[:, 7] = rng.uniform(0, 255, N)
    e.sr[:, 0:9] = q; e.sr[:, 9:18] = v; e.sr[:, 18:26] = ctrl
    e.run(["STEP_K"], k=1)
    m = O.Model(M)
    for i in range(N):
        d = O.Data(m)
        d.qpos[:] = q[i]; d.qvel[:] = v[i]; d.ctrl[:] = ctrl[i]
        d.step()
        assert np.abs(e.sr[i, :9] - d.qpos).max() < 1e-13
        assert np.abs(e.sr[i, 9:18] - d.qvel).max() < 1e-11
        assert np.abs(e.wsf("M", 81, i) - d.qM).max() < 1e-13
        assert np.abs(e.wsf("bias", 9, i) - d.qfrc_bias).max() < 1e-11


def test_workload_trajectory_and_lane_order_independence(fr3):
    M, F, verts = fr3
    N, T = 6, 20
    acts = H.workload_actions(N, T, seed=0)
    m = O.Model(M)
    _, _, ref = O.bench_env_steps(m, O.robot_cfg(M), O.gripper_cfg(M), acts, 2, 10, True, np.deg2rad(5), H.JLOW, H.JHIGH,
                                  want_obs=True)
    outs = []
    for rev in (False, True):
        e = _emu(M, F, verts, N, reverse=rev)
        e.run(RESET, k=1)
        traj = []
        for t in range(T):
            if t > 0 and t % 10 == 0:
                e.run(RESET, k=1)
            e.run(["ACT_JOINTS_REL", "ACT_GRIPPER_BIN", "STEP_K", "OBS"], k=17, act_joints=acts[:, t, :7].copy(),
                  act_gripper=acts[:, t, 7].copy(), max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH)
            traj.append(e.obs.copy())
            # 1e-6: reset states sit exactly on the finger joint limit (see test_gpu_parity.py header)
            assert np.abs(e.obs[:, :14] - ref[:, t, :14]).max() < 1e-6
            assert np.array_equal(e.obs[:, 20], ref[:, t, 20])
        outs.append(np.array(traj))
    # forward vs reverse iteration may differ only by the rounding of reordered reductions (cost sums);
    # a real lane-order dependence (one item reading what another item writes) shows up as a gross difference
    diff = np.abs(outs[0][:, :, :14] - outs[1][:, :, :14]).max()
    print("forward/reverse max diff", diff)
    assert diff < 1e-6, "a parallel-for depends on lane order"


def test_floor_collision_contact_indexing_exact(fr3):
    """north_star: contact-pair indexing is bit-exact. The device code's exported contact list (count, geom ids in
    mjModel numbering, order) equals the oracle's mjData.contact on every sample of the floor-collision trajectory, and
    equals the committed floor_pairs fixture; dist / pos / normal agree to rounding."""
    import os
    M, F, verts = fr3
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_vectors.npz"))
    e = _emu(M, F, verts, 1)
    e.enable_contact_export(cap=8)
    mm, s = H.oracle_sim(M)
    tgt = np.array([0, 1.78, 0, -1.45, 0, 0, 0.0])
    e.run(RESET[:-1], k=1); s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    e.run(["SET_JOINTS"], act_joints=tgt[None]); s.set_joint_position(tgt)
    hit = 0
    for it in range(80):
        e.run(["STEP_K"], k=5); s.step(5)
        ncon = int(e.si[0, 14])
        assert ncon == int(s.data.ncon[0]) == int(e.contact_n[0]) and int(e.si[0, 15]) == int(s.data.nefc[0])
        ref_pairs = s.data.int("contact_geom").reshape(-1, 2)
        assert np.array_equal(e.contact_geom[0, :ncon], ref_pairs), (it, e.contact_geom[0, :ncon], ref_pairs)
        assert (e.contact_geom[0, ncon:] == -1).all()
        assert np.array_equal(e.contact_geom[0, :6], G["floor_pairs"][it]), it
        if ncon:
            hit += 1
            ref = s.data.real("contact_real").reshape(-1, 7)
            assert np.abs(e.contact_real[0, :ncon] - ref).max() < 1e-7, it
        assert np.abs(e.sr[0, :9] - s.data.qpos).max() < 1e-6
    assert hit > 10
    # collision flags after a converge call agree
    e.run(["STEP_CONV"]); s.step_until_convergence()
    assert bool(e.si[0, 6]) == s.is_converged() and int(e.si[0, 7]) == s.convergence_steps()
    assert bool(e.si[0, 1]) == s.robot_state()["collision"] and bool(e.si[0, 5]) == s.gripper_state()["collision"]
    assert s.robot_state()["collision"]


def test_step_until_convergence_counts(fr3):
    M, F, verts = fr3
    rng = np.random.default_rng(3)
    N = 4
    tg = H.Q_HOME + rng.uniform(-0.15, 0.15, (N, 7))
    e = _emu(M, F, verts, N)
    e.run(RESET[:-1], k=1)
    e.run(["SET_JOINTS", "STEP_CONV"], act_joints=tg, max_conv=500)
    for i in range(N):
        mm, s = H.oracle_sim(M)
        s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
        s.set_joint_position(tg[i]); s.step_until_convergence()
        assert int(e.si[i, 7]) == s.convergence_steps() and bool(e.si[i, 6]) == s.is_converged()
        assert np.abs(e.sr[i, :9] - s.data.qpos).max() < 1e-7  # starts from a reset state (finger on its limit)
        st = s.robot_state()
        assert bool(e.si[i, 2]) == st["is_moving"] and bool(e.si[i, 3]) == st["is_arrived"]


def test_pick_up_scene_cube_contacts(fr3):
    """fr3_simple_pick_up: free joint, box-plane contacts, elliptic cones + noslip in the device code."""
    M = H.scene("fr3_simple_pick_up")
    F, verts = devmodel.build_device_fields(M, H.robot_ns(), H.gripper_ns())
    e = _emu(M, F, verts, 1)
    m = O.Model(M)
    d = O.Data(m)
    q0 = M["qpos0"].copy(); q0[:7] = H.Q_HOME
    ctrl = np.zeros(8); ctrl[:7] = H.Q_HOME
    d.qpos[:] = q0; d.ctrl[:] = ctrl
    e.sr[0, :16] = q0; e.sr[0, 31:39] = ctrl
    for it in range(30):
        e.run(["STEP_K"], k=10); d.step(10)
        assert int(e.si[0, 14]) == int(d.ncon[0])
        assert np.abs(e.sr[0, :16] - d.qpos).max() < 1e-7, it
    assert int(d.ncon[0]) >= 1


def test_reduced_layout_hand_over_is_bit_exact(fr3):
    """Environments that outgrow the reduced workspace layout (fast_maxcon / fast_maxefc) are finished by the
    full-capacity pass; the result must equal running everything in the full layout, bit for bit, for STEP_K and
    for step_until_convergence."""
    M, F, verts = fr3
    tgt = np.tile(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]), (3, 1))  # drives the arm into the floor
    tgt[1] = H.Q_HOME + 0.1                                          # this one never touches anything
    outs = []
    for use_reduced in (True, False):
        e = _emu(M, F, verts, 3, use_reduced=use_reduced)
        assert e.has_reduced == use_reduced
        e.run(RESET[:-1], k=1)
        e.run(["SET_JOINTS"], act_joints=tgt)
        handed = 0
        for it in range(60):
            e.run(["STEP_K", "OBS"], k=7)
            handed += e.handed_over
        e.run(["STEP_CONV", "OBS"], max_conv=120)
        handed += e.handed_over
        outs.append((e.sr.copy(), e.sd.copy(), e.si.copy(), e.obs.copy(), e.info.copy(), handed))
    assert outs[0][5] > 5 and outs[1][5] == 0, "the floor contacts must exceed the reduced capacity"
    assert int(outs[0][2][0, 14]) >= 1  # still in contact at the end
    for a, b in zip(outs[0][:5], outs[1][:5]):
        assert np.array_equal(a, b)
    assert not outs[0][2][:, 20].any()  # RCSB_I_RESUME cleared everywhere


def test_separation_budgets_detect_contact_on_the_same_step(fr3):
    """One long launch (the collision groups' separation budgets are only reset at launch start) driving the arm into
    the floor: the first contact must appear on exactly the step the oracle (full collision pass every step) sees it,
    otherwise the trajectories separate."""
    M, F, verts = fr3
    tgt = np.array([0, 1.78, 0, -1.45, 0, 0, 0.0])
    for k in (150, 230, 300, 420):
        e = _emu(M, F, verts, 1)
        mm, s = H.oracle_sim(M)
        e.run(RESET[:-1], k=1); s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
        e.run(["SET_JOINTS", "STEP_K"], k=k, act_joints=tgt[None]); s.set_joint_position(tgt); s.step(k)
        assert int(e.si[0, 14]) == int(s.data.ncon[0]), k
        assert np.abs(e.sr[0, :9] - s.data.qpos).max() < 1e-7, k
        assert np.abs(e.sr[0, 9:18] - s.data.qvel).max() < 1e-5, k


def test_xarm7_friction_loss_rows_and_cylinder(fr3):
    """xarm7_empty_world: 7 friction-loss rows on every step (solved by zone verification, no Newton iteration),
    pyramidal cone option, a cylinder pedestal in the collision set, no gripper / tendon / equality."""
    M = H.scene("xarm7_empty_world")
    F, verts = devmodel.build_device_fields(M, H.xarm_robot_ns(), None)
    N = 16
    e = _emu(M, F, verts, N)
    rng = np.random.default_rng(11)
    q = H.XARM_Q_HOME + rng.uniform(-0.3, 0.3, (N, 7))
    v = rng.uniform(-0.5, 0.5, (N, 7)); v[::3] = 0  # zero velocity: friction rows in their quadratic (sticking) zone
    ctrl = q + rng.uniform(-0.1, 0.1, (N, 7))
    e.sr[:, 0:7] = q; e.sr[:, 7:14] = v; e.sr[:, 14:21] = ctrl
    m = O.Model(M)
    ds = []
    for i in range(N):
        d = O.Data(m); d.qpos[:] = q[i]; d.qvel[:] = v[i]; d.ctrl[:] = ctrl[i]; ds.append(d)
    for it in range(12):
        e.run(["STEP_K"], k=5)
        for i, d in enumerate(ds):
            d.step(5)
            assert int(e.si[i, 15]) == int(d.nefc[0]) >= 7
            assert np.abs(e.sr[i, 0:7] - d.qpos).max() < 1e-9, (it, i)
            assert np.abs(e.sr[i, 7:14] - d.qvel).max() < 1e-7, (it, i)


def test_grasp_and_lift_contacts_and_state():
    """fr3_simple_pick_up, config C3's defining behaviour: close the gripper on the cube and raise it 12 cm. The finger pads
    and the cube are boxes (fr3_0.xml:146-150, fr3_simple_pick_up/scene.xml:31): box-box face clipping gives 4 points per
    small pad (36 contacts with the 4 floor points), elliptic cones + noslip. Contact count, geom pair ids and order are
    compared exactly with the oracle at every sample, the state to 1e-6, and the cube must end up lifted."""
    M = H.scene("fr3_simple_pick_up")
    F, verts = devmodel.build_device_fields(M, H.robot_ns(H.FRANKA_HAND_TCP), H.gripper_ns())
    e = _emu(M, F, verts, 1)
    e.enable_contact_export(cap=40)
    m, s = H.oracle_sim(M, tcp=H.FRANKA_HAND_TCP)
    e.run(RESET[:-1], k=1); s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    peak, worst = 0, 0.0
    for qt, w, n in H.grasp_and_lift_script(M):
        e.run(["SET_JOINTS", "SET_GRIPPER"], act_joints=qt[None], act_gripper=np.array([w]))
        s.set_joint_position(qt); s.gripper_set_normalized_width(w)
        for _ in range(n // 50):
            e.run(["STEP_K"], k=50); s.step(50)
            ncon = int(s.data.ncon[0])
            peak = max(peak, ncon)
            assert int(e.contact_n[0]) == ncon == int(e.si[0, 14])
            assert np.array_equal(e.contact_geom[0, :ncon], s.data.int("contact_geom").reshape(-1, 2))
            worst = max(worst, np.abs(e.sr[0, :16] - s.data.qpos).max())
            assert np.abs(e.sr[0, :16] - s.data.qpos).max() < 1e-6
            assert int(e.si[0, 17]) == 0  # RCSB_I_WARN: nothing dropped
    print("grasp: peak ncon", peak, "worst |dq|", worst)
    assert peak >= 30, peak
    assert s.data.qpos[11] > 0.12 and e.sr[0, 11] > 0.12  # the cube went up with the gripper


def test_xarm7_tabletop_brick_contacts_friction_loss_pyramidal():
    """Synthetic config C4 (tools/scenes/xarm7_tabletop.xml): xArm7 + table + one free duplo brick. Every step carries 7
    friction-loss rows, 4 box-box contacts brick-table as pyramidal cones (16 rows, regulariser 2 mu^2 R) and the
    block-diagonal (arm | brick) solver path; random joint targets. Contact pairs exact, state 1e-7."""
    M = H.scene("xarm7_tabletop")
    F, verts = devmodel.build_device_fields(M, H.xarm_robot_ns(), None)
    N = 4
    e = _emu(M, F, verts, N)
    e.enable_contact_export(cap=16)
    assert e.has_reduced
    rng = np.random.default_rng(7)
    q0 = np.tile(M["qpos0"], (N, 1)).astype(float)
    q0[:, :7] = H.XARM_Q_HOME + rng.uniform(-0.2, 0.2, (N, 7))
    q0[:, 7:9] += rng.uniform(-0.03, 0.03, (N, 2))
    q0[1, 9] += 0.01                                   # this brick is dropped from 1 cm
    ctrl = q0[:, :7] + rng.uniform(-0.1, 0.1, (N, 7))
    e.sr[:, :14] = q0; e.sr[:, 27:34] = ctrl
    m = O.Model(M)
    ds = []
    for i in range(N):
        d = O.Data(m); d.qpos[:] = q0[i]; d.ctrl[:] = ctrl[i]; ds.append(d)
    seen = 0
    for it in range(20):
        e.run(["STEP_K"], k=10)
        for i, d in enumerate(ds):
            d.step(10)
            n = int(d.ncon[0])
            seen = max(seen, n)
            assert int(e.contact_n[i]) == n and int(e.si[i, 15]) == int(d.nefc[0]), (it, i)
            assert np.array_equal(e.contact_geom[i, :n], d.int("contact_geom").reshape(-1, 2))
            assert np.abs(e.sr[i, :14] - d.qpos).max() < 1e-7, (it, i)
            assert np.abs(e.sr[i, 14:27] - d.qvel).max() < 1e-5, (it, i)
    assert seen == 4 and int(e.si[:, 17].max()) == 0
