"""GPU parity proper: CUDA kernels (through the C ABI) vs the CPU oracle on identical inputs.

Tolerances (float64 on both sides; the two implementations differ structurally -- fused body tree,
mask-based recursions, warp-parallel reductions -- so agreement is to rounding, not bitwise):
  qpos / qvel after one step from identical state .... 1e-12 abs
  trajectories from generic states ..................... 1e-9 abs
  trajectories through reset() states .................. 1e-6 abs: after reset the fingers sit EXACTLY on their
      joint limit (qpos0 = 0 = range[0]) with zero actuator force, so the sign of 1e-17 rounding noise decides
      whether the limit row activates on the first substeps, and an activating limit row applies a finite
      velocity-proportional force (MuJoCo's soft-constraint reference acceleration is discontinuous at dist = 0).
      The effect is ~1e-7 on finger velocity and <1e-8 on arm joints; it is a property of the model, present in
      MuJoCo itself, not of either implementation.
  contact pair indexing (geom ids, count) ............. exact
"""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from rcs_b200 import _lib, batch
    M = H.scene()
    dm = batch.DeviceModel(M, H.robot_ns(), H.gripper_ns())
    return M, dm, _lib, batch


def test_library_is_cuda(setup):
    M, dm, _lib, batch = setup
    assert _lib.lib().rcsb_real_bytes() == 8
    assert torch.cuda.is_available()


def test_single_step_parity_random_states(setup):
    M, dm, _lib, batch = setup
    N = 256
    b = batch.Batch(dm, N)
    rng = np.random.default_rng(1)
    q = np.zeros((N, 9)); v = np.zeros((N, 9)); ctrl = np.zeros((N, 8))
    q[:, :7] = H.Q_HOME + rng.uniform(-0.4, 0.4, (N, 7))
    q[:, 7] = q[:, 8] = rng.uniform(0.001, 0.039, N)
    v[:, :7] = rng.uniform(-1, 1, (N, 7)); v[:, 7] = v[:, 8] = rng.uniform(-0.05, 0.05, N)
    ctrl[:, :7] = q[:, :7] + rng.uniform(-0.2, 0.2, (N, 7)); ctrl[:, 7] = rng.uniform(0, 255, N)
    b.qpos.copy_(torch.as_tensor(q)); b.qvel.copy_(torch.as_tensor(v)); b.ctrl.copy_(torch.as_tensor(ctrl))
    b.run(_lib.STEP_K, k=1)
    torch.cuda.synchronize()
    gq, gv, gw = b.qpos.cpu().numpy(), b.qvel.cpu().numpy(), b.qacc_warmstart.cpu().numpy()
    m = O.Model(M)
    for e in range(N):
        d = O.Data(m)
        d.qpos[:] = q[e]; d.qvel[:] = v[e]; d.ctrl[:] = ctrl[e]
        d.step()
        assert np.abs(gq[e] - d.qpos).max() < 1e-12
        assert np.abs(gv[e] - d.qvel).max() < 1e-11
        assert np.abs(gw[e] - d.qacc_warmstart).max() < 1e-7 * max(1.0, np.abs(d.qacc_warmstart).max())


def test_env_workload_trajectory_parity(setup):
    """The benchmark workload (JOINTS relative, binary gripper, async 17 substeps, reset every 10 steps)
    for 16 envs x 20 steps against the oracle's env loop: observations must agree."""
    M, dm, _lib, batch = setup
    N, T = 16, 20
    acts = H.workload_actions(N, T, seed=0)
    m = O.Model(M)
    sec, psteps, ref_obs = O.bench_env_steps(m, O.robot_cfg(M), O.gripper_cfg(M), acts, nthreads=4, episode_len=10,
                                             async_control=True, joint_low=H.JLOW, joint_high=H.JHIGH, want_obs=True)
    b = batch.Batch(dm, N)
    reset_ops = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
    step_ops = _lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS
    b.run(reset_ops, k=1, want_obs=True)
    worst = 0.0
    for t in range(T):
        if t > 0 and t % 10 == 0:
            b.run(reset_ops, k=1, want_obs=True)
        aj = torch.as_tensor(acts[:, t, :7].copy(), device=b.dev)
        ag = torch.as_tensor(acts[:, t, 7].copy(), device=b.dev)
        b.run(step_ops, k=17, act_joints=aj, act_gripper=ag, max_mov=np.deg2rad(5), jlow=H.JLOW, jhigh=H.JHIGH, want_obs=True)
        g = b.obs.cpu().numpy()
        # tquat, joints, gripper: direct compare; xyzrpy: compare only away from the Eigen euler branch cut
        err = np.abs(g[:, :14] - ref_obs[:, t, :14]).max()
        worst = max(worst, err)
        assert err < 1e-6, (t, err)
        assert np.array_equal(g[:, 20], ref_obs[:, t, 20])
        assert np.abs(g[:, 21] - ref_obs[:, t, 21]).max() < 1e-6  # info["gripper_width"], SimGripper.cpp:93-106
        gi = b.info.cpu().numpy()
        # info: collision, ik_success, is_sim_converged, is_grasped | robot / gripper collision (envs/sim.py:60-66,125-131)
        assert np.array_equal(gi[:, [0, 1, 2, 3, 5, 6]], ref_obs[:, t, 22:28].astype(np.int32)), t
        safe = np.abs(np.abs(ref_obs[:, t, 19]) - np.pi / 2) < 1.4  # yaw well inside (0, pi)
        safe &= (ref_obs[:, t, 19] > 0.05) & (ref_obs[:, t, 19] < np.pi - 0.05)
        assert np.abs(g[safe, 14:20] - ref_obs[safe, t, 14:20]).max(initial=0) < 1e-5
    print("worst obs error", worst)


def test_generic_state_trajectory_parity(setup):
    """170 physics steps from generic states (fingers open, away from every limit): 1e-9."""
    M, dm, _lib, batch = setup
    N = 32
    rng = np.random.default_rng(11)
    b = batch.Batch(dm, N)
    q = np.zeros((N, 9)); ctrl = np.zeros((N, 8))
    q[:, :7] = H.Q_HOME + rng.uniform(-0.3, 0.3, (N, 7)); q[:, 7] = q[:, 8] = 0.02
    ctrl[:, :7] = q[:, :7] + rng.uniform(-0.1, 0.1, (N, 7)); ctrl[:, 7] = 255 * 0.5
    b.qpos.copy_(torch.as_tensor(q)); b.ctrl.copy_(torch.as_tensor(ctrl))
    b.run(_lib.STEP_K, k=170)
    gq, gv = b.qpos.cpu().numpy(), b.qvel.cpu().numpy()
    m = O.Model(M)
    for e in range(N):
        d = O.Data(m)
        d.qpos[:] = q[e]; d.ctrl[:] = ctrl[e]
        d.step(170)
        assert np.abs(gq[e] - d.qpos).max() < 1e-9
        assert np.abs(gv[e] - d.qvel).max() < 1e-8


def test_floor_collision_contacts_exact(setup):
    """Arm driven into the floor (the reference's own collision test pose, python/tests/test_sim_envs.py:347-360):
    north_star's bit-exact contact-pair indexing. The exported contact list (count, geom ids in mjModel numbering,
    order) equals the oracle's mjData.contact at every sample and the committed floor_pairs fixture; contact geometry
    agrees to 1e-7, state to 1e-6; SimRobotState.collision / SimGripperState.collision as sampled by the condition
    callbacks (SimRobot.cpp:172-182, SimGripper.cpp:108-130) are identical."""
    import os
    M, dm, _lib, batch = setup
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_vectors.npz"))
    N = 4
    b = batch.Batch(dm, N)
    b.enable_contact_export(cap=8)
    m, s = H.oracle_sim(M)
    tgt = np.array([0, 1.78, 0, -1.45, 0, 0, 0.0])
    s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    b.run(_lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K, k=1)
    s.set_joint_position(tgt)
    b.run(_lib.SET_JOINTS, act_joints=torch.as_tensor(np.tile(tgt, (N, 1)), device=b.dev))
    hits = 0
    for it in range(90):
        s.step(5)
        b.run(_lib.STEP_K, k=5)
        ncon = int(b.si[0, 14].item())
        assert ncon == int(s.data.ncon[0]) == int(b.contact_n[0]), it
        cg = b.contact_geom.cpu().numpy()
        ref_pairs = s.data.int("contact_geom").reshape(-1, 2)
        for e in range(N):
            assert np.array_equal(cg[e, :ncon], ref_pairs), (it, e)
            assert (cg[e, ncon:] == -1).all()
        assert np.array_equal(cg[0, :6], G["floor_pairs"][it]), it
        if ncon:
            hits += 1
            ref = s.data.real("contact_real").reshape(-1, 7)
            assert np.abs(b.contact_real[0, :ncon].cpu().numpy() - ref).max() < 1e-7, it
        assert np.abs(b.qpos[0].cpu().numpy() - s.data.qpos).max() < 1e-6, it
        assert np.abs(b.qvel[0].cpu().numpy() - s.data.qvel).max() < 1e-4, it
    assert hits > 10
    # the condition callbacks sample the collision flags; a hit ends step_until_convergence early (sim.cpp:52-55)
    s.step_until_convergence()
    b.run(_lib.STEP_CONV | _lib.OBS, max_convergence_steps=500, want_obs=True)
    si, info = b.si.cpu().numpy(), b.info.cpu().numpy()
    rs, gs = s.robot_state(), s.gripper_state()
    assert rs["collision"]
    for e in range(N):
        assert bool(si[e, 1]) == rs["collision"] and bool(si[e, 5]) == gs["collision"]
        assert bool(si[e, 2]) == rs["is_moving"] and bool(si[e, 3]) == rs["is_arrived"]
        assert bool(si[e, 6]) == s.is_converged() and int(si[e, 7]) == s.convergence_steps()
        assert bool(info[e, 0]) == (rs["collision"] or gs["collision"]) and bool(info[e, 5]) == rs["collision"]
        assert bool(info[e, 6]) == gs["collision"]
        assert abs(b.obs[e, 21].item() - s.gripper_get_normalized_width()) < 1e-6


def test_step_until_convergence_parity(setup):
    M, dm, _lib, batch = setup
    N = 8
    rng = np.random.default_rng(3)
    b = batch.Batch(dm, N)
    b.run(_lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K, k=1)
    tg = H.Q_HOME + rng.uniform(-0.3, 0.3, (N, 7))
    b.run(_lib.SET_JOINTS | _lib.STEP_CONV, max_convergence_steps=500, act_joints=torch.as_tensor(tg, device=b.dev))
    steps = b.si[:, 7].cpu().numpy(); conv = b.si[:, 6].cpu().numpy()
    for e in range(N):
        m, s = H.oracle_sim(M)
        s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
        s.set_joint_position(tg[e]); s.step_until_convergence()
        assert s.convergence_steps() == steps[e]
        assert s.is_converged() == bool(conv[e])
        assert np.abs(b.qpos[e].cpu().numpy() - s.data.qpos).max() < 1e-9


@pytest.mark.parametrize("lanes", ["8", "1"])
def test_ik_parity(setup, monkeypatch, lanes):
    """Batched Pin::inverse against the oracle: iteration counts exact, q to 1e-9, for both mappings of the solver --
    8 lanes per environment (batches up to 8192) and one environment per thread -- including unreachable targets
    (1000 iterations, no solution) and an orientation error of ~179.5 degrees (log3's near-pi branch)."""
    M, dm, _lib, batch = setup
    monkeypatch.setenv("RCSB_IK_LANES", lanes)
    N = 67  # not a multiple of the 4 environments a warp holds
    rng = np.random.default_rng(5)
    b = batch.Batch(dm, N)
    m = O.Model(M)
    site = O.robot_cfg(M).attachment_site
    poses = np.zeros((N, 7)); q0 = np.tile(H.Q_HOME, (N, 1))
    for e in range(N):
        qt = H.Q_HOME + rng.uniform(-0.3, 0.3, 7)
        poses[e] = O.ik_forward(m, site, 9, qt)
    poses[3, :3] = [2.0, 0, 0.5]                    # out of reach: runs to the iteration cap
    flip = O.pose_mul(poses[5], [0, 0, 0, np.sin(0.49875 * np.pi), 0, 0, np.cos(0.49875 * np.pi)])  # 179.55 deg about x
    poses[5] = flip
    q, ok, it = b.ik_inverse(torch.as_tensor(poses, device=b.dev), torch.as_tensor(q0, device=b.dev))
    q, ok, it = q.cpu().numpy(), ok.cpu().numpy(), it.cpu().numpy()
    assert not ok[3] and it[3] == 1000
    for e in range(N):
        qr, itr = O.ik_inverse(m, site, 9, poses[e], q0[e])
        assert (qr is not None) == bool(ok[e]), e
        assert itr == it[e], (e, itr, it[e])
        if qr is not None:
            assert np.abs(qr - q[e]).max() < 1e-9, e


def _floor_run(batch, _lib, dm, N=96):
    b = batch.Batch(dm, N)
    tgt = np.tile(np.array([0, 1.78, 0, -1.45, 0, 0, 0.0]), (N, 1))
    tgt[1::3] = H.Q_HOME + 0.1  # every third env stays clear of the floor
    b.run(_lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K, k=1)
    b.run(_lib.SET_JOINTS, act_joints=torch.as_tensor(tgt, device=b.dev))
    for it in range(60):
        b.run(_lib.STEP_K | _lib.OBS, k=7, want_obs=True)
    b.run(_lib.STEP_CONV | _lib.OBS, max_convergence_steps=120, want_obs=True)
    torch.cuda.synchronize()
    return b, [t.cpu().numpy().copy() for t in (b.sr, b.sd, b.si, b.obs, b.info)]


def test_reduced_layout_hand_over_and_kernel_variants_bit_exact(setup, monkeypatch):
    """(1) envs that outgrow the reduced workspace layout are finished by the full-capacity launch: identical, bit for
    bit, to a model without a reduced layout; (2) the shape-specialised kernels equal the generic kernel bit for bit."""
    M, dm, _lib, batch = setup
    b0, ref = _floor_run(batch, _lib, dm)
    occ = b0.occupancy()
    assert occ["variant"] == "fr3_reduced" and occ["variant_full"] == "fr3_full", occ
    assert int(ref[2][:, 14].max()) >= 1, "floor contacts expected"
    assert not ref[2][:, 20].any()  # RCSB_I_RESUME cleared
    dm_full = batch.DeviceModel(M, H.robot_ns(), H.gripper_ns(), fast_maxcon=0)
    b1, full = _floor_run(batch, _lib, dm_full)
    assert b1.occupancy()["warps_per_cta"] < occ["warps_per_cta"]
    for a, c in zip(ref, full):
        assert np.array_equal(a, c)
    monkeypatch.setenv("RCSB_VARIANT", "generic")
    b2, gen = _floor_run(batch, _lib, dm)
    assert b2.occupancy()["variant"] == "generic"
    for a, c in zip(ref, gen):
        assert np.array_equal(a, c)


def test_xarm7_scene_parity(setup):
    """xarm7_empty_world through the generic kernel variant: friction-loss rows on every dof, pyramidal cones, a cylinder
    in the collision set, no gripper."""
    _, _, _lib, batch = setup
    M = H.scene("xarm7_empty_world")
    dm = batch.DeviceModel(M, H.xarm_robot_ns(), None)
    N = 64
    b = batch.Batch(dm, N)
    assert b.occupancy()["variant"] == "generic"
    rng = np.random.default_rng(11)
    q = H.XARM_Q_HOME + rng.uniform(-0.3, 0.3, (N, 7))
    v = rng.uniform(-0.5, 0.5, (N, 7)); v[::3] = 0
    ctrl = q + rng.uniform(-0.1, 0.1, (N, 7))
    b.qpos.copy_(torch.as_tensor(q)); b.qvel.copy_(torch.as_tensor(v)); b.ctrl.copy_(torch.as_tensor(ctrl))
    m = O.Model(M)
    for it in range(6):
        b.run(_lib.STEP_K, k=10)
        gq, gv = b.qpos.cpu().numpy(), b.qvel.cpu().numpy()
        for i in range(0, N, 7):
            d = O.Data(m); d.qpos[:] = q[i]; d.qvel[:] = v[i]; d.ctrl[:] = ctrl[i]
            d.step(10 * (it + 1))
            assert np.abs(gq[i] - d.qpos).max() < 1e-9, (it, i)
            assert np.abs(gv[i] - d.qvel).max() < 1e-7, (it, i)
    # the reference-style env surface on the second robot type (examples/xarm7/xarm7_env_joint_control.py)
    import rcs_b200
    from rcs_b200 import common, sim
    from rcs_b200.envs.base import ControlMode
    from rcs_b200.envs.creators import SimEnvCreator
    cfg = sim.SimRobotConfig()
    cfg.actuators = [f"act{i}" for i in range(1, 8)]; cfg.joints = [f"joint{i}" for i in range(1, 8)]
    cfg.base = "base"; cfg.robot_type = common.RobotType.XArm7; cfg.attachment_site = "attachment_site"
    cfg.arm_collision_geoms = []
    cfg.mjcf_scene_path = rcs_b200.scenes["xarm7_empty_world"].mjb
    cfg.kinematic_model_path = rcs_b200.scenes["xarm7_empty_world"].mjcf_robot
    env = SimEnvCreator()(ControlMode.JOINTS, cfg, gripper_cfg=None, max_relative_movement=float(np.deg2rad(5)),
                          sim_cfg=sim.SimConfig(async_control=True), num_envs=32)
    obs0, _ = env.reset()
    q0 = obs0["joints"].clone()  # observations are views of the batch's packed buffer
    assert np.allclose(q0.cpu().numpy(), H.XARM_Q_HOME, atol=2e-3)
    for t in range(5):
        obs, _, _, trunc, info = env.step(env.action_space.sample())
    assert bool(info["ik_success"].all()) and not bool(info["collision"].any())
    assert float((obs["joints"] - q0).abs().max()) > 1e-3


def test_results_do_not_depend_on_batch_size_or_mask(setup):
    """Size-independent properties at BASELINE sizes: environment e's trajectory is bit-identical whether it runs alone, in
    a ragged batch (N not a multiple of the warps per CTA) or in a 4096+ batch spread over every SM (second round of
    the static env->warp map); environments excluded by the mask keep every bit of their state."""
    M, dm, _lib, batch = setup
    T = 3
    sizes = (1, 29, 4097)
    acts = H.workload_actions(max(sizes), T, seed=3)
    reset = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K | _lib.OBS
    step = _lib.ACT_JOINTS_REL | _lib.ACT_GRIPPER_BIN | _lib.STEP_K | _lib.OBS
    outs = {}
    for n in sizes:
        b = batch.Batch(dm, n)
        b.run(reset, k=1, want_obs=True)
        for t in range(T):
            b.run(step, k=17, act_joints=torch.as_tensor(acts[:n, t, :7].copy(), device=b.dev),
                  act_gripper=torch.as_tensor(acts[:n, t, 7].copy(), device=b.dev), max_mov=np.deg2rad(5), jlow=H.JLOW,
                  jhigh=H.JHIGH, want_obs=True)
        torch.cuda.synchronize()
        outs[n] = (b.sr.cpu().numpy().copy(), b.obs.cpu().numpy().copy(), b.si.cpu().numpy().copy())
    for n in sizes[:-1]:
        for a, c in zip(outs[n], outs[sizes[-1]]):
            assert np.array_equal(a, c[:n]), n
    # mask: only every third environment steps
    n = 96
    b = batch.Batch(dm, n)
    b.run(reset, k=1, want_obs=True)
    before = [t.clone() for t in (b.sr, b.sd, b.si)]
    mask = torch.zeros(n, dtype=torch.uint8, device=b.dev); mask[::3] = 1
    b.run(step, k=17, act_joints=torch.as_tensor(acts[:n, 0, :7].copy(), device=b.dev),
          act_gripper=torch.as_tensor(acts[:n, 0, 7].copy(), device=b.dev), mask=mask, max_mov=np.deg2rad(5), jlow=H.JLOW,
          jhigh=H.JHIGH, want_obs=True)
    torch.cuda.synchronize()
    keep = (mask == 0).cpu().numpy()
    for x0, x1 in zip(before, (b.sr, b.sd, b.si)):
        assert np.array_equal(x0.cpu().numpy()[keep], x1.cpu().numpy()[keep])
    assert not np.array_equal(before[0].cpu().numpy()[::3], b.sr.cpu().numpy()[::3])


def test_kernels_match_committed_step_vectors(setup):
    """tests/golden/step_vectors.npz (see tests/golden/make_fixtures.py) through the C ABI, all three scenes; the contact
    counts along the committed floor-collision trajectory are exact."""
    import os
    _, _, _lib, batch = setup
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_vectors.npz"))
    for ci in range(len([k for k in G.files if k.endswith("_scene")])):
        scene, k = str(G[f"c{ci}_scene"]), int(G[f"c{ci}_k"])
        M = H.scene(scene)
        rc, gc = (H.xarm_robot_ns(), None) if scene.startswith("xarm7") else (H.robot_ns(), H.gripper_ns())
        dm = batch.DeviceModel(M, rc, gc)
        q = G[f"c{ci}_qpos"]
        b = batch.Batch(dm, len(q))
        b.qpos.copy_(torch.as_tensor(q)); b.qvel.copy_(torch.as_tensor(G[f"c{ci}_qvel"])); b.ctrl.copy_(torch.as_tensor(G[f"c{ci}_ctrl"]))
        b.run(_lib.STEP_K, k=k)
        tol_q, tol_v = (1e-12, 1e-10) if k == 1 else (1e-8, 1e-6)
        assert np.abs(b.qpos.cpu().numpy() - G[f"c{ci}_qpos_out"]).max() < tol_q, scene
        assert np.abs(b.qvel.cpu().numpy() - G[f"c{ci}_qvel_out"]).max() < tol_v, scene
        assert np.array_equal(b.si[:, 14].cpu().numpy(), G[f"c{ci}_ncon_out"]), scene
    dm = setup[1]
    b = batch.Batch(dm, 1)
    b.run(_lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K, k=1)
    b.run(_lib.SET_JOINTS, act_joints=torch.as_tensor(np.array([[0, 1.78, 0, -1.45, 0, 0, 0.0]]), device=b.dev))
    for it in range(len(G["floor_ncon"])):
        b.run(_lib.STEP_K, k=5)
        assert int(b.si[0, 14]) == int(G["floor_ncon"][it]), it
        assert np.abs(b.qpos[0].cpu().numpy() - G["floor_qpos"][it]).max() < 1e-6, it


def test_grasp_and_lift_parity(setup):
    """Config C3's defining behaviour on the GPU: close the gripper on the cube of fr3_simple_pick_up and raise it 12 cm
    (>= 2000 physics steps). Finger pads and cube are boxes: box-box face clipping gives 4 points per small pad (36
    contacts with the 4 floor points), elliptic cones + noslip, the reduced layout (4 resting contacts) hands the grasp
    over to the full layout. Against the oracle at every sample: contact count, geom pair ids and order exact, state to
    1e-6, gripper width / is_grasped / collision flags equal; nothing may be dropped (RCSB_I_WARN == 0)."""
    _, _, _lib, batch = setup
    M = H.scene("fr3_simple_pick_up")
    dm = batch.DeviceModel(M, H.robot_ns(H.FRANKA_HAND_TCP), H.gripper_ns())
    N = 3
    b = batch.Batch(dm, N)
    occ = b.occupancy()
    assert occ["variant"] == "fr3_pickup", occ
    b.enable_contact_export(cap=40)
    m, s = H.oracle_sim(M, tcp=H.FRANKA_HAND_TCP)
    reset = _lib.GRIPPER_RESET | _lib.SIM_RESET | _lib.ROBOT_RESET | _lib.ENV_RESET_FLAGS | _lib.STEP_K
    b.run(reset, k=1); s.gripper_reset(); s.reset(); s.robot_reset(); s.step(1)
    peak, worst = 0, 0.0
    for qt, w, n in H.grasp_and_lift_script(M):
        b.run(_lib.SET_JOINTS | _lib.SET_GRIPPER, act_joints=torch.as_tensor(np.tile(qt, (N, 1)), device=b.dev),
              act_gripper=torch.full((N,), w, dtype=torch.float64, device=b.dev))
        s.set_joint_position(qt); s.gripper_set_normalized_width(w)
        for _ in range(n // 50):
            b.run(_lib.STEP_K | _lib.OBS, k=50, want_obs=True); s.step(50)
            ncon = int(s.data.ncon[0])
            peak = max(peak, ncon)
            cn, cg = b.contact_n.cpu().numpy(), b.contact_geom.cpu().numpy()
            q, si, obs, info = b.qpos.cpu().numpy(), b.si.cpu().numpy(), b.obs.cpu().numpy(), b.info.cpu().numpy()
            ref_pairs = s.data.int("contact_geom").reshape(-1, 2)
            for e in range(N):
                assert cn[e] == ncon == si[e, 14]
                assert np.array_equal(cg[e, :ncon], ref_pairs)
                assert si[e, 17] == 0
                assert abs(obs[e, 21] - s.gripper_get_normalized_width()) < 1e-6
                gw = s.gripper_get_normalized_width()
                assert bool(info[e, 3]) == (0.01 < gw < 0.99)  # info["is_grasped"], envs/sim.py:130
            worst = max(worst, np.abs(q - s.data.qpos).max())
            assert np.abs(q - s.data.qpos).max() < 1e-6
    print("grasp: peak ncon", peak, "worst |dq|", worst)
    assert peak >= 30
    assert s.data.qpos[11] > 0.12 and (b.qpos[:, 11] > 0.12).all()
    assert 0.3 < s.gripper_get_normalized_width() < 0.5 and bool(b.info[0, 3])  # held open by the 32 mm cube


def test_xarm7_tabletop_parity_and_arm_brick_push(setup):
    """Synthetic config C4 (tools/scenes/xarm7_tabletop.xml) on the GPU: friction-loss rows on all 7 joints, the brick
    resting on the table through 4 box-box points as pyramidal cones, reduced layout + hand-over; then the arm is driven
    into the brick (mesh-box contacts couple the two kinematic trees: dense solver path). Contact pairs exact, state
    1e-6 against the oracle."""
    _, _, _lib, batch = setup
    M = H.scene("xarm7_tabletop")
    dm = batch.DeviceModel(M, H.xarm_robot_ns(), None)
    N = 8
    b = batch.Batch(dm, N)
    b.enable_contact_export(cap=24)
    rng = np.random.default_rng(7)
    q0 = np.tile(M["qpos0"], (N, 1)).astype(float)
    q0[:, :7] = H.XARM_Q_HOME + rng.uniform(-0.2, 0.2, (N, 7))
    q0[:, 7:9] += rng.uniform(-0.03, 0.03, (N, 2))
    ctrl = q0[:, :7] + rng.uniform(-0.1, 0.1, (N, 7))
    # environment 0 reaches down to the brick: joint targets found by the oracle's IK for a point just beside the brick
    m = O.Model(M)
    site = M["site_names"].index("attachment_site")
    sol, _ = O.ik_inverse(m, site, 7, [q0[0, 7] - 0.05, q0[0, 8], 0.225, 1, 0, 0, 0], H.XARM_Q_HOME)
    assert sol is not None
    sol2, _ = O.ik_inverse(m, site, 7, [q0[0, 7] + 0.08, q0[0, 8], 0.225, 1, 0, 0, 0], sol)
    assert sol2 is not None
    q0[0, :7] = sol[:7]; ctrl[0] = sol2[:7]
    b.qpos.copy_(torch.as_tensor(q0)); b.ctrl.copy_(torch.as_tensor(ctrl))
    ds = []
    for i in range(N):
        d = O.Data(m); d.qpos[:] = q0[i]; d.ctrl[:] = ctrl[i]; ds.append(d)
    coupled = 0
    names = M["geom_names"]
    for it in range(30):
        b.run(_lib.STEP_K, k=10)
        cn, cg, q, si = b.contact_n.cpu().numpy(), b.contact_geom.cpu().numpy(), b.qpos.cpu().numpy(), b.si.cpu().numpy()
        for i, d in enumerate(ds):
            d.step(10)
            n = int(d.ncon[0])
            ref = d.int("contact_geom").reshape(-1, 2)
            assert cn[i] == n and si[i, 17] == 0, (it, i)
            assert np.array_equal(cg[i, :n], ref), (it, i)
            assert np.abs(q[i] - d.qpos).max() < 1e-6, (it, i)
            coupled += any(("duplo" in names[g1]) != ("duplo" in names[g2]) and not {names[g1], names[g2]} & {"table", "floor"}
                           for g1, g2 in ref)
    assert coupled > 0, "the arm of environment 0 should have touched the brick"
    assert np.abs(q[0, 7:9] - q0[0, 7:9]).max() > 5e-3  # and pushed it
