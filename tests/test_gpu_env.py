"""Port of the reference's behavioural tests (/root/reference/python/tests/test_sim_envs.py) against the mirrored
API on the CUDA backend, for num_envs = 1 (reference semantics) and num_envs > 1 (batched extension)."""
import numpy as np
import pytest
import torch

import helpers  # noqa: F401

pytestmark = pytest.mark.gpu


def _mk(control_mode, num_envs=1, gripper=True, max_rel=None, async_control=False):
    from rcs_b200 import sim
    from rcs_b200.envs.creators import SimEnvCreator
    from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
    cfg = default_sim_robot_cfg("fr3_empty_world")
    return SimEnvCreator()(control_mode, cfg, gripper_cfg=default_sim_gripper_cfg() if gripper else None,
                           sim_cfg=sim.SimConfig(async_control=async_control), max_relative_movement=max_rel,
                           num_envs=num_envs)


def test_double_reset_and_zero_action_joints():  # test_sim_envs.py:304-331
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.JOINTS, gripper=False)
    env.reset()
    obs0, _ = env.reset()
    obs, _, _, _, info = env.step({"joints": obs0["joints"].clone()})
    assert bool(info["ik_success"][0])
    assert np.allclose(obs["joints"].cpu().numpy(), obs0["joints"].cpu().numpy(), atol=0.01, rtol=0)


def test_non_zero_action_joints():  # test_sim_envs.py:333-345
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.JOINTS, gripper=False)
    obs0, _ = env.reset()
    tgt = obs0["joints"] + torch.tensor([[0.1, 0.1, 0.1, 0.1, -0.1, -0.1, 0.1]], dtype=torch.float64, device=obs0["joints"].device)
    obs, _, _, _, info = env.step({"joints": tgt})
    assert bool(info["ik_success"][0])
    assert np.allclose(obs["joints"].cpu().numpy(), tgt.cpu().numpy(), atol=0.01, rtol=0)


def test_collision_joints():  # test_sim_envs.py:347-360
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.JOINTS, gripper=True)
    env.reset()
    act = {"joints": torch.tensor([[0, 1.78, 0, -1.45, 0, 0, 0]], dtype=torch.float64), "gripper": torch.tensor([1.0], dtype=torch.float64)}
    _, _, _, _, info = env.step(act)
    assert bool(info["collision"][0]) and bool(info["ik_success"][0])


def test_cartesian_tquat_move_x():  # test_sim_envs.py:202-221: +0.2 m in x reached within is_close(0.1 rad, 0.01 m)
    from rcs_b200 import common
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.CARTESIAN_TQuat, gripper=False)
    obs0, _ = env.reset()
    t = obs0["tquat"].clone()
    t[:, 0] += 0.2
    obs, _, _, trunc, info = env.step({"tquat": t})
    assert bool(info["ik_success"][0])
    a, b = obs["tquat"][0].cpu().numpy(), t[0].cpu().numpy()
    assert common.Pose(translation=a[:3], quaternion=a[3:]).is_close(common.Pose(translation=b[:3], quaternion=b[3:]), eps_r=0.1, eps_t=0.01)


def test_cartesian_collision_and_ik_failure():  # test_sim_envs.py:252-271 (floor) and SimRobot.cpp:149-154
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.CARTESIAN_TQuat, gripper=True)
    obs0, _ = env.reset()
    t = obs0["tquat"].clone()
    t[:, 2] = -0.05
    _, _, _, _, info = env.step({"tquat": t, "gripper": torch.tensor([1.0], dtype=torch.float64)})
    assert bool(info["collision"][0]) and bool(info["ik_success"][0])
    env2 = _mk(ControlMode.CARTESIAN_TQuat, gripper=False)
    obs0, _ = env2.reset()
    far = obs0["tquat"].clone(); far[:, 0] = 2.0
    obs, _, _, trunc, info = env2.step({"tquat": far})
    assert not bool(info["ik_success"][0]) and bool(trunc[0])
    assert np.allclose(obs["joints"].cpu().numpy(), obs0["joints"].cpu().numpy(), atol=1e-3)  # ctrl untouched


def test_direct_api_matches_reference_semantics():
    """examples/fr3/fr3_direct_control.py:62-200 call pattern: set_cartesian_position -> step_until_convergence."""
    from rcs_b200 import common, sim
    from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
    cfg = default_sim_robot_cfg("fr3_empty_world")
    simulation = sim.Sim(cfg.mjcf_scene_path)
    ik = sim.Pin(cfg.kinematic_model_path, cfg.attachment_site, urdf=False)
    robot = sim.SimRobot(simulation, ik, cfg)
    gripper = sim.SimGripper(simulation, default_sim_gripper_cfg())
    simulation.reset(); robot.reset(); simulation.step(1)
    p0 = robot.get_cartesian_position()
    assert isinstance(p0, common.Pose)
    robot.set_cartesian_position(p0 * common.Pose(translation=np.array([0.05, 0, 0])))
    simulation.step_until_convergence()
    st = robot.get_state()
    assert st.ik_success and not st.collision
    p1 = robot.get_cartesian_position()
    assert abs((p1.translation() - p0.translation())[2]) < 0.02
    with pytest.raises(ValueError):
        gripper.set_normalized_width(1.5)
    gripper.open(); simulation.step(200)
    assert gripper.get_normalized_width() > 0.5
    with pytest.raises(RuntimeError, match="No geom named"):
        bad = default_sim_robot_cfg("fr3_empty_world"); bad.arm_collision_geoms = ["nope"]
        sim.SimRobot(sim.Sim(cfg.mjcf_scene_path), None, bad)
    q = ik.inverse(p0, robot.get_joint_position())
    assert q is not None and q.shape == (9,)
    assert ik.forward(q[:7]).is_close(p0, 1e-3, 1e-3)


def test_batched_relative_joint_env_async():
    """The benchmark wiring (examples/fr3/fr3_env_joint_control.py:34-41) with num_envs = 512, async 30 Hz."""
    from rcs_b200.envs.base import ControlMode
    env = _mk(ControlMode.JOINTS, num_envs=512, gripper=True, max_rel=float(np.deg2rad(5)), async_control=True)
    obs, _ = env.reset()
    assert obs["joints"].shape == (512, 7) and obs["tquat"].shape == (512, 7) and obs["xyzrpy"].shape == (512, 6)
    q_prev = obs["joints"].clone()
    for _ in range(5):
        act = env.action_space.sample()
        obs, rew, term, trunc, info = env.step(act)
        assert float((obs["joints"] - q_prev).abs().max()) < np.deg2rad(5) + 0.05
        q_prev = obs["joints"].clone()
    assert obs["gripper"].shape == (512,) and info["gripper_width"].shape == (512,)
    assert not bool(term.any()) and float(rew.abs().max()) == 0.0
    h_a = torch.zeros((512, 8), dtype=torch.float64).pin_memory(); h_a[:, 7] = 1
    q_dev = obs["joints"].clone()
    ho = env.step_host(h_a)                                    # zero relative action, gripper open, through host buffers
    assert ho.shape == (512, 30) and np.isfinite(ho.numpy()).all()
    oh, ih, trunc_h = env.unpack(ho)
    assert float((oh["joints"] - q_dev.cpu()).abs().max()) < 0.05 and bool(ih["ik_success"].all()) and not bool(trunc_h.any())
    assert torch.equal(oh["joints"], env.sim.batch.qpos[:, :7].cpu())  # the host block is what the device state says


def test_step_host_pinned_and_pageable_paths_agree():
    """rcsb_env_step_host with page-locked buffers (the kernel reads the actions and writes the observation rows through
    the mapped host pointers) and with pageable buffers (staged copies): bit-identical rows, equal to the device-resident
    step_packed on the same actions."""
    from rcs_b200.envs.base import ControlMode
    N = 300
    envs = [_mk(ControlMode.JOINTS, num_envs=N, gripper=True, max_rel=np.deg2rad(5), async_control=True) for _ in range(3)]
    for e in envs:
        e.reset()
    g = torch.Generator().manual_seed(3)
    for t in range(4):
        a = torch.cat([(torch.rand((N, 7), dtype=torch.float64, generator=g) * 2 - 1) * np.deg2rad(5),
                       torch.randint(0, 2, (N, 1), generator=g).to(torch.float64)], dim=1).contiguous()
        pinned = envs[0].step_host(a.clone().pin_memory()).clone()
        b1 = envs[1].sim.batch
        ops, cfg = envs[1]._step_ops()
        pageable = torch.zeros((N, 30), dtype=torch.float64)
        assert not a.is_pinned() and not pageable.is_pinned()
        b1.step_host(ops, envs[1]._substeps(), cfg.max_convergence_steps, a, float(envs[1].max_mov), envs[1].jlow, envs[1].jhigh, pageable)
        dev = envs[2].step_packed({"joints": a[:, :7].cuda(), "gripper": a[:, 7].cuda()}).cpu()
        assert torch.equal(pinned, pageable)
        assert torch.equal(pinned, dev)


@pytest.mark.parametrize("mode_name", ["CARTESIAN_TRPY", "CARTESIAN_TQuat"])
def test_relative_cartesian_actions_match_reference_math(mode_name):
    """RelativeActionSpace (base.py:490-578, LAST_STEP) on the device: offset clipping (translation length, rotation
    angle by slerp), composition with the current pose, workspace clip, then SimRobot::set_cartesian_position. The
    expected targets are rebuilt with the host Pose class (pinned by tests/test_pose.py) and the batched IK entry point
    (pinned against the oracle in test_gpu_parity.py)."""
    from rcs_b200 import common
    from rcs_b200.envs.base import ControlMode
    N = 48
    mode = getattr(ControlMode, mode_name)
    max_t, max_r = 0.2, np.deg2rad(45)
    env = _mk(mode, num_envs=N, gripper=False, max_rel=(max_t, max_r), async_control=True)
    obs, _ = env.reset()
    b = env.sim.batch
    rng = np.random.default_rng(5)
    xyz = rng.uniform(-0.03, 0.03, (N, 3)); xyz[::4] *= 10          # every fourth offset exceeds the 0.2 m cap
    rpy = rng.uniform(-0.1, 0.1, (N, 3)); rpy[1::4] *= 12            # some exceed the 45 degree cap
    if mode == ControlMode.CARTESIAN_TRPY:
        a = np.concatenate([xyz, rpy], axis=1); key = "xyzrpy"
    else:
        quat = np.stack([common.Pose(translation=np.zeros(3), rpy_vector=r).rotation_q() for r in rpy])
        a = np.concatenate([xyz, quat], axis=1); key = "tquat"
    tq0 = obs["tquat"].cpu().numpy()
    q_now = b.qpos[:, :7].clone().contiguous()
    poses = np.zeros((N, 7))
    for e in range(N):
        origin = common.Pose(translation=tq0[e, :3], quaternion=tq0[e, 3:])
        if mode == ControlMode.CARTESIAN_TRPY:
            off = common.Pose(translation=a[e, :3], rpy_vector=a[e, 3:])
        else:
            off = common.Pose(translation=a[e, :3], quaternion=a[e, 3:])
        off = off.limit_translation_length(max_t).limit_rotation_angle(max_r)
        t = np.clip(origin.translation() + off.translation(), [-0.855, -0.855, 0], [0.855, 0.855, 1.188])
        if mode == ControlMode.CARTESIAN_TRPY:
            tgt = common.Pose(translation=t, rpy_vector=(off * origin).rotation_rpy().as_vector())
        else:
            tgt = common.Pose(translation=t, quaternion=(off * origin).rotation_q())
        poses[e, :3] = tgt.translation(); poses[e, 3:] = tgt.rotation_q()
    q_exp, ok, _ = b.ik_inverse(torch.as_tensor(poses, device=b.dev), q_now)
    assert int(ok.sum()) > N // 2
    obs1, _, _, _, info = env.step({key: torch.as_tensor(a, device=b.dev)})
    okn = ok.cpu().numpy().astype(bool)
    assert np.array_equal(info["ik_success"].cpu().numpy(), okn)
    ctrl = b.ctrl[:, :7].cpu().numpy()
    assert np.abs(ctrl[okn] - q_exp.cpu().numpy()[okn, :7]).max() < 1e-9
    # RobotEnv.step dedupe (base.py:268-287): an absolute action equal to the previous one (atol 1e-3) sends no command.
    # With LAST_STEP a repeated zero offset yields the current pose twice once the robot has stopped moving.
    zero = np.zeros_like(a)
    if mode == ControlMode.CARTESIAN_TQuat:
        zero[:, 6] = 1
    za = torch.as_tensor(zero, device=b.dev)
    for _ in range(40):
        env.step({key: za})
    ctrl_before = b.ctrl[:, :7].clone()
    env.step({key: za})
    assert (b.ctrl[:, :7] - ctrl_before).abs().max() < 1e-3


def test_pick_up_task_env_random_cube_and_reward():
    """FR3SimplePickUpSimEnvCreator (creators.py:192-224) with RandomCubePos and PickCubeSuccessWrapper
    (envs/sim.py:359-431) on the batched backend: cube re-placed in the +-0.1 m square around the iso pose, reward in
    [0, 1], success only with the cube above 0.15 + 0.852 m and the gripper closed."""
    from rcs_b200.envs.creators import FR3SimplePickUpSimEnvCreator
    N = 32
    env = FR3SimplePickUpSimEnvCreator()(num_envs=N)
    obs, _ = env.reset()
    b = env.sim.batch
    box = b.qpos[:, 9:16].cpu().numpy()
    iso = env.unwrapped.robot.to_pose_in_world_coordinates(
        __import__("rcs_b200").common.Pose(translation=np.array([0.498, 0.0, 0.226]))).translation()
    assert np.all(np.abs(box[:, 0] - iso[0]) <= 0.1) and np.all(np.abs(box[:, 1] - iso[1]) <= 0.1)
    assert np.allclose(box[:, 2], 0.0144) and np.all(np.abs(box[:, 3]) <= 1) and np.allclose(box[:, 6], 1)
    assert len(np.unique(box[:, 0])) > N // 2  # per-env draws
    gen = torch.Generator(device=b.dev).manual_seed(0)
    for t in range(4):
        a = (torch.rand((N, 6), dtype=torch.float64, device=b.dev, generator=gen) * 2 - 1) * torch.tensor([0.01] * 3 + [0.05] * 3, device=b.dev)
        g = torch.randint(0, 2, (N,), device=b.dev, generator=gen).to(torch.float64)
        obs, reward, term, trunc, info = env.step({"xyzrpy": a, "gripper": g})
    assert reward.shape == (N,) and float(reward.min()) >= 0 and float(reward.max()) <= 1
    assert not bool(term.any()) and not bool(info["success"].any())
    assert bool(info["ik_success"].all())
    assert int(b.si[:, 14].min()) >= 1  # the cube rests on the floor: contacts in every env
    # lift the cube by hand above the success height and close the gripper
    b.qpos[:, 11] = 0.15 + 0.852 + 0.05
    b.qvel[:, 9:15] = 0
    obs, reward, term, trunc, info = env.step({"xyzrpy": torch.zeros((N, 6), dtype=torch.float64, device=b.dev),
                                               "gripper": torch.zeros((N,), dtype=torch.float64, device=b.dev)})
    assert bool(term.all()) and bool(info["success"].all()) and np.allclose(reward.cpu().numpy(), 1.0)


def test_checkpoint_resume_is_bit_exact_and_fullphysics_state_round_trips():
    """save_checkpoint / load_checkpoint resume bit-identically; get_state / set_state use the mjSTATE_FULLPHYSICS layout
    (time | qpos | qvel) of the reference's GUI bridge (src/sim/gui.h:20)."""
    from rcs_b200.envs.base import ControlMode
    N = 64
    env = _mk(ControlMode.JOINTS, num_envs=N, gripper=True, max_rel=float(np.deg2rad(5)), async_control=True)
    env.reset()
    gen = torch.Generator(device=env.sim.batch.dev).manual_seed(9)
    acts = [{"joints": (torch.rand((N, 7), dtype=torch.float64, device=env.sim.batch.dev, generator=gen) * 2 - 1) * np.deg2rad(5),
             "gripper": torch.randint(0, 2, (N,), device=env.sim.batch.dev, generator=gen).to(torch.float64)} for _ in range(6)]
    for a in acts[:3]:
        env.step(a)
    ck = env.sim.save_checkpoint()
    for a in acts[3:]:
        env.step(a)
    b = env.sim.batch
    first = [t.clone() for t in (b.sr, b.sd, b.si, b.obs)]
    env.sim.load_checkpoint(ck)
    for a in acts[3:]:
        env.step(a)
    for x, y in zip(first, (b.sr, b.sd, b.si, b.obs)):
        assert torch.equal(x, y)
    st = env.sim.get_state()
    assert st.shape == (N, 1 + 9 + 9) and torch.equal(st[:, 0], b.time) and torch.equal(st[:, 1:10], b.qpos)
    env2 = _mk(ControlMode.JOINTS, num_envs=N, gripper=True, max_rel=float(np.deg2rad(5)), async_control=True)
    env2.reset()
    env2.sim.set_state(st)
    assert torch.equal(env2.sim.get_state(), st)


def test_configured_origin_and_continuous_gripper():
    """RelativeTo.CONFIGURED_ORIGIN for joint control (base.py:479-488: offsets are relative to the pose at reset and may
    grow by at most max_mov per step) and the continuous GripperWrapper mode (base.py:710-735)."""
    from rcs_b200 import sim
    from rcs_b200.envs.base import ControlMode, RelativeTo
    from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
    from rcs_b200.envs.vector import SimVectorEnv
    N = 8
    cfg = default_sim_robot_cfg("fr3_empty_world")
    simulation = sim.Sim(cfg.mjcf_scene_path, sim.SimConfig(async_control=True), num_envs=N)
    robot = sim.SimRobot(simulation, sim.Pin(cfg.kinematic_model_path, cfg.attachment_site, urdf=False), cfg)
    gripper = sim.SimGripper(simulation, default_sim_gripper_cfg())
    mm = float(np.deg2rad(5))
    env = SimVectorEnv(simulation, robot, gripper, ControlMode.JOINTS, mm, RelativeTo.CONFIGURED_ORIGIN, binary_gripper=False)
    obs, _ = env.reset()
    b = simulation.batch
    origin = obs["joints"].clone()
    off = torch.tensor([0.2, -0.2, 0.01, 0.0, 0.0, 0.0, 0.03], dtype=torch.float64, device=b.dev).repeat(N, 1)
    width = torch.full((N,), 0.37, dtype=torch.float64, device=b.dev)
    for t in range(1, 4):  # the 0.2 rad offsets are approached in steps of max_mov, the small ones are reached at once
        obs, _, _, _, info = env.step({"joints": off, "gripper": width})
        exp = origin + torch.minimum(torch.maximum(off, torch.full_like(off, -t * mm)), torch.full_like(off, t * mm))
        assert torch.allclose(b.ctrl[:, :7], exp, atol=1e-12)
        assert torch.allclose(b.ctrl[:, 7], width * 255)
    # continuous mode: the observation is the measured normalised width, not the last command
    assert torch.equal(obs["gripper"], info["gripper_width"]) and float(obs["gripper"].min()) >= 0


def test_pin_inverse_default_tcp_offset_is_identity_with_a_configured_tcp():
    """Kinematics::inverse(pose, q0, tcp_offset = Identity) (rcs.cpp:289-300): with a robot configured with
    FrankaHandTCPOffset, get_ik().inverse(p, q0) must drive the attachment frame to p itself (forward(inverse(p)) == p),
    and inverse(p, q0, cfg_tcp) must equal what set_cartesian_position uses; tensor poses take the same path."""
    import rcs_b200
    from rcs_b200 import common, sim
    from helpers import O
    import helpers as H
    s = sim.Sim(rcs_b200.scenes["fr3_empty_world"].mjb, sim.SimConfig(async_control=True), num_envs=4)
    cfg = sim.SimRobotConfig(); cfg.add_id("0")
    cfg.tcp_offset = common.Pose(pose_matrix=common.FrankaHandTCPOffset())
    ik = sim.Pin()
    robot = sim.SimRobot(s, ik, cfg)
    M = H.scene()
    m = O.Model(M); site = O.robot_cfg(M).attachment_site
    qt = H.Q_HOME + np.array([0.1, -0.1, 0.05, 0.1, -0.05, 0.1, 0.0])
    flange = O.ik_forward(m, site, 9, qt)                       # pose of the attachment frame at qt
    goal = common.Pose(translation=flange[:3], quaternion=flange[3:])
    q = robot.get_ik().inverse(torch.as_tensor(np.tile(goal.as7(), (4, 1))), H.Q_HOME)[0][0].cpu().numpy()
    qr, _ = O.ik_inverse(m, site, 9, goal.as7(), H.Q_HOME)      # oracle, identity tcp
    assert np.abs(q - qr).max() < 1e-9
    assert robot.get_ik().forward(q[:7]).is_close(goal, 1e-3, 1e-3)
    # explicit tcp_offset equal to the configured one: the goal is the TCP pose
    tcp_goal = goal * cfg.tcp_offset
    q2 = robot.get_ik().inverse(torch.as_tensor(np.tile(tcp_goal.as7(), (4, 1))), H.Q_HOME, cfg.tcp_offset)[0][0].cpu().numpy()
    assert np.abs(q2 - qr).max() < 1e-7


@pytest.mark.parametrize("mode_name", ["CARTESIAN_TRPY", "CARTESIAN_TQuat"])
def test_cartesian_configured_origin_matches_reference_math(mode_name):
    """RelativeTo.CONFIGURED_ORIGIN for Cartesian control (base.py:443-467, 490-578) on the device: the origin is the pose
    at reset(); the offset may move by at most (max translation, max rotation) per step relative to the LAST clipped
    offset (pose_diff = action * last^-1, clipped, re-applied). Expected targets are rebuilt step by step with the host
    Pose class and solved with the batched IK entry point."""
    from rcs_b200 import common
    from rcs_b200.envs.base import ControlMode, RelativeTo
    from rcs_b200 import sim
    from rcs_b200.envs.creators import SimEnvCreator
    from rcs_b200.envs.utils import default_sim_robot_cfg
    N = 24
    mode = getattr(ControlMode, mode_name)
    max_t, max_r = 0.05, np.deg2rad(10)
    env = SimEnvCreator()(mode, default_sim_robot_cfg("fr3_empty_world"), gripper_cfg=None, sim_cfg=sim.SimConfig(async_control=True),
                          max_relative_movement=(max_t, max_r), relative_to=RelativeTo.CONFIGURED_ORIGIN, num_envs=N)
    obs, _ = env.reset()
    b = env.sim.batch
    tq0 = obs["tquat"].cpu().numpy().copy()
    origins = [common.Pose(translation=tq0[e, :3], quaternion=tq0[e, 3:]) for e in range(N)]
    last = [None] * N
    rng = np.random.default_rng(9)
    key = "xyzrpy" if mode == ControlMode.CARTESIAN_TRPY else "tquat"
    for step in range(4):
        xyz = rng.uniform(-0.08, 0.08, (N, 3))   # mostly beyond the 5 cm per-step cap
        rpy = rng.uniform(-0.3, 0.3, (N, 3))     # mostly beyond the 10 degree cap
        if mode == ControlMode.CARTESIAN_TRPY:
            a = np.concatenate([xyz, rpy], axis=1)
        else:
            a = np.concatenate([xyz, np.stack([common.Pose(rpy_vector=r).rotation_q() for r in rpy])], axis=1)
        poses = np.zeros((N, 7))
        for e in range(N):
            act = common.Pose(translation=a[e, :3], rpy_vector=a[e, 3:]) if mode == ControlMode.CARTESIAN_TRPY else \
                common.Pose(translation=a[e, :3], quaternion=a[e, 3:])
            if last[e] is None:
                off = act.limit_translation_length(max_t).limit_rotation_angle(max_r)
            else:
                diff = act * last[e].inverse()
                off = diff.limit_translation_length(max_t).limit_rotation_angle(max_r) * last[e]
            last[e] = off
            t = np.clip(origins[e].translation() + off.translation(), [-0.855, -0.855, 0], [0.855, 0.855, 1.188])
            if mode == ControlMode.CARTESIAN_TRPY:
                tgt = common.Pose(translation=t, rpy_vector=(off * origins[e]).rotation_rpy().as_vector())
            else:
                tgt = common.Pose(translation=t, quaternion=(off * origins[e]).rotation_q())
            poses[e, :3] = tgt.translation(); poses[e, 3:] = tgt.rotation_q()
        q_now = b.qpos[:, :7].clone().contiguous()
        q_exp, ok, _ = b.ik_inverse(torch.as_tensor(poses, device=b.dev), q_now)
        _, _, _, _, info = env.step({key: torch.as_tensor(a, device=b.dev)})
        okn = ok.cpu().numpy().astype(bool)
        assert okn.sum() > N // 2 and np.array_equal(info["ik_success"].cpu().numpy(), okn), step
        ctrl = b.ctrl[:, :7].cpu().numpy()
        assert np.abs(ctrl[okn] - q_exp.cpu().numpy()[okn, :7]).max() < 1e-8, step


def test_mixed_fleet_steps_groups_concurrently_and_matches_separate_envs():
    """FleetVectorEnv: an FR3 group and an xArm7 group stepped on their own streams give exactly
    the observations of the same envs stepped alone."""
    from rcs_b200 import sim, workloads as WL
    from rcs_b200.envs.base import ControlMode
    from rcs_b200.envs.creators import SimEnvCreator
    from rcs_b200.envs.fleet import FleetVectorEnv
    from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
    cfg = sim.SimConfig(async_control=True)
    mk = {"fr3": lambda: SimEnvCreator()(ControlMode.JOINTS, default_sim_robot_cfg("fr3_empty_world"), gripper_cfg=default_sim_gripper_cfg(),
                                         sim_cfg=cfg, max_relative_movement=float(np.deg2rad(5)), num_envs=96),
          "xarm7": lambda: SimEnvCreator()(ControlMode.JOINTS, WL.xarm7_robot_cfg(), gripper_cfg=None, sim_cfg=cfg,
                                           max_relative_movement=float(np.deg2rad(5)), num_envs=64)}
    fleet = FleetVectorEnv(mk)
    alone = {k: f() for k, f in mk.items()}
    obs, _ = fleet.reset()
    for k, e in alone.items():
        o, _ = e.reset()
        assert torch.equal(o["joints"], obs[k]["joints"])
    torch.manual_seed(0)
    for _ in range(4):
        acts = fleet.sample_actions()
        res = fleet.step(acts)
        torch.cuda.synchronize()
        for k, e in alone.items():
            o, _, _, _, info = e.step(acts[k])
            assert torch.equal(o["joints"], res[k][0]["joints"]) and torch.equal(o["tquat"], res[k][0]["tquat"])
            assert torch.equal(info["ik_success"], res[k][4]["ik_success"])
    assert res["fr3"][0]["joints"].shape == (96, 7) and res["xarm7"][0]["joints"].shape == (64, 7)


def test_collision_guard_vetoes_per_environment_and_multi_robot_wrapper():
    """CollisionGuard (envs/sim.py:156-287) on vector envs: the shadow env executes every action first; environments whose
    action drives the arm into the floor are vetoed -- masked out of the launch (state bit-identical), last observation
    returned with terminated = truncated = True -- while the others step. MultiRobotWrapper (base.py:310-355) steps a dict of
    envs and ORs the flags."""
    from rcs_b200.envs.base import ControlMode
    from rcs_b200.envs.guard import CollisionGuard, MultiRobotWrapper
    N = 12
    env = _mk(ControlMode.JOINTS, num_envs=N, gripper=False)            # sync control: step_until_convergence
    shadow = _mk(ControlMode.JOINTS, num_envs=N, gripper=False)
    guard = CollisionGuard(env, env.sim, shadow, check_home_collision=True)
    obs, _ = guard.reset()
    q0 = obs["joints"].clone()
    ok = q0 + 0.05
    obs, rew, term, trunc, info = guard.step({"joints": ok})
    assert not bool(term.any()) and not bool(info["guard_collision"].any())
    assert float((obs["joints"] - ok).abs().max()) < 0.01
    before = env.sim.batch.sr.clone()
    act = obs["joints"].clone() + 0.03
    floor = torch.tensor([0, 1.78, 0, -1.45, 0, 0, 0.0], dtype=torch.float64, device=act.device)  # test_sim_envs.py:347-360
    act[::3] = floor
    obs2, rew, term, trunc, info = guard.step({"joints": act})
    coll = info["guard_collision"].cpu().numpy()
    assert coll[::3].all() and not coll[1::3].any() and not coll[2::3].any()
    assert torch.equal(term.cpu(), torch.as_tensor(coll)) and bool(trunc[::3].all())
    assert torch.equal(env.sim.batch.sr[::3], before[::3]), "vetoed environments must not change at all"
    assert torch.equal(obs2["joints"][::3], obs["joints"][::3])        # they report their last observation
    assert float((obs2["joints"][1::3] - act[1::3]).abs().max()) < 0.01  # the others moved
    multi = MultiRobotWrapper({"a": _mk(ControlMode.JOINTS, num_envs=4, gripper=False, async_control=True),
                               "b": _mk(ControlMode.JOINTS, num_envs=4, gripper=False, async_control=True)})
    o, i = multi.reset()
    o, r, t, tr, i = multi.step({"a": {"joints": o["a"]["joints"] + 0.02}, "b": {"joints": o["b"]["joints"] - 0.02}})
    assert set(o) == {"a", "b"} and r.shape == (4,) and not bool(t.any()) and "truncated" in i["a"]
