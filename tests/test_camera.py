"""SimCameraSet depth frames: the CUDA ray-caster against the numpy oracle (oracle/depth_oracle.py) on
the same states -- uint16 millimetres, conventions of python/rcs/camera/sim.py:45-115. Integer output: pixels must be
equal; the two sides evaluate the same float64 expressions in a different order (FMA contraction, kinematics through
rotation matrices vs quaternions), so a pixel whose depth lies within rounding of a millimetre boundary or whose ray
grazes a silhouette may differ: at most 0.2 % of the pixels, and never by more than 1 mm away from silhouettes."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import O
from oracle import depth_oracle as D


def test_oracle_depth_conventions_closed_form():
    """A camera 2.08 m above the floor looking straight down (bird_eye_cam of fr3_simple_pick_up is nearly that): the
    centre pixel reads the camera height in mm; a synthetic box scene gives the box top; far plane where nothing is hit."""
    M = H.scene("fr3_simple_pick_up")
    m = O.Model(M); d = O.Data(m); d.qpos[:7] = H.Q_HOME; d.forward()
    cid = M["cam_names"].index("bird_eye_cam")
    img = D.render_depth(M, d.geom_xpos, d.geom_xmat, d.xpos, d.xmat, cid, 64, 48)
    assert img.dtype == np.uint16 and img.shape == (48, 64)
    h = M["cam_pos"][cid][2]
    assert abs(int(img[2, 2]) - 1000 * h) < 40        # a corner pixel sees the floor (slightly longer path is still the z depth)
    assert img.min() < 1000 * h - 300                 # the robot rises towards the camera
    side = D.render_depth(M, d.geom_xpos, d.geom_xmat, d.xpos, d.xmat, M["cam_names"].index("right_side"), 32, 32)
    assert side.max() == 50000                        # rays above the horizon hit nothing: far = 50 m x extent 1


@pytest.mark.gpu
@pytest.mark.parametrize("scene,cam,res", [("fr3_simple_pick_up", "bird_eye_cam", (96, 64)), ("fr3_simple_pick_up", "wrist_0", (64, 64)),
                                           ("fr3_simple_pick_up", "side_view", (80, 60)), ("fr3_empty_world", "wrist_0", (48, 48)),
                                           ("fr3_simple_pick_up", "bird_eye_cam", (136, 100))])  # 63 tiles: two blocks per image, ragged edges
def test_depth_kernel_matches_oracle(scene, cam, res, monkeypatch):
    """float rays (the default) on every case; the (96, 64) and (136, 100) cases also with double rays (RCSB_DEPTH_F64=1)"""
    for f64 in ((False, True) if res in ((96, 64), (136, 100)) else (False,)):
        if f64:
            monkeypatch.setenv("RCSB_DEPTH_F64", "1")
        else:
            monkeypatch.delenv("RCSB_DEPTH_F64", raising=False)
        _depth_case(scene, cam, res)


def _depth_case(scene, cam, res):
    import rcs_b200
    from rcs_b200 import sim
    from rcs_b200.camera import SimCameraConfig, SimCameraSet
    M = H.scene(scene)
    N = 5
    s = sim.Sim(rcs_b200.scenes[scene].mjb, sim.SimConfig(async_control=True), num_envs=N)
    cfg = sim.SimRobotConfig(); cfg.add_id("0")
    robot = sim.SimRobot(s, sim.Pin(), cfg)
    gcfg = sim.SimGripperConfig(); gcfg.add_id("0")
    sim.SimGripper(s, gcfg)
    rng = np.random.default_rng(3)
    q = np.tile(M["qpos0"], (N, 1)).astype(float)
    q[:, :7] = H.Q_HOME + rng.uniform(-0.4, 0.4, (N, 7)); q[:, 7] = q[:, 8] = rng.uniform(0, 0.04, N)
    if scene == "fr3_simple_pick_up":
        q[:, 9:11] += rng.uniform(-0.1, 0.1, (N, 2))
    s.batch.qpos.copy_(torch.as_tensor(q))
    W, Hh = res
    cams = SimCameraSet(s, {"c": SimCameraConfig(cam, 30, W, Hh)}, physical_units=True)
    before = s.batch.sr.clone()
    fs = cams.get_latest_frames()
    assert torch.equal(before, s.batch.sr), "rendering must not change the state"
    img = fs.frames["c"].camera.depth.data.cpu().numpy()
    assert img.shape == (N, Hh, W, 1) and img.dtype == np.uint16
    assert fs.frames["c"].camera.color is None
    m = O.Model(M)
    cid = M["cam_names"].index(cam)
    bad = tot = 0
    for e in range(N):
        d = O.Data(m); d.qpos[:] = q[e]; d.forward()
        ref = D.render_depth(M, d.geom_xpos, d.geom_xmat, d.xpos, d.xmat, cid, W, Hh).astype(np.int32)
        got = img[e, :, :, 0].astype(np.int32)
        diff = np.abs(got - ref)
        bad += int((diff > 0).sum()); tot += diff.size
        assert (diff > 1).mean() < 0.002, (e, (diff > 1).mean())   # silhouette pixels only
        assert got.min() < got.max()                                # the image shows something
        # extrinsics: world -> camera with z in front (camera/sim.py:109-119)
        E = fs.frames["c"].camera.depth.extrinsics[e].cpu().numpy()
        Rb = d.xmat.reshape(-1, 3, 3)[M["cam_bodyid"][cid]]; pb = d.xpos.reshape(-1, 3)[M["cam_bodyid"][cid]]
        pc = pb + Rb @ M["cam_pos"][cid]
        assert np.abs(E @ np.append(pc, 1) - np.array([0, 0, 0, 1])).max() < 1e-9
    assert bad / tot < 0.002, bad / tot
    K = fs.frames["c"].camera.depth.intrinsics
    assert np.isclose(K[0, 0], 0.5 * Hh / np.tan(np.pi * 45 / 360)) and np.isclose(K[0, 2], (W - 1) / 2)


@pytest.mark.gpu
def test_env_with_cameras_returns_depth_frames():
    """SimEnvCreator(..., cameras=...) (creators.py:92-96): obs["frames"][name]["depth"]["data"] for every environment."""
    from rcs_b200 import sim
    from rcs_b200.camera import SimCameraConfig
    from rcs_b200.envs.base import ControlMode
    from rcs_b200.envs.creators import SimEnvCreator
    from rcs_b200.envs.utils import default_sim_gripper_cfg, default_sim_robot_cfg
    cams = {"wrist": SimCameraConfig("wrist_0", 30, 64, 48), "bird_eye": SimCameraConfig("bird_eye_cam", 30, 64, 48)}
    env = SimEnvCreator()(ControlMode.JOINTS, default_sim_robot_cfg("fr3_simple_pick_up"), gripper_cfg=default_sim_gripper_cfg(),
                          sim_cfg=sim.SimConfig(async_control=True), cameras=cams, max_relative_movement=float(np.deg2rad(5)), num_envs=16)
    obs, info = env.reset()
    d0 = obs["frames"]["wrist"]["depth"]["data"].clone()
    assert d0.shape == (16, 48, 64, 1) and d0.dtype == torch.uint16 and info["camera_available"]
    for _ in range(3):
        obs, _, _, _, info = env.step(env.action_space.sample())
    d1 = obs["frames"]["wrist"]["depth"]["data"]
    assert (d1.to(torch.int32) - d0.to(torch.int32)).abs().max() > 0   # the wrist camera moved with the arm
    assert obs["frames"]["bird_eye"]["depth"]["data"].shape == (16, 48, 64, 1)
