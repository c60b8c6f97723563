"""Multi-point pair functions [3P: mjc_BoxBox, mjc_PlaneConvex] on a synthetic two-box scene compiled from MJCF in the
test: closed-form answers for the oracle, and the device code (host emulation) against the oracle over random poses --
contact count, order and geometry."""
import numpy as np
import pytest

import helpers as H
from helpers import O
from rcs_b200 import devmodel, mjcf

XML = """<mujoco model="two boxes">
  <option cone="elliptic" impratio="10" noslip_iterations="3"/>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 0.05"/>
    <geom name="A" type="box" size="0.5 0.5 0.5" pos="0 0 0.5"/>
    <body name="B" pos="0 0 1.24">
      <joint type="free" name="Bj"/>
      <geom name="Bg" type="box" size="%s" density="100"/>
    </body>
  </worldbody>
</mujoco>"""


def _scene(tmp_path, size="0.25 0.25 0.25"):
    p = tmp_path / "scene.xml"
    p.write_text(XML % size)
    return mjcf.compile_mjcf(str(p))


def _contacts(d, q):
    d.qpos[:] = q
    d.forward()
    n = int(d.ncon[0])
    return n, d.int("contact_geom").reshape(-1, 2)[:n].copy(), d.real("contact_real").reshape(-1, 7)[:n].copy()


def _quat(axis, ang):
    a = np.asarray(axis, float) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * a])


def test_box_box_closed_form(tmp_path):
    M = _scene(tmp_path)
    d = O.Data(O.Model(M))
    # face on face, 1 cm deep: the four bottom corners of B, normal from A (geom[0]) to B, position half-way
    n, g, r = _contacts(d, [0, 0, 1.24, 1, 0, 0, 0])
    assert n == 4 and (g == [1, 2]).all()
    assert np.allclose(r[:, 0], -0.01) and np.allclose(r[:, 4:], [0, 0, 1]) and np.allclose(r[:, 3], 0.995)
    assert sorted(map(tuple, np.round(r[:, 1:3], 9))) == sorted([(0.25, 0.25), (-0.25, 0.25), (-0.25, -0.25), (0.25, -0.25)])
    # rotated 45 degrees about z: the corners move onto the axes
    n, g, r = _contacts(d, np.concatenate([[0, 0, 1.24], _quat([0, 0, 1], np.pi / 4)]))
    assert n == 4 and np.allclose(np.sort(np.abs(r[:, 1:3]).max(axis=1)), 0.25 * np.sqrt(2))
    # hanging over A's edge: the incident face is clipped by the reference rectangle (5 points, none outside A)
    n, g, r = _contacts(d, np.concatenate([[0.45, 0, 1.24], _quat([0, 0, 1], np.pi / 4)]))
    assert n == 5 and r[:, 1].max() <= 0.5 + 1e-12 and np.isclose(r[:, 1].max(), 0.5)
    # tilted 10 degrees about y: only the lower edge of B's bottom face is within the margin -> 2 points with equal depth
    n, g, r = _contacts(d, np.concatenate([[0, 0, 1.28], _quat([0, 1, 0], np.deg2rad(10))]))
    assert n == 2 and np.isclose(r[0, 0], r[1, 0]) and r[0, 0] < 0
    # separated: nothing; exactly touching: nothing (dist = 0 is not a constraint, see oracle/mj_collision.c)
    assert _contacts(d, [0, 0, 1.26, 1, 0, 0, 0])[0] == 0
    assert _contacts(d, [0, 0, 1.25, 1, 0, 0, 0])[0] == 0
    # edge against edge: B's lowest edge (after a 45 degree roll) pitched by 30 degrees crosses A's top edge x = 0.5, z = 1
    # (direction y) 5 mm deep: one contact, normal = unit(a_y x b_edge) = (sin 30, 0, cos 30) from A to B
    th = np.deg2rad(30)
    Rx = np.array([[1, 0, 0], [0, np.sqrt(.5), -np.sqrt(.5)], [0, np.sqrt(.5), np.sqrt(.5)]])
    Ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    R = Ry @ Rx
    nrm = np.array([np.sin(th), 0, np.cos(th)])
    centre = np.array([0.5, 0, 1.0]) - 0.005 * nrm - R @ np.array([0, -0.25, -0.25])
    n, g, r = _contacts(d, np.concatenate([centre, mjcf.mat_to_quat(R)]))
    assert n == 1 and np.isclose(r[0, 0], -0.005)
    assert np.allclose(r[0, 4:], nrm, atol=1e-9) and np.allclose(r[0, 1:4], [0.5, 0, 1.0] - 0.0025 * nrm, atol=1e-9)


def test_box_box_bigger_incident_face_gives_eight_points(tmp_path):
    M = _scene(tmp_path, "0.6 0.6 0.1")
    d = O.Data(O.Model(M))
    n, g, r = _contacts(d, np.concatenate([[0, 0, 1.095], _quat([0, 0, 1], np.pi / 4)]))
    assert n == 8 and np.allclose(r[:, 0], -0.005)


@pytest.mark.parametrize("size", ["0.25 0.25 0.25", "0.6 0.6 0.1", "0.05 0.3 0.02"])
def test_device_box_box_and_plane_box_equal_the_oracle(tmp_path, size):
    """Random poses of B around A: the device code (host emulation) lists the same contacts in the same order."""
    from emu.emu import Emu
    M = _scene(tmp_path, size)
    F, verts = devmodel.build_device_fields(M, None, None)
    d = O.Data(O.Model(M))
    rng = np.random.default_rng(5)
    N = 400
    q = np.zeros((N, 7))
    for i in range(N):
        face = rng.integers(0, 3)
        p = rng.uniform(-0.7, 0.7, 3); p[2] += 0.5
        p[face] = (0.5 if face < 2 else 1.0) + rng.uniform(-0.05, 0.3)
        qq = rng.normal(size=4) if i % 3 else _quat(rng.normal(size=3), rng.choice([0, np.pi / 2, 0.01]))
        q[i] = np.concatenate([p, qq / np.linalg.norm(qq)])
    e = Emu(F, verts, N, graph=devmodel.build_mesh_graph(M))
    e.enable_contact_export(cap=16)
    e.sr[:, :7] = q
    e.run(["STEP_K"], k=1)
    hits = multi = 0
    for i in range(N):
        n, g, r = _contacts(d, q[i])
        assert int(e.contact_n[i]) == n, (i, q[i])
        assert np.array_equal(e.contact_geom[i, :n], g)
        if n:
            hits += 1
            multi += n > 2
            assert np.abs(e.contact_real[i, :n] - r).max() < 1e-9, (i, q[i])
    assert hits > 40 and multi > 10, (hits, multi)
