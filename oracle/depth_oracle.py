"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the SimCameraSet depth conventions for the parity tests of the CUDA
ray-caster (see rcs_oracle.h for who may import this).

The reference renders with OpenGL (/root/reference/src/sim/camera.cpp:100-140) and converts the z-buffer in
/root/reference/python/rcs/camera/sim.py:45-95: metres = near / (1 - d (1 - near / far)), near / far = vis.map.znear /
zfar x stat.extent, then (metres x DEPTH_SCALE).astype(uint16), rows flipped to top-down; intrinsics fx = fy =
0.5 H / tan(fovy pi / 360), principal point ((W - 1) / 2, (H - 1) / 2) (camera/sim.py:97-107). No rasteriser exists
here (PARITY UNPINNED against OpenGL by construction): this oracle casts, in float64 and fully vectorised, one ray per pixel
centre against the COLLIDABLE geoms of the compiled scene -- the geometry the CUDA kernel documents it renders -- with the
geom / camera poses taken from the physics oracle (rcso kinematics)."""
from __future__ import annotations

import numpy as np

ZNEAR, ZFAR, DEPTH_SCALE = 0.01, 50.0, 1000
PLANE, SPHERE, CAPSULE, CYLINDER, BOX, MESH = 0, 2, 3, 5, 6, 7


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _ray_geom(M, g, o, d, tmax):
    """o [3], d [P, 3] in the geom frame; returns t [P] (inf = miss)."""
    typ, sz = int(M["geom_type"][g]), M["geom_size"][g]
    P = d.shape[0]
    t = np.full(P, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        if typ == PLANE:
            ok = d[:, 2] < 0
            t[ok] = -o[2] / d[ok, 2]
        elif typ == SPHERE:
            a = (d * d).sum(1); b = d @ o; c = o @ o - sz[0] ** 2; disc = b * b - a * c
            ok = disc >= 0
            t[ok] = (-b[ok] - np.sqrt(disc[ok])) / a[ok]
        elif typ == BOX:
            t0, t1 = np.zeros(P), tmax.copy()
            for k in range(3):
                nz = d[:, k] != 0
                ta = np.where(nz, (-sz[k] - o[k]) / d[:, k], -np.inf); tb = np.where(nz, (sz[k] - o[k]) / d[:, k], np.inf)
                lo, hi = np.minimum(ta, tb), np.maximum(ta, tb)
                t0 = np.where(nz, np.maximum(t0, lo), t0); t1 = np.where(nz, np.minimum(t1, hi), t1)
                if o[k] < -sz[k] or o[k] > sz[k]:
                    t1 = np.where(nz, t1, -1.0)
            ok = t0 <= t1
            t[ok] = t0[ok]
        elif typ in (CAPSULE, CYLINDER):
            r, hl = sz[0], sz[1]
            a = d[:, 0] ** 2 + d[:, 1] ** 2; b = o[0] * d[:, 0] + o[1] * d[:, 1]; c = o[0] ** 2 + o[1] ** 2 - r * r
            disc = b * b - a * c
            ts = np.where((a > 0) & (disc >= 0), (-b - np.sqrt(np.maximum(disc, 0))) / a, np.inf)
            z = o[2] + ts * d[:, 2]
            best = np.where((ts >= 0) & (z >= -hl) & (z <= hl), ts, np.inf)
            for s in (-1, 1):
                if typ == CAPSULE:
                    oz = o[2] - s * hl
                    A = a + d[:, 2] ** 2; B = b + oz * d[:, 2]; C = o[0] ** 2 + o[1] ** 2 + oz * oz - r * r; disc = B * B - A * C
                    ts = np.where(disc >= 0, (-B - np.sqrt(np.maximum(disc, 0))) / A, np.inf)
                    ok = (ts >= 0) & (s * (o[2] + ts * d[:, 2]) >= hl) & (ts < best)
                else:
                    ts = np.where(d[:, 2] != 0, (s * hl - o[2]) / d[:, 2], np.inf)
                    x = o[0] + ts * d[:, 0]; y = o[1] + ts * d[:, 1]
                    ok = (ts >= 0) & (s * d[:, 2] < 0) & (x * x + y * y <= r * r) & (ts < best)
                best = np.where(ok, ts, best)
            t = best
        elif typ == MESH:
            a0, n = int(M["geom_faceadr"][g]), int(M["geom_facenum"][g])
            if n > 0:
                pl = np.asarray(M["mesh_face"][a0:a0 + n])
                den = d @ pl[:, :3].T                      # [P, n]
                num = -(pl[:, :3] @ o + pl[:, 3])          # [n]
                tt = num[None, :] / den
                t0 = np.max(np.where(den < 0, tt, 0.0), axis=1, initial=0.0)
                t1 = np.min(np.where(den > 0, tt, np.inf), axis=1)
                t1 = np.minimum(t1, tmax)
                t1 = np.where(((den == 0) & (num[None, :] < 0)).any(axis=1), -1.0, t1)
                ok = t0 <= t1
                t[ok] = t0[ok]
    t = np.where((t < 0) | (t > tmax), np.inf, t)
    return t


def render_depth(M, geom_xpos, geom_xmat, body_xpos, body_xmat, cam_id, width, height, physical_units=True):
    """uint16 [height, width] depth image of camera cam_id, top row first. Poses are the oracle's (rcso_data fields)."""
    extent = float(M.get("stat_extent", 1.0))
    near, far = ZNEAR * extent, ZFAR * extent
    b = int(M["cam_bodyid"][cam_id])
    Rb = np.asarray(body_xmat).reshape(-1, 3, 3)[b]; pb = np.asarray(body_xpos).reshape(-1, 3)[b]
    pc = pb + Rb @ M["cam_pos"][cam_id]
    Rc = Rb @ _quat_to_mat(np.asarray(M["cam_quat"][cam_id]) / np.linalg.norm(M["cam_quat"][cam_id]))
    f = 0.5 * height / np.tan(float(M["cam_fovy"][cam_id]) * np.pi / 360)
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    dc = np.stack([(u + 0.5 - 0.5 * width) / f, -(v + 0.5 - 0.5 * height) / f, -np.ones_like(u, dtype=float)], axis=-1).reshape(-1, 3)
    dw = dc @ Rc.T
    best = np.full(dw.shape[0], far)
    col = [g for g in range(M["ngeom"]) if M["geom_contype"][g] or M["geom_conaffinity"][g]]
    used = {int(g) for pr in M["pair_geom"] for g in pr}
    for g in (g for g in col if g in used):
        Rg = np.asarray(geom_xmat).reshape(-1, 3, 3)[g]; pg = np.asarray(geom_xpos).reshape(-1, 3)[g]
        og = Rg.T @ (pc - pg)
        dg = dw @ Rg
        t = _ray_geom(M, g, og, dg, best)
        best = np.minimum(best, t)
    z = np.maximum(best, near)
    val = z * DEPTH_SCALE if physical_units else (1 - near / z) / (1 - near / far) * DEPTH_SCALE
    return np.clip(val, 0, 65535).astype(np.uint16).reshape(height, width)
