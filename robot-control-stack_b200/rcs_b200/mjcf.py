"""MJCF -> compiled rigid-body model (the role `mujoco.MjModel.from_xml_path` plays for the
reference, /root/reference/python/rcs/sim/sim.py:47-50, and the build-time `.mjb` compile in
/root/reference/cmake/compile_scenes.cmake:19).

libmujoco is not available offline, so this module restates the subset of MuJoCo 3.2.6 compiler
semantics [3P] that the shipped scenes use (assets/fr3/mjcf/fr3_0.xml, fr3_common.xml,
assets/xarm7/mjcf/xarm7.xml, assets/scenes/*/scene.xml):

  include, compiler(angle, meshdir, autolimits), option, nested default classes + childclass,
  body/inertial/joint(hinge, slide, free)/geom(plane, box, capsule, mesh)/site/camera,
  fixed tendons, joint equalities, position + general(affine) actuators, inheritrange,
  mesh assets (binary/ascii STL, OBJ) -> convex hull, mesh inertia, mesh-fitted capsule,
  body inertia inferred from geoms, invweight0 constants (mj_setConst).

The output is a plain dict of numpy arrays named after the mjModel fields they restate.
It runs on the host only (model load time); nothing here is on the per-step path.
"""
from __future__ import annotations

import math
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

# geom type enum follows mjtGeom [3P]
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
_GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)
_JNT_TYPES = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}
TRN_JOINT, TRN_TENDON = 0, 3  # mjTRN_JOINT, mjTRN_TENDON [3P]
MINVAL = 1e-15


# ----------------------------------------------------------------------------- small math
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0 branch-stable."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def axisangle_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    s = math.sin(angle / 2)
    return np.array([math.cos(angle / 2), axis[0] * s, axis[1] * s, axis[2] * s])


def euler_quat(e, seq="xyz"):
    """MuJoCo eulerseq: lower case = intrinsic (rotating frame) [3P]."""
    q = np.array([1.0, 0, 0, 0])
    for ang, c in zip(e, seq):
        ax = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[c.lower()]
        r = axisangle_quat(ax, ang)
        q = quat_mul(q, r) if c.islower() else quat_mul(r, q)
    return q


def _principal(I):
    """Symmetric 3x3 -> (eigenvalues descending, right-handed rotation R with I = R diag R^T)."""
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    return w, V


# ----------------------------------------------------------------------------- meshes
def _load_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else 0
    if len(data) == 84 + 50 * ntri:
        arr = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri,
                            offset=84)
        tris = arr["v"].astype(np.float64)
    else:  # ascii
        vs = []
        for line in data.decode("ascii", "ignore").splitlines():
            p = line.split()
            if len(p) == 4 and p[0] == "vertex":
                vs.append([float(p[1]), float(p[2]), float(p[3])])
        tris = np.array(vs, dtype=np.float64).reshape(-1, 3, 3)
    verts, inv = np.unique(tris.reshape(-1, 3), axis=0, return_inverse=True)
    return verts, inv.reshape(-1, 3)


def _load_obj(path):
    vs, faces = [], []
    with open(path, "r") as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(p[1]), float(p[2]), float(p[3])])
            elif p[0] == "f":
                idx = [int(t.split("/")[0]) for t in p[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.array(vs, dtype=np.float64), np.array(faces, dtype=np.int64)


def mesh_mass_properties(verts, faces):
    """Unit-density volume, centre of mass and inertia about the COM by signed tetrahedra
    against the vertex centroid (MuJoCo `inertia="exact"` style; `legacy` uses |volume| per
    face, identical for the closed, outward-oriented meshes used here) [3P]."""
    c0 = verts.mean(axis=0)
    a = verts[faces[:, 0]] - c0
    b = verts[faces[:, 1]] - c0
    c = verts[faces[:, 2]] - c0
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
    if vol6.sum() < 0:
        vol6 = -vol6
    if np.any(vol6 < 0) and abs(vol6.sum()) < 0.5 * np.abs(vol6).sum():
        vol6 = np.abs(vol6)  # badly oriented soup: fall back to the legacy |volume| rule
    vol = vol6.sum() / 6.0
    com = (vol6[:, None] * (a + b + c) / 4.0).sum(axis=0) / (6.0 * vol) + c0
    # second moments about c0 via canonical tetrahedron formula
    P = np.zeros((3, 3))
    for (u, v) in ((a, a), (b, b), (c, c)):
        P += np.einsum("i,ij,ik->jk", vol6, u, v) * 2
    for (u, v) in ((a, b), (a, c), (b, c)):
        P += np.einsum("i,ij,ik->jk", vol6, u, v) + np.einsum("i,ij,ik->jk", vol6, v, u)
    P /= 120.0  # covariance integral  int x x^T dV  about c0
    d = com - c0
    P -= vol * np.outer(d, d)
    I = np.trace(P) * np.eye(3) - P
    return vol, com, I


MAX_HULL_VERTS = 256


def convex_hull_vertices(verts):
    """Convex hull vertex set of a mesh (MuJoCo collides meshes as their hull) [3P].

    Hulls with more than MAX_HULL_VERTS vertices (the 22 659-vertex camera mount) are reduced to
    the vertices that are extreme along 2048 quasi-uniform directions: the device support
    function is an exhaustive arg-max over the vertex list, not MuJoCo's hill climb, so the list
    must stay small. The reduced polytope is inscribed in the true hull (documented deviation,
    sub-0.1 mm for the shipped meshes)."""
    from scipy.spatial import ConvexHull

    hull = ConvexHull(verts)
    idx = np.sort(hull.vertices)
    hv = verts[idx].copy()
    if len(hv) > MAX_HULL_VERTS:
        n = 2048
        k = np.arange(n) + 0.5
        phi = np.arccos(1 - 2 * k / n)
        th = math.pi * (1 + 5 ** 0.5) * k
        dirs = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
        dirs = np.concatenate([dirs, np.eye(3), -np.eye(3)])
        best = np.unique(np.argmax(hv @ dirs.T, axis=0))
        if len(best) > MAX_HULL_VERTS:
            # keep the MAX_HULL_VERTS most frequently extreme ones
            cnt = np.bincount(np.argmax(hv @ dirs.T, axis=0), minlength=len(hv))
            best = np.sort(np.argsort(-cnt)[:MAX_HULL_VERTS])
        hv = hv[best]
    return hv


def hull_edge_graph(hv):
    """Edge graph of the convex hull of the vertex set hv (all of them hull vertices): for every vertex the ascending
    list of vertices it shares a (triangulated) hull facet with. MuJoCo builds the same graph with its own qhull at
    compile time and uses it for plane-mesh multi-contacts (mjc_PlaneConvex) [3P]."""
    from scipy.spatial import ConvexHull

    nbr = [set() for _ in range(len(hv))]
    if len(hv) >= 4:
        try:
            hull = ConvexHull(hv, qhull_options="Qt")
            for tri in hull.simplices:
                for a in tri:
                    for b in tri:
                        if a != b:
                            nbr[a].add(int(b))
        except Exception:  # degenerate (flat) vertex set: no graph, single-point contacts
            pass
    return [sorted(x) for x in nbr]


def hull_face_planes(hv):
    """Supporting planes of the convex hull of hv: rows (n, d) with n . x + d <= 0 inside, unit normals, coplanar
    (triangulated) facets merged. Used by the depth ray-caster (ray / convex polytope clipping)."""
    from scipy.spatial import ConvexHull

    if len(hv) < 4:
        return np.zeros((0, 4))
    try:
        eq = ConvexHull(hv).equations
    except Exception:
        return np.zeros((0, 4))
    out = []
    for e in eq:
        e = e / np.linalg.norm(e[:3])
        if not any(np.abs(e - o).max() < 1e-7 for o in out):
            out.append(e)
    return np.array(out)


def mesh_face_arrays(M):
    """Pooled hull face planes: (mesh_face[nface, 4], geom_faceadr[ngeom], geom_facenum[ngeom])."""
    adr = -np.ones(M["ngeom"], dtype=np.int32)
    num = np.zeros(M["ngeom"], dtype=np.int32)
    pool, n = [], 0
    for g in range(M["ngeom"]):
        a, k = int(M["geom_vertadr"][g]), int(M["geom_vertnum"][g])
        if k <= 0:
            continue
        pl = hull_face_planes(np.asarray(M["mesh_vert"][a:a + k]))
        adr[g], num[g] = n, len(pl)
        pool.append(pl)
        n += len(pl)
    return (np.concatenate(pool, axis=0) if pool else np.zeros((0, 4))), adr, num


def mesh_graph_arrays(M):
    """CSR form of hull_edge_graph over the pooled hull vertices: (adr[nmeshvert + 1], nbr[...] local vertex ids)."""
    nvert = len(M["mesh_vert"])
    adr = np.zeros(nvert + 1, dtype=np.int32)
    out = []
    done = {}
    for g in range(M["ngeom"]):
        a, k = int(M["geom_vertadr"][g]), int(M["geom_vertnum"][g])
        if k <= 0 or a in done:
            continue
        done[a] = True
        for i, lst in enumerate(hull_edge_graph(np.asarray(M["mesh_vert"][a:a + k]))):
            adr[a + i + 1] = len(lst)
            out.extend(lst)
    adr = np.cumsum(adr).astype(np.int32)
    return adr, np.array(out if out else [0], dtype=np.int32)


# ----------------------------------------------------------------------------- defaults
_DEFAULT_TAGS = ("joint", "geom", "site", "camera", "mesh", "material", "tendon", "equality", "general", "position",
                 "motor", "velocity", "light", "pair")
_ACT_TAGS = ("general", "position", "motor", "velocity")


class _Defaults:
    def __init__(self):
        self.classes = {"main": {t: {} for t in _DEFAULT_TAGS}}
        self.parent = {"main": None}

    def parse(self, elem, parent="main", top=True):
        name = elem.get("class", "main" if top else None)
        if name is None:
            raise ValueError("nested <default> needs a class")
        if name not in self.classes:
            self.classes[name] = {t: dict(self.classes[parent][t]) for t in _DEFAULT_TAGS}
            self.parent[name] = parent
        for child in elem:
            if child.tag == "default":
                self.parse(child, name, top=False)
            elif child.tag in _DEFAULT_TAGS:
                tag = child.tag
                attrib = dict(child.attrib)
                if tag == "position":
                    # the <position> shortcut writes gain/bias types into the class's single
                    # actuator default record, which <general> elements of the class then inherit
                    # (mjXReader::OneActuator) [3P]
                    kp = float(attrib.get("kp", self.classes[name][tag].get("kp", 1)))
                    kv = float(attrib.get("kv", self.classes[name][tag].get("kv", 0)))
                    attrib.update(gaintype="fixed", biastype="affine", gainprm=f"{kp} 0 0",
                                  biasprm=f"0 {-kp} {-kv}")
                self.classes[name][tag].update(attrib)
                if tag in _ACT_TAGS:  # actuator shortcuts share one default record [3P]
                    for t in _ACT_TAGS:
                        if t != tag:
                            self.classes[name][t].update(attrib)
        # children declared before later attribute lines already copied; MJCF order is top-down so fine

    def get(self, cls, tag):
        return dict(self.classes[cls or "main"].get(tag, {}))


# ----------------------------------------------------------------------------- compiler
def _expand_includes(root, base_dir):
    def rec(elem):
        i = 0
        while i < len(elem):
            ch = elem[i]
            if ch.tag == "include":
                inc_path = os.path.join(base_dir, ch.get("file"))
                inc = ET.parse(inc_path).getroot()
                rec(inc)
                elem.remove(ch)
                for k, sub in enumerate(list(inc)):
                    elem.insert(i + k, sub)
                i += len(inc)
            else:
                rec(ch)
                i += 1

    rec(root)


def compile_mjcf(path: str) -> dict:
    path = os.path.realpath(path)
    base_dir = os.path.dirname(path)
    root = ET.parse(path).getroot()
    _expand_includes(root, base_dir)

    # ---- compiler / option
    comp = {}
    for e in root.findall("compiler"):
        comp.update(e.attrib)
    if comp.get("angle", "degree") != "radian":
        ang = math.pi / 180.0
    else:
        ang = 1.0
    meshdir = os.path.join(base_dir, comp.get("meshdir", ""))
    eulerseq = comp.get("eulerseq", "xyz")
    opt = {}
    for e in root.findall("option"):
        opt.update(e.attrib)
    M = {}
    M["opt_timestep"] = float(opt.get("timestep", 0.002))
    M["opt_gravity"] = _vec(opt.get("gravity"), 3, (0, 0, -9.81))
    M["opt_integrator"] = opt.get("integrator", "Euler")
    M["opt_cone"] = opt.get("cone", "pyramidal")
    M["opt_impratio"] = float(opt.get("impratio", 1))
    M["opt_noslip_iterations"] = int(opt.get("noslip_iterations", 0))
    M["opt_noslip_tolerance"] = float(opt.get("noslip_tolerance", 1e-6))
    M["opt_tolerance"] = float(opt.get("tolerance", 1e-8))
    M["opt_iterations"] = int(opt.get("iterations", 100))
    M["opt_ls_iterations"] = int(opt.get("ls_iterations", 50))
    M["opt_ls_tolerance"] = float(opt.get("ls_tolerance", 0.01))
    M["opt_solver"] = opt.get("solver", "Newton")
    if M["opt_integrator"] not in ("implicitfast", "Euler"):
        raise NotImplementedError(f"integrator {M['opt_integrator']}")

    defaults = _Defaults()
    for e in root.findall("default"):
        defaults.parse(e)

    # ---- mesh assets
    meshes = {}
    for asset in root.findall("asset"):
        for m in asset.findall("mesh"):
            attrs = defaults.get(m.get("class"), "mesh")
            attrs.update(m.attrib)
            fname = attrs["file"]
            name = attrs.get("name", os.path.splitext(os.path.basename(fname))[0])
            meshes[name] = {"file": os.path.join(meshdir, fname), "scale": _vec(attrs.get("scale"), 3, (1, 1, 1))}

    mesh_cache = {}

    def get_mesh(name):
        if name not in mesh_cache:
            info = meshes[name]
            f = info["file"]
            verts, faces = _load_stl(f) if f.lower().endswith(".stl") else _load_obj(f)
            verts = verts * info["scale"]
            if np.prod(info["scale"]) < 0:
                faces = faces[:, ::-1]
            mesh_cache[name] = (verts, faces)
        return mesh_cache[name]

    # ---- containers
    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), inertial=None,
                   gravcomp=0.0, childclass=None)]
    joints, geoms, sites, cams = [], [], [], []

    def orient(attrs):
        if "quat" in attrs:
            q = _vec(attrs["quat"], 4)
            return q / np.linalg.norm(q)
        if "euler" in attrs:
            return euler_quat(_vec(attrs["euler"], 3) * ang, eulerseq)
        if "axisangle" in attrs:
            a = _vec(attrs["axisangle"], 4)
            return axisangle_quat(a[:3] / np.linalg.norm(a[:3]), a[3] * ang)
        if "xyaxes" in attrs:
            a = _vec(attrs["xyaxes"], 6)
            x = a[:3] / np.linalg.norm(a[:3])
            y = a[3:] - x * np.dot(x, a[3:])
            y /= np.linalg.norm(y)
            return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))
        if "zaxis" in attrs:
            z = _vec(attrs["zaxis"], 3)
            z /= np.linalg.norm(z)
            c = np.cross([0, 0, 1.0], z)
            s = np.linalg.norm(c)
            if s < 1e-10:
                return np.array([1.0, 0, 0, 0]) if z[2] > 0 else np.array([0, 1.0, 0, 0])
            return axisangle_quat(c / s, math.atan2(s, z[2]))
        return np.array([1.0, 0, 0, 0])

    def add_geom(e, bid, childclass):
        cls = e.get("class", childclass)
        a = defaults.get(cls, "geom")
        a.update(e.attrib)
        gtype = _GEOM_TYPES[a.get("type", "sphere")]
        g = dict(name=a.get("name", ""), body=bid, type=gtype, pos=_vec(a.get("pos"), 3, (0, 0, 0)), quat=orient(a),
                 contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                 condim=int(a.get("condim", 3)), priority=int(a.get("priority", 0)),
                 friction=_vec(a.get("friction"), None, (1, 0.005, 0.0001)),
                 solref=_vec(a.get("solref"), 2, (0.02, 1)), solimp=_vec(a.get("solimp"), None, (0.9, 0.95, 0.001, 0.5, 2)),
                 solmix=float(a.get("solmix", 1)), margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                 group=int(a.get("group", 0)), mesh=None, hull=None)
        if g["friction"].size < 3:
            g["friction"] = np.concatenate([g["friction"], np.array([1, 0.005, 0.0001])[g["friction"].size:]])
        if g["solimp"].size < 5:
            g["solimp"] = np.concatenate([g["solimp"], np.array([0.9, 0.95, 0.001, 0.5, 2])[g["solimp"].size:]])
        size = _vec(a.get("size"), None, (0, 0, 0))
        size = np.concatenate([size, np.zeros(3 - size.size)]) if size.size < 3 else size[:3]
        mass_attr = a.get("mass")
        density = float(a.get("density", 1000))
        vol, com_l, I_l = 0.0, np.zeros(3), np.zeros((3, 3))  # mass properties in the geom frame, unit density
        if "mesh" in a:
            verts, faces = get_mesh(a["mesh"])
            mvol, mcom, mI = mesh_mass_properties(verts, faces)
            if gtype == GEOM_MESH:
                g["mesh"] = a["mesh"]
                g["hull"] = convex_hull_vertices(verts)
                vol, com_l, I_l = mvol, mcom, mI
                size = np.zeros(3)
            else:
                # primitive fitted to the mesh's equivalent-inertia box, placed at mesh COM in its
                # principal frame (mjCMesh::FitGeom, fitaabb=false) [3P]
                w, R = _principal(mI)
                mm = mvol
                bx = np.array([math.sqrt(max(6 * (w[1] + w[2] - w[0]) / mm, 0)) / 2,
                               math.sqrt(max(6 * (w[0] + w[2] - w[1]) / mm, 0)) / 2,
                               math.sqrt(max(6 * (w[0] + w[1] - w[2]) / mm, 0)) / 2])
                qm = mat_to_quat(R)
                g["pos"] = g["pos"] + quat_to_mat(g["quat"]) @ mcom
                g["quat"] = quat_mul(g["quat"], qm)
                if gtype == GEOM_CAPSULE:
                    r = (bx[0] + bx[1]) / 2
                    size = np.array([r, max(0.0, bx[2] - r / 2), 0])
                elif gtype == GEOM_BOX:
                    size = bx
                elif gtype == GEOM_SPHERE:
                    size = np.array([bx.mean(), 0, 0])
                else:
                    raise NotImplementedError("mesh fit for geom type")
        elif "fromto" in a:
            ft = _vec(a["fromto"], 6)
            d = ft[3:] - ft[:3]
            L = np.linalg.norm(d)
            g["pos"] = 0.5 * (ft[:3] + ft[3:])
            z = d / L
            c = np.cross([0, 0, 1.0], z)
            s = np.linalg.norm(c)
            g["quat"] = np.array([1.0, 0, 0, 0]) if s < 1e-12 else axisangle_quat(c / s, math.atan2(s, z[2]))
            size = np.array([size[0], L / 2, 0])
        g["size"] = size
        if gtype != GEOM_MESH:
            if gtype == GEOM_BOX:
                vol = 8 * size[0] * size[1] * size[2]
                I_l = vol / 3 * np.diag([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2,
                                         size[0] ** 2 + size[1] ** 2])
            elif gtype == GEOM_SPHERE:
                vol = 4 / 3 * math.pi * size[0] ** 3
                I_l = 0.4 * vol * size[0] ** 2 * np.eye(3)
            elif gtype == GEOM_CAPSULE:
                r, h = size[0], size[1]
                vc = math.pi * r * r * 2 * h
                vs = 4 / 3 * math.pi * r ** 3
                vol = vc + vs
                izz = vc * r * r / 2 + vs * 0.4 * r * r
                ixx = vc * (3 * r * r + 4 * h * h) / 12 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
                I_l = np.diag([ixx, ixx, izz])
            elif gtype == GEOM_CYLINDER:
                r, h = size[0], size[1]
                vol = math.pi * r * r * 2 * h
                I_l = vol * np.diag([(3 * r * r + 4 * h * h) / 12] * 2 + [r * r / 2])
            elif gtype == GEOM_PLANE:
                vol = 0.0
        if mass_attr is not None:
            gm = float(mass_attr)
        else:
            gm = density * vol
        g["mass"] = gm
        g["I_unit"] = I_l * (gm / vol) if vol > 0 else np.zeros((3, 3))
        g["com_local"] = com_l
        # bounding radius about the geom frame origin
        if gtype == GEOM_MESH:
            g["rbound"] = float(np.linalg.norm(g["hull"], axis=1).max())
        elif gtype == GEOM_BOX:
            g["rbound"] = float(np.linalg.norm(size))
        elif gtype == GEOM_CAPSULE:
            g["rbound"] = float(size[0] + size[1])
        elif gtype == GEOM_SPHERE:
            g["rbound"] = float(size[0])
        elif gtype == GEOM_CYLINDER:
            g["rbound"] = float(math.hypot(size[0], size[1]))
        else:
            g["rbound"] = 0.0
        geoms.append(g)

    def add_body(e, parent, childclass):
        childclass = e.get("childclass", childclass)
        bid = len(bodies)
        b = dict(name=e.get("name", ""), parent=parent, pos=_vec(e.get("pos"), 3, (0, 0, 0)), quat=orient(e.attrib),
                 gravcomp=float(e.get("gravcomp", 0)), inertial=None, childclass=childclass)
        bodies.append(b)
        for ch in e:
            if ch.tag == "inertial":
                iq = orient(ch.attrib)
                if "fullinertia" in ch.attrib:
                    f = _vec(ch.get("fullinertia"), 6)
                    I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                    w, R = _principal(I)
                    iq, di = quat_mul(iq, mat_to_quat(R)), w
                else:
                    di = _vec(ch.get("diaginertia"), 3, (0, 0, 0))
                b["inertial"] = dict(pos=_vec(ch.get("pos"), 3, (0, 0, 0)), quat=iq, mass=float(ch.get("mass")),
                                     inertia=di)
            elif ch.tag == "joint" or ch.tag == "freejoint":
                cls = ch.get("class", childclass)
                a = defaults.get(cls, "joint") if ch.tag == "joint" else {}
                a.update(ch.attrib)
                jt = _JNT_TYPES[a.get("type", "hinge")] if ch.tag == "joint" else JNT_FREE
                rng = _vec(a.get("range"), 2, (0, 0))
                if jt == JNT_HINGE:
                    rng = rng * ang
                limited = a.get("limited", "auto")
                if limited == "auto":
                    lim = ("range" in a) and comp.get("autolimits", "true") == "true"
                else:
                    lim = limited == "true"
                afr = _vec(a.get("actuatorfrcrange"), 2, (0, 0))
                afl = a.get("actuatorfrclimited", "auto")
                afl = ("actuatorfrcrange" in a) if afl == "auto" else afl == "true"
                ax = _vec(a.get("axis"), 3, (0, 0, 1))
                joints.append(dict(name=a.get("name", ""), body=bid, type=jt, pos=_vec(a.get("pos"), 3, (0, 0, 0)),
                                   axis=ax / max(np.linalg.norm(ax), MINVAL), range=rng, limited=lim,
                                   armature=float(a.get("armature", 0)), damping=float(a.get("damping", 0)),
                                   frictionloss=float(a.get("frictionloss", 0)), stiffness=float(a.get("stiffness", 0)),
                                   ref=float(a.get("ref", 0)) * (ang if jt == JNT_HINGE else 1.0),
                                   margin=float(a.get("margin", 0)),
                                   solref=_vec(a.get("solreflimit"), 2, (0.02, 1)),
                                   solimp=_vec(a.get("solimplimit"), 5, (0.9, 0.95, 0.001, 0.5, 2)),
                                   actfrcrange=afr, actfrclimited=afl,
                                   actgravcomp=a.get("actuatorgravcomp", "false") == "true"))
            elif ch.tag == "geom":
                add_geom(ch, bid, childclass)
            elif ch.tag == "site":
                cls = ch.get("class", childclass)
                a = defaults.get(cls, "site")
                a.update(ch.attrib)
                sites.append(dict(name=a.get("name", ""), body=bid, pos=_vec(a.get("pos"), 3, (0, 0, 0)), quat=orient(a)))
            elif ch.tag == "camera":
                cls = ch.get("class", childclass)
                a = defaults.get(cls, "camera")
                a.update(ch.attrib)
                cams.append(dict(name=a.get("name", ""), body=bid, pos=_vec(a.get("pos"), 3, (0, 0, 0)), quat=orient(a),
                                 fovy=float(a.get("fovy", 45))))
        for ch in e:
            if ch.tag == "body":
                add_body(ch, bid, childclass)
        return bid

    for wb in root.findall("worldbody"):
        for ch in wb:
            if ch.tag == "geom":
                add_geom(ch, 0, None)
            elif ch.tag == "site":
                a = defaults.get(ch.get("class"), "site")
                a.update(ch.attrib)
                sites.append(dict(name=a.get("name", ""), body=0, pos=_vec(a.get("pos"), 3, (0, 0, 0)), quat=orient(a)))
            elif ch.tag == "camera":
                a = dict(ch.attrib)
                cams.append(dict(name=a.get("name", ""), body=0, pos=_vec(a.get("pos"), 3, (0, 0, 0)), quat=orient(a),
                                 fovy=float(a.get("fovy", 45))))
        for ch in wb:
            if ch.tag == "body":
                add_body(ch, 0, None)

    # MuJoCo orders geoms/joints/sites by body id (depth-first body order == our append order for
    # bodies; elements of one body are contiguous) [3P]
    geoms.sort(key=lambda g: g["body"])
    joints.sort(key=lambda j: j["body"])
    sites.sort(key=lambda s: s["body"])

    nbody, njnt, ngeom, nsite = len(bodies), len(joints), len(geoms), len(sites)

    # ---- dof / qpos addressing
    qposadr, dofadr = [], []
    nq = nv = 0
    for j in joints:
        qposadr.append(nq)
        dofadr.append(nv)
        if j["type"] == JNT_FREE:
            nq, nv = nq + 7, nv + 6
        elif j["type"] == JNT_BALL:
            raise NotImplementedError("ball joints")
        else:
            nq, nv = nq + 1, nv + 1
    M.update(nq=nq, nv=nv, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite)

    # ---- bodies
    M["body_names"] = [b["name"] for b in bodies]
    M["body_parentid"] = np.array([b["parent"] for b in bodies], dtype=np.int32)
    M["body_pos"] = np.array([b["pos"] for b in bodies])
    M["body_quat"] = np.array([b["quat"] for b in bodies])
    M["body_gravcomp"] = np.array([b["gravcomp"] for b in bodies])
    body_jntnum = np.zeros(nbody, dtype=np.int32)
    body_jntadr = -np.ones(nbody, dtype=np.int32)
    for ji, j in enumerate(joints):
        if body_jntnum[j["body"]] == 0:
            body_jntadr[j["body"]] = ji
        body_jntnum[j["body"]] += 1
    M["body_jntnum"], M["body_jntadr"] = body_jntnum, body_jntadr
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = -np.ones(nbody, dtype=np.int32)
    for ji, j in enumerate(joints):
        nd = 6 if j["type"] == JNT_FREE else 1
        if body_dofnum[j["body"]] == 0:
            body_dofadr[j["body"]] = dofadr[ji]
        body_dofnum[j["body"]] += nd
    M["body_dofnum"], M["body_dofadr"] = body_dofnum, body_dofadr
    weldid = np.zeros(nbody, dtype=np.int32)
    rootid = np.zeros(nbody, dtype=np.int32)
    for i in range(1, nbody):
        p = bodies[i]["parent"]
        weldid[i] = i if body_jntnum[i] > 0 else weldid[p]
        rootid[i] = i if p == 0 else rootid[p]
    M["body_weldid"], M["body_rootid"] = weldid, rootid

    # inertial: explicit or inferred from geoms
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    body_inertia = np.zeros((nbody, 3))
    for i, b in enumerate(bodies):
        if b["inertial"] is not None:
            body_mass[i] = b["inertial"]["mass"]
            body_ipos[i] = b["inertial"]["pos"]
            body_iquat[i] = b["inertial"]["quat"]
            body_inertia[i] = b["inertial"]["inertia"]
        elif i > 0 and weldid[i] != 0:
            gs = [g for g in geoms if g["body"] == i and g["mass"] > 0]
            if gs:
                m = sum(g["mass"] for g in gs)
                coms = [g["pos"] + quat_to_mat(g["quat"]) @ g["com_local"] for g in gs]
                com = sum(g["mass"] * c for g, c in zip(gs, coms)) / m
                I = np.zeros((3, 3))
                for g, c in zip(gs, coms):
                    R = quat_to_mat(g["quat"])
                    d = c - com
                    I += R @ g["I_unit"] @ R.T + g["mass"] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                w, R = _principal(I)
                body_mass[i], body_ipos[i], body_iquat[i], body_inertia[i] = m, com, mat_to_quat(R), w
        # static (world-welded) bodies: mass is dynamically irrelevant; left at zero (documented deviation
        # from MuJoCo, which would infer a mass from the visual meshes of fr3_link0)
    M["body_mass"], M["body_ipos"], M["body_iquat"], M["body_inertia"] = body_mass, body_ipos, body_iquat, body_inertia

    # ---- joints / dofs
    M["jnt_names"] = [j["name"] for j in joints]
    M["jnt_type"] = np.array([j["type"] for j in joints], dtype=np.int32)
    M["jnt_bodyid"] = np.array([j["body"] for j in joints], dtype=np.int32)
    M["jnt_qposadr"] = np.array(qposadr, dtype=np.int32)
    M["jnt_dofadr"] = np.array(dofadr, dtype=np.int32)
    M["jnt_pos"] = np.array([j["pos"] for j in joints]).reshape(njnt, 3)
    M["jnt_axis"] = np.array([j["axis"] for j in joints]).reshape(njnt, 3)
    M["jnt_range"] = np.array([j["range"] for j in joints]).reshape(njnt, 2)
    M["jnt_limited"] = np.array([j["limited"] for j in joints], dtype=np.int32)
    M["jnt_margin"] = np.array([j["margin"] for j in joints])
    M["jnt_solref"] = np.array([j["solref"] for j in joints]).reshape(njnt, 2)
    M["jnt_solimp"] = np.array([j["solimp"] for j in joints]).reshape(njnt, 5)
    M["jnt_actfrcrange"] = np.array([j["actfrcrange"] for j in joints]).reshape(njnt, 2)
    M["jnt_actfrclimited"] = np.array([j["actfrclimited"] for j in joints], dtype=np.int32)
    M["jnt_actgravcomp"] = np.array([j["actgravcomp"] for j in joints], dtype=np.int32)
    dof_jntid, dof_bodyid, dof_parentid = [], [], []
    dof_arm, dof_damp, dof_floss = [], [], []
    last_dof_of_body = {}
    for ji, j in enumerate(joints):
        nd = 6 if j["type"] == JNT_FREE else 1
        for k in range(nd):
            d = dofadr[ji] + k
            dof_jntid.append(ji)
            dof_bodyid.append(j["body"])
            # parent dof: previous dof on this body, else last dof of nearest ancestor with dofs
            if k > 0 or (j["body"] in last_dof_of_body):
                par = d - 1 if k > 0 else last_dof_of_body[j["body"]]
            else:
                par = -1
                p = bodies[j["body"]]["parent"]
                while p != 0:
                    if p in last_dof_of_body:
                        par = last_dof_of_body[p]
                        break
                    p = bodies[p]["parent"]
            dof_parentid.append(par)
            last_dof_of_body[j["body"]] = d
            dof_arm.append(j["armature"])
            dof_damp.append(j["damping"])
            dof_floss.append(j["frictionloss"])
    M["dof_jntid"] = np.array(dof_jntid, dtype=np.int32)
    M["dof_bodyid"] = np.array(dof_bodyid, dtype=np.int32)
    M["dof_parentid"] = np.array(dof_parentid, dtype=np.int32)
    M["dof_armature"] = np.array(dof_arm)
    M["dof_damping"] = np.array(dof_damp)
    M["dof_frictionloss"] = np.array(dof_floss)
    if any(j["stiffness"] != 0 for j in joints):
        raise NotImplementedError("joint springs")

    # qpos0
    qpos0 = np.zeros(nq)
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            if b["parent"] != 0:
                raise NotImplementedError("free joint below a non-world body")
            qpos0[qposadr[ji]:qposadr[ji] + 3] = b["pos"]
            qpos0[qposadr[ji] + 3:qposadr[ji] + 7] = b["quat"]
        else:
            qpos0[qposadr[ji]] = j["ref"]
    M["qpos0"] = qpos0

    # ---- geoms / sites / cameras
    M["geom_names"] = [g["name"] for g in geoms]
    M["geom_type"] = np.array([g["type"] for g in geoms], dtype=np.int32)
    M["geom_bodyid"] = np.array([g["body"] for g in geoms], dtype=np.int32)
    M["geom_contype"] = np.array([g["contype"] for g in geoms], dtype=np.int32)
    M["geom_conaffinity"] = np.array([g["conaffinity"] for g in geoms], dtype=np.int32)
    M["geom_condim"] = np.array([g["condim"] for g in geoms], dtype=np.int32)
    M["geom_priority"] = np.array([g["priority"] for g in geoms], dtype=np.int32)
    M["geom_size"] = np.array([g["size"] for g in geoms]).reshape(ngeom, 3)
    M["geom_pos"] = np.array([g["pos"] for g in geoms]).reshape(ngeom, 3)
    M["geom_quat"] = np.array([g["quat"] for g in geoms]).reshape(ngeom, 4)
    M["geom_friction"] = np.array([g["friction"][:3] for g in geoms]).reshape(ngeom, 3)
    M["geom_solref"] = np.array([g["solref"] for g in geoms]).reshape(ngeom, 2)
    M["geom_solimp"] = np.array([g["solimp"] for g in geoms]).reshape(ngeom, 5)
    M["geom_solmix"] = np.array([g["solmix"] for g in geoms])
    M["geom_margin"] = np.array([g["margin"] for g in geoms])
    M["geom_gap"] = np.array([g["gap"] for g in geoms])
    M["geom_rbound"] = np.array([g["rbound"] for g in geoms])
    M["geom_group"] = np.array([g["group"] for g in geoms], dtype=np.int32)
    # hull vertex pool (only collidable mesh geoms keep their hull)
    vert_adr = -np.ones(ngeom, dtype=np.int32)
    vert_num = np.zeros(ngeom, dtype=np.int32)
    pool = []
    n = 0
    for gi, g in enumerate(geoms):
        if g["type"] == GEOM_MESH and (g["contype"] or g["conaffinity"]):
            vert_adr[gi], vert_num[gi] = n, len(g["hull"])
            pool.append(g["hull"])
            n += len(g["hull"])
    M["geom_vertadr"], M["geom_vertnum"] = vert_adr, vert_num
    M["mesh_vert"] = np.concatenate(pool, axis=0) if pool else np.zeros((0, 3))
    M["mesh_graphadr"], M["mesh_graph"] = mesh_graph_arrays(M)
    M["mesh_face"], M["geom_faceadr"], M["geom_facenum"] = mesh_face_arrays(M)
    # local AABB (centre, half-size) in the geom frame for the mid-phase box test
    aabb = np.zeros((ngeom, 6))
    for gi, g in enumerate(geoms):
        if g["type"] == GEOM_MESH and g["hull"] is not None:
            lo, hi = g["hull"].min(axis=0), g["hull"].max(axis=0)
            aabb[gi, :3], aabb[gi, 3:] = (lo + hi) / 2, (hi - lo) / 2
        elif g["type"] == GEOM_BOX:
            aabb[gi, 3:] = g["size"]
        elif g["type"] == GEOM_CAPSULE:
            aabb[gi, 3:] = [g["size"][0], g["size"][0], g["size"][0] + g["size"][1]]
        elif g["type"] == GEOM_SPHERE:
            aabb[gi, 3:] = g["size"][0]
        elif g["type"] == GEOM_CYLINDER:
            aabb[gi, 3:] = [g["size"][0], g["size"][0], g["size"][1]]
    M["geom_aabb"] = aabb
    # bounding sphere about the local AABB centre (tighter than geom_rbound, which is about the geom origin)
    bs = np.zeros((ngeom, 4))
    for gi, g in enumerate(geoms):
        bs[gi, :3] = aabb[gi, :3]
        if g["type"] == GEOM_MESH and g["hull"] is not None:
            bs[gi, 3] = float(np.linalg.norm(g["hull"] - aabb[gi, :3], axis=1).max())
        elif g["type"] == GEOM_CAPSULE:
            bs[gi, 3] = g["size"][0] + g["size"][1]
        elif g["type"] == GEOM_SPHERE:
            bs[gi, 3] = g["size"][0]
        else:
            bs[gi, 3] = float(np.linalg.norm(aabb[gi, 3:]))
    M["geom_bsphere"] = bs

    M["site_names"] = [s["name"] for s in sites]
    M["site_bodyid"] = np.array([s["body"] for s in sites], dtype=np.int32)
    M["site_pos"] = np.array([s["pos"] for s in sites]).reshape(nsite, 3)
    M["site_quat"] = np.array([s["quat"] for s in sites]).reshape(nsite, 4)
    M["cam_names"] = [c["name"] for c in cams]
    M["cam_bodyid"] = np.array([c["body"] for c in cams], dtype=np.int32)
    M["cam_pos"] = np.array([c["pos"] for c in cams]).reshape(len(cams), 3)
    M["cam_quat"] = np.array([c["quat"] for c in cams]).reshape(len(cams), 4)
    M["cam_fovy"] = np.array([c["fovy"] for c in cams])

    # ---- tendons (fixed only)
    tendons = []
    for t in root.findall("tendon"):
        for f in t.findall("fixed"):
            coef = np.zeros(nv)
            for jj in f.findall("joint"):
                ji = M["jnt_names"].index(jj.get("joint"))
                coef[dofadr[ji]] = float(jj.get("coef"))
            tendons.append(dict(name=f.get("name", ""), coef=coef))
        if t.findall("spatial"):
            raise NotImplementedError("spatial tendons")
    M["ntendon"] = len(tendons)
    M["tendon_names"] = [t["name"] for t in tendons]
    M["tendon_coef"] = np.array([t["coef"] for t in tendons]).reshape(len(tendons), nv)

    # ---- equality (joint only)
    eqs = []
    for eq in root.findall("equality"):
        for e in eq:
            if e.tag != "joint":
                raise NotImplementedError(f"equality {e.tag}")
            a = defaults.get(e.get("class"), "equality")
            a.update(e.attrib)
            j1 = M["jnt_names"].index(a["joint1"])
            j2 = M["jnt_names"].index(a["joint2"]) if "joint2" in a else -1
            poly = _vec(a.get("polycoef"), 5, (0, 1, 0, 0, 0))
            eqs.append(dict(j1=j1, j2=j2, poly=poly, solref=_vec(a.get("solref"), 2, (0.02, 1)),
                            solimp=_vec(a.get("solimp"), None, (0.9, 0.95, 0.001, 0.5, 2)),
                            active=a.get("active", "true") == "true"))
    for e in eqs:
        if e["solimp"].size < 5:
            e["solimp"] = np.concatenate([e["solimp"], np.array([0.9, 0.95, 0.001, 0.5, 2])[e["solimp"].size:]])
    M["neq"] = len(eqs)
    M["eq_obj1id"] = np.array([e["j1"] for e in eqs], dtype=np.int32)
    M["eq_obj2id"] = np.array([e["j2"] for e in eqs], dtype=np.int32)
    M["eq_polycoef"] = np.array([e["poly"] for e in eqs]).reshape(len(eqs), 5)
    M["eq_solref"] = np.array([e["solref"] for e in eqs]).reshape(len(eqs), 2)
    M["eq_solimp"] = np.array([e["solimp"] for e in eqs]).reshape(len(eqs), 5)
    M["eq_active0"] = np.array([e["active"] for e in eqs], dtype=np.int32)

    # ---- actuators
    acts = []
    for ac in root.findall("actuator"):
        for e in ac:
            if e.tag not in ("position", "general", "motor"):
                raise NotImplementedError(f"actuator {e.tag}")
            a = defaults.get(e.get("class"), e.tag)
            a.update(e.attrib)
            gain = np.zeros(3)
            bias = np.zeros(3)
            gp = _vec(a.get("gainprm"), None, (1, 0, 0))
            gain[:min(3, gp.size)] = gp[:3]
            bp = _vec(a.get("biasprm"), None, (0, 0, 0))
            bias[:min(3, bp.size)] = bp[:3]
            biastype = a.get("biastype", "none")
            if e.tag == "position":
                kp = float(a.get("kp", 1))
                kv = float(a.get("kv", 0))
                gain[:] = (kp, 0, 0)
                bias[:] = (0, -kp, -kv)
                biastype = "affine"
            elif e.tag == "motor":
                gain[:] = (1, 0, 0)
                bias[:] = 0
                biastype = "none"
            if a.get("gaintype", "fixed") != "fixed" or a.get("dyntype", "none") != "none":
                raise NotImplementedError("actuator gaintype/dyntype")
            if biastype == "none":
                bias[:] = 0
            if "joint" in a:
                trntype, trnid = TRN_JOINT, M["jnt_names"].index(a["joint"])
            elif "tendon" in a:
                trntype, trnid = TRN_TENDON, M["tendon_names"].index(a["tendon"])
            else:
                raise NotImplementedError("actuator transmission")
            gear = _vec(a.get("gear"), None, (1,))[0]
            ctrlrange = _vec(a.get("ctrlrange"), 2, (0, 0))
            ctrllimited = a.get("ctrllimited", "auto")
            ctrllimited = ("ctrlrange" in a) if ctrllimited == "auto" else ctrllimited == "true"
            inherit = float(a.get("inheritrange", 0)) if e.tag == "position" else 0.0
            if inherit > 0 and not ("ctrlrange" in e.attrib):
                if trntype != TRN_JOINT:
                    raise NotImplementedError("inheritrange on tendon")
                r = M["jnt_range"][trnid]
                mean, half = 0.5 * (r[0] + r[1]), 0.5 * (r[1] - r[0]) * inherit
                ctrlrange = np.array([mean - half, mean + half])
                ctrllimited = True
            forcerange = _vec(a.get("forcerange"), 2, (0, 0))
            forcelimited = a.get("forcelimited", "auto")
            forcelimited = ("forcerange" in a) if forcelimited == "auto" else forcelimited == "true"
            acts.append(dict(name=a.get("name", ""), trntype=trntype, trnid=trnid, gear=gear, gain=gain, bias=bias,
                             ctrlrange=ctrlrange, ctrllimited=ctrllimited, forcerange=forcerange,
                             forcelimited=forcelimited))
    nu = len(acts)
    M["nu"] = nu
    M["actuator_names"] = [a["name"] for a in acts]
    M["actuator_trntype"] = np.array([a["trntype"] for a in acts], dtype=np.int32)
    M["actuator_trnid"] = np.array([a["trnid"] for a in acts], dtype=np.int32)
    M["actuator_gear"] = np.array([a["gear"] for a in acts])
    M["actuator_gainprm"] = np.array([a["gain"] for a in acts]).reshape(nu, 3)
    M["actuator_biasprm"] = np.array([a["bias"] for a in acts]).reshape(nu, 3)
    M["actuator_ctrlrange"] = np.array([a["ctrlrange"] for a in acts]).reshape(nu, 2)
    M["actuator_ctrllimited"] = np.array([a["ctrllimited"] for a in acts], dtype=np.int32)
    M["actuator_forcerange"] = np.array([a["forcerange"] for a in acts]).reshape(nu, 2)
    M["actuator_forcelimited"] = np.array([a["forcelimited"] for a in acts], dtype=np.int32)

    # ---- statistic (used by camera near/far only)
    st = root.find("statistic")
    M["stat_extent"] = float(st.get("extent", 1)) if st is not None else 1.0

    _set_const(M)
    _collision_pairs(M)
    return M


# ----------------------------------------------------------------------------- mj_setConst subset
def _kin0(M):
    """Forward kinematics at qpos0 (world frames of bodies, joints, inertial frames)."""
    nbody = M["nbody"]
    xpos = np.zeros((nbody, 3))
    xquat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    for i in range(1, nbody):
        p = M["body_parentid"][i]
        if M["body_jntnum"][i] and M["jnt_type"][M["body_jntadr"][i]] == JNT_FREE:
            a = M["jnt_qposadr"][M["body_jntadr"][i]]
            xpos[i] = M["qpos0"][a:a + 3]
            q = M["qpos0"][a + 3:a + 7]
            xquat[i] = q / np.linalg.norm(q)
        else:
            xpos[i] = xpos[p] + quat_to_mat(xquat[p]) @ M["body_pos"][i]
            xquat[i] = quat_mul(xquat[p], M["body_quat"][i])
            # hinge/slide at qpos0==ref contribute identity
    return xpos, xquat


def _set_const(M):
    """dof_invweight0, body_invweight0, tendon_invweight0, body_subtreemass, stat_meaninertia
    (mj_setConst / set0 in engine_setconst.c) [3P]."""
    nbody, nv = M["nbody"], M["nv"]
    xpos, xquat = _kin0(M)
    # dense mass matrix at qpos0 through body Jacobians (independent of the CRBA in oracle/kernels)
    Mm = np.zeros((nv, nv))
    jacs = []
    for b in range(nbody):
        R = quat_to_mat(xquat[b])
        com = xpos[b] + R @ M["body_ipos"][b]
        Jp, Jr = np.zeros((3, nv)), np.zeros((3, nv))
        bb = b
        while bb != 0:
            for k in range(M["body_jntnum"][bb]):
                ji = M["body_jntadr"][bb] + k
                d = M["jnt_dofadr"][ji]
                Rb = quat_to_mat(xquat[bb])
                t = M["jnt_type"][ji]
                if t == JNT_FREE:
                    Jp[:, d:d + 3] = np.eye(3)
                    for a in range(3):
                        ax = Rb[:, a]
                        Jr[:, d + 3 + a] = ax
                        Jp[:, d + 3 + a] = np.cross(ax, com - xpos[bb])
                else:
                    ax = Rb @ M["jnt_axis"][ji]
                    anchor = xpos[bb] + Rb @ M["jnt_pos"][ji]
                    if t == JNT_HINGE:
                        Jr[:, d] = ax
                        Jp[:, d] = np.cross(ax, com - anchor)
                    else:
                        Jp[:, d] = ax
            bb = M["body_parentid"][bb]
        jacs.append((Jp, Jr))
        Ri = quat_to_mat(quat_mul(xquat[b], M["body_iquat"][b]))
        Iw = Ri @ np.diag(M["body_inertia"][b]) @ Ri.T
        Mm += M["body_mass"][b] * Jp.T @ Jp + Jr.T @ Iw @ Jr
    Mm += np.diag(M["dof_armature"])
    Minv = np.linalg.inv(Mm) if nv else np.zeros((0, 0))
    dof_inv = np.zeros(nv)
    for ji in range(M["njnt"]):
        d = M["jnt_dofadr"][ji]
        if M["jnt_type"][ji] == JNT_FREE:
            dof_inv[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            dof_inv[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            dof_inv[d] = Minv[d, d]
    M["dof_invweight0"] = dof_inv
    binv = np.zeros((nbody, 2))
    for b in range(1, nbody):
        Jp, Jr = jacs[b]
        if M["body_weldid"][b] == 0:
            continue
        binv[b, 0] = max(MINVAL, np.trace(Jp @ Minv @ Jp.T) / 3)
        binv[b, 1] = max(MINVAL, np.trace(Jr @ Minv @ Jr.T) / 3)
    M["body_invweight0"] = binv
    M["tendon_invweight0"] = np.array([c @ Minv @ c for c in M["tendon_coef"]]) if M["ntendon"] else np.zeros(0)
    M["stat_meaninertia"] = float(np.mean(np.diag(Mm))) if nv else 1.0
    sub = M["body_mass"].copy()
    for i in range(nbody - 1, 0, -1):
        sub[M["body_parentid"][i]] += sub[i]
    M["body_subtreemass"] = sub
    M["qM0"] = Mm


def _collision_pairs(M):
    """Static candidate geom pairs: contype/conaffinity, same-body, weld and parent filters
    (mj_collision broadphase filters, engine_collision_driver.c) [3P]. Sorted (g1<g2), lexicographic."""
    pairs = []
    w = M["body_weldid"]
    par = M["body_parentid"]
    for g1 in range(M["ngeom"]):
        for g2 in range(g1 + 1, M["ngeom"]):
            c1, a1, c2, a2 = M["geom_contype"][g1], M["geom_conaffinity"][g1], M["geom_contype"][g2], \
                M["geom_conaffinity"][g2]
            if not ((c1 & a2) or (c2 & a1)):
                continue
            b1, b2 = M["geom_bodyid"][g1], M["geom_bodyid"][g2]
            if b1 == b2:
                continue
            w1, w2 = w[b1], w[b2]
            if w1 == w2:
                continue
            wp1, wp2 = w[par[w1]], w[par[w2]]
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            if M["geom_type"][g1] == GEOM_PLANE and M["geom_type"][g2] == GEOM_PLANE:
                continue
            pairs.append((g1, g2))
    M["pair_geom"] = np.array(pairs, dtype=np.int32).reshape(-1, 2)


# ----------------------------------------------------------------------------- (de)serialisation
def save_model(M: dict, path: str):
    out = {}
    for k, v in M.items():
        if isinstance(v, list):
            out[k] = np.array(v, dtype=object) if v and not isinstance(v[0], str) else np.array(v, dtype="U64")
        elif isinstance(v, str):
            out[k] = np.array(v)
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(path, **out)


def load_model(path: str) -> dict:
    z = np.load(path, allow_pickle=False)
    M = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind == "U":
            M[k] = v.tolist() if v.ndim else str(v)
        elif v.ndim == 0:
            M[k] = v.item()
        else:
            M[k] = v
    return M
