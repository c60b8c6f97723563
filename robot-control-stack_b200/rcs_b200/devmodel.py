"""Compiled scene (rcs_b200.mjcf) -> fused device model fields (csrc/rcsb_types.h: RcsbModel).

Fusion: every body without a joint is folded into its nearest moving ancestor (or the world), so the
device kinematic tree has exactly one body per joint. Composite mass properties, geom / site poses
and gravity-compensation centres are re-expressed in the moving body's frame. Tree recursions are
replaced by ancestor / descendant bit masks. The name->id resolution of SimRobot::init_ids
(/root/reference/src/sim/SimRobot.cpp:52-94) and of the SimGripper constructor
(/root/reference/src/sim/SimGripper.cpp:12-39) happens here and raises the same errors.
"""
from __future__ import annotations

import numpy as np

from . import mjcf
from .mjcf import JNT_FREE, quat_mul, quat_to_mat, mat_to_quat

FAST_MAXCON_DEFAULT = 1  # contact capacity of the reduced workspace layout (scenes without free bodies)
FAST_MAXCON_FREE = 4     # ... of scenes with free bodies: an object resting on the floor (plane-box: 4 points)
FAST_LIMIT_ROWS = 4      # simultaneously active joint-limit rows the reduced layout holds
# per-env contact capacity of the full layout (overflow is counted in the warn flag); Sim(maxcon=...) overrides.
# 16: the four small pads of each finger meeting (4 box-box pairs x 4 points); 40: a grasped box (8 pad pairs x 4 points)
# plus its 4 floor contacts
MAXCON_DEFAULT = 16
MAXCON_FREE = 40
ROLE_ARM, ROLE_GRIPPER, ROLE_FINGER, ROLE_IGNORED = 1, 2, 4, 8


def _name_id(names, name, kind):
    try:
        return names.index(name)
    except ValueError:
        raise RuntimeError(f"No {kind} named {name}") from None


def build_mesh_graph(M: dict):
    """Hull edge graph (mjcf.mesh_graph_arrays) restricted to the device vertex pool of build_device_fields:
    (adr[nvert + 1], nbr[...]) with neighbour ids local to each geom's hull."""
    col = [g for g in range(M["ngeom"]) if M["geom_contype"][g] or M["geom_conaffinity"][g]]
    used = {int(g) for pr in M["pair_geom"] for g in pr}
    gadr, gnbr = [0], []
    for g in (g for g in col if g in used):
        a, k = int(M["geom_vertadr"][g]), int(M["geom_vertnum"][g])
        for v in range(a, a + max(k, 0)):
            if "mesh_graphadr" in M:
                gnbr.extend(int(x) for x in M["mesh_graph"][int(M["mesh_graphadr"][v]):int(M["mesh_graphadr"][v + 1])])
            gadr.append(len(gnbr))
    if len(gadr) == 1:
        gadr.append(0)  # the vertex pool holds one dummy vertex when the scene has no meshes
    return np.array(gadr, dtype=np.int32), np.array(gnbr if gnbr else [0], dtype=np.int32)[:max(len(gnbr), 1)]


def _collidable(M: dict):
    col = [g for g in range(M["ngeom"]) if M["geom_contype"][g] or M["geom_conaffinity"][g]]
    used = {int(g) for pr in M["pair_geom"] for g in pr}
    return [g for g in col if g in used]


def build_mesh_faces(M: dict):
    """Hull face planes restricted to the device geom list: (planes[nface, 4], faceadr[ng], facenum[ng])."""
    col = _collidable(M)
    adr, num, pool, n = [], [], [], 0
    for g in col:
        k = int(M["geom_facenum"][g]) if "geom_facenum" in M else 0
        if k > 0:
            a = int(M["geom_faceadr"][g])
            adr.append(n); num.append(k); pool.append(np.asarray(M["mesh_face"][a:a + k], dtype=np.float64)); n += k
        else:
            adr.append(0); num.append(0)
    planes = np.concatenate(pool, axis=0) if pool else np.zeros((1, 4))
    return np.ascontiguousarray(planes), np.array(adr, dtype=np.int32), np.array(num, dtype=np.int32)


def fold_frame(M: dict, body: int, pos, quat):
    """A frame given in (original) body `body` re-expressed in the device model's body tree: returns (moving body index or
    -1 for the world, position, row-major rotation matrix). Jointless bodies are folded into their nearest moving ancestor
    exactly as build_device_fields folds them."""
    parent = M["body_parentid"]
    moving = [b for b in range(M["nbody"]) if M["body_jntnum"][b] > 0]
    mb_of = {b: i for i, b in enumerate(moving)}
    p, q = np.asarray(pos, dtype=np.float64), np.asarray(quat, dtype=np.float64)
    b = int(body)
    while b != 0 and b not in mb_of:  # climb through the welded chain, composing the static offsets
        p = M["body_pos"][b] + quat_to_mat(M["body_quat"][b]) @ p
        q = quat_mul(M["body_quat"][b], q)
        b = int(parent[b])
    q = q / np.linalg.norm(q)
    return (mb_of[b] if b != 0 else -1), p, quat_to_mat(q)


def build_device_fields(M: dict, robot_cfg=None, gripper_cfg=None, maxcon: int | None = None,
                        fast_maxcon: int | None = None) -> tuple[dict, np.ndarray]:
    """Returns ({field: (np.ndarray, is_real)}, mesh_vert[nvert,3]).

    robot_cfg: object with joints, actuators, arm_collision_geoms, attachment_site, base, tcp_offset (7,: xyz+xyzw),
               q_home, joint_rotational_tolerance, seconds_between_callbacks, register_convergence_callback, ik_nq
    gripper_cfg: object with actuator, joint, collision_geoms, collision_geoms_fingers, ignored_collision_geoms,
               epsilon_inner/outer, seconds_between_callbacks, max/min_actuator_width, max/min_joint_width; or None
    """
    nbody, nv, nq, nu = M["nbody"], M["nv"], M["nq"], M["nu"]
    parent = M["body_parentid"]
    # ---- moving bodies: those with a joint
    moving = [b for b in range(nbody) if M["body_jntnum"][b] > 0]
    for b in moving:
        if M["body_jntnum"][b] != 1:
            raise NotImplementedError("bodies with more than one joint")
    mb_of = {b: i for i, b in enumerate(moving)}
    nb = len(moving)

    def owner(b):  # nearest moving ancestor-or-self (orig body id), -1 for world-welded
        while b != 0 and b not in mb_of:
            b = parent[b]
        return b if b != 0 else -1

    # pose of every body relative to its owner's frame, at qpos0 (static offsets only)
    rel_pos = np.zeros((nbody, 3))
    rel_quat = np.tile(np.array([1.0, 0, 0, 0]), (nbody, 1))
    own = np.array([owner(b) for b in range(nbody)])
    for b in range(1, nbody):
        if b in mb_of:
            continue  # identity relative to itself
        p = parent[b]
        # pose in owner's frame = pose_of_parent_in_owner * local
        if p == 0 or (p in mb_of):
            pp, pq = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            pp, pq = rel_pos[p], rel_quat[p]
        rel_pos[b] = pp + quat_to_mat(pq) @ M["body_pos"][b]
        rel_quat[b] = quat_mul(pq, M["body_quat"][b])
    F = {}

    def put(name, arr, real):
        F[name] = (np.ascontiguousarray(arr, dtype=np.float64 if real else np.int32), real)

    b_parent, b_pos, b_quat = [], [], []
    for b in moving:
        p = parent[b]
        po = own[p] if p != 0 else -1
        # frame of b in the owner-of-parent's frame
        if p == 0 or p in mb_of:
            pp, pq = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            pp, pq = rel_pos[p], rel_quat[p]
        b_pos.append(pp + quat_to_mat(pq) @ M["body_pos"][b])
        b_quat.append(quat_mul(pq, M["body_quat"][b]))
        b_parent.append(mb_of[po] if po >= 0 else -1)
    jid = [M["body_jntadr"][b] for b in moving]
    put("b_parent", b_parent, False)
    put("b_jtype", [M["jnt_type"][j] for j in jid], False)
    put("b_qadr", [M["jnt_qposadr"][j] for j in jid], False)
    put("b_dadr", [M["jnt_dofadr"][j] for j in jid], False)
    put("b_ndof", [6 if M["jnt_type"][j] == JNT_FREE else 1 for j in jid], False)
    put("b_pos", b_pos, True)
    put("b_quat", [q / np.linalg.norm(q) for q in b_quat], True)
    put("b_jpos", [M["jnt_pos"][j] for j in jid], True)
    put("b_jaxis", [M["jnt_axis"][j] for j in jid], True)
    # roots: kinematic trees (MuJoCo body_rootid = child of world)
    roots = sorted({int(M["body_rootid"][b]) for b in moving})
    root_of = {r: i for i, r in enumerate(roots)}
    put("b_root", [root_of[int(M["body_rootid"][b])] for b in moving], False)
    # masks
    anc = []
    for i, b in enumerate(moving):
        mask, k = 0, i
        while k >= 0:
            mask |= 1 << k
            k = b_parent[k]
        anc.append(mask)
    desc = [sum(1 << j for j in range(nb) if (anc[j] >> i) & 1) for i in range(nb)]
    dofs_of = [list(range(M["jnt_dofadr"][j], M["jnt_dofadr"][j] + (6 if M["jnt_type"][j] == JNT_FREE else 1))) for j in jid]
    dofmask = [sum(1 << d for k in range(nb) if (anc[i] >> k) & 1 for d in dofs_of[k]) for i in range(nb)]
    put("b_ancmask", np.array(anc, dtype=np.uint32).view(np.int32), False)
    put("b_descmask", np.array(desc, dtype=np.uint32).view(np.int32), False)
    put("b_dofmask", np.array(dofmask, dtype=np.uint32).view(np.int32), False)
    # composite mass properties of each moving body with its welded static descendants
    b_mass, b_ipos, b_inertia, b_gcmass, b_gcpos = [], [], [], [], []
    for b in moving:
        members = [x for x in range(nbody) if own[x] == b]
        ms = np.array([M["body_mass"][x] for x in members])
        coms = [rel_pos[x] + quat_to_mat(rel_quat[x]) @ M["body_ipos"][x] for x in members]
        mt = ms.sum()
        com = sum(mm * cc for mm, cc in zip(ms, coms)) / mt if mt > 0 else np.zeros(3)
        I = np.zeros((3, 3))
        for x, mm, cc in zip(members, ms, coms):
            R = quat_to_mat(quat_mul(rel_quat[x], M["body_iquat"][x]))
            d = cc - com
            I += R @ np.diag(M["body_inertia"][x]) @ R.T + mm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        b_mass.append(mt)
        b_ipos.append(com)
        b_inertia.append([I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])
        gm = np.array([M["body_mass"][x] * M["body_gravcomp"][x] for x in members])
        b_gcmass.append(gm.sum())
        b_gcpos.append(sum(g * cc for g, cc in zip(gm, coms)) / gm.sum() if gm.sum() != 0 else np.zeros(3))
    put("b_mass", b_mass, True)
    put("b_ipos", b_ipos, True)
    put("b_inertia", b_inertia, True)
    put("b_gcmass", b_gcmass, True)
    put("b_gcpos", b_gcpos, True)
    rmass = np.zeros(len(roots))
    for i, b in enumerate(moving):
        rmass[root_of[int(M["body_rootid"][b])]] += b_mass[i]
    put("r_invmass", 1.0 / np.maximum(rmass, 1e-300), True)
    # ---- dofs
    d_body = np.zeros(nv, dtype=np.int32)
    for i in range(nb):
        for d in dofs_of[i]:
            d_body[d] = i
    jn = M["dof_jntid"]
    put("d_body", d_body, False)
    put("d_qadr", [M["jnt_qposadr"][jn[d]] + (d - M["jnt_dofadr"][jn[d]]) for d in range(nv)], False)
    put("d_limited", [int(M["jnt_limited"][jn[d]] and M["jnt_type"][jn[d]] != JNT_FREE) for d in range(nv)], False)
    put("d_actfrclimited", [M["jnt_actfrclimited"][jn[d]] for d in range(nv)], False)
    put("d_actgravcomp", [M["jnt_actgravcomp"][jn[d]] for d in range(nv)], False)
    premask, dotzero, ancd = [], [], []
    for d in range(nv):
        i = d_body[d]
        j = jid[i]
        anc_dofs = sum(1 << x for k in range(nb) if (anc[i] >> k) & 1 and k != i for x in dofs_of[k])
        if M["jnt_type"][j] == JNT_FREE:
            a = d - M["jnt_dofadr"][j]
            dotzero.append(int(a < 3))
            premask.append(anc_dofs | sum(1 << (M["jnt_dofadr"][j] + t) for t in range(3)))
            ancd.append(anc_dofs | sum(1 << x for x in dofs_of[i] if x <= d))
        else:
            dotzero.append(0)
            premask.append(anc_dofs)
            ancd.append(anc_dofs | (1 << d))
    put("d_dotzero", dotzero, False)
    put("d_premask", np.array(premask, dtype=np.uint32).view(np.int32), False)
    put("d_ancmask", np.array(ancd, dtype=np.uint32).view(np.int32), False)
    put("d_armature", M["dof_armature"], True)
    put("d_damping", M["dof_damping"], True)
    put("d_frictionloss", M["dof_frictionloss"], True)
    put("d_invweight0", M["dof_invweight0"], True)
    put("d_range", [M["jnt_range"][jn[d]] for d in range(nv)], True)
    put("d_margin", [M["jnt_margin"][jn[d]] for d in range(nv)], True)
    put("d_solref", [M["jnt_solref"][jn[d]] for d in range(nv)], True)
    put("d_solimp", [M["jnt_solimp"][jn[d]] for d in range(nv)], True)
    put("d_actfrcrange", [M["jnt_actfrcrange"][jn[d]] for d in range(nv)], True)
    put("qpos0", M["qpos0"], True)
    # ---- collidable geoms
    col = [g for g in range(M["ngeom"]) if M["geom_contype"][g] or M["geom_conaffinity"][g]]
    used = sorted({int(g) for pr in M["pair_geom"] for g in pr})
    col = [g for g in col if g in used]
    cg_of = {g: i for i, g in enumerate(col)}
    g_pos, g_quat, g_body = [], [], []
    for g in col:
        b = int(M["geom_bodyid"][g])
        o = own[b]
        if b in mb_of:
            pp, pq = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            pp, pq = rel_pos[b], rel_quat[b]
        g_pos.append(pp + quat_to_mat(pq) @ M["geom_pos"][g])
        q = quat_mul(pq, M["geom_quat"][g])
        g_quat.append(q / np.linalg.norm(q))
        g_body.append(mb_of[o] if o >= 0 else -1)
    put("g_body", g_body, False)
    put("g_type", [M["geom_type"][g] for g in col], False)
    # vertex pool restricted to used meshes
    vadr, vnum, pool, n = [], [], [], 0
    for g in col:
        if M["geom_vertnum"][g] > 0:
            a, k = int(M["geom_vertadr"][g]), int(M["geom_vertnum"][g])
            vadr.append(n)
            vnum.append(k)
            pool.append(M["mesh_vert"][a:a + k])
            n += k
        else:
            vadr.append(0)
            vnum.append(0)
    verts = np.concatenate(pool, axis=0) if pool else np.zeros((1, 3))
    put("g_vertadr", vadr, False)
    put("g_vertnum", vnum, False)
    put("g_origid", col, False)
    put("g_condim", [M["geom_condim"][g] for g in col], False)
    put("g_priority", [M["geom_priority"][g] for g in col], False)
    put("g_pos", g_pos, True)
    put("g_quat", g_quat, True)
    # bounding volume: sphere about the local AABB centre (broad phase) and the oriented AABB itself (mid phase)
    put("g_bpos", [gp + quat_to_mat(gq) @ M["geom_aabb"][g][:3] for g, gp, gq in zip(col, g_pos, g_quat)], True)
    put("g_rbound", [M["geom_bsphere"][g][3] for g in col], True)
    put("g_rbound0", [M["geom_rbound"][g] for g in col], True)
    for f in ("size", "aabb", "friction", "solref", "solimp", "solmix", "margin", "gap"):
        put("g_" + f, [M["geom_" + f][g] for g in col], True)
    put("g_invweight", [M["body_invweight0"][M["geom_bodyid"][g]][0] for g in col], True)
    pairs = []
    for g1, g2 in M["pair_geom"]:
        g1, g2 = int(g1), int(g2)
        if M["geom_type"][g1] > M["geom_type"][g2]:
            g1, g2 = g2, g1
        pairs.append((cg_of[g1], cg_of[g2]))
    put("pair", np.array(pairs, dtype=np.int32).reshape(-1, 2), False)
    # ---- tendons / equalities / actuators
    put("t_coef", np.pad(M["tendon_coef"], ((0, 0), (0, 16 - nv))) if M["ntendon"] else np.zeros((0, 16)), True)
    put("e_dof1", [M["jnt_dofadr"][j] for j in M["eq_obj1id"]], False)
    put("e_dof2", [M["jnt_dofadr"][j] if j >= 0 else -1 for j in M["eq_obj2id"]], False)
    put("e_active", M["eq_active0"], False)
    put("e_poly", M["eq_polycoef"], True)
    put("e_solref", M["eq_solref"], True)
    put("e_solimp", M["eq_solimp"], True)
    put("a_trntype", M["actuator_trntype"], False)
    put("a_trnid", [M["jnt_dofadr"][t] if tt == mjcf.TRN_JOINT else t
                    for t, tt in zip(M["actuator_trnid"], M["actuator_trntype"])], False)
    put("a_ctrllimited", M["actuator_ctrllimited"], False)
    put("a_forcelimited", M["actuator_forcelimited"], False)
    # implicitfast damping derivative: fold never-clamped joint actuators into a per-dof constant
    kvdiag, special = np.zeros(nv), []
    for a in range(nu):
        if M["actuator_trntype"][a] == mjcf.TRN_JOINT and not M["actuator_forcelimited"][a]:
            kvdiag[M["jnt_dofadr"][M["actuator_trnid"][a]]] += M["actuator_biasprm"][a][2] * M["actuator_gear"][a] ** 2
        else:
            special.append(a)
    put("d_kvdiag", kvdiag, True)
    put("n_special", [len(special)], False)
    put("a_special", special if special else [0], False)
    put("a_gear", M["actuator_gear"], True)
    put("a_gain", M["actuator_gainprm"][:, 0], True)
    put("a_bias", M["actuator_biasprm"], True)
    put("a_ctrlrange", M["actuator_ctrlrange"], True)
    put("a_forcerange", M["actuator_forcerange"], True)
    # ---- robot / gripper device layer
    roles = np.zeros(len(col), dtype=np.int32)
    if robot_cfg is not None:
        for nme in robot_cfg.arm_collision_geoms:
            gid = _name_id(M["geom_names"], nme, "geom")
            if gid in cg_of:
                roles[cg_of[gid]] |= ROLE_ARM
        sid = _name_id(M["site_names"], robot_cfg.attachment_site, "site")
        bid = _name_id(M["body_names"], robot_cfg.base, "body")
        jids = [_name_id(M["jnt_names"], nme, "joint") for nme in robot_cfg.joints]
        aids = [_name_id(M["actuator_names"], nme, "actuator") for nme in robot_cfg.actuators]
        sb = int(M["site_bodyid"][sid])
        so = own[sb]
        if sb in mb_of:
            pp, pq = np.zeros(3), np.array([1.0, 0, 0, 0])
        else:
            pp, pq = rel_pos[sb], rel_quat[sb]
        put("rb_njoints", [len(jids)], False)
        put("rb_qadr", [M["jnt_qposadr"][j] for j in jids], False)
        put("rb_act", aids, False)
        put("rb_site_body", [mb_of[so] if so >= 0 else -1], False)
        put("rb_register_convergence", [int(getattr(robot_cfg, "register_convergence_callback", True))], False)
        put("rb_ik_nq", [int(getattr(robot_cfg, "ik_nq", min(nq, 9)))], False)
        put("rb_site_pos", pp + quat_to_mat(pq) @ M["site_pos"][sid], True)
        sq = quat_mul(pq, M["site_quat"][sid])
        put("rb_site_quat", sq / np.linalg.norm(sq), True)
        if own[bid] != -1:
            raise NotImplementedError("robot base on a moving body")
        put("rb_base_pos", rel_pos[bid], True)
        put("rb_base_quat", rel_quat[bid], True)
        put("rb_tcp_offset", np.asarray(robot_cfg.tcp_offset, dtype=np.float64), True)
        put("rb_q_home", np.asarray(robot_cfg.q_home, dtype=np.float64), True)
        put("rb_joint_tol", [robot_cfg.joint_rotational_tolerance], True)
        put("rb_cb_period", [robot_cfg.seconds_between_callbacks], True)
    put("gr_enabled", [int(gripper_cfg is not None)], False)
    if gripper_cfg is not None:
        put("gr_act", [_name_id(M["actuator_names"], gripper_cfg.actuator, "actuator")], False)
        put("gr_qadr", [M["jnt_qposadr"][_name_id(M["jnt_names"], gripper_cfg.joint, "joint")]], False)
        for nme in gripper_cfg.collision_geoms:
            gid = _name_id(M["geom_names"], nme, "geom")
            if gid in cg_of:
                roles[cg_of[gid]] |= ROLE_GRIPPER
        for nme in gripper_cfg.collision_geoms_fingers:
            gid = _name_id(M["geom_names"], nme, "geom")
            if gid in cg_of:
                roles[cg_of[gid]] |= ROLE_FINGER
        for nme in gripper_cfg.ignored_collision_geoms:
            gid = _name_id(M["geom_names"], nme, "geom")
            if gid in cg_of:
                roles[cg_of[gid]] |= ROLE_IGNORED
        for f, v in (("gr_eps_inner", gripper_cfg.epsilon_inner), ("gr_eps_outer", gripper_cfg.epsilon_outer),
                     ("gr_cb_period", gripper_cfg.seconds_between_callbacks),
                     ("gr_max_act", gripper_cfg.max_actuator_width), ("gr_min_act", gripper_cfg.min_actuator_width),
                     ("gr_max_joint", gripper_cfg.max_joint_width), ("gr_min_joint", gripper_cfg.min_joint_width)):
            put(f, [v], True)
    put("g_role", roles, False)
    # ---- sizes / options
    has_free = bool((np.asarray(M["jnt_type"]) == 0).any())
    mc = int(maxcon if maxcon is not None else (MAXCON_FREE if has_free else MAXCON_DEFAULT))
    nlim = int(sum(F["d_limited"][0]))
    nfl = int((M["dof_frictionloss"] > 0).sum())
    rows_per_con = 3 if M["opt_cone"] == "elliptic" else 4
    for f, v in (("nq", nq), ("nv", nv), ("nu", nu), ("nb", nb), ("ng", len(col)), ("npair", len(pairs)),
                 ("nt", M["ntendon"]), ("neq", M["neq"]), ("nroot", len(roots)), ("nmeshvert", len(verts)),
                 ("cone_elliptic", int(M["opt_cone"] == "elliptic")),
                 ("implicitfast", int(M["opt_integrator"] == "implicitfast")), ("iterations", M["opt_iterations"]),
                 ("ls_iterations", M["opt_ls_iterations"]), ("noslip_iterations", M["opt_noslip_iterations"]),
                 ("maxcon", mc), ("maxefc", M["neq"] + nfl + nlim + rows_per_con * mc)):
        put(f, [v], False)
    # Reduced-capacity workspace layout (rcsb_types.h: fast_maxcon): environments whose contacts / active limits fit it
    # run at several times the warps per SM; the others are finished by a second launch in the full layout. Scenes with
    # free bodies rest on contacts all the time: their reduced layout holds the resting contacts of one object.
    fmc = int(fast_maxcon) if fast_maxcon is not None else (FAST_MAXCON_FREE if has_free else FAST_MAXCON_DEFAULT)
    fmc = min(fmc, mc)
    put("fast_maxcon", [fmc], False)
    put("fast_maxefc", [M["neq"] + nfl + min(nlim, FAST_LIMIT_ROWS) + rows_per_con * fmc if fmc > 0 else 0], False)
    for f, v in (("timestep", M["opt_timestep"]), ("impratio", M["opt_impratio"]), ("tolerance", M["opt_tolerance"]),
                 ("ls_tolerance", M["opt_ls_tolerance"]), ("noslip_tolerance", M["opt_noslip_tolerance"]),
                 ("meaninertia", M["stat_meaninertia"])):
        put(f, [v], True)
    put("gravity", M["opt_gravity"], True)
    return F, np.ascontiguousarray(verts, dtype=np.float64)
