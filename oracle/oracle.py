"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/librcs_oracle.so).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference).
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_INT_FIELDS = ["body_parentid", "body_rootid", "body_weldid", "body_jntnum", "body_jntadr", "body_dofnum",
               "body_dofadr", "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_limited",
               "jnt_actfrclimited", "jnt_actgravcomp", "dof_jntid", "dof_bodyid", "dof_parentid", "geom_type",
               "geom_bodyid", "geom_condim", "geom_priority", "geom_vertadr", "geom_vertnum", "pair_geom", "mesh_graphadr", "mesh_graph",
               "site_bodyid", "eq_obj1id", "eq_obj2id", "eq_active0", "actuator_trntype", "actuator_trnid",
               "actuator_ctrllimited", "actuator_forcelimited"]
_REAL_FIELDS = ["body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_gravcomp",
                "body_invweight0", "jnt_pos", "jnt_axis", "jnt_range", "jnt_margin", "jnt_solref", "jnt_solimp",
                "jnt_actfrcrange", "dof_armature", "dof_damping", "dof_frictionloss", "dof_invweight0", "qpos0",
                "geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_solref", "geom_solimp", "geom_solmix",
                "geom_margin", "geom_gap", "geom_rbound", "geom_aabb", "geom_bsphere", "mesh_vert", "site_pos", "site_quat",
                "tendon_coef", "tendon_invweight0", "eq_polycoef", "eq_solref", "eq_solimp", "actuator_gear",
                "actuator_gainprm", "actuator_biasprm", "actuator_ctrlrange", "actuator_forcerange"]


class RobotCfg(C.Structure):
    _fields_ = [("njoints", C.c_int), ("joint_qposadr", C.c_int * 8), ("actuator_id", C.c_int * 8),
                ("attachment_site", C.c_int), ("base_body", C.c_int), ("ncgeom", C.c_int), ("cgeom", C.c_int * 16),
                ("q_home", C.c_double * 8), ("joint_rotational_tolerance", C.c_double),
                ("seconds_between_callbacks", C.c_double), ("tcp_offset", C.c_double * 7),
                ("register_convergence_callback", C.c_int), ("ik_nq", C.c_int)]


class GripperCfg(C.Structure):
    _fields_ = [("enabled", C.c_int), ("actuator_id", C.c_int), ("joint_qposadr", C.c_int), ("ncgeom", C.c_int),
                ("cgeom", C.c_int * 8), ("ncfgeom", C.c_int), ("cfgeom", C.c_int * 4), ("nignored", C.c_int),
                ("ignored", C.c_int * 8), ("epsilon_inner", C.c_double), ("epsilon_outer", C.c_double),
                ("seconds_between_callbacks", C.c_double), ("max_actuator_width", C.c_double),
                ("min_actuator_width", C.c_double), ("max_joint_width", C.c_double), ("min_joint_width", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "librcs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.rcso_model_new.restype = vp
        L.rcso_model_free.argtypes = [vp]
        L.rcso_model_set_int.argtypes = [vp, C.c_char_p, ip, C.c_int]
        L.rcso_model_set_real.argtypes = [vp, C.c_char_p, dp, C.c_int]
        L.rcso_model_finalize.argtypes = [vp]
        L.rcso_data_new.restype = vp
        L.rcso_data_new.argtypes = [vp]
        L.rcso_data_free.argtypes = [vp]
        L.rcso_data_real.restype = dp
        L.rcso_data_real.argtypes = [vp, C.c_char_p, ip]
        L.rcso_data_int.restype = ip
        L.rcso_data_int.argtypes = [vp, C.c_char_p, ip]
        for f in ("rcso_reset_data", "rcso_step1", "rcso_step2", "rcso_step", "rcso_forward"):
            getattr(L, f).argtypes = [vp, vp]
        L.rcso_sim_new.restype = vp
        L.rcso_sim_new.argtypes = [vp, C.POINTER(RobotCfg), C.POINTER(GripperCfg)]
        L.rcso_sim_free.argtypes = [vp]
        L.rcso_sim_data.restype = vp
        L.rcso_sim_data.argtypes = [vp]
        L.rcso_sim_set_config.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.rcso_sim_step.argtypes = [vp, C.c_int]
        for f in ("rcso_sim_step_until_convergence", "rcso_sim_reset", "rcso_robot_reset", "rcso_gripper_reset"):
            getattr(L, f).argtypes = [vp]
        L.rcso_sim_is_converged.argtypes = [vp]
        L.rcso_sim_convergence_steps.argtypes = [vp]
        L.rcso_robot_set_joint_position.argtypes = [vp, dp]
        L.rcso_robot_get_joint_position.argtypes = [vp, dp]
        L.rcso_robot_get_cartesian_position.argtypes = [vp, dp]
        L.rcso_robot_set_cartesian_position.argtypes = [vp, dp]
        L.rcso_robot_state.argtypes = [vp, ip, ip, ip, ip, dp, dp]
        L.rcso_gripper_set_normalized_width.argtypes = [vp, C.c_double, C.c_double]
        L.rcso_gripper_get_normalized_width.restype = C.c_double
        L.rcso_gripper_get_normalized_width.argtypes = [vp]
        L.rcso_gripper_is_grasped.argtypes = [vp]
        L.rcso_gripper_state.argtypes = [vp, dp, ip, dp, ip]
        L.rcso_ik_inverse.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, ip]
        L.rcso_ik_forward.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int, dp, dp]
        L.rcso_pose_mul.argtypes = [dp, dp, dp]
        L.rcso_pose_inverse.argtypes = [dp, dp]
        L.rcso_pose_from_rpy.argtypes = [dp, dp, dp]
        L.rcso_pose_from_matrix.argtypes = [dp, dp, dp]
        L.rcso_pose_xyzrpy.argtypes = [dp, dp]
        L.rcso_pose_rotation_m.argtypes = [dp, dp]
        L.rcso_pose_total_angle.restype = C.c_double
        L.rcso_pose_total_angle.argtypes = [dp]
        L.rcso_pose_limit_rotation_angle.argtypes = [dp, C.c_double, dp]
        L.rcso_pose_limit_translation_length.argtypes = [dp, C.c_double, dp]
        L.rcso_pose_interpolate.argtypes = [dp, dp, C.c_double, dp]
        L.rcso_pose_is_close.argtypes = [dp, dp, C.c_double, C.c_double]
        L.rcso_bench_env_steps.restype = C.c_double
        L.rcso_bench_env_steps.argtypes = [vp, C.POINTER(RobotCfg), C.POINTER(GripperCfg), C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, dp, C.c_double, dp, dp, C.POINTER(C.c_longlong), dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Model:
    """Oracle model built from the compiled-scene dict of rcs_b200.mjcf (passed in by the caller: the
    oracle itself does not import the product package)."""

    def __init__(self, M: dict):
        L = lib()
        self.M = M
        self.ptr = L.rcso_model_new()
        sizes = np.array([M["nq"], M["nv"], M["nu"], M["nbody"], M["njnt"], M["ngeom"], M["nsite"], M["ntendon"],
                          M["neq"], len(M["pair_geom"]), len(M["mesh_vert"])], dtype=np.int32)
        assert L.rcso_model_set_int(self.ptr, b"sizes", _ip(sizes), 11) == 0
        oi = np.array([M["opt_iterations"], M["opt_ls_iterations"], M["opt_noslip_iterations"],
                       1 if M["opt_cone"] == "elliptic" else 0, 1 if M["opt_integrator"] == "implicitfast" else 0],
                      dtype=np.int32)
        assert L.rcso_model_set_int(self.ptr, b"opt_int", _ip(oi), 5) == 0
        orr = np.array([M["opt_timestep"], *M["opt_gravity"], M["opt_impratio"], M["opt_tolerance"],
                        M["opt_noslip_tolerance"], M["opt_ls_tolerance"], M["stat_meaninertia"]], dtype=np.float64)
        assert L.rcso_model_set_real(self.ptr, b"opt_real", _dp(orr), 9) == 0
        for f in _INT_FIELDS:
            a = np.ascontiguousarray(np.asarray(M[f]), dtype=np.int32).ravel()
            assert L.rcso_model_set_int(self.ptr, f.encode(), _ip(a), a.size) == 0, f
        for f in _REAL_FIELDS:
            a = np.ascontiguousarray(np.asarray(M[f]), dtype=np.float64).ravel()
            assert L.rcso_model_set_real(self.ptr, f.encode(), _dp(a), a.size) == 0, f
        rc = L.rcso_model_finalize(self.ptr)
        if rc != 0:
            raise RuntimeError(f"oracle model finalize failed ({rc})")

    def __del__(self):
        try:
            lib().rcso_model_free(self.ptr)
        except Exception:
            pass


class Data:
    def __init__(self, model: Model, ptr=None):
        self.model = model
        self._own = ptr is None
        self.ptr = lib().rcso_data_new(model.ptr) if ptr is None else ptr

    def __del__(self):
        if self._own:
            try:
                lib().rcso_data_free(self.ptr)
            except Exception:
                pass

    def real(self, name: str) -> np.ndarray:
        n = C.c_int(0)
        p = lib().rcso_data_real(self.ptr, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(max(n.value, 0),))

    def int(self, name: str) -> np.ndarray:
        n = C.c_int(0)
        p = lib().rcso_data_int(self.ptr, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(max(n.value, 0),)).copy()

    def __getattr__(self, name):
        if name in ("model", "ptr", "_own"):
            raise AttributeError(name)
        try:
            return self.real(name)
        except KeyError:
            return self.int(name)

    @property
    def time(self):
        return float(self.real("time")[0])

    def step(self, k=1):
        for _ in range(k):
            lib().rcso_step(self.model.ptr, self.ptr)

    def step1(self):
        lib().rcso_step1(self.model.ptr, self.ptr)

    def step2(self):
        lib().rcso_step2(self.model.ptr, self.ptr)

    def forward(self):
        lib().rcso_forward(self.model.ptr, self.ptr)

    def reset(self):
        lib().rcso_reset_data(self.model.ptr, self.ptr)


def robot_cfg(M: dict, idx="0", tcp_offset=None, register_convergence_callback=True, robot="fr3") -> RobotCfg:
    """Name->id resolution of SimRobot::init_ids (/root/reference/src/sim/SimRobot.cpp:52-94)."""
    rc = RobotCfg()
    sfx = f"_{idx}" if idx is not None else ""
    if robot == "fr3":
        jn = [f"fr3_joint{i}{sfx}" for i in range(1, 8)]
        cg = [f"fr3_link{i}_collision{sfx}" for i in range(8)]
        q_home = [0.0, -np.pi / 4, 0.0, -3 * np.pi / 4, 0.0, np.pi / 2, np.pi / 4]
    else:
        raise NotImplementedError(robot)
    rc.njoints = 7
    for i, n in enumerate(jn):
        rc.joint_qposadr[i] = int(M["jnt_qposadr"][M["jnt_names"].index(n)])
        rc.actuator_id[i] = M["actuator_names"].index(n)
        rc.q_home[i] = q_home[i]
    rc.attachment_site = M["site_names"].index(f"attachment_site{sfx}")
    rc.base_body = M["body_names"].index(f"base{sfx}")
    rc.ncgeom = len(cg)
    for i, n in enumerate(cg):
        rc.cgeom[i] = M["geom_names"].index(n)
    rc.joint_rotational_tolerance = 0.05 * (np.pi / 180.0)
    rc.seconds_between_callbacks = 0.1
    t = [0, 0, 0, 0, 0, 0, 1.0] if tcp_offset is None else list(tcp_offset)
    for i in range(7):
        rc.tcp_offset[i] = t[i]
    rc.register_convergence_callback = int(register_convergence_callback)
    rc.ik_nq = 9
    return rc


def gripper_cfg(M: dict, idx="0", enabled=True) -> GripperCfg:
    """SimGripper ctor name resolution (/root/reference/src/sim/SimGripper.cpp:12-39, SimGripper.h:15-45)."""
    gc = GripperCfg()
    gc.enabled = int(enabled)
    sfx = f"_{idx}"
    if enabled:
        gc.actuator_id = M["actuator_names"].index(f"actuator8{sfx}")
        gc.joint_qposadr = int(M["jnt_qposadr"][M["jnt_names"].index(f"finger_joint1{sfx}")])
        cg = [f"hand_c{sfx}", f"d435i_collision{sfx}", f"finger_0_left{sfx}", f"finger_0_right{sfx}"]
        cf = [f"finger_0_left{sfx}", f"finger_0_right{sfx}"]
        gc.ncgeom, gc.ncfgeom, gc.nignored = len(cg), len(cf), 0
        for i, n in enumerate(cg):
            gc.cgeom[i] = M["geom_names"].index(n)
        for i, n in enumerate(cf):
            gc.cfgeom[i] = M["geom_names"].index(n)
    gc.epsilon_inner = gc.epsilon_outer = 0.005
    gc.seconds_between_callbacks = 0.05
    gc.max_actuator_width, gc.min_actuator_width = 255.0, 0.0
    gc.max_joint_width, gc.min_joint_width = 0.04, 0.0
    return gc


class Sim:
    """Oracle Sim + SimRobot + SimGripper (/root/reference/src/sim/*.cpp) for one environment."""

    def __init__(self, model: Model, rc: RobotCfg, gc: GripperCfg | None = None):
        self.model, self.rc, self.gc = model, rc, gc
        self.ptr = lib().rcso_sim_new(model.ptr, C.byref(rc), C.byref(gc) if gc is not None else None)
        self.data = Data(model, lib().rcso_sim_data(self.ptr))

    def __del__(self):
        try:
            lib().rcso_sim_free(self.ptr)
        except Exception:
            pass

    def set_config(self, async_control=False, frequency=30, max_convergence_steps=500):
        lib().rcso_sim_set_config(self.ptr, int(async_control), frequency, max_convergence_steps)

    def step(self, k):
        lib().rcso_sim_step(self.ptr, k)

    def step_until_convergence(self):
        lib().rcso_sim_step_until_convergence(self.ptr)

    def is_converged(self):
        return bool(lib().rcso_sim_is_converged(self.ptr))

    def convergence_steps(self):
        return lib().rcso_sim_convergence_steps(self.ptr)

    def reset(self):
        lib().rcso_sim_reset(self.ptr)

    def robot_reset(self):
        lib().rcso_robot_reset(self.ptr)

    def gripper_reset(self):
        lib().rcso_gripper_reset(self.ptr)

    def set_joint_position(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        lib().rcso_robot_set_joint_position(self.ptr, _dp(q))

    def get_joint_position(self):
        q = np.zeros(8)
        lib().rcso_robot_get_joint_position(self.ptr, _dp(q))
        return q[:self.rc.njoints].copy()

    def get_cartesian_position(self):
        p = np.zeros(7)
        lib().rcso_robot_get_cartesian_position(self.ptr, _dp(p))
        return p

    def set_cartesian_position(self, pose7):
        p = np.ascontiguousarray(pose7, dtype=np.float64)
        return bool(lib().rcso_robot_set_cartesian_position(self.ptr, _dp(p)))

    def robot_state(self):
        a = [C.c_int(0) for _ in range(4)]
        prev, tgt = np.zeros(8), np.zeros(8)
        lib().rcso_robot_state(self.ptr, *[C.byref(x) for x in a], _dp(prev), _dp(tgt))
        return dict(ik_success=bool(a[0].value), collision=bool(a[1].value), is_moving=bool(a[2].value),
                    is_arrived=bool(a[3].value), previous_angles=prev[:7].copy(), target_angles=tgt[:7].copy())

    def gripper_set_normalized_width(self, w, force=0.0):
        if lib().rcso_gripper_set_normalized_width(self.ptr, w, force) != 0:
            raise ValueError("width must be between 0 and 1, force must be positive")

    def gripper_get_normalized_width(self):
        return lib().rcso_gripper_get_normalized_width(self.ptr)

    def gripper_is_grasped(self):
        return bool(lib().rcso_gripper_is_grasped(self.ptr))

    def gripper_state(self):
        lcw, lw, mv, col = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
        lib().rcso_gripper_state(self.ptr, C.byref(lcw), C.byref(mv), C.byref(lw), C.byref(col))
        return dict(last_commanded_width=lcw.value, is_moving=bool(mv.value), last_width=lw.value,
                    collision=bool(col.value))


def ik_inverse(model: Model, site: int, nq_model: int, pose7, q0, tcp_offset7=(0, 0, 0, 0, 0, 0, 1.0)):
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    tcp = np.ascontiguousarray(tcp_offset7, dtype=np.float64)
    out = np.zeros(64)
    it = C.c_int(0)
    ok = lib().rcso_ik_inverse(model.ptr, site, nq_model, _dp(pose7), _dp(q0), q0.size, _dp(tcp), _dp(out), C.byref(it))
    return (out[:nq_model].copy() if ok else None), it.value


def ik_forward(model: Model, site: int, nq_model: int, q0, tcp_offset7=(0, 0, 0, 0, 0, 0, 1.0)):
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    tcp = np.ascontiguousarray(tcp_offset7, dtype=np.float64)
    out = np.zeros(7)
    lib().rcso_ik_forward(model.ptr, site, nq_model, _dp(q0), q0.size, _dp(tcp), _dp(out))
    return out


def log3(R):
    """pinocchio::log3 as the IK uses it (test hook)"""
    R = np.ascontiguousarray(R, dtype=np.float64).ravel()
    w = np.zeros(3)
    lib().rcso_log3.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib().rcso_log3(_dp(R), _dp(w))
    return w


def _p7(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def pose_mul(a, b):
    o = np.zeros(7)
    lib().rcso_pose_mul(_dp(_p7(a)), _dp(_p7(b)), _dp(o))
    return o


def pose_inverse(a):
    o = np.zeros(7)
    lib().rcso_pose_inverse(_dp(_p7(a)), _dp(o))
    return o


def pose_from_rpy(xyz, rpy):
    o = np.zeros(7)
    lib().rcso_pose_from_rpy(_dp(_p7(xyz)), _dp(_p7(rpy)), _dp(o))
    return o


def pose_from_matrix(R, xyz):
    o = np.zeros(7)
    lib().rcso_pose_from_matrix(_dp(_p7(np.asarray(R).reshape(9))), _dp(_p7(xyz)), _dp(o))
    return o


def pose_xyzrpy(a):
    o = np.zeros(6)
    lib().rcso_pose_xyzrpy(_dp(_p7(a)), _dp(o))
    return o


def pose_rotation_m(a):
    o = np.zeros(9)
    lib().rcso_pose_rotation_m(_dp(_p7(a)), _dp(o))
    return o.reshape(3, 3)


def pose_total_angle(a):
    return lib().rcso_pose_total_angle(_dp(_p7(a)))


def pose_limit_rotation_angle(a, m):
    o = np.zeros(7)
    lib().rcso_pose_limit_rotation_angle(_dp(_p7(a)), m, _dp(o))
    return o


def pose_limit_translation_length(a, m):
    o = np.zeros(7)
    lib().rcso_pose_limit_translation_length(_dp(_p7(a)), m, _dp(o))
    return o


def pose_interpolate(a, b, t):
    o = np.zeros(7)
    lib().rcso_pose_interpolate(_dp(_p7(a)), _dp(_p7(b)), t, _dp(o))
    return o


def pose_is_close(a, b, eps_r=1e-8, eps_t=1e-8):
    return bool(lib().rcso_pose_is_close(_dp(_p7(a)), _dp(_p7(b)), eps_r, eps_t))


def bench_env_steps(model: Model, rc: RobotCfg, gc: GripperCfg, actions: np.ndarray, nthreads: int, episode_len=10,
                    async_control=True, max_mov=np.deg2rad(5), joint_low=None, joint_high=None, want_obs=False):
    """actions: [nenv, nsteps, 8] (7 relative joint moves + gripper). Returns (seconds, physics_steps, obs|None)."""
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    nenv, nsteps, _ = actions.shape
    low = np.ascontiguousarray(joint_low, dtype=np.float64)
    high = np.ascontiguousarray(joint_high, dtype=np.float64)
    ps = C.c_longlong(0)
    obs = np.zeros((nenv, nsteps, 28)) if want_obs else None  # rcs_glue.c ENV_OBS_STRIDE
    sec = lib().rcso_bench_env_steps(model.ptr, C.byref(rc), C.byref(gc), nenv, nthreads, nsteps, episode_len,
                                     int(async_control), _dp(actions), float(max_mov), _dp(low), _dp(high),
                                     C.byref(ps), _dp(obs) if want_obs else None)
    return sec, ps.value, obs
