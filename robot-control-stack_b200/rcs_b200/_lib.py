"""ctypes binding of the CUDA library csrc/librcsb.so (C ABI: include/rcsb.h).

There is no CPU fallback: if the library is missing or no CUDA device is usable, loading or the first
device call raises. The library is built in-tree by `make -C robot-control-stack_b200/csrc`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCSB_LIB_PATH") or os.path.abspath(os.path.join(_HERE, "..", "csrc", "librcsb.so"))
_LIB = None


class RcsbError(RuntimeError):
    pass


# op bits (include/rcsb.h)
GRIPPER_RESET, SIM_RESET, ROBOT_RESET, ENV_RESET_FLAGS = 1, 2, 4, 8
ACT_JOINTS_REL, ACT_JOINTS_ABS, ACT_GRIPPER_BIN = 16, 32, 64
SET_JOINTS, SET_GRIPPER, SET_JOINTS_HARD = 128, 256, 512
STEP_K, STEP_CONV, OBS = 1024, 2048, 4096
ACT_GRIPPER_CONT = 8192


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RcsbError(f"CUDA library not built: {LIB_PATH} (run `make -C robot-control-stack_b200/csrc`); "
                        "this backend has no CPU path")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, cp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p
    L.rcsb_last_error.restype = cp
    L.rcsb_launch_count.restype = C.c_longlong
    L.rcsb_model_new.restype = vp
    L.rcsb_model_free.argtypes = [vp]
    L.rcsb_model_set_int.argtypes = [vp, cp, ip, C.c_int]
    L.rcsb_model_set_real.argtypes = [vp, cp, dp, C.c_int]
    L.rcsb_model_set_mesh_vertices.argtypes = [vp, dp, C.c_int]
    L.rcsb_model_set_mesh_graph.argtypes = [vp, ip, C.c_int, ip, C.c_int]
    L.rcsb_model_set_mesh_faces.argtypes = [vp, dp, C.c_int, ip, ip, C.c_int]
    L.rcsb_camera_depth.argtypes = [vp, C.c_int, dp, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp]
    L.rcsb_body_frames.argtypes = [vp, vp]
    L.rcsb_model_finalize.argtypes = [vp]
    L.rcsb_model_upload.argtypes = [vp, C.c_int]
    L.rcsb_model_dims.argtypes = [vp, ip, ip, ip, ip, ip]
    L.rcsb_model_offsets.argtypes = [vp, ip, ip, ip, ip, ip]
    L.rcsb_model_workspace_bytes.argtypes = [vp, ip, ip, ip]
    L.rcsb_batch_new.restype = vp
    L.rcsb_batch_new.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.rcsb_batch_free.argtypes = [vp]
    L.rcsb_batch_init_state.argtypes = [vp]
    L.rcsb_batch_set_contact_export.argtypes = [vp, vp, vp, vp, C.c_int]
    L.rcsb_batch_run.argtypes = [vp, C.c_uint, C.c_int, C.c_int, vp, vp, vp, C.c_double, dp, dp, vp, vp]
    L.rcsb_batch_run_host.argtypes = [vp, C.c_uint, C.c_int, C.c_int, vp, vp, C.c_double, dp, dp, vp, vp]
    L.rcsb_env_step_host.argtypes = [vp, C.c_uint, C.c_int, C.c_int, vp, C.c_double, dp, dp, vp]
    L.rcsb_sim_step.argtypes = [vp, C.c_int]
    L.rcsb_sim_step_until_convergence.argtypes = [vp, C.c_int]
    for f in ("rcsb_sim_reset", "rcsb_robot_reset", "rcsb_gripper_reset"):
        getattr(L, f).argtypes = [vp]
    for f in ("rcsb_robot_set_joint_position", "rcsb_robot_set_joints_hard", "rcsb_gripper_set_normalized_width",
              "rcsb_robot_set_cartesian_position"):
        getattr(L, f).argtypes = [vp, vp]
    L.rcsb_env_get_obs.argtypes = [vp, vp, vp]
    L.rcsb_ik_inverse.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rcsb_env_cartesian_action.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double]
    L.rcsb_env_cartesian_action_origin.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    L.rcsb_kernel_occupancy.argtypes = [vp, ip, ip, ip]
    L.rcsb_kernel_variant.argtypes = [vp, C.c_int]
    L.rcsb_kernel_variant.restype = cp
    _LIB = L
    return L


def check(rc: int):
    if rc != 0:
        raise RcsbError(f"rcsb error {rc}: {lib().rcsb_last_error().decode()}")


EXPORTS = ["rcsb_last_error", "rcsb_version", "rcsb_real_bytes", "rcsb_model_new", "rcsb_model_free",
           "rcsb_model_set_int", "rcsb_model_set_real", "rcsb_model_set_mesh_vertices", "rcsb_model_set_mesh_graph", "rcsb_model_set_mesh_faces", "rcsb_model_finalize", "rcsb_camera_depth", "rcsb_body_frames", "rcsb_batch_info", "rcsb_batch_read_row",
           "rcsb_robot_set_cartesian_position_host", "rcsb_ik_inverse_host",
           "rcsb_model_upload", "rcsb_model_dims", "rcsb_model_offsets", "rcsb_model_workspace_bytes", "rcsb_batch_new", "rcsb_batch_free",
           "rcsb_batch_init_state", "rcsb_batch_set_contact_export", "rcsb_batch_run", "rcsb_batch_run_host", "rcsb_env_step_host", "rcsb_sim_step",
           "rcsb_sim_step_until_convergence", "rcsb_sim_reset", "rcsb_robot_set_joint_position",
           "rcsb_robot_set_joints_hard", "rcsb_robot_reset", "rcsb_gripper_set_normalized_width",
           "rcsb_gripper_reset", "rcsb_env_get_obs", "rcsb_ik_inverse", "rcsb_robot_set_cartesian_position", "rcsb_env_cartesian_action", "rcsb_env_cartesian_action_origin",
           "rcsb_launch_count", "rcsb_kernel_occupancy", "rcsb_kernel_variant", "rcsb_debug_stage_cycles", "rcsb_debug_stage_trace"]
