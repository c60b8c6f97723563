"""TEST INFRASTRUCTURE ONLY: ctypes driver of tests/emu/libemu.so (host emulation of the device code)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "robot-control-stack_b200", "csrc")


def build(reverse=False):
    so = os.path.join(_HERE, "libemu_rev.so" if reverse else "libemu.so")
    srcs = [os.path.join(_HERE, "emu.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unused-function", "-o", so, os.path.join(_HERE, "emu.cpp")]
        if reverse:
            cmd.insert(1, "-DRCSB_EMU_REVERSE=1")
        subprocess.check_call(cmd)
    return so


OPS = dict(GRIPPER_RESET=1, SIM_RESET=2, ROBOT_RESET=4, ENV_RESET_FLAGS=8, ACT_JOINTS_REL=16, ACT_JOINTS_ABS=32,
           ACT_GRIPPER_BIN=64, SET_JOINTS=128, SET_GRIPPER=256, SET_JOINTS_HARD=512, STEP_K=1024, STEP_CONV=2048, OBS=4096, ACT_GRIPPER_CONT=8192)


class Emu:
    def __init__(self, fields, verts, N, reverse=False, use_reduced=True, graph=None):
        L = C.CDLL(build(reverse))
        self.L = L
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.emu_model_new.restype = vp
        L.emu_model_set_int.argtypes = [vp, C.c_char_p, ip, C.c_int]
        L.emu_model_set_real.argtypes = [vp, C.c_char_p, dp, C.c_int]
        L.emu_model_finalize.argtypes = [vp, C.c_int]
        L.emu_nsr.argtypes = [vp]
        L.emu_offset.argtypes = [vp, C.c_char_p, C.c_int]
        L.emu_has_reduced.argtypes = [vp]
        L.emu_run.argtypes = [vp, dp, dp, dp, ip, C.c_int, C.c_uint, C.c_int, C.c_int, dp, dp, C.c_void_p, C.c_double,
                              dp, dp, dp, ip, dp, ip]
        self.m = L.emu_model_new()
        for name, (arr, is_real) in fields.items():
            a = np.ascontiguousarray(arr).ravel()
            if is_real:
                rc = L.emu_model_set_real(self.m, name.encode(), a.ctypes.data_as(dp), a.size)
            else:
                rc = L.emu_model_set_int(self.m, name.encode(), a.ctypes.data_as(ip), a.size)
            assert rc == 0, (name, rc)
        assert L.emu_model_finalize(self.m, int(use_reduced)) == 0
        self.has_reduced = bool(L.emu_has_reduced(self.m))
        sz = (C.c_int * 6)()
        L.emu_sizes(sz)
        self.S_TAIL, self.D_TAIL, self.I_TAIL, self.OBS_DIM, self.INFO_DIM, self.real_bytes = list(sz)
        assert self.real_bytes == 8
        self.nsr = L.emu_nsr(self.m)
        self.N = N
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        # hull edge graph [nvert + 1 offsets | neighbour lists] (devmodel.build_mesh_graph); None = single-point plane-mesh
        self.graph = None if graph is None else np.ascontiguousarray(np.concatenate([graph[0], graph[1][:graph[0][-1]]]), dtype=np.int32)
        self.sr = np.zeros((N, self.nsr))
        self.sd = np.zeros((N, self.D_TAIL))
        self.si = np.zeros((N, self.I_TAIL), dtype=np.int32)
        self.si[:, 0] = 1  # ik_success = true (SimRobotState default)
        self.obs = np.zeros((N, self.OBS_DIM))
        self.info = np.zeros((N, self.INFO_DIM), dtype=np.int32)
        self.ws = np.zeros((N, self.off("ws_reals")))
        self.layout = np.zeros(N, dtype=np.int32)  # 1 = the env last ran in the reduced layout
        self.handed_over = 0                        # envs the reduced layout passed to the full one in the last run

    def enable_contact_export(self, cap=8):
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        self.contact_n = np.zeros(self.N, dtype=np.int32)
        self.contact_geom = np.full((self.N, cap, 2), -1, dtype=np.int32)
        self.contact_real = np.zeros((self.N, cap, 7))
        self.contact_cap = cap

    def off(self, name, reduced=0):
        return self.L.emu_offset(self.m, name.encode(), int(reduced))

    def run(self, ops, k=0, max_conv=500, act_joints=None, act_gripper=None, max_mov=0.0, jlow=None, jhigh=None):
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        code = 0
        for o in ops:
            code |= OPS[o]
        aj = np.ascontiguousarray(act_joints, dtype=np.float64) if act_joints is not None else None
        ag = np.ascontiguousarray(act_gripper, dtype=np.float64) if act_gripper is not None else None
        self.L.emu_set_mesh_graph.argtypes = [ip]
        self.L.emu_set_mesh_graph(self.graph.ctypes.data_as(ip) if self.graph is not None else None)
        self.L.emu_set_contact_export.argtypes = [ip, ip, dp, C.c_int]
        if getattr(self, "contact_cap", 0):  # the library keeps one global export target: set it for this instance's run
            self.L.emu_set_contact_export(self.contact_n.ctypes.data_as(ip), self.contact_geom.ctypes.data_as(ip),
                                          self.contact_real.ctypes.data_as(dp), self.contact_cap)
        else:
            self.L.emu_set_contact_export(None, None, None, 0)
        lo = np.zeros(8); hi = np.zeros(8)
        if jlow is not None:
            lo[:len(jlow)] = jlow; hi[:len(jhigh)] = jhigh
        self.handed_over = self.L.emu_run(self.m, self.verts.ctypes.data_as(dp), self.sr.ctypes.data_as(dp), self.sd.ctypes.data_as(dp),
                       self.si.ctypes.data_as(ip), self.N, code, k, max_conv,
                       aj.ctypes.data_as(dp) if aj is not None else None, ag.ctypes.data_as(dp) if ag is not None else None,
                       None, float(max_mov), lo.ctypes.data_as(dp), hi.ctypes.data_as(dp), self.obs.ctypes.data_as(dp),
                       self.info.ctypes.data_as(ip), self.ws.ctypes.data_as(dp), self.layout.ctypes.data_as(ip))

    def wsf(self, name, n, env=0):
        o = self.off(name, self.layout[env])
        return self.ws[env, o:o + n]
