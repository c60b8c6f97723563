"""Low-level batched device objects over the C ABI: DeviceModel (replaces mjModel) and Batch (replaces
N x mjData + Sim + SimRobot + SimGripper). PyTorch is used only to own device memory and streams; every
state-changing operation is a kernel of csrc/librcsb.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, devmodel


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class DeviceModel:
    def __init__(self, M: dict, robot_cfg=None, gripper_cfg=None, maxcon: int | None = None, device: int = 0,
                 fast_maxcon: int | None = None):
        L = _lib.lib()
        self.M = M
        self.fields, self.verts = devmodel.build_device_fields(M, robot_cfg, gripper_cfg, maxcon, fast_maxcon)
        self.ptr = L.rcsb_model_new()
        for name, (arr, is_real) in self.fields.items():
            a = np.ascontiguousarray(arr).ravel()
            if is_real:
                _lib.check(L.rcsb_model_set_real(self.ptr, name.encode(), _dp(a), a.size))
            else:
                _lib.check(L.rcsb_model_set_int(self.ptr, name.encode(), _ip(a), a.size))
        _lib.check(L.rcsb_model_set_mesh_vertices(self.ptr, _dp(self.verts), len(self.verts)))
        gadr, gnbr = devmodel.build_mesh_graph(M)
        _lib.check(L.rcsb_model_set_mesh_graph(self.ptr, _ip(gadr), len(gadr), _ip(gnbr), int(gadr[-1])))
        planes, fadr, fnum = devmodel.build_mesh_faces(M)
        _lib.check(L.rcsb_model_set_mesh_faces(self.ptr, _dp(planes), int(fnum.sum()), _ip(fadr), _ip(fnum), len(fadr)))
        _lib.check(L.rcsb_model_finalize(self.ptr))
        d = [C.c_int(0) for _ in range(5)]
        _lib.check(L.rcsb_model_dims(self.ptr, *[C.byref(x) for x in d]))
        self.nsr, self.nsd, self.nsi, self.obs_dim, self.info_dim = [x.value for x in d]
        o = [C.c_int(0) for _ in range(5)]
        _lib.check(L.rcsb_model_offsets(self.ptr, *[C.byref(x) for x in o]))
        self.o_qpos, self.o_qvel, self.o_ctrl, self.o_warm, self.o_tail = [x.value for x in o]
        self.nq, self.nv, self.nu = M["nq"], M["nv"], M["nu"]
        self.njoints = int(self.fields["rb_njoints"][0][0]) if "rb_njoints" in self.fields else 0
        self.device = device
        if L.rcsb_real_bytes() != 8:
            raise _lib.RcsbError("float64 build of librcsb.so expected")
        _lib.check(L.rcsb_model_upload(self.ptr, device))  # raises without a CUDA device: no CPU path

    def __del__(self):
        try:
            _lib.lib().rcsb_model_free(self.ptr)
        except Exception:
            pass


class Batch:
    """N environments on one GPU. State lives in three torch tensors (sr, sd, si) that the kernels
    read and write in place; column views of them back the reference-style getters."""

    def __init__(self, model: DeviceModel, n_envs: int, stream: torch.cuda.Stream | None = None):
        self.model, self.n = model, n_envs
        self.dev = torch.device("cuda", model.device)
        self.sr = torch.zeros((n_envs, model.nsr), dtype=torch.float64, device=self.dev)
        self.sd = torch.zeros((n_envs, model.nsd), dtype=torch.float64, device=self.dev)
        self.si = torch.zeros((n_envs, model.nsi), dtype=torch.int32, device=self.dev)
        self.stream = stream
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self.ptr = _lib.lib().rcsb_batch_new(model.ptr, n_envs, self.sr.data_ptr(), self.sd.data_ptr(), self.si.data_ptr(), sp)
        if not self.ptr:
            raise _lib.RcsbError(_lib.lib().rcsb_last_error().decode())
        _lib.check(_lib.lib().rcsb_batch_init_state(self.ptr))
        self.obs = torch.zeros((n_envs, model.obs_dim), dtype=torch.float64, device=self.dev)
        self.info = torch.zeros((n_envs, model.info_dim), dtype=torch.int32, device=self.dev)

    def __del__(self):
        try:
            _lib.lib().rcsb_batch_free(self.ptr)
        except Exception:
            pass

    def enable_contact_export(self, cap: int = 8, with_geometry: bool = True):
        """Keep every environment's contact list (mjData.ncon / contact[i].geom, SimRobot.cpp:172-182) after each
        stepping launch in self.contact_n [n], self.contact_geom [n, cap, 2] and self.contact_real [n, cap, 7]."""
        self.contact_n = torch.zeros((self.n,), dtype=torch.int32, device=self.dev)
        self.contact_geom = torch.full((self.n, cap, 2), -1, dtype=torch.int32, device=self.dev)
        self.contact_real = torch.zeros((self.n, cap, 7), dtype=torch.float64, device=self.dev) if with_geometry else None
        _lib.check(_lib.lib().rcsb_batch_set_contact_export(
            self.ptr, self.contact_n.data_ptr(), self.contact_geom.data_ptr(),
            self.contact_real.data_ptr() if with_geometry else None, cap))

    # ---- column views
    @property
    def qpos(self):
        return self.sr[:, self.model.o_qpos:self.model.o_qpos + self.model.nq]

    @property
    def qvel(self):
        return self.sr[:, self.model.o_qvel:self.model.o_qvel + self.model.nv]

    @property
    def ctrl(self):
        return self.sr[:, self.model.o_ctrl:self.model.o_ctrl + self.model.nu]

    @property
    def qacc_warmstart(self):
        return self.sr[:, self.model.o_warm:self.model.o_warm + self.model.nv]

    @property
    def time(self):
        return self.sd[:, 0]

    def run(self, ops: int, k: int = 0, max_convergence_steps: int = 500, act_joints: torch.Tensor | None = None,
            act_gripper: torch.Tensor | None = None, mask: torch.Tensor | None = None, max_mov: float = 0.0,
            jlow=None, jhigh=None, want_obs: bool = False, fresh_obs: bool = False, obs_out: torch.Tensor | None = None):
        """fresh_obs: pack the observation / info into newly allocated tensors (self.obs / self.info are rebound to
        them), so that results returned to a caller are never overwritten by a later launch. obs_out: a contiguous
        [n, obs_dim] float64 device tensor that receives the packed observation rows instead (e.g. this rank's slice of
        the all-gather buffer); self.obs is rebound to it."""
        if obs_out is not None:
            assert obs_out.shape == self.obs.shape and obs_out.dtype == torch.float64 and obs_out.is_contiguous() and obs_out.device == self.dev
            self.obs, want_obs = obs_out, True
        elif want_obs and fresh_obs:
            self.obs = torch.empty_like(self.obs)
        if want_obs and fresh_obs:
            self.info = torch.empty_like(self.info)
        for t in (act_joints, act_gripper):
            if t is not None:
                assert t.dtype == torch.float64 and t.is_contiguous() and t.device == self.dev
        if mask is not None:
            assert mask.dtype == torch.uint8 and mask.is_contiguous() and mask.device == self.dev
        lo = np.ascontiguousarray(jlow, dtype=np.float64) if jlow is not None else None
        hi = np.ascontiguousarray(jhigh, dtype=np.float64) if jhigh is not None else None
        _lib.check(_lib.lib().rcsb_batch_run(
            self.ptr, ops, k, max_convergence_steps, act_joints.data_ptr() if act_joints is not None else None,
            act_gripper.data_ptr() if act_gripper is not None else None, mask.data_ptr() if mask is not None else None,
            float(max_mov), _dp(lo) if lo is not None else None, _dp(hi) if hi is not None else None,
            self.obs.data_ptr() if want_obs else None, self.info.data_ptr() if want_obs else None))

    def run_host(self, ops: int, k: int, max_convergence_steps: int, act_joints_host: torch.Tensor | None,
                 act_gripper_host: torch.Tensor | None, max_mov: float, jlow, jhigh, obs_host: torch.Tensor | None,
                 info_host: torch.Tensor | None):
        """env.step() through host buffers (pinned CPU tensors): H2D, kernel, D2H, synchronise."""
        lo = np.ascontiguousarray(jlow, dtype=np.float64) if jlow is not None else None
        hi = np.ascontiguousarray(jhigh, dtype=np.float64) if jhigh is not None else None
        _lib.check(_lib.lib().rcsb_batch_run_host(
            self.ptr, ops, k, max_convergence_steps, act_joints_host.data_ptr() if act_joints_host is not None else None,
            act_gripper_host.data_ptr() if act_gripper_host is not None else None, float(max_mov),
            _dp(lo) if lo is not None else None, _dp(hi) if hi is not None else None,
            obs_host.data_ptr() if obs_host is not None else None, info_host.data_ptr() if info_host is not None else None))

    def step_host(self, ops: int, k: int, max_convergence_steps: int, act_host: torch.Tensor, max_mov: float, jlow, jhigh,
                  obs_host: torch.Tensor):
        """env.step() through one packed pinned block each way (rcsb_env_step_host): act_host [n, njoints + 1],
        obs_host [n, obs_dim] (info flags in the last 8 columns)."""
        key = (id(jlow), id(jhigh))
        if getattr(self, "_host_limits_key", None) != key:  # the limits are the same arrays on every step of an env
            lo = np.ascontiguousarray(jlow, dtype=np.float64) if jlow is not None else None
            hi = np.ascontiguousarray(jhigh, dtype=np.float64) if jhigh is not None else None
            self._host_limits = (lo, hi, _dp(lo) if lo is not None else None, _dp(hi) if hi is not None else None, jlow, jhigh)
            self._host_limits_key = key
        _, _, plo, phi, _, _ = self._host_limits
        _lib.check(_lib.lib().rcsb_env_step_host(self.ptr, ops, k, max_convergence_steps, act_host.data_ptr(), float(max_mov),
                                                 plo, phi, obs_host.data_ptr()))

    def body_frames(self) -> torch.Tensor:
        """[n, nb, 12] world frames (position, row-major rotation) of the moving bodies at the current qpos."""
        nb = int(self.model.fields["nb"][0][0])
        out = torch.empty((self.n, nb, 12), dtype=torch.float64, device=self.dev)
        _lib.check(_lib.lib().rcsb_body_frames(self.ptr, out.data_ptr()))
        return out

    def camera_depth(self, cam_body: int, cam_pos, cam_rot, fovy_deg: float, width: int, height: int, znear: float, zfar: float,
                     physical_units: bool, out: torch.Tensor | None = None, cam_frames: torch.Tensor | None = None) -> torch.Tensor:
        """Depth frame of every environment for one camera (rcsb_camera_depth): [n, height, width] uint16. cam_frames
        ([n, 12] float64, optional) receives the camera's world frame of every environment."""
        if out is None:
            out = torch.empty((self.n, height, width), dtype=torch.uint16, device=self.dev)
        assert out.shape == (self.n, height, width) and out.dtype == torch.uint16 and out.is_contiguous()
        p = np.ascontiguousarray(cam_pos, dtype=np.float64)
        r = np.ascontiguousarray(cam_rot, dtype=np.float64).reshape(9)
        _lib.check(_lib.lib().rcsb_camera_depth(self.ptr, int(cam_body), _dp(p), _dp(r), float(fovy_deg), int(width), int(height),
                                                float(znear), float(zfar), int(bool(physical_units)), out.data_ptr(),
                                                cam_frames.data_ptr() if cam_frames is not None else None))
        return out

    def ik_inverse(self, pose: torch.Tensor, q0: torch.Tensor):
        nqm = int(self.model.fields["rb_ik_nq"][0][0])
        q = torch.zeros((self.n, nqm), dtype=torch.float64, device=self.dev)
        ok = torch.zeros((self.n,), dtype=torch.int32, device=self.dev)
        it = torch.zeros((self.n,), dtype=torch.int32, device=self.dev)
        _lib.check(_lib.lib().rcsb_ik_inverse(self.ptr, pose.data_ptr(), q0.data_ptr(), q.data_ptr(), ok.data_ptr(), it.data_ptr()))
        return q, ok, it

    def set_cartesian_position(self, pose: torch.Tensor):
        _lib.check(_lib.lib().rcsb_robot_set_cartesian_position(self.ptr, pose.data_ptr()))

    def occupancy(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.lib().rcsb_kernel_occupancy(self.ptr, C.byref(a), C.byref(b), C.byref(c))
        L = _lib.lib()
        return dict(warps_per_cta=a.value, smem_bytes=b.value, grid=c.value,
                    variant=L.rcsb_kernel_variant(self.ptr, 0).decode(), variant_full=L.rcsb_kernel_variant(self.ptr, 1).decode())
