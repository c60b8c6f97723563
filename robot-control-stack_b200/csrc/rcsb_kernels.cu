// CUDA kernels + C ABI (include/rcsb.h) of the batched rigid-body backend. sm_100a only.
//
// Execution model: persistent CTAs (one per SM), W warps each, one environment per warp at a time.
// The hot part of the model is staged once per CTA into shared memory with a TMA bulk copy
// (cp.async.bulk + mbarrier); each warp owns a private shared-memory workspace that holds the
// environment's state and every intermediate of the physics step for all substeps of a launch, so
// HBM is touched once per launch per environment (row in, row out). Fixed-substep launches map
// environments to warps statically and keep the warps of a CTA aligned with barriers (instruction
// cache); step_until_convergence launches use the same map with a CTA-wide vote per step (or, with
// RCSB_CONV_VOTE=0, pull environment indices from a global atomic counter). Every launch runs in the
// reduced workspace layout first and hands the environments that outgrow it to a second launch in
// the full layout (rcsb_types.h: fast_maxcon). The kernel itself exists once per shape variant
// (rcsb_variant.cuh); this file holds the generic variant, the IK kernels and the host side.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rcsb.h"
#include "rcsb_layout.h"
#include "rcsb_ctx.cuh"

#include "rcsb_stage.cuh"

// ------------------------------------------------------------------ kernel variants
#define RCSB_VARIANT_NS rcsb_generic
#define RCSB_KERNEL rcsb_k_run
#include "rcsb_variant.cuh"
#undef RCSB_VARIANT_NS
#undef RCSB_KERNEL

#ifndef RCSB_NO_FIXED_VARIANTS
// shape-specialised variants live in their own translation units (rcsb_k_fr3_*.cu) so that they compile in parallel;
// the profiling build (-DRCSB_SINGLE_TU) pulls them in here so that they share rcsb_stage_cycles
#ifdef RCSB_SINGLE_TU
#include "rcsb_k_fr3_reduced.cu"
#include "rcsb_k_fr3_full.cu"
#include "rcsb_k_fr3_pickup.cu"
#include "rcsb_k_xarm7_tabletop.cu"
#else
#define RCSB_DECLARE_VARIANT(ns)                                                                                        \
  namespace ns {                                                                                                        \
  void launch(int, int, size_t, cudaStream_t, const RcsbModel*, const real*, const int*, real*, double*, int*, const RcsbLaunch&, int*, size_t); \
  cudaError_t set_smem(size_t);                                                                                         \
  int max_warps();                                                                                                      \
  RcsbShape shape();                                                                                                    \
  }
RCSB_DECLARE_VARIANT(rcsb_fr3_reduced)
RCSB_DECLARE_VARIANT(rcsb_fr3_full)
RCSB_DECLARE_VARIANT(rcsb_fr3_pickup)
RCSB_DECLARE_VARIANT(rcsb_xarm7_tabletop)
#endif
#endif

typedef void (*rcsb_launch_fn)(int, int, size_t, cudaStream_t, const RcsbModel*, const real*, const int*, real*, double*, int*,
                               const RcsbLaunch&, int*, size_t);
typedef cudaError_t (*rcsb_smem_fn)(size_t);
struct RcsbVariant { const char* name; int fixed; RcsbShape shape; rcsb_launch_fn launch; rcsb_smem_fn set_smem; int max_warps; };
static bool shape_equal(const RcsbShape& a, const RcsbShape& b) { return memcmp(&a, &b, sizeof(RcsbShape)) == 0; }
// the most specialised variant compiled for this model's shape (the generic one always matches)
static RcsbVariant pick_variant(const RcsbModel& h) {
  RcsbShape s = rcsb_model_shape(&h);
  const char* force = getenv("RCSB_VARIANT");  // "generic" disables the fixed-shape kernels (testing / tuning)
  if (!(force && !strcmp(force, "generic"))) {
#ifndef RCSB_NO_FIXED_VARIANTS
    if (shape_equal(s, rcsb_fr3_reduced::shape())) return {"fr3_reduced", 1, s, rcsb_fr3_reduced::launch, rcsb_fr3_reduced::set_smem, rcsb_fr3_reduced::max_warps()};
    if (shape_equal(s, rcsb_fr3_full::shape())) return {"fr3_full", 1, s, rcsb_fr3_full::launch, rcsb_fr3_full::set_smem, rcsb_fr3_full::max_warps()};
    if (shape_equal(s, rcsb_fr3_pickup::shape())) return {"fr3_pickup", 1, s, rcsb_fr3_pickup::launch, rcsb_fr3_pickup::set_smem, rcsb_fr3_pickup::max_warps()};
    if (shape_equal(s, rcsb_xarm7_tabletop::shape())) return {"xarm7_tabletop", 1, s, rcsb_xarm7_tabletop::launch, rcsb_xarm7_tabletop::set_smem, rcsb_xarm7_tabletop::max_warps()};
#endif
  }
  return {"generic", 0, s, rcsb_generic::launch, rcsb_generic::set_smem, rcsb_generic::max_warps()};
}

// IK kernel: one environment per thread
#define MD(f) (m.f)
#define LAY (m.lay)
namespace rcsb_generic {
#include "rcsb_ik.cuh"
}
using rcsb_generic::ik_env;
using rcsb_generic::cart_action_env;
using rcsb_generic::CartOrigin;
// Pin::inverse for every environment, one environment per thread (rcsb_ik.cuh): large batches
__global__ void __launch_bounds__(128)
rcsb_k_ik(const RcsbModel* __restrict__ gm, const real* __restrict__ pose, const real* __restrict__ q0, real* __restrict__ q_out,
          int* __restrict__ success, int* __restrict__ iters, int N, int apply, real* __restrict__ sr, int* __restrict__ si) {
  const RcsbModel* sm = stage_model(gm);
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < N; env += gridDim.x * blockDim.x)
    ik_env(sm, nullptr, 0, env, pose, q0, q_out, success, iters, apply, sr, si);
}
__global__ void __launch_bounds__(128)
rcsb_k_cart_action(const RcsbModel* __restrict__ gm, const real* __restrict__ act, int kind, int relative, real max_trans, real max_rot,
                   int N, real* __restrict__ sr, int* __restrict__ si, CartOrigin co) {
  const RcsbModel* sm = stage_model(gm);
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < N; env += gridDim.x * blockDim.x)
    cart_action_env(sm, env, act, kind, relative, max_trans, max_rot, sr, si, co);
}
// the same two entry points with 8 lanes per environment (4 environments per warp): small and medium batches
__global__ void __launch_bounds__(128)
rcsb_k_ik8(const RcsbModel* __restrict__ gm, int nch, const real* __restrict__ pose, const real* __restrict__ q0,
           real* __restrict__ q_out, int* __restrict__ success, int* __restrict__ iters, int N, int apply, real* __restrict__ sr,
           int* __restrict__ si) {
  const RcsbModel* sm = stage_model(gm);
  const int g = threadIdx.x & 7, env = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, valid = env < N;
  rcsb_generic::ik_env8(sm, nch, valid ? env : 0, valid, g, pose, q0, q_out, success, iters, apply, sr, si);
}
__global__ void __launch_bounds__(128)
rcsb_k_cart_action8(const RcsbModel* __restrict__ gm, int nch, const real* __restrict__ act, int kind, int relative, real max_trans,
                    real max_rot, int N, real* __restrict__ sr, int* __restrict__ si, CartOrigin co) {
  const RcsbModel* sm = stage_model(gm);
  const int g = threadIdx.x & 7, env = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, valid = env < N;
  rcsb_generic::cart_action_env8(sm, nch, valid ? env : 0, valid, g, act, kind, relative, max_trans, max_rot, sr, si, co);
}
#undef MD
#undef LAY

// ------------------------------------------------------------------ host side
static thread_local std::string g_err;
static long long g_launches = 0;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_OK(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(RCSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// every entry point that touches the device selects the model's device and puts the caller's current device back
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = prev == device ? cudaSuccess : cudaSetDevice(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEVICE_OK(device)   \
  DeviceGuard guard_(device); \
  if (guard_.err != cudaSuccess) return fail(RCSB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard_.err))

struct rcsb_model {
  RcsbModel h;        // full-capacity layout
  RcsbModel hr;       // reduced-capacity layout (has_reduced)
  bool has_reduced = false;
  RcsbModel* d_model_r = nullptr;
  std::vector<real> verts;
  std::vector<int> vgraph;  // hull edge graph: [nmeshvert + 1] offsets, then the neighbour lists
  int* d_vgraph = nullptr;
  bool finalized = false;
  int device = -1;
  RcsbModel* d_model = nullptr;
  real* d_verts = nullptr;
};
struct rcsb_batch {
  rcsb_model* m;
  int n;
  real* sr; double* sd; int* si;
  cudaStream_t stream;
  int* d_counter = nullptr;   // [0] env cursor phase 0, [1] overflow cursor phase 1, [2] overflow count
  int* d_overflow = nullptr;  // [n] overflow list
  int bar_groups = 1, conv_vote = 1;
  int warps = 0, grid = 0, lockstep = 0x010;  // barrier mask of fixed-substep launches (rcsb_warp.cuh)
  size_t smem = 0, ws_bytes = 0;
  RcsbVariant var, var_full;          // kernel variants of the two phases
  int warps_full = 0, grid_full = 0;  // phase 1 (full layout) launch shape when a reduced layout exists
  size_t smem_full = 0, ws_bytes_full = 0;
  // staging for the host-buffer path
  real *d_act_joints = nullptr, *d_act_gripper = nullptr, *d_obs = nullptr, *d_act_packed = nullptr;
  int* d_info = nullptr;
  real *h_act = nullptr, *h_obs = nullptr;
  int act_jstride = 0, act_gstride = 0;  // non-default action row strides of the next launch (packed host path)
  // optional contact export (rcsb_batch_set_contact_export)
  int *con_n = nullptr, *con_geom = nullptr, con_cap = 0;
  real* con_real = nullptr;
};

extern "C" {
const char* rcsb_last_error(void) { return g_err.c_str(); }
int rcsb_version(void) { return 100; }
int rcsb_real_bytes(void) { return (int)sizeof(real); }
long long rcsb_launch_count(void) { return g_launches; }

rcsb_model* rcsb_model_new(void) {
  rcsb_model* m = new rcsb_model();
  memset(&m->h, 0, sizeof(RcsbModel));
  return m;
}
void rcsb_model_free(rcsb_model* m) {
  if (!m) return;
  if (m->d_model) cudaFree(m->d_model);
  if (m->d_model_r) cudaFree(m->d_model_r);
  if (m->d_verts) cudaFree(m->d_verts);
  if (m->d_vgraph) cudaFree(m->d_vgraph);
  delete m;
}
int rcsb_model_set_int(rcsb_model* m, const char* field, const int* v, int n) {
  int rc = rcsb_model_set_field(&m->h, field, v, n, 0);
  if (rc == -1) return fail(RCSB_ERR_FIELD, std::string("unknown int field ") + field);
  if (rc == -2) return fail(RCSB_ERR_SIZE, std::string("bad size/type for field ") + field);
  return RCSB_OK;
}
int rcsb_model_set_real(rcsb_model* m, const char* field, const double* v, int n) {
  int rc = rcsb_model_set_field(&m->h, field, v, n, 1);
  if (rc == -1) return fail(RCSB_ERR_FIELD, std::string("unknown real field ") + field);
  if (rc == -2) return fail(RCSB_ERR_SIZE, std::string("bad size/type for field ") + field);
  return RCSB_OK;
}
int rcsb_model_set_mesh_vertices(rcsb_model* m, const double* xyz, int nvert) {
  m->verts.resize((size_t)3 * (nvert > 0 ? nvert : 1));
  for (int i = 0; i < 3 * nvert; i++) m->verts[i] = (real)xyz[i];
  return RCSB_OK;
}
int rcsb_model_set_mesh_graph(rcsb_model* m, const int* adr, int nadr, const int* nbr, int nnbr) {
  if (!m || !adr || nadr < 1 || (nnbr > 0 && !nbr)) return fail(RCSB_ERR_ARG, "bad mesh graph");
  if (adr[0] != 0 || adr[nadr - 1] != nnbr) return fail(RCSB_ERR_SIZE, "mesh graph offsets do not match the neighbour list");
  m->vgraph.assign(adr, adr + nadr);
  m->vgraph.insert(m->vgraph.end(), nbr, nbr + nnbr);
  return RCSB_OK;
}
int rcsb_model_finalize(rcsb_model* m) {
  if (rcsb_model_finalize_layout(&m->h) != 0) return fail(RCSB_ERR_MODEL, "model dimensions out of range");
  if (!m->vgraph.empty() && ((int)m->vgraph.size() < m->h.nmeshvert + 1 || m->vgraph[m->h.nmeshvert] + m->h.nmeshvert + 1 != (int)m->vgraph.size()))
    return fail(RCSB_ERR_SIZE, "mesh graph does not cover the vertex pool");
  if ((m->h.lay.nsr * sizeof(real)) % 16 != 0) return fail(RCSB_ERR_MODEL, "state row is not 16-byte granular");
  m->h.cap_reduced = 0;
  m->has_reduced = rcsb_model_make_reduced(&m->h, &m->hr) != 0;
  if (m->has_reduced) {
    int need = m->h.neq;
    for (int j = 0; j < m->h.nv; j++) need += m->h.d_frictionloss[j] > 0;
    if (m->hr.maxefc < need + 2) return fail(RCSB_ERR_MODEL, "fast_maxefc must cover the equality and friction-loss rows plus 2");
  }
  m->finalized = true;
  return RCSB_OK;
}
int rcsb_model_upload(rcsb_model* m, int device) {
  if (!m->finalized) return fail(RCSB_ERR_MODEL, "rcsb_model_finalize not called");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RCSB_ERR_CUDA, "no CUDA device: this backend has no CPU execution path");
  DEVICE_OK(device);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(RCSB_ERR_CUDA, "sm_100a (B200) device required");
  std::vector<unsigned char> padded(RCSB_MODEL_BYTES, 0);
  memcpy(padded.data(), &m->h, sizeof(RcsbModel));
  CUDA_OK(cudaMalloc(&m->d_model, RCSB_MODEL_BYTES));
  CUDA_OK(cudaMemcpy(m->d_model, padded.data(), RCSB_MODEL_BYTES, cudaMemcpyHostToDevice));
  if (m->has_reduced) {
    memcpy(padded.data(), &m->hr, sizeof(RcsbModel));
    CUDA_OK(cudaMalloc(&m->d_model_r, RCSB_MODEL_BYTES));
    CUDA_OK(cudaMemcpy(m->d_model_r, padded.data(), RCSB_MODEL_BYTES, cudaMemcpyHostToDevice));
  }
  if (m->verts.empty()) m->verts.resize(3);
  CUDA_OK(cudaMalloc(&m->d_verts, m->verts.size() * sizeof(real)));
  CUDA_OK(cudaMemcpy(m->d_verts, m->verts.data(), m->verts.size() * sizeof(real), cudaMemcpyHostToDevice));
  if (!m->vgraph.empty()) {
    CUDA_OK(cudaMalloc(&m->d_vgraph, m->vgraph.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(m->d_vgraph, m->vgraph.data(), m->vgraph.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  m->device = device;
  return RCSB_OK;
}
int rcsb_model_dims(const rcsb_model* m, int* nsr, int* nsd, int* nsi, int* obs_dim, int* info_dim) {
  if (!m->finalized) return fail(RCSB_ERR_MODEL, "rcsb_model_finalize not called");
  if (nsr) *nsr = m->h.lay.nsr;
  if (nsd) *nsd = RCSB_D_TAIL;
  if (nsi) *nsi = RCSB_I_TAIL;
  if (obs_dim) *obs_dim = RCSB_OBS_DIM;
  if (info_dim) *info_dim = RCSB_INFO_DIM;
  return RCSB_OK;
}
int rcsb_model_offsets(const rcsb_model* m, int* o_qpos, int* o_qvel, int* o_ctrl, int* o_warm, int* o_tail) {
  if (!m->finalized) return fail(RCSB_ERR_MODEL, "rcsb_model_finalize not called");
  if (o_qpos) *o_qpos = m->h.lay.o_q;
  if (o_qvel) *o_qvel = m->h.lay.o_v;
  if (o_ctrl) *o_ctrl = m->h.lay.o_ctrl;
  if (o_warm) *o_warm = m->h.lay.o_warm;
  if (o_tail) *o_tail = m->h.lay.o_rcs;
  return RCSB_OK;
}

rcsb_batch* rcsb_batch_new(rcsb_model* m, int n_envs, void* sr, void* sd, void* si, void* stream) {
  if (!m || !m->d_model) { fail(RCSB_ERR_MODEL, "model not uploaded (rcsb_model_upload)"); return nullptr; }
  if (n_envs <= 0 || !sr || !sd || !si) { fail(RCSB_ERR_ARG, "bad batch arguments"); return nullptr; }
  rcsb_batch* b = new rcsb_batch();
  b->m = m; b->n = n_envs; b->sr = (real*)sr; b->sd = (double*)sd; b->si = (int*)si; b->stream = (cudaStream_t)stream;
  DeviceGuard guard_(m->device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, m->device);
  size_t avail = prop.sharedMemPerBlockOptin;
  int cap = RCSB_MAX_WARPS;
  if (const char* ov = getenv("RCSB_WARPS")) {  // tuning / profiling knob: fewer warps per CTA
    int o = atoi(ov);
    if (o >= 1 && o < cap) cap = o;
  }
  if (const char* ov = getenv("RCSB_CONV_VOTE")) b->conv_vote = atoi(ov) != 0;  // 0 = dynamic scheduling for step_until_convergence
  if (const char* ov = getenv("RCSB_BAR_GROUPS")) { int g = atoi(ov); if (g >= 1 && g <= 15) b->bar_groups = g; }
  if (const char* ov = getenv("RCSB_LOCKSTEP")) {  // tuning knob: 0 none, 1 every stage, 2 once per step, 0x.. explicit mask
    long v = strtol(ov, nullptr, 0);
    b->lockstep = v == 1 ? RCSB_LOCKSTEP_ALL : (v == 2 ? RCSB_LOCKSTEP_STEP : (int)v);
  }
  b->var = pick_variant(m->has_reduced ? m->hr : m->h);
  b->var_full = pick_variant(m->h);
  auto shape = [&](const RcsbModel& h, const RcsbVariant& var, size_t& ws, int& warps, size_t& smem, int& grid) {
    ws = rcsb_ws_bytes(&h);
    warps = (int)((avail - RCSB_SMEM_HEADER) / ws);
    if (warps > cap) warps = cap;
    if (warps > var.max_warps) warps = var.max_warps;  // the kernel's launch bounds
    if (warps < 1) return false;
    smem = RCSB_SMEM_HEADER + (size_t)warps * ws;
    grid = prop.multiProcessorCount;
    int need = (n_envs + warps - 1) / warps;
    if (grid > need) grid = need;
    return true;
  };
  bool ok = shape(m->has_reduced ? m->hr : m->h, b->var, b->ws_bytes, b->warps, b->smem, b->grid);
  if (ok && m->has_reduced) ok = shape(m->h, b->var_full, b->ws_bytes_full, b->warps_full, b->smem_full, b->grid_full);
  if (!ok) { fail(RCSB_ERR_MODEL, "workspace does not fit in shared memory"); delete b; return nullptr; }
  const size_t smem_max = b->smem > b->smem_full ? b->smem : b->smem_full;  // both phases may use the same kernel
  if (b->var.set_smem(smem_max) != cudaSuccess || (m->has_reduced && b->var_full.set_smem(smem_max) != cudaSuccess) ||
      cudaFuncSetAttribute(rcsb_k_ik, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCSB_SMEM_HEADER) != cudaSuccess ||
      cudaFuncSetAttribute(rcsb_k_cart_action, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCSB_SMEM_HEADER) != cudaSuccess ||
      cudaFuncSetAttribute(rcsb_k_ik8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCSB_SMEM_HEADER) != cudaSuccess ||
      cudaFuncSetAttribute(rcsb_k_cart_action8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCSB_SMEM_HEADER) != cudaSuccess ||
      cudaMalloc(&b->d_counter, 4 * sizeof(int)) != cudaSuccess || cudaMalloc(&b->d_overflow, (size_t)n_envs * sizeof(int)) != cudaSuccess) {
    fail(RCSB_ERR_CUDA, std::string("batch setup: ") + cudaGetErrorString(cudaGetLastError()));
    delete b;
    return nullptr;
  }
  return b;
}
void rcsb_batch_free(rcsb_batch* b) {
  if (!b) return;
  cudaFree(b->d_counter); cudaFree(b->d_overflow); cudaFree(b->d_act_joints); cudaFree(b->d_act_gripper); cudaFree(b->d_act_packed); cudaFree(b->d_obs); cudaFree(b->d_info);
  if (b->h_act) cudaFreeHost(b->h_act);
  if (b->h_obs) cudaFreeHost(b->h_obs);
  delete b;
}
int rcsb_debug_stage_cycles(unsigned long long* out16) {
#ifdef RCSB_STAGE_TIMING
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(out16, rcsb_stage_cycles, 16 * sizeof(unsigned long long)));
  unsigned long long z[16] = {0};
  CUDA_OK(cudaMemcpyToSymbol(rcsb_stage_cycles, z, sizeof(z)));
  return RCSB_OK;
#else
  for (int i = 0; i < 16; i++) out16[i] = 0;
  return fail(RCSB_ERR_ARG, "library built without RCSB_STAGE_TIMING");
#endif
}
int rcsb_model_workspace_bytes(const rcsb_model* m, int* reduced_bytes, int* full_bytes, int* smem_header_bytes) {
  if (!m->finalized) return fail(RCSB_ERR_MODEL, "rcsb_model_finalize not called");
  if (reduced_bytes) *reduced_bytes = m->has_reduced ? (int)rcsb_ws_bytes(&m->hr) : 0;
  if (full_bytes) *full_bytes = (int)rcsb_ws_bytes(&m->h);
  if (smem_header_bytes) *smem_header_bytes = (int)RCSB_SMEM_HEADER;
  return RCSB_OK;
}
const char* rcsb_kernel_variant(rcsb_batch* b, int phase) { return phase == 0 ? b->var.name : b->var_full.name; }
int rcsb_kernel_occupancy(rcsb_batch* b, int* warps_per_cta, int* smem_bytes, int* grid) {
  if (warps_per_cta) *warps_per_cta = b->warps;
  if (smem_bytes) *smem_bytes = (int)b->smem;
  if (grid) *grid = b->grid;
  return RCSB_OK;
}

int rcsb_batch_run(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const void* act_joints_dev,
                   const void* act_gripper_dev, const unsigned char* mask_dev, double max_mov, const double* jlow,
                   const double* jhigh, void* obs_dev, int* info_dev) {
  if (!b) return fail(RCSB_ERR_ARG, "null batch");
  const unsigned needs_joints = RCSB_OP_ACT_JOINTS_REL | RCSB_OP_ACT_JOINTS_ABS | RCSB_OP_SET_JOINTS | RCSB_OP_SET_JOINTS_HARD;
  if ((ops & needs_joints) && !act_joints_dev) return fail(RCSB_ERR_ARG, "act_joints_dev required for the requested ops");
  if ((ops & (RCSB_OP_ACT_GRIPPER_BIN | RCSB_OP_ACT_GRIPPER_CONT | RCSB_OP_SET_GRIPPER)) && !act_gripper_dev)
    return fail(RCSB_ERR_ARG, "act_gripper_dev required for the requested ops");
  if ((ops & RCSB_OP_ACT_JOINTS_REL) && (!jlow || !jhigh)) return fail(RCSB_ERR_ARG, "joint limits required");
  RcsbLaunch L;
  memset(&L, 0, sizeof(L));
  L.N = b->n; L.ops = ops; L.lockstep = b->lockstep; L.bar_groups = b->bar_groups; L.conv_vote = b->conv_vote; L.k = k; L.max_convergence_steps = max_convergence_steps;
  L.act_joints = (const real*)act_joints_dev; L.act_gripper = (const real*)act_gripper_dev; L.mask = mask_dev;
  L.act_jstride = b->act_jstride > 0 ? b->act_jstride : b->m->h.rb_njoints; L.act_gstride = b->act_gstride > 0 ? b->act_gstride : 1;
  L.max_mov = (real)max_mov;
  for (int i = 0; i < b->m->h.rb_njoints && i < RCSB_MAXJ; i++) {
    L.jlow[i] = jlow ? (real)jlow[i] : 0;
    L.jhigh[i] = jhigh ? (real)jhigh[i] : 0;
  }
  L.obs = (real*)obs_dev; L.info = info_dev;
  L.con_n = b->con_n; L.con_geom = b->con_geom; L.con_real = b->con_real; L.con_cap = b->con_cap;
  DEVICE_OK(b->m->device);
  CUDA_OK(cudaMemsetAsync(b->d_counter, 0, 4 * sizeof(int), b->stream));
  L.phase = 0; L.overflow_list = b->d_overflow; L.overflow_count = b->d_counter + 2;
  const bool two = b->m->has_reduced;
  b->var.launch(b->grid, b->warps * 32, b->smem, b->stream, two ? b->m->d_model_r : b->m->d_model, b->m->d_verts, b->m->d_vgraph, b->sr, b->sd, b->si,
                L, b->d_counter, b->ws_bytes);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  if (two && (ops & (RCSB_OP_STEP_K | RCSB_OP_STEP_CONV))) {  // finishes the environments that outgrew the reduced layout
    L.phase = 1; L.lockstep = 0;
    b->var_full.launch(b->grid_full, b->warps_full * 32, b->smem_full, b->stream, b->m->d_model, b->m->d_verts, b->m->d_vgraph, b->sr, b->sd, b->si, L,
                       b->d_counter + 1, b->ws_bytes_full);
    g_launches++;
    CUDA_OK(cudaGetLastError());
  }
  return RCSB_OK;
}

int rcsb_batch_set_contact_export(rcsb_batch* b, int* ncon_dev, int* geom_dev, void* real_dev, int cap) {
  if (!b) return fail(RCSB_ERR_ARG, "null batch");
  if (ncon_dev && (!geom_dev || cap < 1)) return fail(RCSB_ERR_ARG, "geom_dev and cap >= 1 required with ncon_dev");
  b->con_n = ncon_dev; b->con_geom = ncon_dev ? geom_dev : nullptr; b->con_real = ncon_dev ? (real*)real_dev : nullptr;
  b->con_cap = ncon_dev ? cap : 0;
  return RCSB_OK;
}
int rcsb_batch_init_state(rcsb_batch* b) {
  DEVICE_OK(b->m->device);
  CUDA_OK(cudaMemsetAsync(b->sr, 0, (size_t)b->n * b->m->h.lay.nsr * sizeof(real), b->stream));
  CUDA_OK(cudaMemsetAsync(b->sd, 0, (size_t)b->n * RCSB_D_TAIL * sizeof(double), b->stream));
  std::vector<int> row(RCSB_I_TAIL, 0), all((size_t)b->n * RCSB_I_TAIL);
  row[RCSB_I_IK_SUCCESS] = 1;  // SimRobotState::ik_success = true (SimRobot.h:53)
  row[RCSB_I_CONVERGED] = 1;   // Sim::converged = true (sim.h:46)
  for (int e = 0; e < b->n; e++) memcpy(&all[(size_t)e * RCSB_I_TAIL], row.data(), sizeof(int) * RCSB_I_TAIL);
  CUDA_OK(cudaMemcpyAsync(b->si, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  // mj_resetData, then the constructors' m_reset() calls (SimRobot.cpp:42, SimGripper.cpp:38)
  int rc = rcsb_batch_run(b, RCSB_OP_SIM_RESET | RCSB_OP_ROBOT_RESET | RCSB_OP_GRIPPER_RESET | RCSB_OP_ENV_RESET_FLAGS, 0, 0, nullptr,
                          nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return RCSB_OK;
}

static int ensure_staging(rcsb_batch* b) {
  if (b->d_act_joints) return RCSB_OK;
  size_t n = (size_t)b->n;
  CUDA_OK(cudaMalloc(&b->d_act_joints, n * RCSB_MAXJ * sizeof(real)));
  CUDA_OK(cudaMalloc(&b->d_act_gripper, n * sizeof(real)));
  CUDA_OK(cudaMalloc(&b->d_act_packed, n * (RCSB_MAXJ + 1) * sizeof(real)));
  CUDA_OK(cudaMalloc(&b->d_obs, n * RCSB_OBS_DIM * sizeof(real)));
  CUDA_OK(cudaMalloc(&b->d_info, n * RCSB_INFO_DIM * sizeof(int)));
  return RCSB_OK;
}
int rcsb_batch_run_host(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const double* act_joints_host,
                        const double* act_gripper_host, double max_mov, const double* jlow, const double* jhigh,
                        double* obs_host, int* info_host) {
  if (sizeof(real) != sizeof(double)) return fail(RCSB_ERR_ARG, "host-buffer path requires a float64 build");
  int rc = ensure_staging(b);
  if (rc) return rc;
  DEVICE_OK(b->m->device);
  int nj = b->m->h.rb_njoints;
  if (act_joints_host)
    CUDA_OK(cudaMemcpyAsync(b->d_act_joints, act_joints_host, (size_t)b->n * nj * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  if (act_gripper_host)
    CUDA_OK(cudaMemcpyAsync(b->d_act_gripper, act_gripper_host, (size_t)b->n * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  rc = rcsb_batch_run(b, ops, k, max_convergence_steps, act_joints_host ? b->d_act_joints : nullptr,
                      act_gripper_host ? b->d_act_gripper : nullptr, nullptr, max_mov, jlow, jhigh, obs_host ? b->d_obs : nullptr,
                      info_host ? b->d_info : nullptr);
  if (rc) return rc;
  if (obs_host)
    CUDA_OK(cudaMemcpyAsync(obs_host, b->d_obs, (size_t)b->n * RCSB_OBS_DIM * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  if (info_host)
    CUDA_OK(cudaMemcpyAsync(info_host, b->d_info, (size_t)b->n * RCSB_INFO_DIM * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return RCSB_OK;
}

int rcsb_env_step_host(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const double* act_host, double max_mov,
                       const double* jlow, const double* jhigh, double* obs_host) {
  if (sizeof(real) != sizeof(double)) return fail(RCSB_ERR_ARG, "host-buffer path requires a float64 build");
  if (!b || !act_host || !obs_host) return fail(RCSB_ERR_ARG, "null argument");
  int rc = ensure_staging(b);
  if (rc) return rc;
  DEVICE_OK(b->m->device);
  const int nj = b->m->h.rb_njoints, stride = nj + 1;
  // one copy in: the packed action block lands in the joint staging array ([n][MAXJ] reals are reserved, nj + 1 <= MAXJ + 1)
  if ((size_t)stride > (size_t)RCSB_MAXJ + 1) return fail(RCSB_ERR_ARG, "too many joints");
  CUDA_OK(cudaMemcpyAsync(b->d_act_packed, act_host, (size_t)b->n * stride * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  b->act_jstride = stride; b->act_gstride = stride;
  rc = rcsb_batch_run(b, ops | RCSB_OP_OBS, k, max_convergence_steps, b->d_act_packed, b->d_act_packed + nj, nullptr, max_mov, jlow, jhigh,
                      b->d_obs, nullptr);
  b->act_jstride = 0; b->act_gstride = 0;
  if (rc) return rc;
  // one copy out: the packed observation block carries the info flags as reals
  CUDA_OK(cudaMemcpyAsync(obs_host, b->d_obs, (size_t)b->n * RCSB_OBS_DIM * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return RCSB_OK;
}

int rcsb_sim_step(rcsb_batch* b, int k) {
  return rcsb_batch_run(b, RCSB_OP_STEP_K, k, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_sim_step_until_convergence(rcsb_batch* b, int max_steps) {
  return rcsb_batch_run(b, RCSB_OP_STEP_CONV, 0, max_steps, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_sim_reset(rcsb_batch* b) {
  return rcsb_batch_run(b, RCSB_OP_SIM_RESET, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_robot_set_joint_position(rcsb_batch* b, const void* q_dev) {
  return rcsb_batch_run(b, RCSB_OP_SET_JOINTS, 0, 0, q_dev, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_robot_set_joints_hard(rcsb_batch* b, const void* q_dev) {
  return rcsb_batch_run(b, RCSB_OP_SET_JOINTS_HARD, 0, 0, q_dev, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_robot_reset(rcsb_batch* b) {
  return rcsb_batch_run(b, RCSB_OP_ROBOT_RESET, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_gripper_set_normalized_width(rcsb_batch* b, const void* width_dev) {
  return rcsb_batch_run(b, RCSB_OP_SET_GRIPPER, 0, 0, nullptr, width_dev, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_gripper_reset(rcsb_batch* b) {
  return rcsb_batch_run(b, RCSB_OP_GRIPPER_RESET, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}
int rcsb_env_get_obs(rcsb_batch* b, void* obs_dev, int* info_dev) {
  return rcsb_batch_run(b, RCSB_OP_OBS, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, obs_dev, info_dev);
}

// Two mappings of the CLIK solver (rcsb_ik.cuh). One environment per thread: a long serial solve per thread (250
// registers, ~5 k instructions per iteration) in blocks of one warp spread over the SMs -- the fewest instructions in
// total, but latency-bound until there are tens of thousands of environments. 8 lanes per environment: ~2.4 x the
// instructions in total, a quarter of the latency; the choice below (RCSB_IK_LANES=1 / 8 overrides it) follows the
// measured crossover. The lanes mapping needs a canonical chain (FR3, xArm7).
static int ik_block_threads(int n) {
  return n >= 128 * 148 * 4 ? 128 : (n >= 64 * 148 * 4 ? 64 : 32);
}
static int canonical_chain(const RcsbModel& m) {  // rcsb_ik.cuh: ik_canonical_chain
  int n = 0;
  for (int b = m.rb_site_body; b >= 0; b = m.b_parent[b]) n++;
  if (n < 1 || n > 8) return 0;
  for (int b = m.rb_site_body, i = n - 1; b >= 0; b = m.b_parent[b], i--)
    if (m.b_jtype[b] == RCSB_JNT_FREE || m.b_dadr[b] != i || m.b_qadr[b] != i) return 0;
  return n;
}
static int ik_lanes_chain(rcsb_batch* b) {  // chain length when this batch uses the 8-lane kernels, else 0
  const int nch = canonical_chain(b->m->h);
  int lanes = b->n <= 8192 ? 8 : 1;
  if (const char* ov = getenv("RCSB_IK_LANES")) lanes = atoi(ov);
  return (lanes == 8 && nch > 0) ? nch : 0;
}
static int launch_ik(rcsb_batch* b, const void* pose_dev, const void* q0_dev, void* q_out_dev, int* success_dev, int* iters_dev,
                     int apply) {
  if (!b || !pose_dev) return fail(RCSB_ERR_ARG, "null argument");
  DEVICE_OK(b->m->device);
  if (const int nch = ik_lanes_chain(b)) {
    const int threads = 128, grid = (b->n * 8 + threads - 1) / threads;
    rcsb_k_ik8<<<grid, threads, RCSB_SMEM_HEADER, b->stream>>>(b->m->d_model, nch, (const real*)pose_dev, (const real*)q0_dev,
                                                              (real*)q_out_dev, success_dev, iters_dev, b->n, apply, b->sr, b->si);
  } else {
    int threads = ik_block_threads(b->n), grid = (b->n + threads - 1) / threads;
    rcsb_k_ik<<<grid, threads, RCSB_SMEM_HEADER, b->stream>>>(b->m->d_model, (const real*)pose_dev, (const real*)q0_dev,
                                                             (real*)q_out_dev, success_dev, iters_dev, b->n, apply, b->sr, b->si);
  }
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return RCSB_OK;
}
int rcsb_env_cartesian_action_origin(rcsb_batch* b, const void* act_dev, int kind, int relative, double max_trans, double max_rot,
                                     const void* origin_dev, void* last_dev, int* have_last_dev) {
  if (!b || !act_dev || (kind != 0 && kind != 1) || relative < 0 || relative > 2) return fail(RCSB_ERR_ARG, "bad argument");
  if (relative == 2 && (!origin_dev || !last_dev || !have_last_dev))
    return fail(RCSB_ERR_ARG, "CONFIGURED_ORIGIN needs the origin / last-offset / have-last device arrays");
  DEVICE_OK(b->m->device);
  CartOrigin co = {(const real*)origin_dev, (real*)last_dev, have_last_dev};
  if (const int nch = ik_lanes_chain(b)) {
    const int threads = 128, grid = (b->n * 8 + threads - 1) / threads;
    rcsb_k_cart_action8<<<grid, threads, RCSB_SMEM_HEADER, b->stream>>>(b->m->d_model, nch, (const real*)act_dev, kind, relative,
                                                                       (real)max_trans, (real)max_rot, b->n, b->sr, b->si, co);
  } else {
    int threads = ik_block_threads(b->n), grid = (b->n + threads - 1) / threads;
    rcsb_k_cart_action<<<grid, threads, RCSB_SMEM_HEADER, b->stream>>>(b->m->d_model, (const real*)act_dev, kind, relative,
                                                                      (real)max_trans, (real)max_rot, b->n, b->sr, b->si, co);
  }
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return RCSB_OK;
}
int rcsb_env_cartesian_action(rcsb_batch* b, const void* act_dev, int kind, int relative, double max_trans, double max_rot) {
  if (relative != 0 && relative != 1) return fail(RCSB_ERR_ARG, "bad argument");
  return rcsb_env_cartesian_action_origin(b, act_dev, kind, relative, max_trans, max_rot, nullptr, nullptr, nullptr);
}
int rcsb_ik_inverse(rcsb_batch* b, const void* pose_dev, const void* q0_dev, void* q_out_dev, int* success_dev, int* iters_dev) {
  if (!q0_dev || !q_out_dev || !success_dev) return fail(RCSB_ERR_ARG, "null argument");
  return launch_ik(b, pose_dev, q0_dev, q_out_dev, success_dev, iters_dev, 0);
}
int rcsb_robot_set_cartesian_position(rcsb_batch* b, const void* pose_dev) {
  return launch_ik(b, pose_dev, nullptr, nullptr, nullptr, nullptr, 1);
}
}  // extern "C"
