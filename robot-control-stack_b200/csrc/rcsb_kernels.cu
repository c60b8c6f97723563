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
// lockstep: where the warps of a CTA re-align inside a physics step (bit i = before stage i, bit 9 = at its end; rcsb_warp.cuh),
// measured per kernel: after the collision stage for the FR3 kernels, at the step's end for the tabletop kernel (+12 %)
struct RcsbVariant { const char* name; int fixed; RcsbShape shape; rcsb_launch_fn launch; rcsb_smem_fn set_smem; int max_warps; int lockstep; };
static bool shape_equal(const RcsbShape& a, const RcsbShape& b) { return memcmp(&a, &b, sizeof(RcsbShape)) == 0; }
// the most specialised variant compiled for this model's shape (the generic one always matches)
static RcsbVariant pick_variant(const RcsbModel& h) {
  RcsbShape s = rcsb_model_shape(&h);
  const char* force = getenv("RCSB_VARIANT");  // "generic" disables the fixed-shape kernels (testing / tuning)
  if (!(force && !strcmp(force, "generic"))) {
#ifndef RCSB_NO_FIXED_VARIANTS
    if (shape_equal(s, rcsb_fr3_reduced::shape())) return {"fr3_reduced", 1, s, rcsb_fr3_reduced::launch, rcsb_fr3_reduced::set_smem, rcsb_fr3_reduced::max_warps(), 0x010};
    if (shape_equal(s, rcsb_fr3_full::shape())) return {"fr3_full", 1, s, rcsb_fr3_full::launch, rcsb_fr3_full::set_smem, rcsb_fr3_full::max_warps(), 0x010};
    if (shape_equal(s, rcsb_fr3_pickup::shape())) return {"fr3_pickup", 1, s, rcsb_fr3_pickup::launch, rcsb_fr3_pickup::set_smem, rcsb_fr3_pickup::max_warps(), 0x010};
    if (shape_equal(s, rcsb_xarm7_tabletop::shape())) return {"xarm7_tabletop", 1, s, rcsb_xarm7_tabletop::launch, rcsb_xarm7_tabletop::set_smem, rcsb_xarm7_tabletop::max_warps(), 0x200};
#endif
  }
  return {"generic", 0, s, rcsb_generic::launch, rcsb_generic::set_smem, rcsb_generic::max_warps(), 0x010};
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

// ------------------------------------------------------------------ depth camera: ray-caster over the collision geoms
// SimCameraSet depth frames (src/sim/camera.cpp:100-140 renders with OpenGL and reads the z-buffer back;
// python/rcs/camera/sim.py:45-115 turns it into metres and uint16 millimetres). Here every pixel casts one ray against
// the collidable geoms of its environment: plane, sphere, capsule, cylinder, box analytically, convex meshes by clipping
// the ray against the supporting planes of the hull. One CTA renders a 16 x 16 tile of one environment's image: the
// geoms' world frames are put together once per CTA in shared memory from the body frames the step kernel exported.
struct RcsbCamera {
  int body;              // moving body the camera rides on (-1: fixed in the world)
  real pos[3], rot[9];   // camera frame in that body's frame (MuJoCo convention: looks along -z, +y up)
  real f, inv_f;         // focal length in pixels: 0.5 * H / tan(fovy / 2), and its reciprocal (no division per ray)
  int W, H;
  real znear, zfar;      // clip planes in metres (mjVisual.map.znear / zfar x mjStatistic.extent)
  int physical_units;    // 1: millimetres of eye-space depth; 0: 1000 x the OpenGL window-space depth in [0, 1]
};
enum { RCSB_CAM_TILE = 16 };
template <typename T>
__device__ __forceinline__ bool ray_geom(const RcsbModel& m, int g, const T* __restrict__ faces, const int* __restrict__ face_adr,
                                         const int* __restrict__ face_num, const T* sz, const T* ab, const T* o, const T* d, T tmax,
                                         T* t_out) {
  // o, d: ray in the geom frame (d not normalised: t is in units of it)
  const int type = m.g_type[g];
  T t = -1;
  if (type == RCSB_GEOM_PLANE) {
    if (d[2] < 0) t = -o[2] / d[2];
  } else if (type == RCSB_GEOM_SPHERE) {
    const T a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], b = o[0] * d[0] + o[1] * d[1] + o[2] * d[2];
    const T cc = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - sz[0] * sz[0], disc = b * b - a * cc;
    if (disc >= 0) t = (-b - sqrt(disc)) / a;
  } else if (type == RCSB_GEOM_BOX) {
    T t0 = 0, t1 = tmax;
    for (int k = 0; k < 3; k++) {
      if (d[k] != 0) {
        const T inv = (T)1 / d[k];
        T ta = (-sz[k] - o[k]) * inv, tb = (sz[k] - o[k]) * inv;
        if (ta > tb) { T s = ta; ta = tb; tb = s; }
        t0 = ta > t0 ? ta : t0; t1 = tb < t1 ? tb : t1;
      } else if (o[k] < -sz[k] || o[k] > sz[k]) t1 = -1;
    }
    if (t0 <= t1) t = t0;
  } else if (type == RCSB_GEOM_CAPSULE || type == RCSB_GEOM_CYLINDER) {
    const T r = sz[0], hl = sz[1];
    // side: infinite cylinder about z, accepted while |z| <= hl
    const T a = d[0] * d[0] + d[1] * d[1], b = o[0] * d[0] + o[1] * d[1], cc = o[0] * o[0] + o[1] * o[1] - r * r;
    T best = (T)1e30;
    if (a > 0) {
      const T disc = b * b - a * cc;
      if (disc >= 0) {
        const T ts = (-b - sqrt(disc)) / a, z = o[2] + ts * d[2];
        if (ts >= 0 && z >= -hl && z <= hl) best = ts;
      }
    }
    for (int s = -1; s <= 1; s += 2) {
      if (type == RCSB_GEOM_CAPSULE) {  // end spheres
        const T oz = o[2] - s * hl;
        const T A = a + d[2] * d[2], B = b + oz * d[2], C = o[0] * o[0] + o[1] * o[1] + oz * oz - r * r, disc = B * B - A * C;
        if (disc >= 0) {
          const T ts = (-B - sqrt(disc)) / A;
          if (ts >= 0 && s * (o[2] + ts * d[2]) >= hl && ts < best) best = ts;
        }
      } else if (d[2] != 0) {           // end caps
        const T ts = (s * hl - o[2]) / d[2], x = o[0] + ts * d[0], y = o[1] + ts * d[1];
        if (ts >= 0 && s * d[2] < 0 && x * x + y * y <= r * r && ts < best) best = ts;
      }
    }
    if (best < (T)1e29) t = best;
  } else if (type == RCSB_GEOM_MESH) {
    {  // the hull lies inside its local AABB: a slab test first (most rays that pass the bounding sphere miss the box);
       // one reciprocal per axis and a little slack - the test only has to be conservative, the planes decide alone
      T t0 = 0, t1 = tmax;
      for (int k = 0; k < 3; k++) {
        if (d[k] != 0) {
          const T inv = (T)1 / d[k];
          T ta = (ab[k] - ab[3 + k] - o[k]) * inv, tb = (ab[k] + ab[3 + k] - o[k]) * inv;
          if (ta > tb) { T sw = ta; ta = tb; tb = sw; }
          t0 = ta > t0 ? ta : t0; t1 = tb < t1 ? tb : t1;
        } else if (o[k] < ab[k] - ab[3 + k] || o[k] > ab[k] + ab[3 + k]) t1 = -1;
      }
      if (t0 > t1 * (sizeof(T) == 4 ? (T)1.00001 : (T)1.000000001) + (sizeof(T) == 4 ? (T)1e-6 : (T)1e-12)) return false;
    }
    // Clip the ray against the hull's face planes. The entry / exit parameters are kept as fractions with positive
    // denominators and compared by cross-multiplication: one division per geom instead of one per face.
    T t0n = 0, t0d = 1, t1n = tmax, t1d = 1;
    const T* pl = faces + 4 * (size_t)face_adr[g];
    const int n = face_num[g];
    bool open = true;
    for (int i = 0; i < n && open; i++) {
      const T nx = __ldg(pl + 4 * i), ny = __ldg(pl + 4 * i + 1), nz = __ldg(pl + 4 * i + 2), dd = __ldg(pl + 4 * i + 3);
      const T den = nx * d[0] + ny * d[1] + nz * d[2], num = -(nx * o[0] + ny * o[1] + nz * o[2] + dd);
      if (den < 0) { if (-num * t0d > t0n * -den) { t0n = -num; t0d = -den; } }   // entering: t = num / den
      else if (den > 0) { if (num * t1d < t1n * den) { t1n = num; t1d = den; } }  // leaving
      else if (num < 0) open = false;
      open = open && t0n * t1d <= t1n * t0d;
    }
    if (n > 0 && open) t = t0n / t0d;
  }
  if (t < 0 || t > tmax) return false;
  *t_out = t;
  return true;
}
// One block renders RCSB_CAM_CHUNK tiles of one environment's image: the geom frames and the camera frame are set up once
// per block (a block per tile spent more time on that prologue than on its 256 rays), every warp culls the geoms of a few
// tiles, then the block walks through its tiles without further barriers.
#define RCSB_CAM_CHUNK 32
// T: arithmetic type of the rays (float by default: the reference's depth buffer is float32 and a float ray is good to a few
// micrometres, three orders below the millimetre the output is quantised to; double on request). Frames and the
// conservative tile culling stay in double.
template <typename T>
__global__ void __launch_bounds__(RCSB_CAM_TILE * RCSB_CAM_TILE)
rcsb_k_depth(const RcsbModel* __restrict__ gm, const T* __restrict__ faces, const int* __restrict__ face_adr,
             const int* __restrict__ face_num, const real* __restrict__ frames, RcsbCamera cam, unsigned short* __restrict__ out, real* __restrict__ cam_frames_out,
             int N) {
  const RcsbModel& m = *gm;
  __shared__ real gfr[RCSB_MAXG][12];   // world frame of every geom: position, row-major rotation
  __shared__ real gbs[RCSB_MAXG][4];    // bounding sphere: centre, radius
  __shared__ real cfr[12];              // camera frame in the world
  __shared__ unsigned tile_geoms[RCSB_CAM_CHUNK];  // per tile: geoms whose bounding sphere can meet one of its rays (planes always)
  __shared__ T rfr[RCSB_MAXG][12], rbs[RCSB_MAXG][4], rsz[RCSB_MAXG][3], rab[RCSB_MAXG][6], rcf[12];  // the rays' copies, in T
  const int tiles_x = (cam.W + RCSB_CAM_TILE - 1) / RCSB_CAM_TILE, tiles_y = (cam.H + RCSB_CAM_TILE - 1) / RCSB_CAM_TILE;
  const int tiles = tiles_x * tiles_y, chunks = (tiles + RCSB_CAM_CHUNK - 1) / RCSB_CAM_CHUNK;
  const int env = blockIdx.x / chunks, tile_lo = (blockIdx.x - env * chunks) * RCSB_CAM_CHUNK;
  const int ntile = tiles - tile_lo < RCSB_CAM_CHUNK ? tiles - tile_lo : RCSB_CAM_CHUNK;
  const int tid = threadIdx.y * RCSB_CAM_TILE + threadIdx.x;
  const real* fr = frames + (size_t)env * m.nb * 12;
  if (tid < m.ng) {
    const int g = tid, b = m.g_body[g];
    real p[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (b >= 0) { for (int i = 0; i < 3; i++) p[i] = fr[12 * b + i]; for (int i = 0; i < 9; i++) R[i] = fr[12 * b + 3 + i]; }
    const real* Rl = m.g_rot[g];
    for (int r = 0; r < 3; r++) {
      gfr[g][r] = p[r] + R[3 * r] * m.g_pos[g][0] + R[3 * r + 1] * m.g_pos[g][1] + R[3 * r + 2] * m.g_pos[g][2];
      gbs[g][r] = p[r] + R[3 * r] * m.g_bpos[g][0] + R[3 * r + 1] * m.g_bpos[g][1] + R[3 * r + 2] * m.g_bpos[g][2];
      for (int k = 0; k < 3; k++) gfr[g][3 + 3 * r + k] = R[3 * r] * Rl[k] + R[3 * r + 1] * Rl[3 + k] + R[3 * r + 2] * Rl[6 + k];
    }
    gbs[g][3] = m.g_rbound[g];
    for (int i = 0; i < 12; i++) rfr[g][i] = (T)gfr[g][i];
    for (int i = 0; i < 4; i++) rbs[g][i] = (T)gbs[g][i];
    for (int i = 0; i < 3; i++) rsz[g][i] = (T)m.g_size[g][i];
    for (int i = 0; i < 6; i++) rab[g][i] = (T)m.g_aabb[g][i];
  }
  if (tid == 64) {
    real p[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (cam.body >= 0) { for (int i = 0; i < 3; i++) p[i] = fr[12 * cam.body + i]; for (int i = 0; i < 9; i++) R[i] = fr[12 * cam.body + 3 + i]; }
    for (int r = 0; r < 3; r++) {
      cfr[r] = p[r] + R[3 * r] * cam.pos[0] + R[3 * r + 1] * cam.pos[1] + R[3 * r + 2] * cam.pos[2];
      for (int k = 0; k < 3; k++) cfr[3 + 3 * r + k] = R[3 * r] * cam.rot[k] + R[3 * r + 1] * cam.rot[3 + k] + R[3 * r + 2] * cam.rot[6 + k];
    }
    if (cam_frames_out && tile_lo == 0)  // the camera's world frame of this environment (extrinsics), written once
      for (int i = 0; i < 12; i++) cam_frames_out[(size_t)env * 12 + i] = cfr[i];
    for (int i = 0; i < 12; i++) rcf[i] = (T)cfr[i];
  }
  __syncthreads();
  // Tile culling, one warp per tile and one lane per geom: the tile's rays lie in a cone about its centre ray (half-angle
  // alpha, from the farthest corner); a sphere of radius r at distance dist subtends asin(r / dist). Geom g can only be
  // hit if the angle between the centre ray and the sphere centre is at most alpha + beta (tested through cosines, with
  // a little slack: the test is conservative, so culling never changes a pixel).
  for (int i = tid >> 5; i < ntile; i += (RCSB_CAM_TILE * RCSB_CAM_TILE) >> 5) {
    const int tile = tile_lo + i, g = tid & 31;
    int keep = 0;
    if (g < m.ng) {
      const real u0 = (tile % tiles_x) * RCSB_CAM_TILE, v0 = (tile / tiles_x) * RCSB_CAM_TILE;
      const real uc = u0 + (real)0.5 * RCSB_CAM_TILE, vc = v0 + (real)0.5 * RCSB_CAM_TILE;
      real dcn[3] = {(uc - (real)0.5 * cam.W) * cam.inv_f, -(vc - (real)0.5 * cam.H) * cam.inv_f, (real)-1};
      const real nc = sqrt(dcn[0] * dcn[0] + dcn[1] * dcn[1] + dcn[2] * dcn[2]);
      real cosa = 1;
      for (int k = 0; k < 4; k++) {
        const real uk = u0 + ((k & 1) ? (real)RCSB_CAM_TILE : (real)0), vk = v0 + ((k & 2) ? (real)RCSB_CAM_TILE : (real)0);
        const real e[3] = {(uk - (real)0.5 * cam.W) * cam.inv_f, -(vk - (real)0.5 * cam.H) * cam.inv_f, (real)-1};
        const real ce = (e[0] * dcn[0] + e[1] * dcn[1] + e[2] * dcn[2]) / (nc * sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]));
        cosa = ce < cosa ? ce : cosa;
      }
      cosa = cosa - (real)1e-9;
      const real sina = sqrt(1 - cosa * cosa > 0 ? 1 - cosa * cosa : (real)0);
      if (m.g_type[g] == RCSB_GEOM_PLANE) keep = 1;
      else {
        real cw[3];  // centre ray in the world
        for (int r = 0; r < 3; r++) cw[r] = (cfr[3 + 3 * r] * dcn[0] + cfr[3 + 3 * r + 1] * dcn[1] + cfr[3 + 3 * r + 2] * dcn[2]) / nc;
        const real cx = gbs[g][0] - cfr[0], cy = gbs[g][1] - cfr[1], cz = gbs[g][2] - cfr[2], rr = gbs[g][3] * (real)1.000001 + (real)1e-9;
        const real dist2 = cx * cx + cy * cy + cz * cz;
        if (dist2 <= rr * rr) keep = 1;  // the camera sits inside the sphere
        else {
          const real dist = sqrt(dist2), sinb = rr / dist, cosb = sqrt(1 - sinb * sinb);
          const real cost = (cx * cw[0] + cy * cw[1] + cz * cw[2]) / dist;
          const real cosab = cosa * cosb - sina * sinb;                 // cos(alpha + beta)
          keep = cost >= cosab;  // alpha, beta < 90 degrees: alpha + beta < 180, where the cosine is monotonic
        }
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if ((tid & 31) == 0) tile_geoms[i] = mask;
  }
  __syncthreads();
  const T ow[3] = {rcf[0], rcf[1], rcf[2]};
  const T inv_f = (T)cam.inv_f, half_w = (T)0.5 * cam.W, half_h = (T)0.5 * cam.H;
  // a float sphere test keeps a sliver of slack so that rounding never drops a geom a double test would keep
  const T slack = sizeof(T) == 4 ? (T)1.0001 : (T)1;
  int tx = tile_lo % tiles_x, ty = tile_lo / tiles_x;  // the tile walk is row-major: no division per tile and thread
  for (int i = 0; i < ntile; i++, tx++) {
    if (tx == tiles_x) { tx = 0; ty++; }
    const int u = tx * RCSB_CAM_TILE + threadIdx.x, v = ty * RCSB_CAM_TILE + threadIdx.y;
    if (u >= cam.W || v >= cam.H) continue;
    // pixel centre -> ray in the camera frame (x right, y up, looking along -z), scaled so that t is the eye-space depth
    const T dc[3] = {(u + (T)0.5 - half_w) * inv_f, -(v + (T)0.5 - half_h) * inv_f, (T)-1};
    T dw[3];
    for (int r = 0; r < 3; r++) dw[r] = rcf[3 + 3 * r] * dc[0] + rcf[3 + 3 * r + 1] * dc[1] + rcf[3 + 3 * r + 2] * dc[2];
    const T dlen2 = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2];
    T tbest = (T)cam.zfar;
    for (unsigned rest = tile_geoms[i]; rest; rest &= rest - 1) {
      const int g = __ffs(rest) - 1;
      if (m.g_type[g] != RCSB_GEOM_PLANE) {  // bounding sphere, tested without a division: |c x d|^2 > r^2 |d|^2 misses it
        const T cx = rbs[g][0] - ow[0], cy = rbs[g][1] - ow[1], cz = rbs[g][2] - ow[2], rr = rbs[g][3] * slack;
        const T cd = cx * dw[0] + cy * dw[1] + cz * dw[2], cc = cx * cx + cy * cy + cz * cz;
        if (cc * dlen2 - cd * cd > rr * rr * dlen2) continue;
        const T behind = cd - tbest * dlen2;  // (closest approach - best hit) x |d|^2
        if (behind > 0 && behind * behind > rr * rr * dlen2) continue;  // entirely behind the best hit
      }
      T og[3], dg[3], rel[3] = {ow[0] - rfr[g][0], ow[1] - rfr[g][1], ow[2] - rfr[g][2]};
      const T* R = &rfr[g][3];
      for (int k = 0; k < 3; k++) {
        og[k] = R[k] * rel[0] + R[3 + k] * rel[1] + R[6 + k] * rel[2];
        dg[k] = R[k] * dw[0] + R[3 + k] * dw[1] + R[6 + k] * dw[2];
      }
      T t;
      if (ray_geom<T>(m, g, faces, face_adr, face_num, rsz[g], rab[g], og, dg, tbest, &t) && t < tbest) tbest = t;
    }
    const T zn = (T)cam.znear, z = tbest < zn ? zn : tbest;  // eye-space depth, clipped like the view frustum
    T val;
    if (cam.physical_units) val = z * (T)1000;                                  // camera/sim.py:65-72 then x DEPTH_SCALE
    else val = (1 - zn / z) / (1 - zn / (T)cam.zfar) * (T)1000;                 // window-space depth x DEPTH_SCALE
    val = val < 0 ? (T)0 : (val > (T)65535 ? (T)65535 : val);
    out[((size_t)env * cam.H + v) * cam.W + u] = (unsigned short)val;           // astype(np.uint16): truncation
  }
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
  std::vector<real> faces;  // hull face planes (n, d) of the collidable mesh geoms, pooled (depth camera)
  std::vector<int> face_adr, face_num;
  real* d_faces = nullptr;
  float* d_faces32 = nullptr;  // the same planes for float rays
  int *d_face_adr = nullptr, *d_face_num = nullptr;
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
  real* d_frames = nullptr;  // [n][nb][12] body frames for the depth camera (RCSB_OP_FRAMES)
  int act_jstride = 0, act_gstride = 0;  // non-default action row strides of the next launch (packed host path)
  // optional contact export (rcsb_batch_set_contact_export)
  int *con_n = nullptr, *con_geom = nullptr, con_cap = 0;
  real* con_real = nullptr;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel, not per batch: batches of different scenes share the generic
// kernel, so the attribute only ever grows
#include <map>
static cudaError_t ensure_smem(rcsb_smem_fn fn, size_t bytes) {
  static std::map<rcsb_smem_fn, size_t> granted;
  size_t& have = granted[fn];
  if (bytes <= have) return cudaSuccess;
  cudaError_t e = fn(bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

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
  if (m->d_faces) cudaFree(m->d_faces);
  if (m->d_faces32) cudaFree(m->d_faces32);
  if (m->d_face_adr) cudaFree(m->d_face_adr);
  if (m->d_face_num) cudaFree(m->d_face_num);
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
int rcsb_model_set_mesh_faces(rcsb_model* m, const double* planes, int nface, const int* geom_faceadr, const int* geom_facenum, int ng) {
  if (!m || nface < 0 || ng < 0 || ng > RCSB_MAXG || (nface > 0 && !planes) || (ng > 0 && (!geom_faceadr || !geom_facenum)))
    return fail(RCSB_ERR_ARG, "bad mesh faces");
  for (int g = 0; g < ng; g++)
    if (geom_facenum[g] < 0 || (geom_facenum[g] > 0 && (geom_faceadr[g] < 0 || geom_faceadr[g] + geom_facenum[g] > nface)))
      return fail(RCSB_ERR_SIZE, "geom face range outside the plane pool");
  m->faces.resize((size_t)4 * (nface > 0 ? nface : 1));
  for (int i = 0; i < 4 * nface; i++) m->faces[i] = (real)planes[i];
  m->face_adr.assign(RCSB_MAXG, 0); m->face_num.assign(RCSB_MAXG, 0);
  for (int g = 0; g < ng; g++) { m->face_adr[g] = geom_facenum[g] > 0 ? geom_faceadr[g] : 0; m->face_num[g] = geom_facenum[g]; }
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
  if (m->face_adr.empty()) { m->faces.assign(4, 0); m->face_adr.assign(RCSB_MAXG, 0); m->face_num.assign(RCSB_MAXG, 0); }
  CUDA_OK(cudaMalloc(&m->d_faces, m->faces.size() * sizeof(real)));
  CUDA_OK(cudaMemcpy(m->d_faces, m->faces.data(), m->faces.size() * sizeof(real), cudaMemcpyHostToDevice));
  {
    std::vector<float> f32(m->faces.begin(), m->faces.end());
    CUDA_OK(cudaMalloc(&m->d_faces32, f32.size() * sizeof(float)));
    CUDA_OK(cudaMemcpy(m->d_faces32, f32.data(), f32.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  CUDA_OK(cudaMalloc(&m->d_face_adr, RCSB_MAXG * sizeof(int)));
  CUDA_OK(cudaMemcpy(m->d_face_adr, m->face_adr.data(), RCSB_MAXG * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&m->d_face_num, RCSB_MAXG * sizeof(int)));
  CUDA_OK(cudaMemcpy(m->d_face_num, m->face_num.data(), RCSB_MAXG * sizeof(int), cudaMemcpyHostToDevice));
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
  if (!getenv("RCSB_LOCKSTEP")) b->lockstep = b->var.lockstep;
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
  // Per-step alignment groups. A launch that gives every warp a single environment (one round) takes as long as its
  // slowest CTA, and a CTA as long as the sum over the steps of its slowest warp: two groups that align separately shorten
  // that (a straggler holds back 13 warps instead of 27; 0.737 -> 0.716 ms for 4096 x fr3_empty_world). With several
  // rounds per warp the CTAs even out anyway and the second code region in the instruction cache costs more than it
  // saves (5.77 -> 5.57 M env-steps/s at 16384), as it does for the generic kernel and for a dozen-warp CTA.
  if (!getenv("RCSB_BAR_GROUPS")) b->bar_groups = (b->warps >= 24 && b->var.fixed && n_envs <= b->grid * b->warps) ? 2 : 1;
  const size_t smem_max = b->smem > b->smem_full ? b->smem : b->smem_full;  // both phases may use the same kernel
  if (ensure_smem(b->var.set_smem, smem_max) != cudaSuccess || (m->has_reduced && ensure_smem(b->var_full.set_smem, smem_max) != cudaSuccess) ||
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
  cudaFree(b->d_counter); cudaFree(b->d_overflow); cudaFree(b->d_act_joints); cudaFree(b->d_act_gripper); cudaFree(b->d_act_packed); cudaFree(b->d_frames); cudaFree(b->d_obs); cudaFree(b->d_info);
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
int rcsb_debug_stage_trace(unsigned* out, int max_steps) {
#ifdef RCSB_STAGE_TIMING
  CUDA_OK(cudaDeviceSynchronize());
  if (out && max_steps > 0) {  // out[max_steps][10][32] stage cycles, then out2[max_steps][32] collision counts
    if (max_steps > RCSB_TRACE_STEPS) max_steps = RCSB_TRACE_STEPS;
    CUDA_OK(cudaMemcpyFromSymbol(out, rcsb_trace, (size_t)max_steps * 10 * 32 * sizeof(unsigned)));
    CUDA_OK(cudaMemcpyFromSymbol(out + (size_t)max_steps * 10 * 32, rcsb_trace_aux, (size_t)max_steps * 32 * sizeof(unsigned)));
    CUDA_OK(cudaMemcpyFromSymbol(out + (size_t)max_steps * 11 * 32, rcsb_trace_col, (size_t)max_steps * 3 * 32 * sizeof(unsigned)));
  }
  static unsigned zt[RCSB_TRACE_STEPS][10][32];
  CUDA_OK(cudaMemcpyToSymbol(rcsb_trace_aux, zt, sizeof(unsigned) * RCSB_TRACE_STEPS * 32));
  CUDA_OK(cudaMemcpyToSymbol(rcsb_trace_col, zt, sizeof(unsigned) * RCSB_TRACE_STEPS * 3 * 32));
  unsigned zs[32] = {0};
  CUDA_OK(cudaMemcpyToSymbol(rcsb_trace, zt, sizeof(zt)));
  CUDA_OK(cudaMemcpyToSymbol(rcsb_trace_step, zs, sizeof(zs)));
  return RCSB_OK;
#else
  (void)out; (void)max_steps;
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
  L.frames = b->d_frames;
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

// device-side alias of a page-locked host buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory), if it is one
static bool mapped_host_pointer(const void* host, void** dev) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
  *dev = a.devicePointer;
  return true;
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
  // A page-locked (device-mapped) action block is read by the kernel itself: every warp fetches its own 64-byte action row
  // over PCIe when it starts, which hides behind the other warps' physics and saves the copy launch. The observation rows
  // always go through the device block and ONE DMA copy: 240-byte rows posted by 4096 warps cross PCIe as small
  // transactions and take 40-70 us longer than the copy engine (measured, tools/diag_e2e.py). A pageable action block takes
  // the staged copy below. RCSB_HOST_ZEROCOPY=0 forces the staged path.
  {
    static const bool zero_copy = !(getenv("RCSB_HOST_ZEROCOPY") && atoi(getenv("RCSB_HOST_ZEROCOPY")) == 0);
    void* act_map = nullptr;
    if (zero_copy && mapped_host_pointer(act_host, &act_map)) {
      b->act_jstride = stride; b->act_gstride = stride;
      rc = rcsb_batch_run(b, ops | RCSB_OP_OBS, k, max_convergence_steps, act_map, (const real*)act_map + nj, nullptr, max_mov, jlow, jhigh,
                          b->d_obs, nullptr);
      b->act_jstride = 0; b->act_gstride = 0;
      if (rc) return rc;
      CUDA_OK(cudaMemcpyAsync(obs_host, b->d_obs, (size_t)b->n * RCSB_OBS_DIM * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
      CUDA_OK(cudaStreamSynchronize(b->stream));
      return RCSB_OK;
    }
  }
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

int rcsb_body_frames(rcsb_batch* b, void* frames_dev) {
  if (!b || !frames_dev) return fail(RCSB_ERR_ARG, "null argument");
  DEVICE_OK(b->m->device);
  real* keep = b->d_frames;
  b->d_frames = (real*)frames_dev;
  int rc = rcsb_batch_run(b, RCSB_OP_FRAMES, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
  b->d_frames = keep;
  return rc;
}
int rcsb_camera_depth(rcsb_batch* b, int cam_body, const double* cam_pos, const double* cam_rot, double fovy_deg, int width, int height,
                      double znear, double zfar, int physical_units, void* out_dev, void* cam_frames_dev) {
  if (!b || !cam_pos || !cam_rot || !out_dev || width < 1 || height < 1 || !(fovy_deg > 0 && fovy_deg < 180) || !(znear > 0 && zfar > znear))
    return fail(RCSB_ERR_ARG, "bad camera argument");
  if (cam_body >= b->m->h.nb) return fail(RCSB_ERR_ARG, "camera body out of range");
  DEVICE_OK(b->m->device);
  if (!b->d_frames) CUDA_OK(cudaMalloc(&b->d_frames, (size_t)b->n * b->m->h.nb * 12 * sizeof(real)));
  int rc = rcsb_batch_run(b, RCSB_OP_FRAMES, 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  RcsbCamera cam;
  cam.body = cam_body < 0 ? -1 : cam_body;
  for (int i = 0; i < 3; i++) cam.pos[i] = (real)cam_pos[i];
  for (int i = 0; i < 9; i++) cam.rot[i] = (real)cam_rot[i];
  cam.f = (real)(0.5 * height / tan(fovy_deg * 3.14159265358979323846 / 360.0));
  cam.inv_f = (real)1 / cam.f;
  cam.W = width; cam.H = height; cam.znear = (real)znear; cam.zfar = (real)zfar; cam.physical_units = physical_units != 0;
  const long long tiles = (long long)((width + RCSB_CAM_TILE - 1) / RCSB_CAM_TILE) * ((height + RCSB_CAM_TILE - 1) / RCSB_CAM_TILE);
  const long long chunks = (tiles + RCSB_CAM_CHUNK - 1) / RCSB_CAM_CHUNK;  // blocks per environment
  if (chunks * b->n > 0x7fffffffLL) return fail(RCSB_ERR_ARG, "image too large for one launch");
  dim3 block(RCSB_CAM_TILE, RCSB_CAM_TILE), grid((unsigned)(chunks * b->n));
  // float rays unless RCSB_DEPTH_F64=1 asks for double ones (the oracle comparison at the last digit)
  const char* f64 = getenv("RCSB_DEPTH_F64");
  if (f64 && atoi(f64) != 0)
    rcsb_k_depth<real><<<grid, block, 0, b->stream>>>(b->m->d_model, b->m->d_faces, b->m->d_face_adr, b->m->d_face_num, b->d_frames, cam,
                                                     (unsigned short*)out_dev, (real*)cam_frames_dev, b->n);
  else
    rcsb_k_depth<float><<<grid, block, 0, b->stream>>>(b->m->d_model, b->m->d_faces32, b->m->d_face_adr, b->m->d_face_num, b->d_frames, cam,
                                                      (unsigned short*)out_dev, (real*)cam_frames_dev, b->n);
  g_launches++;
  CUDA_OK(cudaGetLastError());
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
// ---- host-pointer variants for bindings that own no device memory (the compiled rcs_b200._core module): slow path,
//      synchronous, temporary device buffers per call
struct DevTmp {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  ~DevTmp() { if (p) cudaFree(p); }
};
int rcsb_batch_info(rcsb_batch* b, int* n_envs, int* njoints, int* ik_nq, int* nq, int* nv, int* nu) {
  if (!b) return fail(RCSB_ERR_ARG, "null batch");
  if (n_envs) *n_envs = b->n;
  if (njoints) *njoints = b->m->h.rb_njoints;
  if (ik_nq) *ik_nq = b->m->h.rb_ik_nq < 9 ? b->m->h.rb_ik_nq : 9;
  if (nq) *nq = b->m->h.nq;
  if (nv) *nv = b->m->h.nv;
  if (nu) *nu = b->m->h.nu;
  return RCSB_OK;
}
int rcsb_batch_read_row(rcsb_batch* b, int env, double* sr_row, double* sd_row, int* si_row) {
  if (!b || env < 0 || env >= b->n) return fail(RCSB_ERR_ARG, "bad environment index");
  if (sizeof(real) != sizeof(double)) return fail(RCSB_ERR_ARG, "float64 build required");
  DEVICE_OK(b->m->device);
  CUDA_OK(cudaStreamSynchronize(b->stream));
  const int nsr = b->m->h.lay.nsr;
  if (sr_row) CUDA_OK(cudaMemcpy(sr_row, b->sr + (size_t)env * nsr, nsr * sizeof(real), cudaMemcpyDeviceToHost));
  if (sd_row) CUDA_OK(cudaMemcpy(sd_row, b->sd + (size_t)env * RCSB_D_TAIL, RCSB_D_TAIL * sizeof(double), cudaMemcpyDeviceToHost));
  if (si_row) CUDA_OK(cudaMemcpy(si_row, b->si + (size_t)env * RCSB_I_TAIL, RCSB_I_TAIL * sizeof(int), cudaMemcpyDeviceToHost));
  return RCSB_OK;
}
int rcsb_robot_set_cartesian_position_host(rcsb_batch* b, const double* pose7_host) {
  if (!b || !pose7_host) return fail(RCSB_ERR_ARG, "null argument");
  DEVICE_OK(b->m->device);
  DevTmp p;
  CUDA_OK(p.alloc((size_t)b->n * 7 * sizeof(real)));
  CUDA_OK(cudaMemcpyAsync(p.p, pose7_host, (size_t)b->n * 7 * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  int rc = launch_ik(b, p.p, nullptr, nullptr, nullptr, nullptr, 1);
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return rc;
}
int rcsb_ik_inverse_host(rcsb_batch* b, const double* pose7_host, const double* q0_host, double* q_out_host, int* success_host,
                         int* iters_host) {
  if (!b || !pose7_host || !q0_host || !q_out_host || !success_host) return fail(RCSB_ERR_ARG, "null argument");
  DEVICE_OK(b->m->device);
  const int nj = b->m->h.rb_njoints, nqm = b->m->h.rb_ik_nq < 9 ? b->m->h.rb_ik_nq : 9;
  const size_t n = (size_t)b->n;
  DevTmp p, q0, q, ok, it;
  CUDA_OK(p.alloc(n * 7 * sizeof(real))); CUDA_OK(q0.alloc(n * nj * sizeof(real))); CUDA_OK(q.alloc(n * nqm * sizeof(real)));
  CUDA_OK(ok.alloc(n * sizeof(int))); CUDA_OK(it.alloc(n * sizeof(int)));
  CUDA_OK(cudaMemcpyAsync(p.p, pose7_host, n * 7 * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  CUDA_OK(cudaMemcpyAsync(q0.p, q0_host, n * nj * sizeof(real), cudaMemcpyHostToDevice, b->stream));
  CUDA_OK(cudaMemsetAsync(q.p, 0, n * nqm * sizeof(real), b->stream));
  int rc = launch_ik(b, p.p, q0.p, q.p, (int*)ok.p, (int*)it.p, 0);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(q_out_host, q.p, n * nqm * sizeof(real), cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaMemcpyAsync(success_host, ok.p, n * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (iters_host) CUDA_OK(cudaMemcpyAsync(iters_host, it.p, n * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return RCSB_OK;
}
int rcsb_ik_inverse(rcsb_batch* b, const void* pose_dev, const void* q0_dev, void* q_out_dev, int* success_dev, int* iters_dev) {
  if (!q0_dev || !q_out_dev || !success_dev) return fail(RCSB_ERR_ARG, "null argument");
  return launch_ik(b, pose_dev, q0_dev, q_out_dev, success_dev, iters_dev, 0);
}
int rcsb_robot_set_cartesian_position(rcsb_batch* b, const void* pose_dev) {
  return launch_ik(b, pose_dev, nullptr, nullptr, nullptr, nullptr, 1);
}
}  // extern "C"
