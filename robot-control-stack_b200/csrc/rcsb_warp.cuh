// Warp-level execution vocabulary of the per-environment physics code: one environment per warp,
// lanes cooperate over shared-memory arrays. Every stage is a "parallel for" over independent work
// items followed by a warp barrier; reductions go through shuffles.
//
// RCSB_HOST_EMU compiles the same stage code for a single host "lane" (NLANES == 1, barriers and
// shuffles become no-ops). That build exists only under tests/ to check the stage logic against the
// oracle on a machine without a GPU; the shipped library has no CPU path.
#pragma once
#include <math.h>

#include "rcsb_types.h"

#ifdef RCSB_HOST_EMU
#define RCSB_DEV static inline
#define RCSB_DEV_NOINLINE static
#define RCSB_NLANES 1
#define RCSB_SYNC() ((void)0)
#define RCSB_STAGE_SYNC(i) ((void)0)
#define RCSB_STEP_SYNC() ((void)0)
#ifdef RCSB_EMU_REVERSE  // run every parallel-for backwards: results must not depend on lane order
#define PFOR(i, n) for (int i = (n)-1; i >= 0; --i)
#else
#define PFOR(i, n) for (int i = 0; i < (n); ++i)
#endif
#define PFOR1(i, n) PFOR(i, n)
RCSB_DEV real warp_sum(real x) { return x; }
RCSB_DEV real warp_max(real x) { return x; }
RCSB_DEV int warp_any(int p) { return p; }
RCSB_DEV int warp_all(int p) { return p; }
RCSB_DEV unsigned warp_ballot(int p) { return p ? 1u : 0u; }
RCSB_DEV void warp_argmax(real& v, int& idx) {}
RCSB_DEV real warp_bcast(real x, int src) { return x; }
RCSB_DEV int warp_bcast_i(int x, int src) { return x; }
#define RCSB_LDG(p) (*(p))
#else
#define RCSB_DEV __device__ __forceinline__
// large routines called from several sites: keep one copy (the hot loop already overflows the instruction cache)
#define RCSB_DEV_NOINLINE __device__ __noinline__
#define RCSB_NLANES 32
#define RCSB_SYNC() __syncwarp()
// CTA-wide barrier used only in fixed-substep launches, where every warp of the CTA (including warps without an
// environment, see rcsb_k_run) executes exactly the same number of them
// c.lockstep is a bit mask: bit i (0..8) = CTA barrier before stage i of the physics step, bit 9 = barrier at its end
// The CTA's warps are split into bar_groups groups that align separately (named barriers): a straggler only holds
// back its own group, the other groups fill the issue slots meanwhile.
// One instruction address for every arrival (out-of-line): warps that run an environment and warps that only keep the
// barrier count meet at the same barrier instruction, which is also what compute-sanitizer's synccheck expects.
static __device__ __noinline__ void rcsb_group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
static __device__ __noinline__ int rcsb_cta_vote(int pred) { return __syncthreads_or(pred); }
#define RCSB_GROUP_BARRIER() rcsb_group_barrier(c.bar_id, c.bar_threads)
#define RCSB_STAGE_SYNC(i) do { if ((c.lockstep >> (i)) & 1) RCSB_GROUP_BARRIER(); } while (0)
#define RCSB_STEP_SYNC() RCSB_STAGE_SYNC(9)
#define RCSB_LOCKSTEP_ALL 0x3ff
#define RCSB_LOCKSTEP_STEP 0x200
#define PFOR(i, n) for (int i = (int)(threadIdx.x & 31); i < (n); i += 32)
// same for n <= 32 work items: one guarded pass, no loop
#define PFOR1(i, n) for (int i = (int)(threadIdx.x & 31), once_ = 1; once_ && i < (n); once_ = 0)
RCSB_DEV real warp_sum(real x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
RCSB_DEV real warp_max(real x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    real y = __shfl_xor_sync(0xffffffffu, x, o);
    x = y > x ? y : x;
  }
  return x;
}
RCSB_DEV int warp_any(int p) { return __any_sync(0xffffffffu, p); }
RCSB_DEV int warp_all(int p) { return __all_sync(0xffffffffu, p); }
RCSB_DEV unsigned warp_ballot(int p) { return __ballot_sync(0xffffffffu, p); }
// arg-max with ties resolved to the lowest index (== first maximum of a serial scan)
RCSB_DEV void warp_argmax(real& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    real ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
RCSB_DEV real warp_bcast(real x, int src) { return __shfl_sync(0xffffffffu, x, src); }
RCSB_DEV int warp_bcast_i(int x, int src) { return __shfl_sync(0xffffffffu, x, src); }
#define RCSB_LDG(p) __ldg(p)
#endif

// ---------------------------------------------------------------- scalar / 3-vector helpers
#define RCSB_MINVAL ((real)1e-15)
RCSB_DEV real r_sqrt(real x) { return sqrt(x); }
RCSB_DEV real r_abs(real x) { return fabs(x); }
#ifdef RCSB_HOST_EMU
RCSB_DEV real r_rsqrt(real x) { return (real)1 / sqrt(x); }
#else
RCSB_DEV real r_rsqrt(real x) { return (real)1 / sqrt(x); }  // IEEE path kept: rsqrt() differs from the oracle in the last bits
#endif
RCSB_DEV real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
RCSB_DEV void cross3(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
RCSB_DEV void copy3(real* r, const real* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
RCSB_DEV real norm3(const real* a) { return r_sqrt(dot3(a, a)); }
RCSB_DEV real normalize3(real* a) {
  real n = norm3(a);
  if (n < RCSB_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  real s = (real)1 / n;
  a[0] *= s; a[1] *= s; a[2] *= s;
  return n;
}
RCSB_DEV void mulmat3(real* r, const real* M, const real* v) {  // row-major 3x3
  real x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  real y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  real z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
RCSB_DEV void mulmatT3(real* r, const real* M, const real* v) {
  real x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  real y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  real z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
RCSB_DEV void quat_mul(real* r, const real* a, const real* b) {  // (w,x,y,z)
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
RCSB_DEV void quat_to_mat(real* M, const real* q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
  M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
  M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
RCSB_DEV void quat_normalize(real* q) {
  real n = r_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < RCSB_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  real s = (real)1 / n;
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
RCSB_DEV void make_frame(real* f) {  // f[0:3] normal -> tangents (mju_makeFrame semantics)
  real* x = f; real* y = f + 3; real* z = f + 6;
  normalize3(x);
  if (x[1] > (real)-0.5 && x[1] < (real)0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  real dd = dot3(x, y);
  y[0] -= dd * x[0]; y[1] -= dd * x[1]; y[2] -= dd * x[2];
  normalize3(y);
  cross3(z, x, y);
}
// spatial algebra in MuJoCo's "com frame": motion = [angular; linear], inertia = 10-vector
RCSB_DEV void mul_inert_vec(real* r, const real* I, const real* v) {
  real c[3];
  r[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  r[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  r[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  cross3(c, I + 6, v + 3);
  r[0] += c[0]; r[1] += c[1]; r[2] += c[2];
  cross3(c, I + 6, v);
  r[3] = I[9] * v[3] - c[0]; r[4] = I[9] * v[4] - c[1]; r[5] = I[9] * v[5] - c[2];
}
RCSB_DEV void cross_motion(real* r, const real* v, const real* s) {
  real a[3], b[3], c[3];
  cross3(a, v, s); cross3(b, v, s + 3); cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
RCSB_DEV void cross_force(real* r, const real* v, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
