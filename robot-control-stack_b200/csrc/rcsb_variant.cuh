// One kernel variant: the physics / RCS device code compiled for a shape that is either read from the model at run
// time (generic) or fixed at compile time (RCSB_FIXED_SHAPE: loop bounds, feature switches and every workspace offset
// fold into constants). Included once per variant from rcsb_kernels.cu with
//   RCSB_VARIANT_NS   namespace of the variant
//   RCSB_KERNEL       name of its __global__ entry point
//   RCSB_FIXED_SHAPE  (optional) brace initialiser of an RcsbShape
#if defined(RCSB_FIXED_SHAPE) && !defined(RCSB_SINGLE_TU)
#define RCSB_VARIANT_LINKAGE
#else
#define RCSB_VARIANT_LINKAGE static
#endif
namespace RCSB_VARIANT_NS {
#ifdef RCSB_FIXED_SHAPE
static constexpr RcsbShape kShape = RCSB_FIXED_SHAPE;
static constexpr RcsbLayout kLay = rcsb_make_layout(kShape);
#define MD(f) (kShape.f)
#define LAY kLay
#else
#define MD(f) (m.f)
#define LAY (m.lay)
#endif
#include "rcsb_dynamics.cuh"
#include "rcsb_solver.cuh"
#include "rcsb_env.cuh"

#ifndef RCSB_HOST_EMU
__device__ __forceinline__ Ctx make_ctx(const RcsbModel* sm, const RcsbModel* gm, const real* verts, const int* vgraph, size_t ws_bytes) {
  const RcsbModel& m = *sm;
  (void)m;
  int warp = threadIdx.x >> 5;
  Ctx c;
  c.wb = (uint32_t)(RCSB_SMEM_HEADER + (size_t)warp * ws_bytes);
  c.clkb = c.wb + (uint32_t)((size_t)LAY.ws_reals * sizeof(real));
  c.wib = c.clkb + (uint32_t)((size_t)LAY.ws_doubles * sizeof(double));
  c.gm = gm;
  c.verts = verts;
  c.vgraph = vgraph;
  c.lane = threadIdx.x & 31;
  c.lockstep = 0;
  c.bar_id = 0; c.bar_threads = blockDim.x;
  c.conv_vote = 0;
  return c;
}
// ------------------------------------------------------------------ the per-launch program kernel
#ifndef RCSB_VARIANT_WARPS
#define RCSB_VARIANT_WARPS RCSB_MAX_WARPS  // warps per CTA the variant is compiled for: fewer warps, more registers each
#endif
__global__ void __launch_bounds__(RCSB_VARIANT_WARPS * 32, 1)
RCSB_KERNEL(const RcsbModel* __restrict__ gm, const real* __restrict__ verts, const int* __restrict__ vgraph, real* __restrict__ sr, double* __restrict__ sd,
           int* __restrict__ si, RcsbLaunch L, int* __restrict__ counter, size_t ws_bytes) {
  if (L.phase == 1 && *L.overflow_count == 0) return;  // the common case: nothing outgrew the reduced layout
  const RcsbModel* sm = stage_model(gm);
  const RcsbModel& m = *sm;
  (void)m;
  Ctx c = make_ctx(sm, gm, verts, vgraph, ws_bytes);
  if ((L.ops & RCSB_OP_STEP_K) && L.phase == 0) {
    // Fixed-substep launch: static env -> warp mapping. Every warp of the CTA runs the same number of rounds and
    // hits exactly L.k * RCSB_STAGE_BARRIERS CTA barriers per round (inside run_env_program, or here when it has no environment), which
    // keeps the warps in the same stage of the step so that they share instruction-cache lines.
    c.lockstep = L.lockstep;
    {  // barrier groups: contiguous blocks of warps, sizes differ by at most one warp
      const int W_ = blockDim.x >> 5, G_ = L.bar_groups < 1 ? 1 : (L.bar_groups > W_ ? W_ : L.bar_groups), w_ = threadIdx.x >> 5;
      const int g_ = w_ * G_ / W_, lo_ = (g_ * W_ + G_ - 1) / G_, hi_ = ((g_ + 1) * W_ + G_ - 1) / G_;
      c.bar_id = 1 + g_;
      c.bar_threads = (hi_ - lo_) * 32;
    }
    const int nbar = L.k * __popc((unsigned)L.lockstep);
    const int W = blockDim.x >> 5, per_round = gridDim.x * W;
    const int rounds = (L.N + per_round - 1) / per_round;
    for (int r = 0; r < rounds; r++) {
      int env = r * per_round + (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
      bool valid = env < L.N && !(L.mask && !L.mask[env]);
      if (valid) {
        load_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
        run_env_program(c, L, env);
        store_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
        __syncwarp();
      } else {
        for (int i = 0; i < nbar; i++) RCSB_GROUP_BARRIER();
      }
#ifndef RCSB_HOST_EMU
      // groups that align separately drift apart by up to a step per round; bring them back together between rounds
      if (L.bar_groups > 1 && r + 1 < rounds) __syncthreads();
#endif
    }
    return;
  }
  if ((L.ops & RCSB_OP_STEP_CONV) && L.phase == 0 && L.conv_vote) {
    // step_until_convergence with the static map: the warps of a CTA step together until the last of their environments
    // has converged (a vote per step); warps without an environment in the last round only vote
    c.conv_vote = 1;
    const int W = blockDim.x >> 5, per_round = gridDim.x * W;
    const int rounds = (L.N + per_round - 1) / per_round;
    for (int r = 0; r < rounds; r++) {
      int env = r * per_round + (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
      bool valid = env < L.N && !(L.mask && !L.mask[env]);
      if (valid) {
        load_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
        run_env_program(c, L, env);
        store_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
        __syncwarp();
      } else {
        while (rcsb_cta_vote(0)) {}
      }
    }
    return;
  }
  if (L.phase == 1) {  // environments the reduced layout handed over: dynamic scheduling over the overflow list
    const int n = *L.overflow_count;
    for (;;) {
      int i = 0;
      if (c.lane == 0) i = atomicAdd(counter, 1);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (i >= n) break;
      int env = L.overflow_list[i];
      load_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
      run_env_program(c, L, env);
      store_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
      __syncwarp();
    }
    return;
  }
  for (;;) {
    int env = 0;
    if (c.lane == 0) env = atomicAdd(counter, 1);
    env = __shfl_sync(0xffffffffu, env, 0);
    if (env >= L.N) break;
    if (L.mask && !L.mask[env]) continue;
    load_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
    run_env_program(c, L, env);
    store_env(c, sr + (size_t)env * LAY.nsr, sd + (size_t)env * RCSB_D_TAIL, si + (size_t)env * RCSB_I_TAIL);
    __syncwarp();
  }
}


RCSB_VARIANT_LINKAGE void launch(int grid, int threads, size_t smem, cudaStream_t stream, const RcsbModel* gm, const real* verts,
                   const int* vgraph, real* sr, double* sd, int* si, const RcsbLaunch& L, int* counter, size_t ws_bytes) {
  RCSB_KERNEL<<<grid, threads, smem, stream>>>(gm, verts, vgraph, sr, sd, si, L, counter, ws_bytes);
}
RCSB_VARIANT_LINKAGE cudaError_t set_smem(size_t bytes) {
  return cudaFuncSetAttribute(RCSB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
RCSB_VARIANT_LINKAGE int max_warps() { return RCSB_VARIANT_WARPS; }
#ifdef RCSB_FIXED_SHAPE
RCSB_VARIANT_LINKAGE RcsbShape shape() { return kShape; }
#endif
#endif
#undef MD
#undef LAY
#undef RCSB_VARIANT_LINKAGE
#undef RCSB_VARIANT_WARPS
}  // namespace
