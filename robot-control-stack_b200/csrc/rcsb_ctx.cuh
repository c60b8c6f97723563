// Execution context of one environment's warp, shared by every kernel variant (rcsb_variant.cuh).
#pragma once
#include "rcsb_warp.cuh"

#ifdef RCSB_HOST_EMU
struct Ctx {
  const RcsbModel* md;  // model constants
  real* w;              // real workspace
  int* wi;              // int workspace
 const real* verts;    // convex hull vertex pool
  const int* vgraph;    // hull edge graph: [nmeshvert + 1] offsets, then neighbour lists (local vertex ids); may be null
  double* clk;          // simulation time + callback clocks (always double)
  int lane;
  int lockstep;
};
#define CMODEL(c) (*(c).md)
#define CMODEL_G(c) (*(c).md)
#define CW(c) ((c).w)
#define CWI(c) ((c).wi)
#define CCLK(c) ((c).clk)
#else
// Device: the model sits at the start of the CTA's dynamic shared memory and every warp owns a workspace window in it.
// Ctx carries 32-bit byte offsets into that window, and every access is formed from the rcsb_smem symbol, so the
// compiler addresses shared memory directly (LDS/STS with 32-bit address arithmetic) in every function, inlined or
// not, instead of falling back to generic 64-bit loads.
extern __shared__ __align__(128) unsigned char rcsb_smem[];
struct Ctx {
  const RcsbModel* gm;  // the model in global memory (cold tail: fields after cold_begin)
  const real* verts;    // convex hull vertex pool (global memory, read-only)
  const int* vgraph;    // hull edge graph (global memory): [nmeshvert + 1] offsets, then neighbour lists; may be null
  uint32_t wb;          // this warp's real workspace (byte offset in shared memory)
  uint32_t clkb;        // this warp's simulation time + callback clocks (always double)
  uint32_t wib;         // this warp's int workspace
  int lane;
  int lockstep;         // fixed-substep launch: CTA barriers between stages keep the warps on the same code
  int bar_id, bar_threads;  // named barrier of this warp's group and the number of threads that meet at it
  int conv_vote;        // step_until_convergence with a static env -> warp map: CTA-wide vote per step (run_env_program)
};
#define CMODEL(c) (*(const RcsbModel*)rcsb_smem)
#define CMODEL_G(c) (*(c).gm)
#define CW(c) ((real*)(rcsb_smem + (c).wb))
#define CWI(c) ((int*)(rcsb_smem + (c).wib))
#define CCLK(c) ((double*)(rcsb_smem + (c).clkb))
#endif

// misc int slots in the workspace
enum { MI_NCON = 0, MI_NEFC, MI_NE, MI_NF, MI_NL, MI_HAVE_L, MI_SOLVER_ITER, MI_WARN, MI_OVERFLOW, MI_HAVE_H2, MI_COUPLED, MI_COUNT };

#define WR(name) (CW(c) + LAY.o_##name)
#define WI(name) (CWI(c) + LAY.oi_##name)
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
__device__ unsigned long long rcsb_stage_cycles[16];  // profiling build = one translation unit (RCSB_SINGLE_TU)
// per-warp trace of CTA 0: cycles of stage i (0..8; 9 = whole step incl. barrier waits) of the first RCSB_TRACE_STEPS physics
// steps each warp runs after the last reset of the trace
#define RCSB_TRACE_STEPS 256
__device__ unsigned rcsb_trace[RCSB_TRACE_STEPS][10][32];
__device__ unsigned rcsb_trace_step[32];
__device__ unsigned rcsb_trace_col[RCSB_TRACE_STEPS][3][32];  // collision stage of an event: cycles up to the geom centres | the broad phase | the mid phase
__device__ unsigned rcsb_trace_aux[RCSB_TRACE_STEPS][32];  // collision stage: due groups | broad survivors << 8 | mid survivors << 16
#endif
