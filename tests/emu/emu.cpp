// TEST INFRASTRUCTURE ONLY. Host emulation of the per-warp device code (csrc/*.cuh compiled with
// RCSB_HOST_EMU: one host "lane", barriers/shuffles are no-ops). It lets the CPU-only test tier check
// the kernel *source logic* against the oracle without a GPU. It is never built into, linked by or
// reachable from the shipped library (robot-control-stack_b200/csrc/librcsb.so), which has no CPU path.
#define RCSB_HOST_EMU 1
#include <stdlib.h>
#include <vector>

#include "../../robot-control-stack_b200/csrc/rcsb_layout.h"
#include "../../robot-control-stack_b200/csrc/rcsb_ctx.cuh"
#define RCSB_VARIANT_NS rcsb_generic
#define RCSB_KERNEL unused
#include "../../robot-control-stack_b200/csrc/rcsb_variant.cuh"
using namespace rcsb_generic;

extern "C" {
struct EmuModel { RcsbModel full, reduced; int has_reduced; };
EmuModel* emu_model_new() { return (EmuModel*)calloc(1, sizeof(EmuModel)); }
void emu_model_free(EmuModel* m) { free(m); }
int emu_model_set_int(EmuModel* m, const char* f, const int* v, int n) { return rcsb_model_set_field(&m->full, f, v, n, 0); }
int emu_model_set_real(EmuModel* m, const char* f, const double* v, int n) { return rcsb_model_set_field(&m->full, f, v, n, 1); }
int emu_model_finalize(EmuModel* m, int use_reduced) {
  int rc = rcsb_model_finalize_layout(&m->full);
  m->has_reduced = use_reduced && rcsb_model_make_reduced(&m->full, &m->reduced);
  return rc;
}
int emu_nsr(const EmuModel* m) { return m->full.lay.nsr; }
int emu_has_reduced(const EmuModel* m) { return m->has_reduced; }
int emu_sizes(int* out) { out[0] = RCSB_S_TAIL; out[1] = RCSB_D_TAIL; out[2] = RCSB_I_TAIL; out[3] = RCSB_OBS_DIM; out[4] = RCSB_INFO_DIM; out[5] = (int)sizeof(real); return 0; }

static const int* g_vgraph = nullptr;
void emu_set_mesh_graph(const int* g) { g_vgraph = g; }
static void run_phase(const RcsbModel* m, const real* verts, real* sr, double* sd, int* si, RcsbLaunch L, const int* envs, int n,
                      const unsigned char* mask, real* dbg_ws, int dbg_stride, int* dbg_layout) {
  std::vector<real> w(m->lay.ws_reals);
  std::vector<int> wi(m->lay.ws_ints);
  double clk[RCSB_D_TAIL];
  for (int i = 0; i < n; i++) {
    int e = envs ? envs[i] : i;
    if (!envs && mask && !mask[e]) continue;
    Ctx c = {m, w.data(), wi.data(), verts, g_vgraph, clk, 0, 0};
    load_env(c, sr + (size_t)e * m->lay.nsr, sd + (size_t)e * RCSB_D_TAIL, si + (size_t)e * RCSB_I_TAIL);
    run_env_program(c, L, e);
    store_env(c, sr + (size_t)e * m->lay.nsr, sd + (size_t)e * RCSB_D_TAIL, si + (size_t)e * RCSB_I_TAIL);
    if (dbg_ws) memcpy(dbg_ws + (size_t)e * dbg_stride, w.data(), sizeof(real) * m->lay.ws_reals);
    if (dbg_layout) dbg_layout[e] = m->cap_reduced;
  }
}
// optional contact export, as rcsb_batch_set_contact_export
static int* g_con_n = nullptr; static int* g_con_geom = nullptr; static real* g_con_real = nullptr; static int g_con_cap = 0;
void emu_set_contact_export(int* n, int* geom, real* r, int cap) { g_con_n = n; g_con_geom = geom; g_con_real = r; g_con_cap = cap; }
// run the per-launch program over N environments, serially: reduced layout first (when present), then the full
// layout for the environments that outgrew it -- the two launches of rcsb_batch_run
int emu_run(const EmuModel* em, const real* verts, real* sr, double* sd, int* si, int N, unsigned ops, int k,
            int max_conv, const real* act_joints, const real* act_gripper, const unsigned char* mask, real max_mov,
            const real* jlow, const real* jhigh, real* obs, int* info, real* dbg_ws, int* dbg_layout) {
  RcsbLaunch L;
  memset(&L, 0, sizeof(L));
  L.N = N; L.ops = ops; L.k = k; L.max_convergence_steps = max_conv;
  L.act_joints = act_joints; L.act_gripper = act_gripper; L.mask = mask; L.max_mov = max_mov;
  L.act_jstride = em->full.rb_njoints; L.act_gstride = 1;
  for (int i = 0; i < RCSB_MAXJ; i++) { L.jlow[i] = jlow ? jlow[i] : 0; L.jhigh[i] = jhigh ? jhigh[i] : 0; }
  L.obs = obs; L.info = info;
  L.con_n = g_con_n; L.con_geom = g_con_geom; L.con_real = g_con_real; L.con_cap = g_con_cap;
  std::vector<int> list(N);
  int count = 0;
  L.overflow_list = list.data(); L.overflow_count = &count; L.phase = 0;
  const int stride = em->full.lay.ws_reals;
  run_phase(em->has_reduced ? &em->reduced : &em->full, verts, sr, sd, si, L, nullptr, N, mask, dbg_ws, stride, dbg_layout);
  int handed = count;
  if (em->has_reduced && handed > 0) {
    L.phase = 1;
    std::vector<int> envs(list.begin(), list.begin() + handed);
    run_phase(&em->full, verts, sr, sd, si, L, envs.data(), handed, nullptr, dbg_ws, stride, dbg_layout);
  }
  return handed;
}
// the RcsbShape of a model (what a fixed-shape kernel variant must be compiled for), 19 ints
void emu_shape(const EmuModel* em, int reduced, int* out) {
  RcsbShape s = rcsb_model_shape(reduced ? &em->reduced : &em->full);
  memcpy(out, &s, sizeof(RcsbShape));
}
int emu_offset(const EmuModel* em, const char* name, int reduced) {
  const RcsbModel* m = reduced ? &em->reduced : &em->full;
#define OFF(n) if (!strcmp(name, #n)) return m->lay.o_##n;
  OFF(q) OFF(v) OFF(ctrl) OFF(warm) OFF(bpos) OFF(bquat) OFF(bmat) OFF(rootcom) OFF(cinert) OFF(crb) OFF(cdof)
  OFF(M) OFF(L) OFF(H) OFF(bias) OFF(passive) OFF(gravc) OFF(actfrc) OFF(smooth) OFF(qacc_smooth) OFF(qacc) OFF(qfc)
  OFF(gpos) OFF(con) OFF(J) OFF(efc) OFF(rcs)
#undef OFF
  if (!strcmp(name, "ws_reals")) return m->lay.ws_reals;
  if (!strcmp(name, "maxefc")) return m->maxefc;
  if (!strcmp(name, "maxcon")) return m->maxcon;
  return -1;
}
}
