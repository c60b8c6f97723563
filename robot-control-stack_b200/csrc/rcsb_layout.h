// Host-side helpers shared by the CUDA library and the test-only host emulation build:
// named-field access to RcsbModel and the per-warp workspace layout.
#pragma once
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "rcsb_types.h"

struct RcsbField { const char* name; size_t off; size_t bytes; int kind; };  // kind: 0 int32/uint32, 1 real, 2 uint8
#define RCSB_FI(f) {#f, offsetof(RcsbModel, f), sizeof(((RcsbModel*)0)->f), 0}
#define RCSB_FR(f) {#f, offsetof(RcsbModel, f), sizeof(((RcsbModel*)0)->f), 1}
#define RCSB_FB(f) {#f, offsetof(RcsbModel, f), sizeof(((RcsbModel*)0)->f), 2}

static const RcsbField rcsb_model_fields[] = {
    RCSB_FI(nq), RCSB_FI(nv), RCSB_FI(nu), RCSB_FI(nb), RCSB_FI(ng), RCSB_FI(npair), RCSB_FI(nt), RCSB_FI(neq),
    RCSB_FI(nroot), RCSB_FI(nmeshvert), RCSB_FI(cone_elliptic), RCSB_FI(implicitfast), RCSB_FI(iterations),
    RCSB_FI(ls_iterations), RCSB_FI(noslip_iterations), RCSB_FI(maxcon), RCSB_FI(maxefc), RCSB_FI(fast_maxcon), RCSB_FI(fast_maxefc),
    RCSB_FR(timestep), RCSB_FR(gravity), RCSB_FR(impratio), RCSB_FR(tolerance), RCSB_FR(ls_tolerance),
    RCSB_FR(noslip_tolerance), RCSB_FR(meaninertia),
    RCSB_FI(b_parent), RCSB_FI(b_jtype), RCSB_FI(b_qadr), RCSB_FI(b_dadr), RCSB_FI(b_ndof), RCSB_FI(b_root),
    RCSB_FI(b_ancmask), RCSB_FI(b_descmask), RCSB_FI(b_dofmask),
    RCSB_FR(b_pos), RCSB_FR(b_quat), RCSB_FR(b_jpos), RCSB_FR(b_jaxis), RCSB_FR(b_mass), RCSB_FR(b_ipos),
    RCSB_FR(b_inertia), RCSB_FR(b_gcmass), RCSB_FR(b_gcpos),
    RCSB_FI(d_body), RCSB_FI(d_qadr), RCSB_FI(d_limited), RCSB_FI(d_actfrclimited), RCSB_FI(d_actgravcomp),
    RCSB_FI(d_dotzero), RCSB_FI(d_premask), RCSB_FI(d_ancmask),
    RCSB_FR(d_armature), RCSB_FR(d_damping), RCSB_FR(d_frictionloss), RCSB_FR(d_invweight0), RCSB_FR(d_range),
    RCSB_FR(d_margin), RCSB_FR(d_solref), RCSB_FR(d_solimp), RCSB_FR(d_actfrcrange), RCSB_FR(qpos0), RCSB_FR(r_invmass),
    RCSB_FI(g_body), RCSB_FI(g_type), RCSB_FI(g_vertadr), RCSB_FI(g_vertnum), RCSB_FI(g_origid), RCSB_FI(g_role),
    RCSB_FI(g_condim), RCSB_FI(g_priority),
    RCSB_FR(g_pos), RCSB_FR(g_quat), RCSB_FR(g_bpos), RCSB_FR(g_size), RCSB_FR(g_rbound), RCSB_FR(g_aabb), RCSB_FR(g_friction),
    RCSB_FR(g_solref), RCSB_FR(g_solimp), RCSB_FR(g_solmix), RCSB_FR(g_margin), RCSB_FR(g_gap), RCSB_FR(g_invweight), RCSB_FR(g_rbound0),
    RCSB_FB(pair),
    RCSB_FR(t_coef), RCSB_FI(e_dof1), RCSB_FI(e_dof2), RCSB_FI(e_active), RCSB_FR(e_poly), RCSB_FR(e_solref),
    RCSB_FR(e_solimp),
    RCSB_FI(n_special), RCSB_FI(a_special), RCSB_FR(d_kvdiag), RCSB_FI(a_trntype), RCSB_FI(a_trnid), RCSB_FI(a_ctrllimited), RCSB_FI(a_forcelimited), RCSB_FR(a_gear),
    RCSB_FR(a_gain), RCSB_FR(a_bias), RCSB_FR(a_ctrlrange), RCSB_FR(a_forcerange),
    RCSB_FI(rb_njoints), RCSB_FI(rb_qadr), RCSB_FI(rb_act), RCSB_FI(rb_site_body), RCSB_FI(rb_register_convergence),
    RCSB_FI(rb_ik_nq), RCSB_FR(rb_site_pos), RCSB_FR(rb_site_quat), RCSB_FR(rb_base_pos), RCSB_FR(rb_base_quat),
    RCSB_FR(rb_tcp_offset), RCSB_FR(rb_q_home), RCSB_FR(rb_joint_tol), RCSB_FR(rb_cb_period),
    RCSB_FI(gr_enabled), RCSB_FI(gr_act), RCSB_FI(gr_qadr), RCSB_FR(gr_eps_inner), RCSB_FR(gr_eps_outer),
    RCSB_FR(gr_cb_period), RCSB_FR(gr_max_act), RCSB_FR(gr_min_act), RCSB_FR(gr_max_joint), RCSB_FR(gr_min_joint),
};
#define RCSB_NFIELDS ((int)(sizeof(rcsb_model_fields) / sizeof(rcsb_model_fields[0])))

// Set a model field from host doubles / ints. Reals are always passed as double and converted to `real`.
static inline int rcsb_model_set_field(RcsbModel* m, const char* name, const void* data, int count, int src_is_double) {
  for (int i = 0; i < RCSB_NFIELDS; i++) {
    const RcsbField& f = rcsb_model_fields[i];
    if (strcmp(f.name, name) != 0) continue;
    char* dst = (char*)m + f.off;
    if (f.kind == 1) {
      if (!src_is_double || (size_t)count * sizeof(real) > f.bytes) return -2;
      for (int k = 0; k < count; k++) ((real*)dst)[k] = (real)((const double*)data)[k];
    } else if (f.kind == 0) {
      if (src_is_double || (size_t)count * sizeof(int) > f.bytes) return -2;
      memcpy(dst, data, (size_t)count * sizeof(int));
    } else {
      if (src_is_double || (size_t)count > f.bytes) return -2;
      for (int k = 0; k < count; k++) ((uint8_t*)dst)[k] = (uint8_t)((const int*)data)[k];
    }
    return 0;
  }
  return -1;
}

// Workspace layout. The first nsr reals mirror the env's HBM row: q | v | ctrl | warm | RCS tail.
// Lifetimes within one physics step (stage order: kinematics, com, crb, collision, velocity, make_constraint,
// callbacks, actuation, constraint solve, integrate) decide what may share memory:
//   region K  position-stage results read up to make_constraint: bpos bmat rootcom cinert cdof
//   region U  stage-local scratch, one union: kinematics local frames | crb + crb buffer | geom centres + broad-phase
//             candidate lists | cdofdot cvel cfrc
//   region S  solver / integrator scratch, first written after make_constraint -> aliases K and U
//   persistent across the step: M, force vectors, contacts, constraint rows
#ifdef __CUDACC__
#define RCSB_HD __host__ __device__
#else
#define RCSB_HD
#endif
RCSB_HD constexpr RcsbLayout rcsb_make_layout(const RcsbShape& s) {
  RcsbLayout y = {};
  const int nq = s.nq, nv = s.nv, nu = s.nu, nb = s.nb;
  int o = 0;
#define RCSB_ALLOC(field, n) do { y.field = o; o += (n); } while (0)
#define RCSB_MAX(a, b) ((a) > (b) ? (a) : (b))
  RCSB_ALLOC(o_q, nq); RCSB_ALLOC(o_v, nv); RCSB_ALLOC(o_ctrl, nu); RCSB_ALLOC(o_warm, nv);
  RCSB_ALLOC(o_rcs, RCSB_S_TAIL);
  // collision groups' separation budgets (floats) and the qpos they refer to: they persist across launches, and a row
  // whose qpos was changed from outside (host writes, resets) fails the comparison and starts with every group due
  RCSB_ALLOC(o_cbud, (s.ngrp * (int)sizeof(float) + (int)sizeof(real) - 1) / (int)sizeof(real));
  RCSB_ALLOC(o_cbq, nq);
  // separating-direction cache of the convex narrow phase: RCSB_SEPSLOTS x (pair tag + 1, unit direction). Persistent: a
  // cached direction is re-validated with one support pair before it is trusted, so it may outlive launches and resets
  RCSB_ALLOC(o_sepcache, 4 * RCSB_SEPSLOTS);
  o = (o + 1) & ~1;  // rows stay 16-byte granular in HBM when real is 8 bytes
  y.nsr = o;
  y.o_site = y.o_rcs + RCSB_S_SITEPOS;
  const int k_begin = o;
  RCSB_ALLOC(o_bpos, 3 * nb); RCSB_ALLOC(o_bmat, 9 * nb); RCSB_ALLOC(o_rootcom, 3 * s.nroot);
  RCSB_ALLOC(o_cinert, 10 * nb); RCSB_ALLOC(o_cdof, 6 * nv);
  const int u_begin = o;
  int u_end = u_begin;
  RCSB_ALLOC(o_bquat, 12 * nb);  // kinematics: local rotations [nb][9] + local translations [nb][3]
  u_end = RCSB_MAX(u_end, o); o = u_begin;
  RCSB_ALLOC(o_crb, 10 * nb); RCSB_ALLOC(o_crbbuf, 6 * nv);
  u_end = RCSB_MAX(u_end, o); o = u_begin;
  RCSB_ALLOC(o_gpos, 3 * s.ng); RCSB_ALLOC(o_cand, (2 * RCSB_MAXCAND * (int)sizeof(uint16_t) + (int)sizeof(real) - 1) / (int)sizeof(real));
  RCSB_ALLOC(o_pairfr, 24);  // narrow phase: world frames [p | R] of the pair being tested
  RCSB_ALLOC(o_sup, 48);     // narrow phase: portal of the Minkowski refinement, 5 support records (45) | box-box clip polygons 2 x [8][3]
  u_end = RCSB_MAX(u_end, o); o = u_begin;
  RCSB_ALLOC(o_cdofdot, 6 * nv); RCSB_ALLOC(o_cvel, 6 * nv); RCSB_ALLOC(o_cacc, 6 * nv);
  y.o_gcw = y.o_cacc;  // gravity-compensation wrenches replace the accelerations, body b in the slot of its last dof
  y.o_cfrc = y.o_cdofdot;  // body forces take over the cdof_dot slots once the accelerations are known (nb <= nv)
  u_end = RCSB_MAX(u_end, o);
  const int k_end = u_end;
  o = k_begin;
  RCSB_ALLOC(o_H, nv * nv + nv); RCSB_ALLOC(o_L, nv * nv + nv);
  RCSB_ALLOC(o_grad, nv); RCSB_ALLOC(o_search, nv); RCSB_ALLOC(o_Ma, nv); RCSB_ALLOC(o_Mv, nv);
  RCSB_ALLOC(o_tmp, RCSB_MAX(nv, RCSB_MAXJ) + 2);  // triangular-solve scratch (host emulation), action staging
  RCSB_ALLOC(o_conehess, s.cone_elliptic ? 9 * s.maxcon : 0);
  RCSB_ALLOC(o_noslip, s.noslip_iterations > 0 ? (s.nfl + 2 * s.maxcon) * nv + s.nfl + 4 * s.maxcon : 0);  // M^-1 J^T of every friction row + diagonal blocks
  if (o < k_end) o = k_end;
  RCSB_ALLOC(o_M, nv * nv);
  RCSB_ALLOC(o_bias, nv); RCSB_ALLOC(o_passive, nv); RCSB_ALLOC(o_gravc, nv); RCSB_ALLOC(o_actfrc, nv);
  RCSB_ALLOC(o_smooth, nv); RCSB_ALLOC(o_qacc_smooth, nv); RCSB_ALLOC(o_qacc, nv); RCSB_ALLOC(o_qfc, nv);
  RCSB_ALLOC(o_aforce, nu);
  RCSB_ALLOC(o_con, RCSB_C_REALS * s.maxcon);
  RCSB_ALLOC(o_J, s.maxefc * nv);
  RCSB_ALLOC(o_efc, RCSB_E_NARR * s.maxefc);
  o = (o + 1) & ~1;  // keep the double clock block 16-byte aligned when real is 8 bytes
  y.ws_reals = o;
  y.ws_doubles = RCSB_D_TAIL;
  o = 0;
  RCSB_ALLOC(oi_con, RCSB_CI_INTS * s.maxcon);
  RCSB_ALLOC(oi_efc, RCSB_EI_NARR * s.maxefc);
  RCSB_ALLOC(oi_misc, 11 /* MI_COUNT */ + RCSB_I_TAIL);
  y.ws_ints = (o + 3) & ~3;
#undef RCSB_ALLOC
#undef RCSB_MAX
  return y;
}
// group reach table entries: float32 truncated to its upper 16 bits, rounded towards +inf (the bound stays a bound)
static inline uint16_t rcsb_reach_encode(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  uint16_t h = (uint16_t)(u >> 16);
  if (u & 0xffffu) h++;  // x >= 0: the next representable value up
  return h;
}
static inline void rcsb_host_quat_to_mat(real* M, const real* q) {  // same expression order as quat_to_mat (rcsb_warp.cuh)
  real w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
  M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
  M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
static inline RcsbShape rcsb_model_shape(const RcsbModel* m) {
  RcsbShape s = {m->nq, m->nv, m->nu, m->nb, m->ng, m->npair, m->nt, m->neq, m->nroot, m->maxcon, m->maxefc, m->rb_njoints,
                 m->cone_elliptic, m->implicitfast, m->noslip_iterations, m->cap_reduced, m->gr_enabled, m->ngrp, m->nfl};
  return s;
}
static inline int rcsb_model_finalize_layout(RcsbModel* m) {
  if (m->nq <= 0 || m->nq > RCSB_MAXQ || m->nv <= 0 || m->nv > RCSB_MAXV || m->nu > RCSB_MAXU || m->nb > RCSB_MAXB ||
      m->ng > RCSB_MAXG || m->npair > RCSB_MAXPAIR || m->nt > RCSB_MAXT || m->neq > RCSB_MAXEQ ||
      m->nroot > RCSB_MAXROOT || m->rb_njoints > RCSB_MAXJ || m->maxcon < 1 || m->maxefc < 1)
    return -1;
  // ---- collision groups: one per pair of bodies that owns candidate geom pairs (world = -1), in order of appearance
  {
    int ga[RCSB_MAXGRP], gb[RCSB_MAXGRP];
    m->ngrp = 0;
    for (int p = 0; p < m->npair; p++) {
      int a = m->g_body[m->pair[p][0]], b = m->g_body[m->pair[p][1]], g = -1;
      if (a > b) { int t = a; a = b; b = t; }
      for (int i = 0; i < m->ngrp; i++) if (ga[i] == a && gb[i] == b) { g = i; break; }
      if (g < 0) {
        if (m->ngrp >= RCSB_MAXGRP) return -1;
        g = m->ngrp++;
        ga[g] = a; gb[g] = b;
        m->grp_mask[g] = (a >= 0 ? m->b_dofmask[a] : 0u) ^ (b >= 0 ? m->b_dofmask[b] : 0u);  // dofs on the tree path
      }
      m->pair_grp[p] = (uint8_t)g;
    }
    memset(m->pair_blk_grps, 0, sizeof(m->pair_blk_grps));
    for (int p = 0; p < m->npair; p++) m->pair_blk_grps[p >> 5][m->pair_grp[p] >> 5] |= 1u << (m->pair_grp[p] & 31);
    // E[x]: extent of body x's own collidable geoms about its frame origin; off[x]: bound of |origin of x in its parent|
    real E[RCSB_MAXB], off[RCSB_MAXB];
    for (int x = 0; x < m->nb; x++) {
      real e = 0;
      for (int g = 0; g < m->ng; g++)
        if (m->g_body[g] == x) {
          real r = sqrt(m->g_bpos[g][0] * m->g_bpos[g][0] + m->g_bpos[g][1] * m->g_bpos[g][1] + m->g_bpos[g][2] * m->g_bpos[g][2]) + m->g_rbound[g];
          if (r > e) e = r;
        }
      E[x] = e;
      real o = sqrt(m->b_pos[x][0] * m->b_pos[x][0] + m->b_pos[x][1] * m->b_pos[x][1] + m->b_pos[x][2] * m->b_pos[x][2]);
      o += 2 * sqrt(m->b_jpos[x][0] * m->b_jpos[x][0] + m->b_jpos[x][1] * m->b_jpos[x][1] + m->b_jpos[x][2] * m->b_jpos[x][2]);
      if (m->b_jtype[x] == RCSB_JNT_SLIDE) {
        int j = m->b_dadr[x];
        real lo = fabs(m->d_range[j][0]), hi = fabs(m->d_range[j][1]);
        o += m->d_limited[j] ? (lo > hi ? lo : hi) : (real)10;
      }
      off[x] = o;
    }
    memset(m->grp_reach, 0, sizeof(m->grp_reach));
    for (int g = 0; g < m->ngrp; g++) {
      m->grp_body[g][0] = (int8_t)ga[g]; m->grp_body[g][1] = (int8_t)gb[g];
      for (int j = 0; j < m->nv; j++) {
        if (!((m->grp_mask[g] >> j) & 1u)) continue;
        // the dof is an ancestor dof of exactly one of the two bodies: that body's geoms are what it moves
        const int x = (ga[g] >= 0 && ((m->b_dofmask[ga[g]] >> j) & 1u)) ? ga[g] : gb[g];
        const int bj = m->d_body[j];
        const int translational = m->b_jtype[bj] == RCSB_JNT_SLIDE || (m->b_jtype[bj] == RCSB_JNT_FREE && j - m->b_dadr[bj] < 3);
        real r = 1;
        if (!translational) {
          // joint anchor -> origin of its own body -> down the chain to x -> farthest geom point of x
          r = sqrt(m->b_jpos[bj][0] * m->b_jpos[bj][0] + m->b_jpos[bj][1] * m->b_jpos[bj][1] + m->b_jpos[bj][2] * m->b_jpos[bj][2]) + E[x];
          for (int k = x; k != bj && k >= 0; k = m->b_parent[k]) r += off[k];
        }
        m->grp_reach[g][j] = rcsb_reach_encode((float)(r * 1.000001));
      }
    }
  }
  m->nfl = 0;
  for (int j = 0; j < m->nv; j++) m->nfl += m->d_frictionloss[j] > 0;
  m->lay = rcsb_make_layout(rcsb_model_shape(m));
  for (int a = 0; a < RCSB_MAXU; a++)
    for (int k = 0; k < RCSB_MAXV; k++) {
      real v = 0;
      if (a < m->nu && k < m->nv) {
        if (m->a_trntype[a] == RCSB_TRN_JOINT) v = m->a_trnid[a] == k ? m->a_gear[a] : (real)0;
        else v = m->a_gear[a] * m->t_coef[m->a_trnid[a]][k];
      }
      m->a_moment[a][k] = v;
    }
  for (int b = 0; b < m->nb; b++) rcsb_host_quat_to_mat(m->b_rot[b], m->b_quat[b]);
  for (int g = 0; g < m->ng; g++) rcsb_host_quat_to_mat(m->g_rot[g], m->g_quat[g]);
  rcsb_host_quat_to_mat(m->rb_site_rot, m->rb_site_quat);
  for (int b = 0; b < m->nb; b++) {
    const int da = m->b_dadr[b], nd = m->b_ndof[b], p = m->b_parent[b];
    m->b_lastdof[b] = da + nd - 1;
    for (int a = 0; a < nd; a++) {
      const int j = da + a, up = p >= 0 ? m->b_dadr[p] + m->b_ndof[p] - 1 : -1;
      m->d_parent[j] = a > 0 ? j - 1 : up;
      // free joint: the rotational dofs see the velocity after the three translational ones; hinge / slide: the parent's
      m->d_pre[j] = (m->b_jtype[b] == RCSB_JNT_FREE) ? (a >= 3 ? da + 2 : -1) : up;
    }
  }
  {
    int lo[RCSB_MAXROOT], hi[RCSB_MAXROOT], cnt[RCSB_MAXROOT], ok = 1;
    for (int r = 0; r < RCSB_MAXROOT; r++) { lo[r] = m->nv; hi[r] = 0; cnt[r] = 0; }
    for (int j = 0; j < m->nv; j++) {
      int r = m->b_root[m->d_body[j]];
      if (j < lo[r]) lo[r] = j;
      if (j + 1 > hi[r]) hi[r] = j + 1;
      cnt[r]++;
    }
    for (int r = 0; r < m->nroot; r++) if (cnt[r] != hi[r] - lo[r]) ok = 0;
    for (int j = 0; j < m->nv; j++) {
      int r = m->b_root[m->d_body[j]];
      m->d_tree_lo[j] = ok ? lo[r] : 0;
      m->d_tree_hi[j] = ok ? hi[r] : m->nv;
    }
  }
  {  // trees whose dofs carry no velocity-dependent force: build_integrator_matrix leaves their block of M as it is
    int plain = (1 << m->nroot) - 1;
    for (int j = 0; j < m->nv; j++) {
      int vel = m->d_damping[j] != 0 || (m->implicitfast && m->d_kvdiag[j] != 0);
      if (m->implicitfast)
        for (int sa = 0; sa < m->n_special; sa++) {
          const int a = m->a_special[sa];
          if (m->a_bias[a][2] != 0 && m->a_moment[a][j] != 0) vel = 1;
        }
      if (vel) plain &= ~(1 << m->b_root[m->d_body[j]]);
    }
    m->root_plain = plain;
  }
  for (int i = 0, t = 0; i < RCSB_MAXV; i++)
    for (int j = 0; j <= i; j++, t++) { m->tri_i[t] = (uint8_t)i; m->tri_j[t] = (uint8_t)j; }
  return 0;
}
// The reduced-capacity copy of a finalised model (see RcsbModel::fast_maxcon); returns 0 when there is none.
static inline int rcsb_model_make_reduced(const RcsbModel* full, RcsbModel* out) {
  if (full->fast_maxcon <= 0 || full->fast_maxefc <= 0 || (full->fast_maxcon >= full->maxcon && full->fast_maxefc >= full->maxefc))
    return 0;
  *out = *full;
  if (out->fast_maxcon < out->maxcon) out->maxcon = out->fast_maxcon;
  if (out->fast_maxefc < out->maxefc) out->maxefc = out->fast_maxefc;
  out->cap_reduced = 1;
  return rcsb_model_finalize_layout(out) == 0;
}
static inline size_t rcsb_ws_bytes(const RcsbModel* m) {
  size_t b = (size_t)m->lay.ws_reals * sizeof(real) + (size_t)m->lay.ws_doubles * sizeof(double) + (size_t)m->lay.ws_ints * sizeof(int);
  return (b + 15) & ~(size_t)15;
}
