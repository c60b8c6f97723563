/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h).
 * CPU restatement of the smooth-dynamics half of MuJoCo 3.2.6's mj_step [3P], reached from the
 * reference at /root/reference/src/sim/sim.cpp:110-112 (mj_step1 / mj_step2). Algorithms follow the
 * MuJoCo documentation "Computation" chapter: mj_kinematics, mj_comPos, mj_tendon, mj_crb,
 * mj_factorM, mj_transmission, mj_comVel, mj_passive (+gravcomp), mj_rne, mj_fwdActuation,
 * mj_fwdAcceleration, implicitfast integration. */
#include <stdio.h>

#include "oracle_internal.h"

/* ------------------------------------------------------------------ math */
void rcso_quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
void rcso_quat_to_mat(double* M, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
  M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
  M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
void rcso_quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
void rcso_axisangle_quat(double* q, const double* axis, double angle) {
  double s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
void rcso_rot_vec_quat(double* r, const double* v, const double* q) {
  double M[9];
  rcso_quat_to_mat(M, q);
  mulmat3(r, M, v);
}
void rcso_make_frame(double* f) {
  double* x = f; double* y = f + 3; double* z = f + 6;
  normalize3(x);
  if (x[1] > -0.5 && x[1] < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  double dd = dot3(x, y);
  for (int i = 0; i < 3; i++) y[i] -= dd * x[i];
  normalize3(y);
  cross3(z, x, y);
}
int rcso_chol_factor(double* A, int n) {
  int deficit = 0;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < MINVAL) { s = MINVAL; deficit++; }
    s = sqrt(s);
    A[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / s;
    }
  }
  return deficit;
}
void rcso_chol_solve(const double* L, int n, double* x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* ------------------------------------------------------------------ model */
typedef struct { const char* name; size_t off; int is_int; } field_t;
#define MF(name, isint) {#name, offsetof(struct rcso_model, name), isint}
static const field_t model_fields[] = {
    MF(body_parentid, 1), MF(body_rootid, 1), MF(body_weldid, 1), MF(body_jntnum, 1), MF(body_jntadr, 1),
    MF(body_dofnum, 1), MF(body_dofadr, 1), MF(body_pos, 0), MF(body_quat, 0), MF(body_ipos, 0), MF(body_iquat, 0),
    MF(body_mass, 0), MF(body_inertia, 0), MF(body_gravcomp, 0), MF(body_invweight0, 0),
    MF(jnt_type, 1), MF(jnt_bodyid, 1), MF(jnt_qposadr, 1), MF(jnt_dofadr, 1), MF(jnt_limited, 1),
    MF(jnt_actfrclimited, 1), MF(jnt_actgravcomp, 1), MF(jnt_pos, 0), MF(jnt_axis, 0), MF(jnt_range, 0),
    MF(jnt_margin, 0), MF(jnt_solref, 0), MF(jnt_solimp, 0), MF(jnt_actfrcrange, 0),
    MF(dof_jntid, 1), MF(dof_bodyid, 1), MF(dof_parentid, 1), MF(dof_armature, 0), MF(dof_damping, 0),
    MF(dof_frictionloss, 0), MF(dof_invweight0, 0), MF(qpos0, 0),
    MF(geom_type, 1), MF(geom_bodyid, 1), MF(geom_condim, 1), MF(geom_priority, 1), MF(geom_vertadr, 1),
    MF(geom_vertnum, 1), MF(geom_size, 0), MF(geom_pos, 0), MF(geom_quat, 0), MF(geom_friction, 0),
    MF(geom_solref, 0), MF(geom_solimp, 0), MF(geom_solmix, 0), MF(geom_margin, 0), MF(geom_gap, 0),
    MF(geom_rbound, 0), MF(geom_aabb, 0), MF(geom_bsphere, 0), MF(mesh_vert, 0), MF(mesh_graphadr, 1), MF(mesh_graph, 1),
    MF(pair_geom, 1),
    MF(site_bodyid, 1), MF(site_pos, 0), MF(site_quat, 0), MF(tendon_coef, 0), MF(tendon_invweight0, 0),
    MF(eq_obj1id, 1), MF(eq_obj2id, 1), MF(eq_active0, 1), MF(eq_polycoef, 0), MF(eq_solref, 0), MF(eq_solimp, 0),
    MF(actuator_trntype, 1), MF(actuator_trnid, 1), MF(actuator_ctrllimited, 1), MF(actuator_forcelimited, 1),
    MF(actuator_gear, 0), MF(actuator_gainprm, 0), MF(actuator_biasprm, 0), MF(actuator_ctrlrange, 0),
    MF(actuator_forcerange, 0),
};
#define NMODEL_FIELDS ((int)(sizeof(model_fields) / sizeof(model_fields[0])))

rcso_model* rcso_model_new(void) { return (rcso_model*)calloc(1, sizeof(rcso_model)); }
void rcso_model_free(rcso_model* m) {
  if (!m) return;
  for (int i = 0; i < NMODEL_FIELDS; i++) free(*(void**)((char*)m + model_fields[i].off));
  free(m);
}
int rcso_model_set_int(rcso_model* m, const char* field, const int* v, int n) {
  if (!strcmp(field, "sizes")) { /* nq nv nu nbody njnt ngeom nsite ntendon neq npair nmeshvert */
    if (n != 11) return -1;
    m->nq = v[0]; m->nv = v[1]; m->nu = v[2]; m->nbody = v[3]; m->njnt = v[4]; m->ngeom = v[5]; m->nsite = v[6];
    m->ntendon = v[7]; m->neq = v[8]; m->npair = v[9]; m->nmeshvert = v[10];
    return 0;
  }
  if (!strcmp(field, "opt_int")) { /* iterations ls_iterations noslip_iterations cone_elliptic implicitfast */
    if (n != 5) return -1;
    m->iterations = v[0]; m->ls_iterations = v[1]; m->noslip_iterations = v[2]; m->cone_elliptic = v[3];
    m->integrator_implicitfast = v[4];
    return 0;
  }
  for (int i = 0; i < NMODEL_FIELDS; i++)
    if (model_fields[i].is_int && !strcmp(field, model_fields[i].name)) {
      int** p = (int**)((char*)m + model_fields[i].off);
      free(*p);
      *p = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
      memcpy(*p, v, sizeof(int) * (size_t)n);
      return 0;
    }
  return -1;
}
int rcso_model_set_real(rcso_model* m, const char* field, const double* v, int n) {
  if (!strcmp(field, "opt_real")) { /* timestep g[3] impratio tolerance noslip_tol ls_tol meaninertia */
    if (n != 9) return -1;
    m->timestep = v[0]; m->gravity[0] = v[1]; m->gravity[1] = v[2]; m->gravity[2] = v[3]; m->impratio = v[4];
    m->tolerance = v[5]; m->noslip_tolerance = v[6]; m->ls_tolerance = v[7]; m->meaninertia = v[8];
    return 0;
  }
  for (int i = 0; i < NMODEL_FIELDS; i++)
    if (!model_fields[i].is_int && !strcmp(field, model_fields[i].name)) {
      double** p = (double**)((char*)m + model_fields[i].off);
      free(*p);
      *p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
      memcpy(*p, v, sizeof(double) * (size_t)n);
      return 0;
    }
  return -1;
}
int rcso_model_finalize(rcso_model* m) {
  if (m->nv <= 0 || m->nbody <= 0 || m->timestep <= 0) return -1;
  for (int i = 0; i < NMODEL_FIELDS; i++)
    if (*(void**)((char*)m + model_fields[i].off) == NULL) {
      fprintf(stderr, "rcs_oracle: model field %s missing\n", model_fields[i].name);
      return -2;
    }
  for (int b = 1; b < m->nbody; b++) {
    if (m->body_jntnum[b] > 1) return -3; /* one joint per body in all shipped scenes */
  }
  return 0;
}

/* ------------------------------------------------------------------ data */
typedef struct { const char* name; size_t off; int kind; } dfield_t; /* kind: per-size code */
#define DF(name) {#name, offsetof(struct rcso_data, name), 0}
static const dfield_t data_fields[] = {
    DF(qpos), DF(qvel), DF(ctrl), DF(qacc_warmstart), DF(qacc), DF(xpos), DF(xquat), DF(xmat), DF(xipos),
    DF(ximat), DF(xanchor), DF(xaxis), DF(geom_xpos), DF(geom_xmat), DF(site_xpos), DF(site_xmat),
    DF(subtree_com), DF(cinert), DF(crb), DF(cdof), DF(qM), DF(qLD), DF(ten_length), DF(ten_J),
    DF(actuator_length), DF(actuator_moment), DF(cvel), DF(cdof_dot), DF(cacc), DF(cfrc), DF(actuator_velocity),
    DF(qfrc_bias), DF(qfrc_passive), DF(qfrc_gravcomp), DF(actuator_force), DF(qfrc_actuator), DF(qfrc_smooth),
    DF(qacc_smooth), DF(qfrc_constraint), DF(qDeriv), DF(efc_J),
};
#define NDATA_FIELDS ((int)(sizeof(data_fields) / sizeof(data_fields[0])))
static int data_field_size(const rcso_model* m, const char* f) {
  int nq = m->nq, nv = m->nv, nu = m->nu, nb = m->nbody;
  if (!strcmp(f, "qpos")) return nq;
  if (!strcmp(f, "qvel") || !strcmp(f, "qacc_warmstart") || !strcmp(f, "qacc") || !strcmp(f, "qfrc_bias") ||
      !strcmp(f, "qfrc_passive") || !strcmp(f, "qfrc_gravcomp") || !strcmp(f, "qfrc_actuator") ||
      !strcmp(f, "qfrc_smooth") || !strcmp(f, "qacc_smooth") || !strcmp(f, "qfrc_constraint"))
    return nv;
  if (!strcmp(f, "ctrl") || !strcmp(f, "actuator_length") || !strcmp(f, "actuator_velocity") ||
      !strcmp(f, "actuator_force"))
    return nu;
  if (!strcmp(f, "xpos") || !strcmp(f, "xipos") || !strcmp(f, "subtree_com")) return 3 * nb;
  if (!strcmp(f, "xquat")) return 4 * nb;
  if (!strcmp(f, "xmat") || !strcmp(f, "ximat")) return 9 * nb;
  if (!strcmp(f, "xanchor") || !strcmp(f, "xaxis")) return 3 * m->njnt;
  if (!strcmp(f, "geom_xpos")) return 3 * m->ngeom;
  if (!strcmp(f, "geom_xmat")) return 9 * m->ngeom;
  if (!strcmp(f, "site_xpos")) return 3 * m->nsite;
  if (!strcmp(f, "site_xmat")) return 9 * m->nsite;
  if (!strcmp(f, "cinert") || !strcmp(f, "crb")) return 10 * nb;
  if (!strcmp(f, "cdof") || !strcmp(f, "cdof_dot")) return 6 * nv;
  if (!strcmp(f, "cvel") || !strcmp(f, "cacc") || !strcmp(f, "cfrc")) return 6 * nb;
  if (!strcmp(f, "qM") || !strcmp(f, "qLD") || !strcmp(f, "qDeriv")) return nv * nv;
  if (!strcmp(f, "ten_length")) return m->ntendon;
  if (!strcmp(f, "ten_J")) return m->ntendon * nv;
  if (!strcmp(f, "actuator_moment")) return nu * nv;
  if (!strcmp(f, "efc_J")) return MAXEFC * nv;
  return -1;
}
rcso_data* rcso_data_new(const rcso_model* m) {
  rcso_data* d = (rcso_data*)calloc(1, sizeof(rcso_data));
  d->m = m;
  for (int i = 0; i < NDATA_FIELDS; i++) {
    int n = data_field_size(m, data_fields[i].name);
    *(double**)((char*)d + data_fields[i].off) = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  }
  rcso_reset_data(m, d);
  return d;
}
void rcso_data_free(rcso_data* d) {
  if (!d) return;
  for (int i = 0; i < NDATA_FIELDS; i++) free(*(void**)((char*)d + data_fields[i].off));
  free(d);
}
double* rcso_data_real(rcso_data* d, const char* field, int* n) {
  if (!strcmp(field, "time")) { if (n) *n = 1; return &d->time; }
  if (!strcmp(field, "contact_real")) {
    for (int i = 0; i < d->ncon; i++) {
      double* o = d->contact_flat + 7 * i;
      o[0] = d->contact[i].dist;
      for (int k = 0; k < 3; k++) { o[1 + k] = d->contact[i].pos[k]; o[4 + k] = d->contact[i].frame[k]; }
    }
    if (n) *n = 7 * d->ncon;
    return d->contact_flat;
  }
  static const struct { const char* name; size_t off; } efc[] = {
      {"efc_pos", offsetof(struct rcso_data, efc_pos)}, {"efc_D", offsetof(struct rcso_data, efc_D)},
      {"efc_R", offsetof(struct rcso_data, efc_R)}, {"efc_aref", offsetof(struct rcso_data, efc_aref)},
      {"efc_force", offsetof(struct rcso_data, efc_force)}, {"efc_vel", offsetof(struct rcso_data, efc_vel)},
      {"efc_b", offsetof(struct rcso_data, efc_b)}};
  for (unsigned i = 0; i < sizeof(efc) / sizeof(efc[0]); i++)
    if (!strcmp(field, efc[i].name)) { if (n) *n = d->nefc; return (double*)((char*)d + efc[i].off); }
  for (int i = 0; i < NDATA_FIELDS; i++)
    if (!strcmp(field, data_fields[i].name)) {
      if (n) *n = data_field_size(d->m, field);
      return *(double**)((char*)d + data_fields[i].off);
    }
  return NULL;
}
int* rcso_data_int(rcso_data* d, const char* field, int* n) {
  if (!strcmp(field, "ncon")) { if (n) *n = 1; return &d->ncon; }
  if (!strcmp(field, "nefc")) { if (n) *n = 1; return &d->nefc; }
  if (!strcmp(field, "solver_iter")) { if (n) *n = 1; return &d->solver_iter; }
  if (!strcmp(field, "warnings")) { if (n) *n = 1; return &d->warnings; }
  if (!strcmp(field, "efc_type")) { if (n) *n = d->nefc; return d->efc_type; }
  if (!strcmp(field, "efc_state")) { if (n) *n = d->nefc; return d->efc_state; }
  if (!strcmp(field, "contact_geom")) {
    for (int i = 0; i < d->ncon; i++) { d->contact_geom[2 * i] = d->contact[i].geom[0]; d->contact_geom[2 * i + 1] = d->contact[i].geom[1]; }
    if (n) *n = 2 * d->ncon;
    return d->contact_geom;
  }
  return NULL;
}

/* mj_resetData [3P]: qpos <- qpos0, everything else zero (sim.cpp:118) */
void rcso_reset_data(const rcso_model* m, rcso_data* d) {
  for (int i = 0; i < NDATA_FIELDS; i++) {
    int n = data_field_size(m, data_fields[i].name);
    zero(*(double**)((char*)d + data_fields[i].off), n > 0 ? n : 0);
  }
  if (m->qpos0) memcpy(d->qpos, m->qpos0, sizeof(double) * (size_t)m->nq);
  d->time = 0;
  d->ncon = d->nefc = d->ne = d->nf = d->nl = 0;
  d->solver_iter = 0;
  if (m->nbody) { d->xquat[0] = 1; d->xmat[0] = d->xmat[4] = d->xmat[8] = 1; }
}

/* ------------------------------------------------------------------ mj_kinematics [3P] */
void rcso_kinematics(const rcso_model* m, rcso_data* d) {
  /* normalise free-joint quaternions in qpos */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JNT_FREE) rcso_quat_normalize(d->qpos + m->jnt_qposadr[j] + 3);
  zero(d->xpos, 3); d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  zero(d->xmat, 9); d->xmat[0] = d->xmat[4] = d->xmat[8] = 1;
  zero(d->xipos, 3); memcpy(d->ximat, d->xmat, 9 * sizeof(double));
  for (int i = 1; i < m->nbody; i++) {
    int p = m->body_parentid[i];
    double xpos[3], xquat[4];
    int free_body = m->body_jntnum[i] && m->jnt_type[m->body_jntadr[i]] == JNT_FREE;
    if (free_body) {
      int a = m->jnt_qposadr[m->body_jntadr[i]];
      copy3(xpos, d->qpos + a);
      memcpy(xquat, d->qpos + a + 3, 4 * sizeof(double));
      int j = m->body_jntadr[i];
      copy3(d->xanchor + 3 * j, xpos);
      d->xaxis[3 * j] = 0; d->xaxis[3 * j + 1] = 0; d->xaxis[3 * j + 2] = 1;
    } else {
      double v[3];
      mulmat3(v, d->xmat + 9 * p, m->body_pos + 3 * i);
      for (int k = 0; k < 3; k++) xpos[k] = d->xpos[3 * p + k] + v[k];
      rcso_quat_mul(xquat, d->xquat + 4 * p, m->body_quat + 4 * i);
      for (int jj = 0; jj < m->body_jntnum[i]; jj++) {
        int j = m->body_jntadr[i] + jj;
        double* xanchor = d->xanchor + 3 * j;
        double* xaxis = d->xaxis + 3 * j;
        double R[9];
        rcso_quat_to_mat(R, xquat);
        mulmat3(xaxis, R, m->jnt_axis + 3 * j);
        mulmat3(v, R, m->jnt_pos + 3 * j);
        for (int k = 0; k < 3; k++) xanchor[k] = xpos[k] + v[k];
        double q = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        if (m->jnt_type[j] == JNT_SLIDE) {
          for (int k = 0; k < 3; k++) xpos[k] += xaxis[k] * q;
        } else { /* hinge: rotate about the joint axis through the anchor */
          double qloc[4], qn[4];
          rcso_axisangle_quat(qloc, m->jnt_axis + 3 * j, q);
          rcso_quat_mul(qn, xquat, qloc);
          memcpy(xquat, qn, sizeof(qn));
          rcso_rot_vec_quat(v, m->jnt_pos + 3 * j, xquat);
          for (int k = 0; k < 3; k++) xpos[k] = xanchor[k] - v[k];
        }
      }
    }
    rcso_quat_normalize(xquat);
    copy3(d->xpos + 3 * i, xpos);
    memcpy(d->xquat + 4 * i, xquat, sizeof(xquat));
    rcso_quat_to_mat(d->xmat + 9 * i, xquat);
    double v[3], qi[4];
    mulmat3(v, d->xmat + 9 * i, m->body_ipos + 3 * i);
    for (int k = 0; k < 3; k++) d->xipos[3 * i + k] = xpos[k] + v[k];
    rcso_quat_mul(qi, xquat, m->body_iquat + 4 * i);
    rcso_quat_to_mat(d->ximat + 9 * i, qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double v[3], q[4];
    mulmat3(v, d->xmat + 9 * b, m->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + v[k];
    rcso_quat_mul(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    rcso_quat_to_mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double v[3], q[4];
    mulmat3(v, d->xmat + 9 * b, m->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) d->site_xpos[3 * s + k] = d->xpos[3 * b + k] + v[k];
    rcso_quat_mul(q, d->xquat + 4 * b, m->site_quat + 4 * s);
    rcso_quat_to_mat(d->site_xmat + 9 * s, q);
  }
}

/* ------------------------------------------------------------------ mj_comPos [3P] */
void rcso_com_pos(const rcso_model* m, rcso_data* d) {
  int nb = m->nbody;
  double* mass_sub = (double*)calloc((size_t)nb, sizeof(double));
  for (int i = 0; i < nb; i++) {
    for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] = m->body_mass[i] * d->xipos[3 * i + k];
    mass_sub[i] = m->body_mass[i];
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * i + k];
    mass_sub[p] += mass_sub[i];
  }
  for (int i = 0; i < nb; i++) {
    if (mass_sub[i] < MINVAL) copy3(d->subtree_com + 3 * i, d->xipos + 3 * i);
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] /= mass_sub[i];
  }
  free(mass_sub);
  /* body inertias about the COM of their kinematic tree, world axes: [Ixx Iyy Izz Ixy Ixz Iyz, m*d, m] */
  zero(d->cinert, 10);
  for (int i = 1; i < nb; i++) {
    const double* R = d->ximat + 9 * i;
    const double* I = m->body_inertia + 3 * i;
    double mass = m->body_mass[i], dif[3], T[9];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * i + k] - d->subtree_com[3 * m->body_rootid[i] + k];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        T[3 * r + c] = R[3 * r] * I[0] * R[3 * c] + R[3 * r + 1] * I[1] * R[3 * c + 1] + R[3 * r + 2] * I[2] * R[3 * c + 2];
    double dd = dot3(dif, dif);
    double* ci = d->cinert + 10 * i;
    ci[0] = T[0] + mass * (dd - dif[0] * dif[0]);
    ci[1] = T[4] + mass * (dd - dif[1] * dif[1]);
    ci[2] = T[8] + mass * (dd - dif[2] * dif[2]);
    ci[3] = T[1] - mass * dif[0] * dif[1];
    ci[4] = T[2] - mass * dif[0] * dif[2];
    ci[5] = T[5] - mass * dif[1] * dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2];
    ci[9] = mass;
  }
  /* motion axes of every dof about the tree COM: [angular; linear] */
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3 * m->body_rootid[b] + k] - d->xanchor[3 * j + k];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int a = 0; a < 3; a++) {
        double* c = d->cdof + 6 * (da + a);
        zero(c, 6); c[3 + a] = 1;
      }
      for (int a = 0; a < 3; a++) {
        double* c = d->cdof + 6 * (da + 3 + a);
        double ax[3] = {d->xmat[9 * b + a], d->xmat[9 * b + 3 + a], d->xmat[9 * b + 6 + a]};
        copy3(c, ax);
        cross3(c + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_SLIDE) {
      double* c = d->cdof + 6 * da;
      zero(c, 3); copy3(c + 3, d->xaxis + 3 * j);
    } else {
      double* c = d->cdof + 6 * da;
      copy3(c, d->xaxis + 3 * j);
      cross3(c + 3, d->xaxis + 3 * j, off);
    }
  }
}

/* spatial inertia (10-vector about the tree COM) times motion vector -> force vector */
static void mul_inert_vec(double* r, const double* I, const double* v) {
  double w[3] = {v[0], v[1], v[2]}, l[3] = {v[3], v[4], v[5]}, c[3];
  r[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2];
  r[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2];
  r[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2];
  cross3(c, I + 6, l);
  r[0] += c[0]; r[1] += c[1]; r[2] += c[2];
  cross3(c, I + 6, w);
  r[3] = I[9] * l[0] - c[0]; r[4] = I[9] * l[1] - c[1]; r[5] = I[9] * l[2] - c[2];
}
static void cross_motion(double* r, const double* v, const double* s) {
  double a[3], b[3], c[3];
  cross3(a, v, s);
  cross3(b, v, s + 3);
  cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* v, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* ------------------------------------------------------------------ mj_tendon (fixed) [3P] */
void rcso_tendon(const rcso_model* m, rcso_data* d) {
  for (int t = 0; t < m->ntendon; t++) {
    double L = 0;
    for (int dof = 0; dof < m->nv; dof++) {
      double c = m->tendon_coef[t * m->nv + dof];
      d->ten_J[t * m->nv + dof] = c;
      if (c != 0) L += c * d->qpos[m->jnt_qposadr[m->dof_jntid[dof]]];
    }
    d->ten_length[t] = L;
  }
}

/* ------------------------------------------------------------------ mj_crb + mj_factorM [3P] */
void rcso_crb(const rcso_model* m, rcso_data* d) {
  int nv = m->nv, nb = m->nbody;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * (size_t)nb);
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * i + k];
  }
  zero(d->qM, nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mul_inert_vec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->qM[i * nv + j] = s;
      d->qM[j * nv + i] = s;
    }
    d->qM[i * nv + i] += m->dof_armature[i];
  }
  memcpy(d->qLD, d->qM, sizeof(double) * (size_t)(nv * nv));
  rcso_chol_factor(d->qLD, nv);
}
void rcso_mul_M(const rcso_model* m, const rcso_data* d, double* res, const double* v) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += d->qM[i * nv + j] * v[j];
    res[i] = s;
  }
}

/* ------------------------------------------------------------------ mj_transmission [3P] */
void rcso_transmission(const rcso_model* m, rcso_data* d) {
  int nv = m->nv;
  zero(d->actuator_moment, m->nu * nv);
  for (int i = 0; i < m->nu; i++) {
    double gear = m->actuator_gear[i];
    int id = m->actuator_trnid[i];
    if (m->actuator_trntype[i] == TRN_JOINT) {
      d->actuator_length[i] = gear * d->qpos[m->jnt_qposadr[id]];
      d->actuator_moment[i * nv + m->jnt_dofadr[id]] = gear;
    } else { /* tendon */
      d->actuator_length[i] = gear * d->ten_length[id];
      for (int k = 0; k < nv; k++) d->actuator_moment[i * nv + k] = gear * d->ten_J[id * nv + k];
    }
  }
}

/* translational / rotational Jacobian of a world point rigidly attached to `body` (mj_jac) [3P] */
void rcso_jac_point(const rcso_model* m, const rcso_data* d, int body, const double* point, double* jacp,
                    double* jacr) {
  int nv = m->nv;
  if (jacp) zero(jacp, 3 * nv);
  if (jacr) zero(jacr, 3 * nv);
  double off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[3 * m->body_rootid[body] + k];
  while (body > 0 && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (body == 0) return;
  int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  while (i >= 0) {
    const double* c = d->cdof + 6 * i;
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2 * nv + i] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2 * nv + i] = c[5] + t[2];
    }
    i = m->dof_parentid[i];
  }
}

/* ------------------------------------------------------------------ mj_fwdVelocity [3P] */
void rcso_fwd_velocity(const rcso_model* m, rcso_data* d) {
  int nv = m->nv, nb = m->nbody;
  /* actuator velocity */
  for (int i = 0; i < m->nu; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->actuator_moment[i * nv + k] * d->qvel[k];
    d->actuator_velocity[i] = s;
  }
  /* mj_comVel */
  zero(d->cvel, 6);
  for (int i = 1; i < nb; i++) {
    double cvel[6];
    memcpy(cvel, d->cvel + 6 * m->body_parentid[i], sizeof(cvel));
    for (int jj = 0; jj < m->body_jntnum[i]; jj++) {
      int j = m->body_jntadr[i] + jj, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        zero(d->cdof_dot + 6 * da, 18);
        for (int a = 0; a < 3; a++)
          for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6 * (da + a) + k] * d->qvel[da + a];
        for (int a = 3; a < 6; a++) cross_motion(d->cdof_dot + 6 * (da + a), cvel, d->cdof + 6 * (da + a));
        for (int a = 3; a < 6; a++)
          for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6 * (da + a) + k] * d->qvel[da + a];
      } else {
        cross_motion(d->cdof_dot + 6 * da, cvel, d->cdof + 6 * da);
        for (int k = 0; k < 6; k++) cvel[k] += d->cdof[6 * da + k] * d->qvel[da];
      }
    }
    memcpy(d->cvel + 6 * i, cvel, sizeof(cvel));
  }
  /* mj_passive: damping, gravity compensation */
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  zero(d->qfrc_gravcomp, nv);
  {
    double* jacp = (double*)malloc(sizeof(double) * 3 * (size_t)nv);
    for (int b = 1; b < nb; b++) {
      if (m->body_gravcomp[b] == 0 || m->body_mass[b] == 0) continue;
      double f[3];
      for (int k = 0; k < 3; k++) f[k] = -m->gravity[k] * m->body_mass[b] * m->body_gravcomp[b];
      rcso_jac_point(m, d, b, d->xipos + 3 * b, jacp, NULL);
      for (int i = 0; i < nv; i++) d->qfrc_gravcomp[i] += jacp[i] * f[0] + jacp[nv + i] * f[1] + jacp[2 * nv + i] * f[2];
    }
    free(jacp);
  }
  for (int i = 0; i < nv; i++)
    if (!m->jnt_actgravcomp[m->dof_jntid[i]]) d->qfrc_passive[i] += d->qfrc_gravcomp[i];
  /* mj_rne with zero acceleration: Coriolis, centrifugal, gravity */
  zero(d->cacc, 6);
  d->cacc[3] = -m->gravity[0]; d->cacc[4] = -m->gravity[1]; d->cacc[5] = -m->gravity[2];
  zero(d->cfrc, 6);
  for (int i = 1; i < nb; i++) {
    double* cacc = d->cacc + 6 * i;
    memcpy(cacc, d->cacc + 6 * m->body_parentid[i], 6 * sizeof(double));
    for (int k = 0; k < m->body_dofnum[i]; k++) {
      int dof = m->body_dofadr[i] + k;
      for (int c = 0; c < 6; c++) cacc[c] += d->cdof_dot[6 * dof + c] * d->qvel[dof];
    }
    double Ia[6], Iv[6], vxIv[6];
    mul_inert_vec(Ia, d->cinert + 10 * i, cacc);
    mul_inert_vec(Iv, d->cinert + 10 * i, d->cvel + 6 * i);
    cross_force(vxIv, d->cvel + 6 * i, Iv);
    for (int c = 0; c < 6; c++) d->cfrc[6 * i + c] = Ia[c] + vxIv[c];
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int c = 0; c < 6; c++) d->cfrc[6 * p + c] += d->cfrc[6 * i + c];
  }
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[6 * i + c] * d->cfrc[6 * m->dof_bodyid[i] + c];
    d->qfrc_bias[i] = s;
  }
}

/* ------------------------------------------------------------------ mj_fwdActuation [3P] */
void rcso_fwd_actuation(const rcso_model* m, rcso_data* d) {
  int nv = m->nv;
  zero(d->qfrc_actuator, nv);
  for (int i = 0; i < m->nu; i++) {
    double ctrl = d->ctrl[i];
    if (m->actuator_ctrllimited[i]) {
      if (ctrl < m->actuator_ctrlrange[2 * i]) ctrl = m->actuator_ctrlrange[2 * i];
      if (ctrl > m->actuator_ctrlrange[2 * i + 1]) ctrl = m->actuator_ctrlrange[2 * i + 1];
    }
    const double* g = m->actuator_gainprm + 3 * i;
    const double* b = m->actuator_biasprm + 3 * i;
    double force = g[0] * ctrl + b[0] + b[1] * d->actuator_length[i] + b[2] * d->actuator_velocity[i];
    if (m->actuator_forcelimited[i]) {
      if (force < m->actuator_forcerange[2 * i]) force = m->actuator_forcerange[2 * i];
      if (force > m->actuator_forcerange[2 * i + 1]) force = m->actuator_forcerange[2 * i + 1];
    }
    d->actuator_force[i] = force;
    for (int k = 0; k < nv; k++) d->qfrc_actuator[k] += d->actuator_moment[i * nv + k] * force;
  }
  for (int i = 0; i < nv; i++) {
    int j = m->dof_jntid[i];
    if (m->jnt_actgravcomp[j]) d->qfrc_actuator[i] += d->qfrc_gravcomp[i];
    if (m->jnt_actfrclimited[j]) {
      if (d->qfrc_actuator[i] < m->jnt_actfrcrange[2 * j]) d->qfrc_actuator[i] = m->jnt_actfrcrange[2 * j];
      if (d->qfrc_actuator[i] > m->jnt_actfrcrange[2 * j + 1]) d->qfrc_actuator[i] = m->jnt_actfrcrange[2 * j + 1];
    }
  }
}

/* ------------------------------------------------------------------ mj_fwdAcceleration [3P] */
void rcso_fwd_acceleration(const rcso_model* m, rcso_data* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  memcpy(d->qacc_smooth, d->qfrc_smooth, sizeof(double) * (size_t)nv);
  rcso_chol_solve(d->qLD, nv, d->qacc_smooth);
}

/* ------------------------------------------------------------------ integrator [3P]
 * implicitfast: (M - h*D) qacc+ = qfrc_smooth + qfrc_constraint, D = d(passive+actuator)/dqvel
 * (symmetric, no Coriolis term); Euler with implicit joint damping otherwise. Then mj_advance. */
void rcso_integrate(const rcso_model* m, rcso_data* d) {
  int nv = m->nv;
  double h = m->timestep;
  double* A = (double*)malloc(sizeof(double) * (size_t)(nv * nv));
  double* acc = (double*)malloc(sizeof(double) * (size_t)nv);
  zero(d->qDeriv, nv * nv);
  for (int i = 0; i < nv; i++) d->qDeriv[i * nv + i] = -m->dof_damping[i];
  if (m->integrator_implicitfast) {
    for (int a = 0; a < m->nu; a++) {
      double bv = m->actuator_biasprm[3 * a + 2];
      if (bv == 0) continue;
      if (m->actuator_forcelimited[a] && (d->actuator_force[a] <= m->actuator_forcerange[2 * a] ||
                                          d->actuator_force[a] >= m->actuator_forcerange[2 * a + 1]))
        continue;
      for (int i = 0; i < nv; i++) {
        double mi = d->actuator_moment[a * nv + i];
        if (mi == 0) continue;
        for (int j = 0; j < nv; j++) d->qDeriv[i * nv + j] += bv * mi * d->actuator_moment[a * nv + j];
      }
    }
  }
  for (int i = 0; i < nv * nv; i++) A[i] = d->qM[i] - h * d->qDeriv[i];
  for (int i = 0; i < nv; i++) acc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  rcso_chol_factor(A, nv);
  rcso_chol_solve(A, nv, acc);
  /* mj_advance */
  for (int i = 0; i < nv; i++) d->qvel[i] += h * acc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      /* quaternion integration with body-frame angular velocity */
      double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      double ang = norm3(w) * h;
      if (ang > 0) {
        double ax[3] = {w[0], w[1], w[2]}, dq[4], qn[4];
        normalize3(ax);
        rcso_axisangle_quat(dq, ax, ang);
        rcso_quat_mul(qn, d->qpos + qa + 3, dq);
        memcpy(d->qpos + qa + 3, qn, sizeof(qn));
        rcso_quat_normalize(d->qpos + qa + 3);
      }
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h; /* qacc_warmstart was saved by rcso_fwd_constraint, before noslip */
  free(A);
  free(acc);
}

/* ------------------------------------------------------------------ mj_checkPos/Vel/Acc [3P] */
static int bad(const double* v, int n) {
  for (int i = 0; i < n; i++)
    if (isnan(v[i]) || v[i] > 1e10 || v[i] < -1e10) return 1;
  return 0;
}

/* ------------------------------------------------------------------ step halves */
void rcso_step1(const rcso_model* m, rcso_data* d) {
  if (bad(d->qpos, m->nq) || bad(d->qvel, m->nv)) { int w = d->warnings + 1; rcso_reset_data(m, d); d->warnings = w; }
  rcso_kinematics(m, d);
  rcso_com_pos(m, d);
  rcso_tendon(m, d);
  rcso_crb(m, d);
  rcso_collision(m, d);
  rcso_make_constraint(m, d);
  rcso_transmission(m, d);
  rcso_fwd_velocity(m, d);
}
void rcso_step2(const rcso_model* m, rcso_data* d) {
  rcso_fwd_actuation(m, d);
  rcso_fwd_acceleration(m, d);
  rcso_fwd_constraint(m, d);
  if (bad(d->qacc, m->nv)) {
    int w = d->warnings + 1;
    rcso_reset_data(m, d);
    d->warnings = w;
    rcso_forward(m, d);
  }
  rcso_integrate(m, d);
}
void rcso_forward(const rcso_model* m, rcso_data* d) {
  rcso_kinematics(m, d);
  rcso_com_pos(m, d);
  rcso_tendon(m, d);
  rcso_crb(m, d);
  rcso_collision(m, d);
  rcso_make_constraint(m, d);
  rcso_transmission(m, d);
  rcso_fwd_velocity(m, d);
  rcso_fwd_actuation(m, d);
  rcso_fwd_acceleration(m, d);
  rcso_fwd_constraint(m, d);
}
void rcso_step(const rcso_model* m, rcso_data* d) {
  rcso_step1(m, d);
  rcso_step2(m, d);
}
