/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h). Internal layout of the oracle's model/data. */
#ifndef RCS_ORACLE_INTERNAL_H
#define RCS_ORACLE_INTERNAL_H
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "rcs_oracle.h"

#define MINVAL 1e-15
#define MAXCON 64
#define MAXEFC (16 + 3 * MAXCON)

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { TRN_JOINT = 0, TRN_TENDON = 3 };
enum { CNSTR_EQUALITY = 0, CNSTR_FRICTION_DOF = 1, CNSTR_LIMIT_JOINT = 3, CNSTR_CONTACT_FRICTIONLESS = 5,
       CNSTR_CONTACT_PYRAMIDAL = 6, CNSTR_CONTACT_ELLIPTIC = 7 };
enum { STATE_SATISFIED = 0, STATE_QUADRATIC = 1, STATE_LINEARNEG = 2, STATE_LINEARPOS = 3, STATE_CONE = 4 };

struct rcso_model {
  /* sizes (set through "sizes" int field) */
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ntendon, neq, npair, nmeshvert;
  /* options */
  double timestep, gravity[3], impratio, tolerance, noslip_tolerance, ls_tolerance, meaninertia;
  int iterations, ls_iterations, noslip_iterations, cone_elliptic, integrator_implicitfast;
  /* bodies */
  int *body_parentid, *body_rootid, *body_weldid, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr;
  double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_gravcomp,
      *body_invweight0;
  /* joints, dofs */
  int *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited, *jnt_actfrclimited, *jnt_actgravcomp;
  double *jnt_pos, *jnt_axis, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp, *jnt_actfrcrange;
  int *dof_jntid, *dof_bodyid, *dof_parentid;
  double *dof_armature, *dof_damping, *dof_frictionloss, *dof_invweight0, *qpos0;
  /* geoms */
  int *geom_type, *geom_bodyid, *geom_condim, *geom_priority, *geom_vertadr, *geom_vertnum;
  double *geom_size, *geom_pos, *geom_quat, *geom_friction, *geom_solref, *geom_solimp, *geom_solmix,
      *geom_margin, *geom_gap, *geom_rbound, *geom_aabb, *geom_bsphere, *mesh_vert;
  int *mesh_graphadr, *mesh_graph; /* hull edge graph: neighbours of pooled vertex v are mesh_graph[adr[v] .. adr[v+1]), local ids */
  int* pair_geom;
  /* sites */
  int* site_bodyid;
  double *site_pos, *site_quat;
  /* tendons, equalities, actuators */
  double *tendon_coef, *tendon_invweight0;
  int *eq_obj1id, *eq_obj2id, *eq_active0;
  double *eq_polycoef, *eq_solref, *eq_solimp;
  int *actuator_trntype, *actuator_trnid, *actuator_ctrllimited, *actuator_forcelimited;
  double *actuator_gear, *actuator_gainprm, *actuator_biasprm, *actuator_ctrlrange, *actuator_forcerange;
};

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom[2], efc_address;
} rcso_contact;

struct rcso_data {
  const rcso_model* m;
  double time;
  double *qpos, *qvel, *ctrl, *qacc_warmstart, *qacc;
  /* position-dependent */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos, *site_xmat,
      *subtree_com, *cinert, *crb, *cdof, *qM, *qLD, *ten_length, *ten_J, *actuator_length, *actuator_moment;
  /* velocity-dependent */
  double *cvel, *cdof_dot, *cacc, *cfrc, *actuator_velocity, *qfrc_bias, *qfrc_passive, *qfrc_gravcomp;
  /* acceleration */
  double *actuator_force, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qfrc_constraint, *qDeriv;
  /* contacts and constraints */
  int ncon, nefc, ne, nf, nl;
  rcso_contact contact[MAXCON];
  int contact_geom[2 * MAXCON]; /* flat copy for the python view */
  double contact_flat[7 * MAXCON]; /* dist, pos[3], normal[3] per contact, for the python view */
  double* efc_J; /* MAXEFC x nv */
  double efc_pos[MAXEFC], efc_margin[MAXEFC], efc_frictionloss[MAXEFC], efc_D[MAXEFC], efc_R[MAXEFC],
      efc_aref[MAXEFC], efc_vel[MAXEFC], efc_force[MAXEFC], efc_b[MAXEFC], efc_KBIP[4 * MAXEFC],
      efc_diagApprox[MAXEFC];
  int efc_type[MAXEFC], efc_id[MAXEFC], efc_state[MAXEFC];
  int solver_iter, warnings;
};

/* ---- small math (mj_math.c) ---- */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void copy3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void zero(double* r, int n) { memset(r, 0, sizeof(double) * (size_t)n); }
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
/* row-major 3x3 times vector */
static inline void mulmat3(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmatT3(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
void rcso_quat_mul(double* r, const double* a, const double* b);      /* (w,x,y,z) */
void rcso_quat_to_mat(double* M, const double* q);                    /* row-major */
void rcso_quat_normalize(double* q);
void rcso_axisangle_quat(double* q, const double* axis, double angle);
void rcso_rot_vec_quat(double* r, const double* v, const double* q);
void rcso_make_frame(double* frame); /* frame[0:3] given normal -> fills tangents (mju_makeFrame) */
int rcso_chol_factor(double* A, int n);                               /* in place lower; returns rank deficit */
void rcso_chol_solve(const double* L, int n, double* x);              /* x <- A^{-1} x */

/* pipeline pieces */
void rcso_kinematics(const rcso_model* m, rcso_data* d);
void rcso_com_pos(const rcso_model* m, rcso_data* d);
void rcso_tendon(const rcso_model* m, rcso_data* d);
void rcso_crb(const rcso_model* m, rcso_data* d);
void rcso_transmission(const rcso_model* m, rcso_data* d);
void rcso_collision(const rcso_model* m, rcso_data* d);
void rcso_make_constraint(const rcso_model* m, rcso_data* d);
void rcso_fwd_velocity(const rcso_model* m, rcso_data* d);
void rcso_fwd_actuation(const rcso_model* m, rcso_data* d);
void rcso_fwd_acceleration(const rcso_model* m, rcso_data* d);
void rcso_fwd_constraint(const rcso_model* m, rcso_data* d);
void rcso_integrate(const rcso_model* m, rcso_data* d);
void rcso_jac_point(const rcso_model* m, const rcso_data* d, int body, const double* point, double* jacp,
                    double* jacr);
void rcso_mul_M(const rcso_model* m, const rcso_data* d, double* res, const double* v);

/* convex narrowphase (mj_convex.c): returns number of contacts (0/1), fills dist/pos/normal g1->g2 */
int rcso_convex_convex(const rcso_model* m, const rcso_data* d, int g1, int g2, double margin, double* dist,
                       double* pos, double* normal);

#endif
