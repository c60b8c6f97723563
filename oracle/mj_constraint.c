/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h).
 * CPU restatement of MuJoCo 3.2.6 constraint construction and solution [3P]:
 * mj_makeConstraint (equality -> friction loss -> joint limits -> contacts), mj_makeImpedance
 * (solref/solimp -> K, B, impedance, R, D, aref; elliptic-cone R scaling with impratio),
 * mj_fwdConstraint (warm start choice, Newton solver on the primal problem with exact line
 * search, elliptic cones) and the noslip post-pass (mj_solNoSlip).
 *
 * The problem is strictly convex, so the optimum is unique; this Newton implementation follows
 * MuJoCo's termination rule (improvement or gradient below `tolerance`, scaled by
 * 1/(meaninertia*max(1,nv))) but not its exact iterate sequence. */
#include "oracle_internal.h"

/* ------------------------------------------------------------------ impedance */
static void get_impedance(const double* solimp, double pos, double margin, double* imp, double* impP) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin < 0.0001) dmin = 0.0001; if (dmin > 0.9999) dmin = 0.9999;
  if (dmax < 0.0001) dmax = 0.0001; if (dmax > 0.9999) dmax = 0.9999;
  if (width < MINVAL) width = MINVAL;
  if (mid < 0.0001) mid = 0.0001; if (mid > 0.9999) mid = 0.9999;
  if (power < 1) power = 1;
  double x = fabs(pos - margin) / width, y, yP;
  if (x >= 1) { *imp = dmax; *impP = 0; return; }
  if (x <= 0) { *imp = dmin; *impP = 0; return; }
  if (power == 1) { y = x; yP = 1; }
  else if (x <= mid) { double a = 1 / pow(mid, power - 1); y = a * pow(x, power); yP = power * a * pow(x, power - 1); }
  else { double b = 1 / pow(1 - mid, power - 1); y = 1 - b * pow(1 - x, power); yP = power * b * pow(1 - x, power - 1); }
  *imp = dmin + y * (dmax - dmin);
  *impP = yP * (dmax - dmin) * ((pos - margin) > 0 ? 1 : -1) / width;
}

static int add_row(rcso_data* d, int nv, const double* jac, double pos, double margin, double floss, int type, int id) {
  if (d->nefc >= MAXEFC) { d->warnings++; return -1; }
  int r = d->nefc++;
  memcpy(d->efc_J + (size_t)r * nv, jac, sizeof(double) * (size_t)nv);
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_frictionloss[r] = floss;
  d->efc_type[r] = type; d->efc_id[r] = id;
  return r;
}

void rcso_make_constraint(const rcso_model* m, rcso_data* d) {
  int nv = m->nv;
  double* jac = (double*)calloc((size_t)(3 * nv) * 3, sizeof(double));
  d->nefc = d->ne = d->nf = d->nl = 0;
  /* ---- equality: joint coupling  q1 - q1_0 = poly(q2 - q2_0) ---- */
  for (int e = 0; e < m->neq; e++) {
    if (!m->eq_active0[e]) continue;
    int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
    const double* c = m->eq_polycoef + 5 * e;
    zero(jac, nv);
    double pos;
    double p1 = d->qpos[m->jnt_qposadr[j1]] - m->qpos0[m->jnt_qposadr[j1]];
    if (j2 >= 0) {
      double p2 = d->qpos[m->jnt_qposadr[j2]] - m->qpos0[m->jnt_qposadr[j2]];
      double poly = c[0] + p2 * (c[1] + p2 * (c[2] + p2 * (c[3] + p2 * c[4])));
      double dpoly = c[1] + p2 * (2 * c[2] + p2 * (3 * c[3] + p2 * 4 * c[4]));
      pos = p1 - poly;
      jac[m->jnt_dofadr[j1]] = 1;
      jac[m->jnt_dofadr[j2]] = -dpoly;
    } else {
      pos = p1 - c[0];
      jac[m->jnt_dofadr[j1]] = 1;
    }
    add_row(d, nv, jac, pos, 0, 0, CNSTR_EQUALITY, e);
    d->ne++;
  }
  /* ---- friction loss ---- */
  for (int i = 0; i < nv; i++) {
    if (m->dof_frictionloss[i] <= 0) continue;
    zero(jac, nv);
    jac[i] = 1;
    add_row(d, nv, jac, 0, 0, m->dof_frictionloss[i], CNSTR_FRICTION_DOF, i);
    d->nf++;
  }
  /* ---- joint limits (hinge / slide) ---- */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
    double q = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - q);
      if (dist < margin) {
        zero(jac, nv);
        jac[m->jnt_dofadr[j]] = -side;
        add_row(d, nv, jac, dist, margin, 0, CNSTR_LIMIT_JOINT, j);
        d->nl++;
      }
    }
  }
  /* ---- contacts ---- */
  for (int ci = 0; ci < d->ncon; ci++) {
    rcso_contact* c = &d->contact[ci];
    c->efc_address = -1;
    if (c->dist >= c->includemargin) continue;
    int b1 = m->geom_bodyid[c->geom[0]], b2 = m->geom_bodyid[c->geom[1]];
    double* jp1 = jac; double* jp2 = jac + 3 * nv; double* jd = jac + 6 * nv;
    rcso_jac_point(m, d, b1, c->pos, jp1, NULL);
    rcso_jac_point(m, d, b2, c->pos, jp2, NULL);
    /* rows of the contact frame times (J2 - J1) */
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < nv; k++)
        jd[r * nv + k] = c->frame[3 * r] * (jp2[k] - jp1[k]) + c->frame[3 * r + 1] * (jp2[nv + k] - jp1[nv + k]) +
                         c->frame[3 * r + 2] * (jp2[2 * nv + k] - jp1[2 * nv + k]);
    if (c->dim == 1) {
      c->efc_address = add_row(d, nv, jd, c->dist, c->includemargin, 0, CNSTR_CONTACT_FRICTIONLESS, ci);
    } else if (m->cone_elliptic) {
      c->efc_address = add_row(d, nv, jd, c->dist, c->includemargin, 0, CNSTR_CONTACT_ELLIPTIC, ci);
      for (int r = 1; r < 3 && r < c->dim; r++) add_row(d, nv, jd + r * nv, 0, 0, 0, CNSTR_CONTACT_ELLIPTIC, ci);
    } else { /* pyramidal, condim 3: J_n +- mu_k J_tk */
      double* row = (double*)malloc(sizeof(double) * (size_t)nv);
      for (int r = 1; r < 3 && r < c->dim; r++)
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int k = 0; k < nv; k++) row[k] = jd[k] + sgn * c->friction[r - 1] * jd[r * nv + k];
          int a = add_row(d, nv, row, c->dist, c->includemargin, 0, CNSTR_CONTACT_PYRAMIDAL, ci);
          if (c->efc_address < 0) c->efc_address = a;
        }
      free(row);
    }
  }
  free(jac);

  /* ---- mj_makeImpedance: diagApprox, KBIP, R, D, then reference acceleration ---- */
  double h = m->timestep;
  for (int i = 0; i < d->nefc; i++) {
    int id = d->efc_id[i], type = d->efc_type[i];
    const double *solref, *solimp;
    double diag;
    int first_of_contact = 1;
    if (type == CNSTR_EQUALITY) {
      solref = m->eq_solref + 2 * id; solimp = m->eq_solimp + 5 * id;
      diag = m->dof_invweight0[m->jnt_dofadr[m->eq_obj1id[id]]];
      if (m->eq_obj2id[id] >= 0) diag += m->dof_invweight0[m->jnt_dofadr[m->eq_obj2id[id]]];
    } else if (type == CNSTR_FRICTION_DOF) {
      static const double sr[2] = {0.02, 1}, si[5] = {0.9, 0.95, 0.001, 0.5, 2}; /* dof_solref/solimp defaults */
      solref = sr; solimp = si;
      diag = m->dof_invweight0[id];
    } else if (type == CNSTR_LIMIT_JOINT) {
      solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id;
      diag = m->dof_invweight0[m->jnt_dofadr[id]];
    } else {
      rcso_contact* c = &d->contact[id];
      solref = c->solref; solimp = c->solimp;
      int b1 = m->geom_bodyid[c->geom[0]], b2 = m->geom_bodyid[c->geom[1]];
      diag = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
      first_of_contact = (i == c->efc_address);
      if (type == CNSTR_CONTACT_PYRAMIDAL) diag *= 1 + c->friction[0] * c->friction[0]; /* approx. as MuJoCo */
    }
    if (diag < MINVAL) diag = MINVAL;
    d->efc_diagApprox[i] = diag;
    double imp, impP;
    get_impedance(solimp, d->efc_pos[i], d->efc_margin[i], &imp, &impP);
    double dmax = solimp[1];
    if (dmax < 0.0001) dmax = 0.0001; if (dmax > 0.9999) dmax = 0.9999;
    double K, B;
    if (solref[0] > 0) {
      double tc = solref[0], dr = solref[1];
      if (tc < 2 * h) tc = 2 * h; /* refsafe */
      K = 1 / (dmax * dmax * tc * tc * dr * dr);
      B = 2 / (dmax * tc);
    } else {
      K = -solref[0] / (dmax * dmax);
      B = -solref[1] / dmax;
    }
    if (type == CNSTR_FRICTION_DOF || (type == CNSTR_CONTACT_ELLIPTIC && !first_of_contact)) K = 0;
    d->efc_KBIP[4 * i] = K; d->efc_KBIP[4 * i + 1] = B; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = impP;
    double R = (1 - imp) / imp * diag;
    if (R < MINVAL) R = MINVAL;
    d->efc_R[i] = R;
  }
  /* elliptic cones: friction-row regularisation tied to the normal row through impratio */
  for (int ci = 0; ci < d->ncon; ci++) {
    rcso_contact* c = &d->contact[ci];
    int a = c->efc_address;
    if (a < 0 || d->efc_type[a] != CNSTR_CONTACT_ELLIPTIC || c->dim < 3) continue;
    double ir = m->impratio < MINVAL ? MINVAL : m->impratio;
    d->efc_R[a + 1] = d->efc_R[a] / ir;
    c->mu = c->friction[0] * sqrt(d->efc_R[a + 1] / d->efc_R[a]);
    for (int j = 1; j < c->dim - 1; j++)
      d->efc_R[a + 1 + j] = d->efc_R[a + 1] * c->friction[0] * c->friction[0] / (c->friction[j] * c->friction[j]);
  }
  /* pyramidal cones [3P] (mj_makeImpedance): every edge of the pyramid gets Rpy = 2 mu^2 R_first, with the
   * regularised friction mu = friction[0] / sqrt(impratio) */
  for (int ci = 0; ci < d->ncon; ci++) {
    rcso_contact* c = &d->contact[ci];
    int a = c->efc_address;
    if (a < 0 || d->efc_type[a] != CNSTR_CONTACT_PYRAMIDAL || c->dim < 3) continue;
    double ir = m->impratio < MINVAL ? MINVAL : m->impratio;
    c->mu = c->friction[0] / sqrt(ir);
    double Rpy = 2 * c->mu * c->mu * d->efc_R[a];
    if (Rpy < MINVAL) Rpy = MINVAL;
    for (int j = 0; j < 2 * (c->dim - 1); j++) d->efc_R[a + j] = Rpy;
  }
  for (int i = 0; i < d->nefc; i++) {
    d->efc_D[i] = 1 / d->efc_R[i];
    double vel = 0;
    for (int k = 0; k < nv; k++) vel += d->efc_J[(size_t)i * nv + k] * d->qvel[k];
    d->efc_vel[i] = vel;
    d->efc_aref[i] = -d->efc_KBIP[4 * i + 1] * vel - d->efc_KBIP[4 * i] * d->efc_KBIP[4 * i + 2] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}

/* ------------------------------------------------------------------ constraint cost / force (mj_constraintUpdate)
 * jar = J*qacc - aref. Fills force, state; returns the constraint cost s(jar). If H (nefc-indexed
 * list of dense cone blocks) is requested, cone_hess[ci] gets the 3x3 Hessian wrt jar of contact ci. */
static double constraint_update(const rcso_model* m, rcso_data* d, const double* jar, double* force, int* state,
                                double (*cone_hess)[9]) {
  double cost = 0;
  int i = 0;
  while (i < d->nefc) {
    int type = d->efc_type[i];
    double D = d->efc_D[i];
    if (type == CNSTR_EQUALITY) {
      force[i] = -D * jar[i]; state[i] = STATE_QUADRATIC; cost += 0.5 * D * jar[i] * jar[i]; i++;
    } else if (type == CNSTR_FRICTION_DOF) {
      double f = d->efc_frictionloss[i], R = d->efc_R[i];
      if (jar[i] <= -R * f) { force[i] = f; state[i] = STATE_LINEARNEG; cost += -0.5 * R * f * f - f * jar[i]; }
      else if (jar[i] >= R * f) { force[i] = -f; state[i] = STATE_LINEARPOS; cost += -0.5 * R * f * f + f * jar[i]; }
      else { force[i] = -D * jar[i]; state[i] = STATE_QUADRATIC; cost += 0.5 * D * jar[i] * jar[i]; }
      i++;
    } else if (type == CNSTR_LIMIT_JOINT || type == CNSTR_CONTACT_FRICTIONLESS || type == CNSTR_CONTACT_PYRAMIDAL) {
      if (jar[i] < 0) { force[i] = -D * jar[i]; state[i] = STATE_QUADRATIC; cost += 0.5 * D * jar[i] * jar[i]; }
      else { force[i] = 0; state[i] = STATE_SATISFIED; }
      i++;
    } else { /* elliptic contact, rows i .. i+dim-1 */
      rcso_contact* c = &d->contact[d->efc_id[i]];
      int dim = c->dim;
      double mu = c->mu, U[3], N, T2 = 0, T;
      U[0] = jar[i] * mu;
      for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * c->friction[j - 1]; T2 += U[j] * U[j]; }
      N = U[0]; T = sqrt(T2);
      if (cone_hess) zero(cone_hess[d->efc_id[i]], 9);
      if (N >= mu * T || (T <= 0 && N >= 0)) { /* top zone: inside the dual cone */
        for (int j = 0; j < dim; j++) { force[i + j] = 0; state[i + j] = STATE_SATISFIED; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) { /* bottom zone: full quadratic */
        for (int j = 0; j < dim; j++) {
          force[i + j] = -d->efc_D[i + j] * jar[i + j];
          state[i + j] = STATE_QUADRATIC;
          cost += 0.5 * d->efc_D[i + j] * jar[i + j] * jar[i + j];
        }
      } else { /* middle zone: distance to the cone surface */
        double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        force[i] = -Dm * NmT * mu;
        for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * c->friction[j - 1];
        for (int j = 0; j < dim; j++) state[i + j] = STATE_CONE;
        if (cone_hess) {
          /* Hessian of 0.5*Dm*(N - mu*T)^2 wrt jar: Dm*(g g^T + NmT * d2(N - mu T)) */
          double g[3] = {mu, 0, 0};
          for (int j = 1; j < dim; j++) g[j] = -mu * c->friction[j - 1] * U[j] / T;
          double* H = cone_hess[d->efc_id[i]];
          for (int a = 0; a < dim; a++)
            for (int b = 0; b < dim; b++) H[3 * a + b] = Dm * g[a] * g[b];
          for (int a = 1; a < dim; a++)
            for (int b = 1; b < dim; b++) {
              double fa = c->friction[a - 1], fb = c->friction[b - 1];
              double d2T = fa * fb * ((a == b ? 1.0 : 0.0) / T - U[a] * U[b] / (T * T * T));
              H[3 * a + b] += Dm * NmT * (-mu) * d2T;
            }
        }
      }
      i += dim;
    }
  }
  return cost;
}

/* ------------------------------------------------------------------ Newton solver (mj_solNewton) */
typedef struct {
  int nv, nefc;
  double *Ma, *jar, *grad, *Mgrad, *search, *Mv, *Jv, *H, *force_tmp;
  double cost, gauss;
} ctx_t;

static double total_cost(const rcso_model* m, rcso_data* d, const double* qacc, double* Ma, double* jar, double* force,
                         int* state, double (*cone_hess)[9], double* gauss_out) {
  int nv = m->nv;
  rcso_mul_M(m, d, Ma, qacc);
  for (int i = 0; i < d->nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * qacc[k];
    jar[i] = s - d->efc_aref[i];
  }
  double cost = constraint_update(m, d, jar, force, state, cone_hess);
  double g = 0;
  for (int k = 0; k < nv; k++) g += 0.5 * (Ma[k] - d->qfrc_smooth[k]) * (qacc[k] - d->qacc_smooth[k]);
  if (gauss_out) *gauss_out = g;
  return cost + g;
}

/* phi(alpha) derivatives along search: value, first and second derivative */
static void line_eval(const rcso_model* m, rcso_data* d, const double* jar, const double* Jv, double alpha, double quadGauss0,
                      double quadGauss1, double quadGauss2, double* val, double* d1, double* d2) {
  double v = alpha * alpha * quadGauss2 + alpha * quadGauss1 + quadGauss0;
  double g1 = 2 * alpha * quadGauss2 + quadGauss1, g2 = 2 * quadGauss2;
  int i = 0;
  while (i < d->nefc) {
    int type = d->efc_type[i];
    double D = d->efc_D[i];
    double x = jar[i] + alpha * Jv[i];
    if (type == CNSTR_EQUALITY) {
      v += 0.5 * D * x * x; g1 += D * x * Jv[i]; g2 += D * Jv[i] * Jv[i]; i++;
    } else if (type == CNSTR_FRICTION_DOF) {
      double f = d->efc_frictionloss[i], R = d->efc_R[i];
      if (x <= -R * f) { v += -0.5 * R * f * f - f * x; g1 += -f * Jv[i]; }
      else if (x >= R * f) { v += -0.5 * R * f * f + f * x; g1 += f * Jv[i]; }
      else { v += 0.5 * D * x * x; g1 += D * x * Jv[i]; g2 += D * Jv[i] * Jv[i]; }
      i++;
    } else if (type != CNSTR_CONTACT_ELLIPTIC) {
      if (x < 0) { v += 0.5 * D * x * x; g1 += D * x * Jv[i]; g2 += D * Jv[i] * Jv[i]; }
      i++;
    } else {
      rcso_contact* c = &d->contact[d->efc_id[i]];
      int dim = c->dim;
      double mu = c->mu, U[3], V[3], T2 = 0;
      U[0] = x * mu; V[0] = Jv[i] * mu;
      for (int j = 1; j < dim; j++) {
        U[j] = (jar[i + j] + alpha * Jv[i + j]) * c->friction[j - 1];
        V[j] = Jv[i + j] * c->friction[j - 1];
        T2 += U[j] * U[j];
      }
      double N = U[0], T = sqrt(T2);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          double xj = jar[i + j] + alpha * Jv[i + j], Dj = d->efc_D[i + j];
          v += 0.5 * Dj * xj * xj; g1 += Dj * xj * Jv[i + j]; g2 += Dj * Jv[i + j] * Jv[i + j];
        }
      } else {
        double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        double UV = 0, VV = 0;
        for (int j = 1; j < dim; j++) { UV += U[j] * V[j]; VV += V[j] * V[j]; }
        double T1 = UV / T;                           /* dT/dalpha */
        double T2d = VV / T - UV * UV / (T * T * T);  /* d2T/dalpha2 */
        double N1 = V[0];
        v += 0.5 * Dm * NmT * NmT;
        g1 += Dm * NmT * (N1 - mu * T1);
        g2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2d));
      }
      i += dim;
    }
  }
  *val = v; *d1 = g1; *d2 = g2;
}

/* exact line search on the convex 1-D restriction: safeguarded Newton with bracketing */
static double line_search(const rcso_model* m, rcso_data* d, const double* jar, const double* Jv, double qG0, double qG1,
                          double qG2, double gtol, int maxiter) {
  double v0, d10, d20, v, d1, d2;
  line_eval(m, d, jar, Jv, 0, qG0, qG1, qG2, &v0, &d10, &d20);
  if (d10 >= 0 || d20 <= 0) return 0; /* not a descent direction */
  double lo = 0, hi = -1, dlo = d10, alpha = -d10 / d20;
  (void)dlo;
  for (int it = 0; it < maxiter; it++) {
    line_eval(m, d, jar, Jv, alpha, qG0, qG1, qG2, &v, &d1, &d2);
    if (fabs(d1) < gtol) break;
    if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; }
    double next = (d2 > 0) ? alpha - d1 / d2 : -1;
    if (hi > 0) {
      if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
      if (hi - lo < 1e-15 * (1 + fabs(hi))) { alpha = next; break; }
    } else if (!(next > lo)) {
      next = 2 * alpha + 1e-12;
    }
    alpha = next;
  }
  line_eval(m, d, jar, Jv, alpha, qG0, qG1, qG2, &v, &d1, &d2);
  if (v > v0) return 0;
  return alpha;
}

static void solve_newton(const rcso_model* m, rcso_data* d) {
  int nv = m->nv, nefc = d->nefc;
  double* buf = (double*)calloc((size_t)(8 * nv + 3 * nefc + nv * nv), sizeof(double));
  double *Ma = buf, *grad = Ma + nv, *Mgrad = grad + nv, *search = Mgrad + nv, *Mv = search + nv, *tmp = Mv + nv,
         *jar = tmp + 3 * nv, *Jv = jar + nefc, *force = Jv + nefc, *H = force + nefc;
  static double cone_hess[MAXCON][9];
  double(*ch)[9] = (double(*)[9])malloc(sizeof(double) * 9 * MAXCON);
  (void)cone_hess;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double gauss;
  double cost = total_cost(m, d, d->qacc, Ma, jar, d->efc_force, d->efc_state, ch, &gauss);
  int iter = 0;
  for (; iter < m->iterations;) {
    /* gradient and Hessian at the current point */
    for (int k = 0; k < nv; k++) {
      double s = 0;
      for (int i = 0; i < nefc; i++) s += d->efc_J[(size_t)i * nv + k] * d->efc_force[i];
      d->qfrc_constraint[k] = s;
      grad[k] = Ma[k] - d->qfrc_smooth[k] - s;
    }
    memcpy(H, d->qM, sizeof(double) * (size_t)(nv * nv));
    for (int i = 0; i < nefc;) {
      int st = d->efc_state[i];
      if (st == STATE_QUADRATIC) {
        const double* J = d->efc_J + (size_t)i * nv;
        double D = d->efc_D[i];
        for (int a = 0; a < nv; a++) {
          if (J[a] == 0) continue;
          for (int b = 0; b < nv; b++) H[a * nv + b] += D * J[a] * J[b];
        }
        i++;
      } else if (st == STATE_CONE) {
        rcso_contact* c = &d->contact[d->efc_id[i]];
        const double* Hc = ch[d->efc_id[i]];
        for (int r = 0; r < c->dim; r++)
          for (int s = 0; s < c->dim; s++) {
            double w = Hc[3 * r + s];
            if (w == 0) continue;
            const double *Jr = d->efc_J + (size_t)(i + r) * nv, *Js = d->efc_J + (size_t)(i + s) * nv;
            for (int a = 0; a < nv; a++) {
              if (Jr[a] == 0) continue;
              for (int b = 0; b < nv; b++) H[a * nv + b] += w * Jr[a] * Js[b];
            }
          }
        i += c->dim;
      } else {
        i++;
      }
    }
    rcso_chol_factor(H, nv);
    memcpy(Mgrad, grad, sizeof(double) * (size_t)nv);
    rcso_chol_solve(H, nv, Mgrad);
    for (int k = 0; k < nv; k++) search[k] = -Mgrad[k];
    /* line search */
    rcso_mul_M(m, d, Mv, search);
    for (int i = 0; i < nefc; i++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * search[k];
      Jv[i] = s;
    }
    double qG0 = gauss, qG1 = 0, qG2 = 0, snorm = 0;
    for (int k = 0; k < nv; k++) {
      qG1 += search[k] * (Ma[k] - d->qfrc_smooth[k]);
      qG2 += 0.5 * search[k] * Mv[k];
      snorm += search[k] * search[k];
    }
    snorm = sqrt(snorm);
    if (snorm < MINVAL) break;
    double gtol = m->tolerance * m->ls_tolerance * snorm / scale;
    double alpha = line_search(m, d, jar, Jv, qG0, qG1, qG2, gtol, m->ls_iterations);
    if (alpha == 0) break;
    for (int k = 0; k < nv; k++) d->qacc[k] += alpha * search[k];
    double oldcost = cost;
    cost = total_cost(m, d, d->qacc, Ma, jar, d->efc_force, d->efc_state, ch, &gauss);
    iter++;
    double gnorm = 0;
    for (int k = 0; k < nv; k++) {
      double s = 0;
      for (int i = 0; i < nefc; i++) s += d->efc_J[(size_t)i * nv + k] * d->efc_force[i];
      double g = Ma[k] - d->qfrc_smooth[k] - s;
      gnorm += g * g;
    }
    double improvement = scale * (oldcost - cost), gradient = scale * sqrt(gnorm);
    if (improvement < m->tolerance || gradient < m->tolerance) break;
  }
  d->solver_iter = iter;
  for (int k = 0; k < nv; k++) {
    double s = 0;
    for (int i = 0; i < nefc; i++) s += d->efc_J[(size_t)i * nv + k] * d->efc_force[i];
    d->qfrc_constraint[k] = s;
  }
  free(ch);
  free(buf);
}

/* ------------------------------------------------------------------ noslip (mj_solNoSlip) [3P]
 * Dual Gauss-Seidel sweeps over friction-loss rows and the friction dimensions of contacts with the
 * regulariser R removed and normal forces held fixed; elliptic friction is projected onto the
 * ellipse sum (f_j/mu_j)^2 <= f_n^2 by a Newton search on the KKT multiplier (mju_QCQP2). */
static void qcqp2(const double* A, const double* b, const double* dd, double r, double* res) {
  /* minimise 0.5 x'Ax + x'b  s.t.  sum (x_i/d_i)^2 <= r^2 */
  double As[4] = {A[0] * dd[0] * dd[0], A[1] * dd[0] * dd[1], A[2] * dd[1] * dd[0], A[3] * dd[1] * dd[1]};
  double bs[2] = {b[0] * dd[0], b[1] * dd[1]};
  double la = 0, v0 = 0, v1 = 0;
  for (int it = 0; it < 20; it++) {
    double a00 = As[0] + la, a11 = As[3] + la, det = a00 * a11 - As[1] * As[2];
    if (det < 1e-10) { res[0] = res[1] = 0; return; }
    double P00 = a11 / det, P01 = -As[1] / det, P10 = -As[2] / det, P11 = a00 / det;
    v0 = -P00 * bs[0] - P01 * bs[1];
    v1 = -P10 * bs[0] - P11 * bs[1];
    double val = v0 * v0 + v1 * v1 - r * r;
    if (val < 1e-10) break;
    double pv0 = P00 * v0 + P01 * v1, pv1 = P10 * v0 + P11 * v1;
    double deriv = -2 * (v0 * pv0 + v1 * pv1);
    double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  res[0] = v0 * dd[0]; res[1] = v1 * dd[1];
}

static void solve_noslip(const rcso_model* m, rcso_data* d) {
  int nv = m->nv, nefc = d->nefc;
  int any = 0;
  for (int i = 0; i < nefc; i++)
    if (d->efc_type[i] == CNSTR_FRICTION_DOF || d->efc_type[i] == CNSTR_CONTACT_ELLIPTIC || d->efc_type[i] == CNSTR_CONTACT_PYRAMIDAL) any = 1;
  if (!any) return;
  /* A = J M^-1 J^T (unregularised), b = J qacc_smooth - aref */
  double* MinvJT = (double*)malloc(sizeof(double) * (size_t)nefc * nv);
  double* A = (double*)malloc(sizeof(double) * (size_t)nefc * nefc);
  for (int i = 0; i < nefc; i++) {
    memcpy(MinvJT + (size_t)i * nv, d->efc_J + (size_t)i * nv, sizeof(double) * (size_t)nv);
    rcso_chol_solve(d->qLD, nv, MinvJT + (size_t)i * nv);
  }
  for (int i = 0; i < nefc; i++)
    for (int j = 0; j < nefc; j++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * MinvJT[(size_t)j * nv + k];
      A[(size_t)i * nefc + j] = s;
    }
  for (int i = 0; i < nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qacc_smooth[k];
    d->efc_b[i] = s - d->efc_aref[i];
  }
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double* f = d->efc_force;
  for (int iter = 0; iter < m->noslip_iterations; iter++) {
    double improvement = 0;
    for (int i = d->ne; i < d->ne + d->nf; i++) { /* dry friction rows */
      double res = d->efc_b[i];
      for (int j = 0; j < nefc; j++) res += A[(size_t)i * nefc + j] * f[j];
      double old = f[i], Aii = A[(size_t)i * nefc + i];
      if (Aii < MINVAL) continue;
      f[i] -= res / Aii;
      double fl = d->efc_frictionloss[i];
      if (f[i] < -fl) f[i] = -fl; else if (f[i] > fl) f[i] = fl;
      double df = f[i] - old;
      improvement -= 0.5 * df * df * Aii + df * res;
    }
    for (int ci = 0; ci < d->ncon; ci++) {
      rcso_contact* c = &d->contact[ci];
      int a = c->efc_address;
      if (a < 0 || c->dim < 3) continue;
      if (d->efc_type[a] == CNSTR_CONTACT_ELLIPTIC) {
        int dim = c->dim;
        double fn = f[a], res[2], old[2] = {f[a + 1], f[a + 2]}, Ac[4], bc[2];
        for (int j = 0; j < 2; j++) {
          res[j] = d->efc_b[a + 1 + j];
          for (int k = 0; k < nefc; k++) res[j] += A[(size_t)(a + 1 + j) * nefc + k] * f[k];
        }
        (void)dim;
        for (int r = 0; r < 2; r++)
          for (int s = 0; s < 2; s++) Ac[2 * r + s] = A[(size_t)(a + 1 + r) * nefc + a + 1 + s];
        if (fn < MINVAL) { f[a + 1] = f[a + 2] = 0; }
        else {
          /* bc = res - Ac*old */
          for (int r = 0; r < 2; r++) bc[r] = res[r] - Ac[2 * r] * old[0] - Ac[2 * r + 1] * old[1];
          double det = Ac[0] * Ac[3] - Ac[1] * Ac[2];
          double v[2] = {0, 0};
          if (det > 1e-10) {
            v[0] = -(Ac[3] * bc[0] - Ac[1] * bc[1]) / det;
            v[1] = -(-Ac[2] * bc[0] + Ac[0] * bc[1]) / det;
          }
          double e = v[0] * v[0] / (c->friction[0] * c->friction[0]) + v[1] * v[1] / (c->friction[1] * c->friction[1]);
          if (det <= 1e-10 || e > fn * fn) qcqp2(Ac, bc, c->friction, fn, v);
          f[a + 1] = v[0]; f[a + 2] = v[1];
        }
        double df[2] = {f[a + 1] - old[0], f[a + 2] - old[1]};
        improvement -= 0.5 * (df[0] * (Ac[0] * df[0] + Ac[1] * df[1]) + df[1] * (Ac[2] * df[0] + Ac[3] * df[1])) + df[0] * res[0] + df[1] * res[1];
      }
      /* pyramidal noslip (xArm7 scene has no noslip iterations) is not restated */
    }
    if (improvement * scale < m->noslip_tolerance) break;
  }
  for (int k = 0; k < nv; k++) {
    double s = 0;
    for (int i = 0; i < nefc; i++) s += d->efc_J[(size_t)i * nv + k] * f[i];
    d->qfrc_constraint[k] = s;
  }
  memcpy(d->qacc, d->qfrc_constraint, sizeof(double) * (size_t)nv);
  rcso_chol_solve(d->qLD, nv, d->qacc);
  for (int k = 0; k < nv; k++) d->qacc[k] += d->qacc_smooth[k];
  free(MinvJT);
  free(A);
}

/* ------------------------------------------------------------------ mj_fwdConstraint */
void rcso_fwd_constraint(const rcso_model* m, rcso_data* d) {
  int nv = m->nv, nefc = d->nefc;
  if (nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * (size_t)nv);
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(double) * (size_t)nv);
    zero(d->qfrc_constraint, nv);
    d->solver_iter = 0;
    return;
  }
  /* warm start: keep qacc_warmstart only if its total cost beats qacc_smooth's */
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(nv + 2 * nefc));
  int* st = (int*)malloc(sizeof(int) * (size_t)nefc);
  double cost_warm = total_cost(m, d, d->qacc_warmstart, tmp, tmp + nv, tmp + nv + nefc, st, NULL, NULL);
  double cost_smooth = total_cost(m, d, d->qacc_smooth, tmp, tmp + nv, tmp + nv + nefc, st, NULL, NULL);
  memcpy(d->qacc, cost_warm < cost_smooth ? d->qacc_warmstart : d->qacc_smooth, sizeof(double) * (size_t)nv);
  free(tmp);
  free(st);
  solve_newton(m, d);
  /* [3P] mj_fwdConstraint saves the warm start from the main solver's result, BEFORE the noslip post-pass */
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * (size_t)nv);
  if (m->noslip_iterations > 0) solve_noslip(m, d);
}
