/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h).
 * CPU restatement of MuJoCo 3.2.6 mj_collision [3P] for the geom types of the shipped scenes
 * (plane, box, capsule, convex mesh). The candidate pair list is static (compiled by mjcf.py from
 * contype/conaffinity, same-body, weld and parent filters); per step: bounding-sphere test, then
 * narrowphase. Pair functions: plane-mesh, plane-box, plane-capsule restate mjc_PlaneConvex /
 * mjc_PlaneBox / mjc_PlaneCapsule; all other pairs go through Minkowski Portal Refinement, the
 * algorithm of libccd's ccdMPRPenetration that MuJoCo 3.2.6 calls for convex pairs (mjc_Convex).
 *
 * Stated deviations from MuJoCo (all "parity unpinned"):
 *  - plane-mesh emits the single deepest hull vertex (MuJoCo adds up to 3 more support points);
 *  - box-box uses MPR (one contact) instead of mjc_BoxBox's up-to-8-point clipping;
 *  - contacts are emitted in pair-list order (g1<g2 lexicographic), geom[0] = lower geom type. */
#include "oracle_internal.h"

#define MPR_TOL 1e-6
#define MPR_ITER 50

/* ---- support mapping in world coordinates ---- */
static void support(const rcso_model* m, const rcso_data* d, int g, const double* dir, double* out) {
  const double* R = d->geom_xmat + 9 * g;
  const double* p = d->geom_xpos + 3 * g;
  double dl[3], v[3] = {0, 0, 0};
  mulmatT3(dl, R, dir);
  int type = m->geom_type[g];
  const double* size = m->geom_size + 3 * g;
  if (type == GEOM_MESH) {
    const double* verts = m->mesh_vert + 3 * m->geom_vertadr[g];
    int n = m->geom_vertnum[g], best = 0;
    double bd = -1e300;
    for (int i = 0; i < n; i++) {
      double s = dot3(verts + 3 * i, dl);
      if (s > bd) { bd = s; best = i; }
    }
    copy3(v, verts + 3 * best);
  } else if (type == GEOM_BOX) {
    for (int k = 0; k < 3; k++) v[k] = dl[k] >= 0 ? size[k] : -size[k];
  } else if (type == GEOM_CAPSULE) {
    double n = norm3(dl);
    if (n > MINVAL) for (int k = 0; k < 3; k++) v[k] = size[0] * dl[k] / n;
    v[2] += dl[2] >= 0 ? size[1] : -size[1];
  } else if (type == GEOM_SPHERE) {
    double n = norm3(dl);
    if (n > MINVAL) for (int k = 0; k < 3; k++) v[k] = size[0] * dl[k] / n;
  } else if (type == GEOM_CYLINDER) { /* [3P] mjc_support: rim point in the xy direction, cap by the sign of z */
    double n = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (n > MINVAL) { v[0] = dl[0] / n * size[0]; v[1] = dl[1] / n * size[0]; }
    v[2] = dl[2] > 0 ? size[1] : (dl[2] < 0 ? -size[1] : 0);
  }
  mulmat3(out, R, v);
  out[0] += p[0]; out[1] += p[1]; out[2] += p[2];
}

typedef struct { double v[3], v1[3], v2[3]; } sup_t; /* Minkowski-difference point and its witnesses */
static void mink_support(const rcso_model* m, const rcso_data* d, int g1, int g2, const double* dir, sup_t* s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  support(m, d, g1, dir, s->v1);
  support(m, d, g2, nd, s->v2);
  for (int k = 0; k < 3; k++) s->v[k] = s->v1[k] - s->v2[k];
}
static double tri_dist2(const double* P, const double* A, const double* B, const double* C, double* witness) {
  /* squared distance from P to triangle ABC, closest point in witness */
  double ab[3], ac[3], ap[3];
  for (int k = 0; k < 3; k++) { ab[k] = B[k] - A[k]; ac[k] = C[k] - A[k]; ap[k] = P[k] - A[k]; }
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  double w[3];
  if (d1 <= 0 && d2 <= 0) { copy3(w, A); goto done; }
  double bp[3];
  for (int k = 0; k < 3; k++) bp[k] = P[k] - B[k];
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { copy3(w, B); goto done; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double t = d1 / (d1 - d3); for (int k = 0; k < 3; k++) w[k] = A[k] + t * ab[k]; goto done; }
  double cp[3];
  for (int k = 0; k < 3; k++) cp[k] = P[k] - C[k];
  double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { copy3(w, C); goto done; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double t = d2 / (d2 - d6); for (int k = 0; k < 3; k++) w[k] = A[k] + t * ac[k]; goto done; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) w[k] = B[k] + t * (C[k] - B[k]);
    goto done;
  }
  {
    double den = 1.0 / (va + vb + vc), v = vb * den, u = vc * den;
    for (int k = 0; k < 3; k++) w[k] = A[k] + ab[k] * v + ac[k] * u;
  }
done:;
  double dd[3] = {P[0] - w[0], P[1] - w[1], P[2] - w[2]};
  if (witness) copy3(witness, w);
  return dot3(dd, dd);
}
static void portal_dir(const sup_t* s, double* dir) {
  double a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = s[2].v[k] - s[1].v[k]; b[k] = s[3].v[k] - s[1].v[k]; }
  cross3(dir, a, b);
  normalize3(dir);
}
static int portal_reach_tol(const sup_t* s, const sup_t* v4, const double* dir) {
  double dv1 = dot3(s[1].v, dir), dv2 = dot3(s[2].v, dir), dv3 = dot3(s[3].v, dir), dv4 = dot3(v4->v, dir);
  double dot1 = dv4 - dv1, dot2 = dv4 - dv2, dot3_ = dv4 - dv3;
  double mn = dot1 < dot2 ? dot1 : dot2;
  mn = mn < dot3_ ? mn : dot3_;
  return mn <= MPR_TOL;
}
static void expand_portal(sup_t* s, const sup_t* v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, s[0].v);
  double dt = dot3(s[1].v, v4v0);
  if (dt > 0) {
    dt = dot3(s[2].v, v4v0);
    if (dt > 0) s[1] = *v4; else s[3] = *v4;
  } else {
    dt = dot3(s[3].v, v4v0);
    if (dt > 0) s[2] = *v4; else s[1] = *v4;
  }
}
static void find_pos(const sup_t* s, double* pos) {
  double dir[3], b[4], vec[3], sum;
  portal_dir(s, dir);
  cross3(vec, s[1].v, s[2].v); b[0] = dot3(vec, s[3].v);
  cross3(vec, s[3].v, s[2].v); b[1] = dot3(vec, s[0].v);
  cross3(vec, s[0].v, s[1].v); b[2] = dot3(vec, s[3].v);
  cross3(vec, s[2].v, s[1].v); b[3] = dot3(vec, s[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0) {
    b[0] = 0;
    cross3(vec, s[2].v, s[3].v); b[1] = dot3(vec, dir);
    cross3(vec, s[3].v, s[1].v); b[2] = dot3(vec, dir);
    cross3(vec, s[1].v, s[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 3; k++) { p1[k] += b[i] * s[i].v1[k]; p2[k] += b[i] * s[i].v2[k]; }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p1[k] + p2[k]) / sum;
}

/* returns 1 when penetrating: depth > 0, dir (unit, from g1 into g2), pos */
static int mpr_penetration(const rcso_model* m, const rcso_data* d, int g1, int g2, double* depth, double* dir_out,
                           double* pos) {
  sup_t s[4], v4;
  double dir[3], va[3], vb[3];
  /* ---- discover portal ---- */
  for (int k = 0; k < 3; k++) {
    s[0].v1[k] = d->geom_xpos[3 * g1 + k];
    s[0].v2[k] = d->geom_xpos[3 * g2 + k];
    s[0].v[k] = s[0].v1[k] - s[0].v2[k];
  }
  if (fabs(s[0].v[0]) < 1e-12 && fabs(s[0].v[1]) < 1e-12 && fabs(s[0].v[2]) < 1e-12) s[0].v[0] += 1e-5;
  for (int k = 0; k < 3; k++) dir[k] = -s[0].v[k];
  normalize3(dir);
  mink_support(m, d, g1, g2, dir, &s[1]);
  if (dot3(s[1].v, dir) <= 0) return 0;
  cross3(dir, s[0].v, s[1].v);
  if (dot3(dir, dir) < 1e-24) {
    /* origin on the v0-v1 ray: penetration along that segment */
    *depth = norm3(s[1].v);
    for (int k = 0; k < 3; k++) { dir_out[k] = s[1].v[k]; pos[k] = 0.5 * (s[1].v1[k] + s[1].v2[k]); }
    normalize3(dir_out);
    return 1;
  }
  normalize3(dir);
  mink_support(m, d, g1, g2, dir, &s[2]);
  if (dot3(s[2].v, dir) <= 0) return 0;
  for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
  cross3(dir, va, vb);
  normalize3(dir);
  if (dot3(dir, s[0].v) > 0) {
    sup_t t = s[1]; s[1] = s[2]; s[2] = t;
    for (int k = 0; k < 3; k++) dir[k] = -dir[k];
  }
  for (int it = 0;; it++) {
    if (it > 100) return 0;
    mink_support(m, d, g1, g2, dir, &s[3]);
    if (dot3(s[3].v, dir) <= 0) return 0;
    int cont = 0;
    cross3(va, s[1].v, s[3].v);
    if (dot3(va, s[0].v) < -1e-18) { s[2] = s[3]; cont = 1; }
    if (!cont) {
      cross3(va, s[3].v, s[2].v);
      if (dot3(va, s[0].v) < -1e-18) { s[1] = s[3]; cont = 1; }
    }
    if (!cont) break;
    for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
    cross3(dir, va, vb);
    normalize3(dir);
  }
  /* ---- refine portal ---- */
  for (int it = 0;; it++) {
    portal_dir(s, dir);
    if (dot3(s[1].v, dir) >= 0) break; /* portal encapsulates the origin: shapes intersect */
    mink_support(m, d, g1, g2, dir, &v4);
    if (dot3(v4.v, dir) < 0 || portal_reach_tol(s, &v4, dir) || it >= MPR_ITER) return 0;
    expand_portal(s, &v4);
  }
  /* ---- find penetration ---- */
  for (int it = 0;; it++) {
    portal_dir(s, dir);
    mink_support(m, d, g1, g2, dir, &v4);
    if (portal_reach_tol(s, &v4, dir) || it >= MPR_ITER) {
      double origin[3] = {0, 0, 0}, w[3];
      *depth = sqrt(tri_dist2(origin, s[1].v, s[2].v, s[3].v, w));
      if (*depth < 1e-12) { copy3(dir_out, dir); } else { copy3(dir_out, w); normalize3(dir_out); }
      find_pos(s, pos);
      return 1;
    }
    expand_portal(s, &v4);
  }
}

int rcso_convex_convex(const rcso_model* m, const rcso_data* d, int g1, int g2, double margin, double* dist,
                       double* pos, double* normal) {
  (void)margin; /* margin is 0 in all shipped scenes: only penetrating pairs produce contacts */
  double depth, dir[3];
  if (!mpr_penetration(m, d, g1, g2, &depth, dir, pos)) return 0;
  /* exactly touching pairs (the finger pads at qpos0 after every reset) give depth = +-1e-17 with an
   * arbitrary direction; MuJoCo's analytic box-box reports dist = 0 there, which is excluded from the
   * constraint set (dist >= includemargin). Dropping depth < 1e-12 gives the same constraint set. */
  if (depth < 1e-12) return 0;
  *dist = -depth;
  /* the MPR direction runs from the interior point cA-cB through the origin, i.e. from g1 towards g2:
   * it is the contact normal (translating g2 by depth*dir separates the pair) */
  for (int k = 0; k < 3; k++) normal[k] = dir[k];
  return 1;
}

/* ---- contact parameter mixing (mj_contactParam) [3P] ---- */
static void add_contact(const rcso_model* m, rcso_data* d, int g1, int g2, double dist, const double* pos,
                        const double* normal, double margin, double gap) {
  if (d->ncon >= MAXCON) { d->warnings++; return; }
  rcso_contact* c = &d->contact[d->ncon++];
  c->dist = dist;
  copy3(c->pos, pos);
  copy3(c->frame, normal);
  rcso_make_frame(c->frame);
  c->includemargin = margin - gap;
  c->geom[0] = g1; c->geom[1] = g2;
  c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  double mix;
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  const double *f1 = m->geom_friction + 3 * g1, *f2 = m->geom_friction + 3 * g2;
  double fr[3];
  if (p1 == p2) {
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else mix = s1 < MINVAL ? 0.0 : 1.0;
    for (int k = 0; k < 3; k++) fr[k] = f1[k] > f2[k] ? f1[k] : f2[k];
    c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  } else {
    int g = p1 > p2 ? g1 : g2;
    mix = p1 > p2 ? 1.0 : 0.0;
    for (int k = 0; k < 3; k++) fr[k] = m->geom_friction[3 * g + k];
    c->dim = m->geom_condim[g];
  }
  const double *r1 = m->geom_solref + 2 * g1, *r2 = m->geom_solref + 2 * g2;
  if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) c->solref[k] = mix * r1[k] + (1 - mix) * r2[k];
  else for (int k = 0; k < 2; k++) c->solref[k] = r1[k] < r2[k] ? r1[k] : r2[k];
  for (int k = 0; k < 5; k++) c->solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
  c->friction[0] = fr[0]; c->friction[1] = fr[0]; c->friction[2] = fr[1]; c->friction[3] = fr[2]; c->friction[4] = fr[2];
  for (int k = 0; k < 5; k++) if (c->friction[k] < 1e-5) c->friction[k] = 1e-5; /* mjMINMU */
  c->mu = 0;
  c->efc_address = -1;
}

/* ---- plane pair functions ---- */
static void plane_mesh(const rcso_model* m, rcso_data* d, int gp, int g, double margin, double gap) {
  const double* Rp = d->geom_xmat + 9 * gp;
  double n[3] = {Rp[2], Rp[5], Rp[8]}, nd[3] = {-Rp[2], -Rp[5], -Rp[8]}, v[3], dif[3], pos[3];
  support(m, d, g, nd, v);
  for (int k = 0; k < 3; k++) dif[k] = v[k] - d->geom_xpos[3 * gp + k];
  double dist = dot3(dif, n);
  if (dist > margin) return;
  for (int k = 0; k < 3; k++) pos[k] = v[k] - 0.5 * dist * n[k];
  add_contact(m, d, gp, g, dist, pos, n, margin, gap);
}
static void plane_box(const rcso_model* m, rcso_data* d, int gp, int g, double margin, double gap) {
  const double* Rp = d->geom_xmat + 9 * gp;
  const double* R = d->geom_xmat + 9 * g;
  const double* size = m->geom_size + 3 * g;
  double n[3] = {Rp[2], Rp[5], Rp[8]};
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double loc[3] = {(i & 1 ? size[0] : -size[0]), (i & 2 ? size[1] : -size[1]), (i & 4 ? size[2] : -size[2])};
    double c[3], dif[3], pos[3];
    mulmat3(c, R, loc);
    for (int k = 0; k < 3; k++) { c[k] += d->geom_xpos[3 * g + k]; dif[k] = c[k] - d->geom_xpos[3 * gp + k]; }
    double dist = dot3(dif, n);
    if (dist > margin) continue;
    for (int k = 0; k < 3; k++) pos[k] = c[k] - 0.5 * dist * n[k];
    add_contact(m, d, gp, g, dist, pos, n, margin, gap);
    cnt++;
  }
}
static void plane_capsule(const rcso_model* m, rcso_data* d, int gp, int g, double margin, double gap) {
  const double* Rp = d->geom_xmat + 9 * gp;
  const double* R = d->geom_xmat + 9 * g;
  double n[3] = {Rp[2], Rp[5], Rp[8]};
  double r = m->geom_size[3 * g], hl = m->geom_size[3 * g + 1];
  double ax[3] = {R[2] * hl, R[5] * hl, R[8] * hl};
  for (int s = 0; s < 2; s++) {
    double c[3], dif[3], pos[3];
    for (int k = 0; k < 3; k++) {
      c[k] = d->geom_xpos[3 * g + k] + (s == 0 ? ax[k] : -ax[k]);
      dif[k] = c[k] - d->geom_xpos[3 * gp + k];
    }
    double dist = dot3(dif, n) - r;
    if (dist > margin) continue;
    for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * (r + 0.5 * dist);
    add_contact(m, d, gp, g, dist, pos, n, margin, gap);
  }
}

/* world centre of the geom's bounding volume (local AABB centre) */
static void bv_center(const rcso_model* m, const rcso_data* d, int g, double* c) {
  mulmat3(c, d->geom_xmat + 9 * g, m->geom_aabb + 6 * g);
  c[0] += d->geom_xpos[3 * g]; c[1] += d->geom_xpos[3 * g + 1]; c[2] += d->geom_xpos[3 * g + 2];
}
/* oriented-box separating-axis test on the geoms' local AABBs (mid-phase; exact for box-box). 1 = separated */
static int obb_separated(const rcso_model* m, const rcso_data* d, int g1, int g2, double margin) {
  const double *A = d->geom_xmat + 9 * g1, *B = d->geom_xmat + 9 * g2;
  const double *ha = m->geom_aabb + 6 * g1 + 3, *hb = m->geom_aabb + 6 * g2 + 3;
  double ca[3], cb[3], R[9], AR[9], t[3], dv[3];
  bv_center(m, d, g1, ca); bv_center(m, d, g2, cb);
  for (int k = 0; k < 3; k++) dv[k] = cb[k] - ca[k];
  mulmatT3(t, A, dv);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
      AR[3 * i + j] = fabs(R[3 * i + j]) + 1e-12;
    }
  for (int i = 0; i < 3; i++)
    if (fabs(t[i]) > ha[i] + hb[0] * AR[3 * i] + hb[1] * AR[3 * i + 1] + hb[2] * AR[3 * i + 2] + margin) return 1;
  for (int j = 0; j < 3; j++)
    if (fabs(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) > ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j] + hb[j] + margin) return 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j];
      double rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
      if (fabs(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb + margin) return 1;
    }
  return 0;
}

void rcso_collision(const rcso_model* m, rcso_data* d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom[2 * p], g2 = m->pair_geom[2 * p + 1];
    if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
    double margin = m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2];
    double gap = m->geom_gap[g1] > m->geom_gap[g2] ? m->geom_gap[g1] : m->geom_gap[g2];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    if (t1 == GEOM_PLANE) {
      const double* Rp = d->geom_xmat + 9 * g1;
      double n[3] = {Rp[2], Rp[5], Rp[8]}, dif[3];
      double c2[3];
      bv_center(m, d, g2, c2);
      for (int k = 0; k < 3; k++) dif[k] = c2[k] - d->geom_xpos[3 * g1 + k];
      if (dot3(dif, n) > m->geom_bsphere[4 * g2 + 3] + margin) continue;
      { /* oriented box against the plane */
        const double* B = d->geom_xmat + 9 * g2; const double* hb = m->geom_aabb + 6 * g2 + 3;
        double r = 0;
        for (int j = 0; j < 3; j++) r += hb[j] * fabs(n[0] * B[j] + n[1] * B[3 + j] + n[2] * B[6 + j]);
        if (dot3(dif, n) - r > margin) continue;
      }
      if (t2 == GEOM_MESH) plane_mesh(m, d, g1, g2, margin, gap);
      else if (t2 == GEOM_BOX) plane_box(m, d, g1, g2, margin, gap);
      else if (t2 == GEOM_CAPSULE) plane_capsule(m, d, g1, g2, margin, gap);
      continue;
    }
    double dif[3];
    double c1[3], c2[3];
    bv_center(m, d, g1, c1); bv_center(m, d, g2, c2);
    for (int k = 0; k < 3; k++) dif[k] = c2[k] - c1[k];
    double bound = m->geom_bsphere[4 * g1 + 3] + m->geom_bsphere[4 * g2 + 3] + margin;
    if (dot3(dif, dif) > bound * bound) continue;
    if (obb_separated(m, d, g1, g2, margin)) continue;
    double dist, pos[3], normal[3];
    if (rcso_convex_convex(m, d, g1, g2, margin, &dist, pos, normal)) add_contact(m, d, g1, g2, dist, pos, normal, margin, gap);
  }
}
