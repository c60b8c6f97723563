/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h).
 * CPU restatement of MuJoCo 3.2.6 mj_collision [3P] for the geom types of the shipped scenes
 * (plane, box, capsule, convex mesh). The candidate pair list is static (compiled by mjcf.py from
 * contype/conaffinity, same-body, weld and parent filters); per step: bounding-sphere test, then
 * narrowphase. Pair functions: plane-mesh, plane-box, plane-capsule restate mjc_PlaneConvex /
 * mjc_PlaneBox / mjc_PlaneCapsule; all other pairs go through Minkowski Portal Refinement, the
 * algorithm of libccd's ccdMPRPenetration that MuJoCo 3.2.6 calls for convex pairs (mjc_Convex).
 *
 * Multi-point pair functions [3P]: plane-mesh (mjc_PlaneConvex) emits the deepest hull vertex plus neighbours
 * of it in the hull's edge graph that are within the margin and at least 0.3 * rbound away from the points already
 * taken, 3 contacts at most; box-box (mjc_BoxBox) is the separating-axis test followed by clipping of the incident
 * face against the reference face (up to 8 points) or the closest points of the two edges.
 * Stated deviations from MuJoCo (all "parity unpinned"):
 *  - the hull's edge graph comes from scipy's qhull at scene-compile time with neighbours in ascending vertex order
 *    (MuJoCo walks its own qhull graph in qhull's order);
 *  - box-box follows the classic face-clipping construction, not mjc_BoxBox line by line: same contact count for
 *    face-face / edge-edge configurations, point order and sub-mm positions may differ;
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
  const double* R = d->geom_xmat + 9 * g;
  const double* pg = d->geom_xpos + 3 * g;
  double n[3] = {Rp[2], Rp[5], Rp[8]}, nl[3], nd[3] = {-Rp[2], -Rp[5], -Rp[8]};
  const double* verts = m->mesh_vert + 3 * m->geom_vertadr[g];
  int nvert = m->geom_vertnum[g], best = 0;
  double bd = -1e300;
  mulmatT3(nl, R, nd);
  for (int i = 0; i < nvert; i++) {
    double sdot = dot3(verts + 3 * i, nl);
    if (sdot > bd) { bd = sdot; best = i; }
  }
  double taken[3][3];
  int count = 0;
  const double thr = 0.3 * m->geom_rbound[g], thr2 = thr * thr; /* tolplanemesh */
  /* candidates: the support vertex, then its hull-graph neighbours in ascending order */
  int nb0 = m->mesh_graphadr ? m->mesh_graphadr[m->geom_vertadr[g] + best] : 0;
  int nb1 = m->mesh_graphadr ? m->mesh_graphadr[m->geom_vertadr[g] + best + 1] : 0;
  for (int c = -1; c < nb1 - nb0 && count < 3; c++) {
    int vi = c < 0 ? best : m->mesh_graph[nb0 + c];
    double v[3], dif[3], pos[3];
    mulmat3(v, R, verts + 3 * vi);
    for (int k = 0; k < 3; k++) { v[k] += pg[k]; dif[k] = v[k] - d->geom_xpos[3 * gp + k]; }
    double dist = dot3(dif, n);
    if (dist > margin) { if (c < 0) return; continue; }
    int close = 0;
    for (int t = 0; t < count; t++) {
      double e[3] = {v[0] - taken[t][0], v[1] - taken[t][1], v[2] - taken[t][2]};
      if (dot3(e, e) < thr2) close = 1;
    }
    if (close) continue;
    copy3(taken[count], v);
    count++;
    for (int k = 0; k < 3; k++) pos[k] = v[k] - 0.5 * dist * n[k];
    add_contact(m, d, gp, g, dist, pos, n, margin, gap);
  }
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


/* ---- box-box [3P: mjc_BoxBox, restated as separating axes + face clipping] ----
 * Frames: column k of the row-major 3x3 geom_xmat is the box axis k in world coordinates. */
#define BB_EDGE_FUDGE 1.05 /* an edge-edge axis wins only if it is clearly better than the best face axis */
static int box_box(const rcso_model* m, rcso_data* d, int g1, int g2, double margin, double gap) {
  const double *R1 = d->geom_xmat + 9 * g1, *R2 = d->geom_xmat + 9 * g2;
  const double *p1 = d->geom_xpos + 3 * g1, *p2 = d->geom_xpos + 3 * g2;
  const double *a = m->geom_size + 3 * g1, *b = m->geom_size + 3 * g2;
  double Rel[9], Q[9], t[3], dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  mulmatT3(t, R1, dp);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Rel[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
      Q[3 * i + j] = fabs(Rel[3 * i + j]);
    }
  /* face axes: separation s <= margin on every axis or no contact; keep the axis of least penetration */
  double best = -1e300;
  int code = -1, flip = 0;
  for (int i = 0; i < 3; i++) {
    double s = fabs(t[i]) - (a[i] + b[0] * Q[3 * i] + b[1] * Q[3 * i + 1] + b[2] * Q[3 * i + 2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; flip = t[i] < 0; }
  }
  for (int j = 0; j < 3; j++) {
    double tj = t[0] * Rel[j] + t[1] * Rel[3 + j] + t[2] * Rel[6 + j];
    double s = fabs(tj) - (a[0] * Q[j] + a[1] * Q[3 + j] + a[2] * Q[6 + j] + b[j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; flip = tj < 0; }
  }
  /* edge-edge axes a_i x b_j (normalised) */
  double ebest = -1e300, en[3] = {0, 0, 0};
  int ecode = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double l2 = 1 - Rel[3 * i + j] * Rel[3 * i + j];
      if (l2 < 1e-10) continue; /* parallel edges: covered by the face axes */
      double l = sqrt(l2);
      double proj = t[i2] * Rel[3 * i1 + j] - t[i1] * Rel[3 * i2 + j];
      double ra = a[i1] * Q[3 * i2 + j] + a[i2] * Q[3 * i1 + j];
      double rb = b[j1] * Q[3 * i + j2] + b[j2] * Q[3 * i + j1];
      double s = (fabs(proj) - (ra + rb)) / l;
      if (s > margin) return 0;
      if (s > ebest) {
        ebest = s; ecode = 3 * i + j;
        /* axis in the frame of box 1: e_i x (column j of Rel), pointing from box 1 to box 2 */
        double ax[3] = {0, 0, 0};
        ax[i1] = -Rel[3 * i2 + j] / l; ax[i2] = Rel[3 * i1 + j] / l;
        if (proj < 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
        en[0] = ax[0]; en[1] = ax[1]; en[2] = ax[2];
      }
    }
  /* exactly touching boxes (the finger pads at qpos0 after every reset) have best = +-1e-18: not a contact, as for
   * the convex pairs (MuJoCo lists dist = 0 there, which dist >= includemargin keeps out of the constraint set) */
  if (best > margin - 1e-12) return 0;
  if (ecode >= 0 && ebest * BB_EDGE_FUDGE > best) { /* separations are negative depths: the edge must be clearly shallower */
    /* edge-edge: closest points of the two edge lines */
    int i = ecode / 3, j = ecode % 3;
    double nw[3];
    mulmat3(nw, R1, en); /* world normal, from box 1 to box 2 */
    /* point on the edge of box 1: the vertex most along +n (edge direction left free), of box 2 most along -n */
    double pa[3], pb[3], ua[3] = {R1[i], R1[3 + i], R1[6 + i]}, ub[3] = {R2[j], R2[3 + j], R2[6 + j]};
    copy3(pa, p1); copy3(pb, p2);
    for (int k = 0; k < 3; k++) {
      if (k != i) {
        double ak[3] = {R1[k], R1[3 + k], R1[6 + k]};
        double sg = dot3(nw, ak) > 0 ? a[k] : -a[k];
        for (int c = 0; c < 3; c++) pa[c] += sg * ak[c];
      }
      if (k != j) {
        double bk[3] = {R2[k], R2[3 + k], R2[6 + k]};
        double sg = dot3(nw, bk) > 0 ? -b[k] : b[k];
        for (int c = 0; c < 3; c++) pb[c] += sg * bk[c];
      }
    }
    double w[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    double uaub = dot3(ua, ub), q1 = dot3(ua, w), q2 = -dot3(ub, w), den = 1 - uaub * uaub;
    double alpha = 0, beta = 0;
    if (den > 1e-10) { alpha = (q1 + uaub * q2) / den; beta = (uaub * q1 + q2) / den; }
    double pos[3];
    for (int c = 0; c < 3; c++) pos[c] = 0.5 * ((pa[c] + alpha * ua[c]) + (pb[c] + beta * ub[c]));
    add_contact(m, d, g1, g2, ebest, pos, nw, margin, gap);
    return 1;
  }
  /* ---- face contact: reference face on box 1 (code < 3) or box 2 ---- */
  const double *Rr, *Ri, *pr, *pi, *sr, *si;
  int ax = code < 3 ? code : code - 3;
  if (code < 3) { Rr = R1; Ri = R2; pr = p1; pi = p2; sr = a; si = b; }
  else { Rr = R2; Ri = R1; pr = p2; pi = p1; sr = b; si = a; }
  /* outward normal of the reference face, pointing towards the incident box */
  double sgn = (code < 3) ? (flip ? -1.0 : 1.0) : (flip ? 1.0 : -1.0);
  double nr[3] = {sgn * Rr[ax], sgn * Rr[3 + ax], sgn * Rr[6 + ax]};
  /* incident face: the face of the other box whose outward normal is most opposed to nr */
  int iax = 0;
  double md = -1;
  double isg = 1;
  for (int k = 0; k < 3; k++) {
    double dk = nr[0] * Ri[k] + nr[1] * Ri[3 + k] + nr[2] * Ri[6 + k];
    if (fabs(dk) > md) { md = fabs(dk); iax = k; isg = dk > 0 ? -1.0 : 1.0; }
  }
  int u1 = (iax + 1) % 3, u2 = (iax + 2) % 3, r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  /* the 4 vertices of the incident face (world), counter-clockwise in its own (u1, u2) plane */
  double quad[4][3];
  static const double cs[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  for (int v = 0; v < 4; v++)
    for (int c = 0; c < 3; c++)
      quad[v][c] = pi[c] + isg * si[iax] * Ri[3 * c + iax] + cs[v][0] * si[u1] * Ri[3 * c + u1] + cs[v][1] * si[u2] * Ri[3 * c + u2];
  /* to reference-face coordinates: (x, y) along the reference box axes r1, r2; z = height above the reference face */
  double poly[16][3], tmp[16][3];
  int np = 4;
  for (int v = 0; v < 4; v++) {
    double e[3] = {quad[v][0] - pr[0], quad[v][1] - pr[1], quad[v][2] - pr[2]};
    poly[v][0] = e[0] * Rr[r1] + e[1] * Rr[3 + r1] + e[2] * Rr[6 + r1];
    poly[v][1] = e[0] * Rr[r2] + e[1] * Rr[3 + r2] + e[2] * Rr[6 + r2];
    poly[v][2] = dot3(e, nr) - sr[ax];
  }
  /* Sutherland-Hodgman against the four sides of the reference rectangle */
  for (int side = 0; side < 4; side++) {
    int c = side >> 1;
    double sg = (side & 1) ? -1.0 : 1.0, lim = c == 0 ? sr[r1] : sr[r2];
    int nq = 0;
    for (int v = 0; v < np; v++) {
      const double *A = poly[v], *B = poly[(v + 1) % np];
      double da = lim - sg * A[c], db = lim - sg * B[c];
      if (da >= 0) { copy3(tmp[nq], A); nq++; }
      if ((da >= 0) != (db >= 0)) {
        double f = da / (da - db);
        for (int k = 0; k < 3; k++) tmp[nq][k] = A[k] + f * (B[k] - A[k]);
        nq++;
      }
    }
    np = nq;
    for (int v = 0; v < np; v++) copy3(poly[v], tmp[v]);
    if (np == 0) break;
  }
  /* contacts: clipped points at or below the reference face (within the margin), 8 at most */
  double nw[3] = {nr[0], nr[1], nr[2]};
  if (code >= 3) { nw[0] = -nw[0]; nw[1] = -nw[1]; nw[2] = -nw[2]; } /* contact normal runs from geom 1 to geom 2 */
  int cnt = 0;
  for (int v = 0; v < np && cnt < 8; v++) {
    double dist = poly[v][2];
    if (dist > margin) continue;
    double pos[3];
    for (int c = 0; c < 3; c++)
      pos[c] = pr[c] + poly[v][0] * Rr[3 * c + r1] + poly[v][1] * Rr[3 * c + r2] + (sr[ax] + 0.5 * dist) * nr[c];
    add_contact(m, d, g1, g2, dist, pos, nw, margin, gap);
    cnt++;
  }
  return cnt;
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
    if (t1 == GEOM_BOX && t2 == GEOM_BOX) { box_box(m, d, g1, g2, margin, gap); continue; }
    double dist, pos[3], normal[3];
    if (rcso_convex_convex(m, d, g1, g2, margin, &dist, pos, normal)) add_contact(m, d, g1, g2, dist, pos, normal, margin, gap);
  }
}
