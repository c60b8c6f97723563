/* TEST INFRASTRUCTURE ONLY (see rcs_oracle.h).
 * CPU restatement of RCS's own glue around the physics step:
 *   Pose            /root/reference/src/rcs/Pose.cpp, include/rcs/Pose.h:23-65 (Eigen semantics restated)
 *   Pin IK          /root/reference/src/rcs/Kinematics.cpp:28-81, include/rcs/Kinematics.h:32-35
 *   Sim             /root/reference/src/sim/sim.cpp (callback clocks, step, step_until_convergence, reset)
 *   SimRobot        /root/reference/src/sim/SimRobot.cpp
 *   SimGripper      /root/reference/src/sim/SimGripper.cpp (quirks kept: SURVEY.md appendix B 11-13)
 *   env step/reset  /root/reference/python/rcs/envs/base.py:246-304,469-488,684-735, envs/sim.py:49-76,120-131
 */
#include <pthread.h>
#include <stdio.h>
#include <time.h>

#include "oracle_internal.h"

/* ================================================================== Pose (xyz + quat xyzw) */
typedef struct { double x, y, z, w; } quat_t;
static quat_t q_from(const double* p7) { quat_t q = {p7[3], p7[4], p7[5], p7[6]}; return q; }
static void q_store(double* p7, quat_t q) { p7[3] = q.x; p7[4] = q.y; p7[5] = q.z; p7[6] = q.w; }
static quat_t q_normalized(quat_t q) { /* Eigen QuaternionBase::normalize */
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  if (n > 0) { q.x /= n; q.y /= n; q.z /= n; q.w /= n; }
  return q;
}
static quat_t q_mul(quat_t a, quat_t b) {
  quat_t r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
static quat_t q_conj(quat_t a) { quat_t r = {-a.x, -a.y, -a.z, a.w}; return r; }
static void q_rot(quat_t q, const double* v, double* out) { /* Eigen: v + w*uv + qv x uv, uv = 2 qv x v */
  double qv[3] = {q.x, q.y, q.z}, uv[3], t[3];
  cross3(uv, qv, v);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(t, qv, uv);
  for (int k = 0; k < 3; k++) out[k] = v[k] + q.w * uv[k] + t[k];
}
static void q_to_mat(quat_t q, double* R) { /* Eigen toRotationMatrix, row-major out */
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
         tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static quat_t q_from_mat(const double* m) { /* Eigen Quaternion(Matrix3) ; m row-major */
  quat_t q;
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double qq[3];
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[3 * k + j] - m[3 * j + k]) * t;
    qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q.x = qq[0]; q.y = qq[1]; q.z = qq[2];
  }
  return q;
}
static double q_angular_distance(quat_t a, quat_t b) { /* Eigen angularDistance */
  quat_t d = q_mul(a, q_conj(b));
  double vn = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  return 2 * atan2(vn, fabs(d.w));
}
static quat_t q_slerp(quat_t a, double t, quat_t b) { /* Eigen slerp */
  const double one = 1.0 - 2.220446049250313e-16;
  double dd = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w, ad = fabs(dd), s0, s1;
  if (ad >= one) { s0 = 1 - t; s1 = t; }
  else {
    double th = acos(ad), st = sin(th);
    s0 = sin((1 - t) * th) / st;
    s1 = sin(t * th) / st;
  }
  if (dd < 0) s1 = -s1;
  quat_t r = {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
  return r;
}
static quat_t q_from_rpy(const double* rpy) { /* Rz(yaw) * Ry(pitch) * Rx(roll), Pose.h:37-43 */
  quat_t qz = {0, 0, sin(0.5 * rpy[2]), cos(0.5 * rpy[2])};
  quat_t qy = {0, sin(0.5 * rpy[1]), 0, cos(0.5 * rpy[1])};
  quat_t qx = {sin(0.5 * rpy[0]), 0, 0, cos(0.5 * rpy[0])};
  return q_mul(q_mul(qz, qy), qx);
}
void rcso_pose_mul(const double* a, const double* b, double* out) { /* Pose.cpp:173-178 */
  quat_t qa = q_from(a), qb = q_from(b);
  double t[3];
  q_rot(qa, b, t);
  double o[7] = {t[0] + a[0], t[1] + a[1], t[2] + a[2]};
  q_store(o, q_normalized(q_mul(qa, qb)));
  memcpy(out, o, sizeof(o));
}
void rcso_pose_inverse(const double* a, double* out) { /* Pose.cpp:203-206 */
  quat_t c = q_conj(q_from(a));
  double t[3], o[7];
  q_rot(c, a, t);
  o[0] = -t[0]; o[1] = -t[1]; o[2] = -t[2];
  q_store(o, q_normalized(c));
  memcpy(out, o, sizeof(o));
}
void rcso_pose_from_rpy(const double* xyz, const double* rpy, double* out) {
  copy3(out, xyz);
  q_store(out, q_normalized(q_from_rpy(rpy)));
}
void rcso_pose_from_matrix(const double* R, const double* xyz, double* out) {
  copy3(out, xyz);
  q_store(out, q_normalized(q_from_mat(R)));
}
void rcso_pose_rotation_m(const double* a, double* R) { q_to_mat(q_from(a), R); }
void rcso_pose_xyzrpy(const double* a, double* out6) { /* Pose.cpp:133-138,155-160: eulerAngles(2,1,0) (Eigen 3.4) */
  double m[9];
  q_to_mat(q_from(a), m);
  /* i=2, j=1, k=0, odd */
  double r0 = atan2(m[3 * 1 + 0], m[0]);
  double c2 = sqrt(m[8] * m[8] + m[7] * m[7]);
  double r1;
  if (r0 < 0) {
    r0 += M_PI;
    r1 = atan2(-m[6], -c2);
  } else {
    r1 = atan2(-m[6], c2);
  }
  double s1 = sin(r0), c1 = cos(r0);
  double r2 = atan2(s1 * m[2] - c1 * m[5], c1 * m[4] - s1 * m[1]);
  copy3(out6, a);
  out6[3] = r2; out6[4] = r1; out6[5] = r0; /* roll, pitch, yaw */
}
double rcso_pose_total_angle(const double* a) {
  quat_t id = {0, 0, 0, 1};
  return q_angular_distance(q_from(a), id);
}
void rcso_pose_limit_rotation_angle(const double* a, double max_angle, double* out) { /* Pose.cpp:184-192 */
  double cur = rcso_pose_total_angle(a);
  memmove(out, a, 7 * sizeof(double));
  if (cur > max_angle && max_angle >= 0) {
    quat_t id = {0, 0, 0, 1};
    q_store(out, q_normalized(q_slerp(id, max_angle / cur, q_from(a))));
  }
}
void rcso_pose_limit_translation_length(const double* a, double max_len, double* out) { /* Pose.cpp:193-201 */
  double n = norm3(a);
  memmove(out, a, 7 * sizeof(double));
  if (n > max_len && max_len >= 0) {
    for (int k = 0; k < 3; k++) out[k] = a[k] / n * max_len;
    q_store(out, q_normalized(q_from(a)));
  }
}
void rcso_pose_interpolate(const double* a, const double* b, double progress, double* out) { /* Pose.cpp:140-153 */
  if (progress > 1) progress = 1;
  double o[7];
  for (int k = 0; k < 3; k++) o[k] = a[k] + (b[k] - a[k]) * progress;
  q_store(o, q_normalized(q_slerp(q_from(a), progress, q_from(b))));
  memcpy(out, o, sizeof(o));
}
int rcso_pose_is_close(const double* a, const double* b, double eps_r, double eps_t) { /* Pose.cpp:208-211 */
  double l1 = fabs(a[0] - b[0]) + fabs(a[1] - b[1]) + fabs(a[2] - b[2]);
  return l1 < eps_t && q_angular_distance(q_from(a), q_from(b)) < eps_r;
}

/* ================================================================== Pin IK / FK
 * Kinematics of the site frame for joint vector q (first nq_model qpos entries), in the model's
 * world frame (== robot base for the shipped robot.xml). Returns R (row-major), p and the 6 x nv
 * LOCAL-frame Jacobian [linear; angular] (Pinocchio computeFrameJacobian default) [3P]. */
static void site_fk(const rcso_model* m, int site, int nq_model, const double* q, double* R, double* p, double* J) {
  int chain[64], n = 0;
  for (int b = m->site_bodyid[site]; b > 0; b = m->body_parentid[b]) chain[n++] = b;
  double xpos[3] = {0, 0, 0}, xquat[4] = {1, 0, 0, 0}, Rm[9];
  double anchors[64][3], axes[64][3];
  int jtype[64], jdof[64], nj = 0;
  for (int c = n - 1; c >= 0; c--) {
    int b = chain[c];
    double v[3], qn[4];
    rcso_quat_to_mat(Rm, xquat);
    mulmat3(v, Rm, m->body_pos + 3 * b);
    for (int k = 0; k < 3; k++) xpos[k] += v[k];
    rcso_quat_mul(qn, xquat, m->body_quat + 4 * b);
    memcpy(xquat, qn, sizeof(qn));
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj;
      if (m->jnt_type[j] == JNT_FREE) continue;
      int qa = m->jnt_qposadr[j];
      double qj = (qa < nq_model ? q[qa] : 0.0) - m->qpos0[qa];
      rcso_quat_to_mat(Rm, xquat);
      mulmat3(axes[nj], Rm, m->jnt_axis + 3 * j);
      mulmat3(v, Rm, m->jnt_pos + 3 * j);
      for (int k = 0; k < 3; k++) anchors[nj][k] = xpos[k] + v[k];
      jtype[nj] = m->jnt_type[j];
      jdof[nj] = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_SLIDE) {
        for (int k = 0; k < 3; k++) xpos[k] += axes[nj][k] * qj;
      } else {
        double ql[4];
        rcso_axisangle_quat(ql, m->jnt_axis + 3 * j, qj);
        rcso_quat_mul(qn, xquat, ql);
        memcpy(xquat, qn, sizeof(qn));
        rcso_rot_vec_quat(v, m->jnt_pos + 3 * j, xquat);
        for (int k = 0; k < 3; k++) xpos[k] = anchors[nj][k] - v[k];
      }
      nj++;
    }
    rcso_quat_normalize(xquat);
  }
  double v[3], qs[4];
  rcso_quat_to_mat(Rm, xquat);
  mulmat3(v, Rm, m->site_pos + 3 * site);
  for (int k = 0; k < 3; k++) p[k] = xpos[k] + v[k];
  rcso_quat_mul(qs, xquat, m->site_quat + 4 * site);
  rcso_quat_to_mat(R, qs);
  if (J) {
    zero(J, 6 * nq_model);
    for (int a = 0; a < nj; a++) {
      if (jdof[a] >= nq_model) continue;
      double lin[3], ang[3] = {0, 0, 0}, r[3];
      if (jtype[a] == JNT_HINGE) {
        for (int k = 0; k < 3; k++) r[k] = p[k] - anchors[a][k];
        cross3(lin, axes[a], r);
        copy3(ang, axes[a]);
      } else {
        copy3(lin, axes[a]);
      }
      double l[3], w[3];
      mulmatT3(l, R, lin);
      mulmatT3(w, R, ang);
      for (int k = 0; k < 3; k++) { J[k * nq_model + jdof[a]] = l[k]; J[(3 + k) * nq_model + jdof[a]] = w[k]; }
    }
  }
}
static void log3(const double* R, double* w, double* theta) { /* pinocchio::log3 [3P] */
  double tr = R[0] + R[4] + R[8];
  double ct = 0.5 * (tr - 1);
  if (ct > 1) ct = 1; if (ct < -1) ct = -1;
  double t = acos(ct);
  *theta = t;
  /* Pinocchio 3.7 log3: near pi the antisymmetric part vanishes, so the axis comes from the diagonal,
   * w_k^2 = theta^2 (R_kk - cos theta) / (1 - cos theta), signed by the antisymmetric part (threshold pi - 1e-2);
   * below eps^(1/4) the factor theta / sin(theta) is taken as 1 */
  if (t >= M_PI - 1e-2) {
    double beta = t * t / (1 - ct);
    double a[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    for (int k = 0; k < 3; k++) {
      double v = (R[4 * k] - ct) * beta;
      w[k] = (a[k] > 0 ? 1.0 : -1.0) * (v > 0 ? sqrt(v) : 0.0);
    }
    return;
  }
  double f = 0.5 * (t > 1.220703125e-4 ? t / sin(t) : 1.0);
  w[0] = f * (R[7] - R[5]); w[1] = f * (R[2] - R[6]); w[2] = f * (R[3] - R[1]);
}
void rcso_log3(const double* R9_rowmajor, double* w3) { double t; log3(R9_rowmajor, w3, &t); } /* test hook */
static void log6(const double* R, const double* p, double* out /* [v; w] */) { /* pinocchio::log6 [3P] */
  double w[3], t;
  log3(R, w, &t);
  double alpha, beta, t2 = t * t;
  if (t < 1e-4) { alpha = 1 - t2 / 12 - t2 * t2 / 720; beta = 1.0 / 12 + t2 / 720; }
  else { double st = sin(t), ct = cos(t); alpha = t * st / (2 * (1 - ct)); beta = 1 / t2 - st / (2 * t * (1 - ct)); }
  double wxp[3], wp = dot3(w, p);
  cross3(wxp, w, p);
  for (int k = 0; k < 3; k++) out[k] = alpha * p[k] - 0.5 * wxp[k] + beta * wp * w[k];
  copy3(out + 3, w);
}
static void skew(const double* v, double* S) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
static void jlog3(double t, const double* w, double* Jl) { /* pinocchio::Jlog3 [3P] */
  double alpha, diag;
  if (t < 1e-4) { alpha = 1.0 / 12 + t * t / 720; diag = 0.5 * (2 - t * t / 6); }
  else { double st = sin(t), ct = cos(t); alpha = 1 / (t * t) - st / (2 * t * (1 - ct)); diag = 0.5 * (t * st / (1 - ct)); }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Jl[3 * r + c] = alpha * w[r] * w[c];
  Jl[0] += diag; Jl[4] += diag; Jl[8] += diag;
  double S[9];
  skew(w, S);
  for (int i = 0; i < 9; i++) Jl[i] += 0.5 * S[i];
}
static void jlog6(const double* R, const double* p, double* J6 /* 6x6 row-major, [lin;ang] */) { /* pinocchio::Jlog6 [3P] */
  double w[3], t;
  log3(R, w, &t);
  double TL[9];
  jlog3(t, w, TL);
  double t2 = t * t, beta, beta_dot_over_theta;
  if (t < 1e-4) { beta = 1.0 / 12 + t2 / 720; beta_dot_over_theta = 1.0 / 360; }
  else {
    double st = sin(t), ct = cos(t), tinv = 1 / t, t2inv = tinv * tinv, inv_2_2ct = 1 / (2 * (1 - ct));
    beta = t2inv - st * tinv * inv_2_2ct;
    beta_dot_over_theta = -2 * t2inv * t2inv + (1 + st * tinv) * t2inv * inv_2_2ct;
  }
  double wTp = dot3(w, p), v3[3], B[9], S[9];
  for (int k = 0; k < 3; k++) v3[k] = (beta_dot_over_theta * wTp) * w[k] - (t2 * beta_dot_over_theta + 2 * beta) * p[k];
  skew(p, S);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) B[3 * r + c] = 0.5 * S[3 * r + c] + v3[r] * w[c] + beta * w[r] * p[c];
  B[0] += wTp * beta; B[4] += wTp * beta; B[8] += wTp * beta;
  double TR[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) TR[3 * r + c] = B[3 * r] * TL[c] + B[3 * r + 1] * TL[3 + c] + B[3 * r + 2] * TL[6 + c];
  zero(J6, 36);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      J6[6 * r + c] = TL[3 * r + c];
      J6[6 * r + 3 + c] = TR[3 * r + c];
      J6[6 * (3 + r) + 3 + c] = TL[3 * r + c];
    }
}
static int ldlt6_solve(double* A, double* b) { /* in-place LDL^T, SPD expected */
  double L[36] = {0}, D[6];
  for (int j = 0; j < 6; j++) {
    double s = A[6 * j + j];
    for (int k = 0; k < j; k++) s -= L[6 * j + k] * L[6 * j + k] * D[k];
    D[j] = s;
    if (s == 0) return -1;
    L[6 * j + j] = 1;
    for (int i = j + 1; i < 6; i++) {
      double t = A[6 * i + j];
      for (int k = 0; k < j; k++) t -= L[6 * i + k] * L[6 * j + k] * D[k];
      L[6 * i + j] = t / s;
    }
  }
  for (int i = 0; i < 6; i++) for (int k = 0; k < i; k++) b[i] -= L[6 * i + k] * b[k];
  for (int i = 0; i < 6; i++) b[i] /= D[i];
  for (int i = 5; i >= 0; i--) for (int k = i + 1; k < 6; k++) b[i] -= L[6 * k + i] * b[k];
  return 0;
}

int rcso_ik_inverse(const rcso_model* m, int site, int nq_model, const double* pose7, const double* q0, int nq0,
                    const double* tcp_offset7, double* q_out, int* iters) {
  const double eps = 1e-4, DT = 1e-1, damp = 1e-6;
  const int IT_MAX = 1000;
  double inv_tcp[7], goal[7], Rd[9];
  rcso_pose_inverse(tcp_offset7, inv_tcp);
  rcso_pose_mul(pose7, inv_tcp, goal);
  rcso_pose_rotation_m(goal, Rd);
  double q[64] = {0}, J[6 * 64], Jn[6 * 64];
  for (int i = 0; i < nq0 && i < nq_model; i++) q[i] = q0[i];
  int success = 0, i;
  for (i = 0;; i++) {
    double R[9], p[3];
    site_fk(m, site, nq_model, q, R, p, J);
    /* iMd = oMf^-1 * oMdes */
    double Ri[9], pi[3], dp[3] = {goal[0] - p[0], goal[1] - p[1], goal[2] - p[2]};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Ri[3 * r + c] = R[r] * Rd[c] + R[3 + r] * Rd[3 + c] + R[6 + r] * Rd[6 + c];
    mulmatT3(pi, R, dp);
    double err[6];
    log6(Ri, pi, err);
    double en = 0;
    for (int k = 0; k < 6; k++) en += err[k] * err[k];
    if (sqrt(en) < eps) { success = 1; break; }
    if (i >= IT_MAX) { success = 0; break; }
    /* J <- -Jlog6(iMd^-1) * J */
    double Rinv[9], pinv[3], Jl[36];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rinv[3 * r + c] = Ri[3 * c + r];
    mulmat3(pinv, Rinv, pi);
    for (int k = 0; k < 3; k++) pinv[k] = -pinv[k];
    jlog6(Rinv, pinv, Jl);
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < nq_model; c++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += Jl[6 * r + k] * J[k * nq_model + c];
        Jn[r * nq_model + c] = -s;
      }
    double JJt[36], y[6];
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) {
        double s = 0;
        for (int k = 0; k < nq_model; k++) s += Jn[r * nq_model + k] * Jn[c * nq_model + k];
        JJt[6 * r + c] = s;
      }
    for (int k = 0; k < 6; k++) JJt[7 * k] += damp;
    memcpy(y, err, sizeof(y));
    ldlt6_solve(JJt, y);
    for (int c = 0; c < nq_model; c++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += Jn[k * nq_model + c] * y[k];
      q[c] += -s * DT;
    }
  }
  if (iters) *iters = i;
  if (success) for (int k = 0; k < nq_model; k++) q_out[k] = q[k];
  return success;
}
void rcso_ik_forward(const rcso_model* m, int site, int nq_model, const double* q0, int nq0, const double* tcp_offset7,
                     double* pose7) { /* Kinematics.cpp:70-81 (post-multiplies by tcp_offset.inverse(), sic) */
  double q[64] = {0}, R[9], p[3], f[7], inv_tcp[7];
  for (int i = 0; i < nq0 && i < nq_model; i++) q[i] = q0[i];
  site_fk(m, site, nq_model, q, R, p, NULL);
  rcso_pose_from_matrix(R, p, f);
  rcso_pose_inverse(tcp_offset7, inv_tcp);
  rcso_pose_mul(f, inv_tcp, pose7);
}

/* ================================================================== Sim + SimRobot + SimGripper */
enum { CB_ROBOT_ARRIVED, CB_ROBOT_MOVING, CB_ROBOT_CONVERGENCE, CB_ROBOT_COLLISION, CB_GRIPPER_CONVERGENCE,
       CB_GRIPPER_COLLISION };
typedef struct { int kind; double period, last; int last_return; } cb_t;

struct rcso_sim {
  const rcso_model* m;
  rcso_data* d;
  rcso_robot_cfg rc;
  rcso_gripper_cfg gc;
  /* SimConfig, sim.h:29-34 */
  int async_control, frequency, max_convergence_steps;
  cb_t cbs[4], any_cbs[4], all_cbs[4];
  int ncb, nany, nall;
  int converged, convergence_steps;
  /* SimRobotState, SimRobot.h:49-57 */
  double previous_angles[8], target_angles[16];
  int ik_success, collision, is_moving, is_arrived;
  /* SimGripperState, SimGripper.h:47-52 */
  double g_last_commanded_width, g_last_width;
  int g_is_moving, g_collision;
};

static int in_set(const int* set, int n, int v) {
  for (int i = 0; i < n; i++) if (set[i] == v) return 1;
  return 0;
}
static int run_cb(rcso_sim* s, int kind) {
  rcso_data* d = s->d;
  switch (kind) {
    case CB_ROBOT_ARRIVED: { /* SimRobot.cpp:165-170 */
      double mx = 0;
      for (int i = 0; i < s->rc.njoints; i++) {
        double e = fabs(d->qpos[s->rc.joint_qposadr[i]] - s->target_angles[i]);
        if (e > mx) mx = e;
      }
      s->is_arrived = mx < s->rc.joint_rotational_tolerance;
      return 0;
    }
    case CB_ROBOT_MOVING: { /* SimRobot.cpp:156-163 */
      double mx = 0;
      for (int i = 0; i < s->rc.njoints; i++) {
        double q = d->qpos[s->rc.joint_qposadr[i]];
        double e = fabs(q - s->previous_angles[i]);
        if (e > mx) mx = e;
        s->previous_angles[i] = q;
      }
      s->is_moving = mx > 0.0001;
      return 0;
    }
    case CB_ROBOT_CONVERGENCE: /* SimRobot.cpp:184-191 */
      if (!s->ik_success) return 1;
      return s->is_arrived && !s->is_moving;
    case CB_ROBOT_COLLISION: /* SimRobot.cpp:172-182 */
      s->collision = 0;
      for (int i = 0; i < d->ncon; i++)
        if (in_set(s->rc.cgeom, s->rc.ncgeom, d->contact[i].geom[0]) || in_set(s->rc.cgeom, s->rc.ncgeom, d->contact[i].geom[1])) {
          s->collision = 1;
          break;
        }
      return s->collision;
    case CB_GRIPPER_CONVERGENCE: { /* SimGripper.cpp:143-151 */
      double w = rcso_gripper_get_normalized_width(s);
      s->g_is_moving = fabs(s->g_last_width - w) > 0.001 * (s->gc.max_actuator_width - s->gc.min_actuator_width);
      s->g_last_width = w;
      return !s->g_is_moving;
    }
    case CB_GRIPPER_COLLISION: /* SimGripper.cpp:108-130 */
      s->g_collision = 0;
      for (int i = 0; i < d->ncon; i++) {
        int g0 = d->contact[i].geom[0], g1 = d->contact[i].geom[1];
        if (in_set(s->gc.cfgeom, s->gc.ncfgeom, g0) && in_set(s->gc.cfgeom, s->gc.ncfgeom, g1)) continue;
        if ((in_set(s->gc.cgeom, s->gc.ncgeom, g0) || in_set(s->gc.cgeom, s->gc.ncgeom, g1)) &&
            !(in_set(s->gc.ignored, s->gc.nignored, g1) || in_set(s->gc.ignored, s->gc.nignored, g1))) {
          s->g_collision = 1;
          break;
        }
      }
      return s->g_collision;
  }
  return 0;
}

rcso_sim* rcso_sim_new(const rcso_model* m, const rcso_robot_cfg* rc, const rcso_gripper_cfg* gc) {
  rcso_sim* s = (rcso_sim*)calloc(1, sizeof(rcso_sim));
  s->m = m;
  s->d = rcso_data_new(m);
  s->rc = *rc;
  if (gc) s->gc = *gc;
  s->frequency = 30; s->max_convergence_steps = 500; s->converged = 1;
  s->ik_success = 1;
  /* SimRobot ctor, SimRobot.cpp:27-43 */
  if (rc->register_convergence_callback) {
    s->cbs[s->ncb++] = (cb_t){CB_ROBOT_ARRIVED, rc->seconds_between_callbacks, 0, 0};
    s->cbs[s->ncb++] = (cb_t){CB_ROBOT_MOVING, rc->seconds_between_callbacks, 0, 0};
    s->all_cbs[s->nall++] = (cb_t){CB_ROBOT_CONVERGENCE, rc->seconds_between_callbacks, 0, 0};
  }
  s->any_cbs[s->nany++] = (cb_t){CB_ROBOT_COLLISION, rc->seconds_between_callbacks, 0, 0};
  rcso_robot_reset(s);
  if (gc && gc->enabled) { /* SimGripper ctor, SimGripper.cpp:12-39 */
    s->all_cbs[s->nall++] = (cb_t){CB_GRIPPER_CONVERGENCE, gc->seconds_between_callbacks, 0, 0};
    s->any_cbs[s->nany++] = (cb_t){CB_GRIPPER_COLLISION, gc->seconds_between_callbacks, 0, 0};
    rcso_gripper_reset(s);
  }
  return s;
}
void rcso_sim_free(rcso_sim* s) { if (s) { rcso_data_free(s->d); free(s); } }
rcso_data* rcso_sim_data(rcso_sim* s) { return s->d; }
void rcso_sim_set_config(rcso_sim* s, int async_control, int frequency, int max_convergence_steps) {
  s->async_control = async_control; s->frequency = frequency; s->max_convergence_steps = max_convergence_steps;
}
static void invoke_callbacks(rcso_sim* s) { /* sim.cpp:38-47 */
  for (int i = 0; i < s->ncb; i++) {
    double dt = s->d->time - s->cbs[i].last;
    if (dt > s->cbs[i].period) { run_cb(s, s->cbs[i].kind); s->cbs[i].last = s->d->time; }
  }
}
static void process_condition(rcso_sim* s, cb_t* cbs, int n) { /* sim.cpp:14-23 */
  for (int i = 0; i < n; i++) {
    double dt = s->d->time - cbs[i].last;
    if (dt > cbs[i].period) { cbs[i].last_return = run_cb(s, cbs[i].kind); cbs[i].last = s->d->time; }
  }
}
static int invoke_condition_callbacks(rcso_sim* s) { /* sim.cpp:49-61 */
  process_condition(s, s->any_cbs, s->nany);
  process_condition(s, s->all_cbs, s->nall);
  for (int i = 0; i < s->nany; i++) if (s->any_cbs[i].last_return) return 1;
  for (int i = 0; i < s->nall; i++) if (!s->all_cbs[i].last_return) return 0;
  return 1;
}
void rcso_sim_step(rcso_sim* s, int k) { /* sim.cpp:108-115 */
  for (int i = 0; i < k; i++) {
    rcso_step1(s->m, s->d);
    invoke_callbacks(s);
    rcso_step2(s->m, s->d);
  }
}
void rcso_sim_step_until_convergence(rcso_sim* s) { /* sim.cpp:84-106 */
  s->convergence_steps = 0;
  s->converged = 0;
  for (int i = 0; i < s->nany; i++) s->any_cbs[i].last_return = 0;
  for (int i = 0; i < s->nall; i++) s->all_cbs[i].last_return = 0;
  while (!s->converged && (s->max_convergence_steps == -1 || s->convergence_steps < s->max_convergence_steps)) {
    rcso_sim_step(s, 1);
    s->convergence_steps++;
    s->converged = invoke_condition_callbacks(s);
  }
}
int rcso_sim_is_converged(rcso_sim* s) { return s->converged; }
int rcso_sim_convergence_steps(rcso_sim* s) { return s->convergence_steps; }
void rcso_sim_reset(rcso_sim* s) { /* sim.cpp:117-138 */
  rcso_reset_data(s->m, s->d);
  for (int i = 0; i < s->ncb; i++) s->cbs[i].last = 0;
  for (int i = 0; i < s->nany; i++) s->any_cbs[i].last = 0;
  for (int i = 0; i < s->nall; i++) s->all_cbs[i].last = 0;
}
void rcso_robot_get_joint_position(rcso_sim* s, double* q) {
  for (int i = 0; i < s->rc.njoints; i++) q[i] = s->d->qpos[s->rc.joint_qposadr[i]];
}
static void robot_set_joints_n(rcso_sim* s, const double* q, int n) {
  for (int i = 0; i < n && i < 16; i++) s->target_angles[i] = q[i];
  rcso_robot_get_joint_position(s, s->previous_angles);
  s->is_moving = 1;
  s->is_arrived = 0;
  for (int i = 0; i < s->rc.njoints; i++) s->d->ctrl[s->rc.actuator_id[i]] = q[i];
}
void rcso_robot_set_joint_position(rcso_sim* s, const double* q) { robot_set_joints_n(s, q, s->rc.njoints); }
void rcso_robot_get_cartesian_position(rcso_sim* s, double* pose7) { /* SimRobot.cpp:114-121,207-213; Robot.cpp:5-8 */
  rcso_data* d = s->d;
  double site[7], base[7], binv[7], t[7];
  rcso_pose_from_matrix(d->site_xmat + 9 * s->rc.attachment_site, d->site_xpos + 3 * s->rc.attachment_site, site);
  const double* bq = d->xquat + 4 * s->rc.base_body;
  copy3(base, d->xpos + 3 * s->rc.base_body);
  quat_t q = {bq[1], bq[2], bq[3], bq[0]};
  q_store(base, q_normalized(q));
  rcso_pose_inverse(base, binv);
  rcso_pose_mul(binv, site, t);
  rcso_pose_mul(t, s->rc.tcp_offset, pose7);
}
int rcso_robot_set_cartesian_position(rcso_sim* s, const double* pose7) { /* SimRobot.cpp:145-155 */
  double q0[8], q[64];
  rcso_robot_get_joint_position(s, q0);
  int ok = rcso_ik_inverse(s->m, s->rc.attachment_site, s->rc.ik_nq, pose7, q0, s->rc.njoints, s->rc.tcp_offset, q, NULL);
  if (ok) { s->ik_success = 1; robot_set_joints_n(s, q, s->rc.ik_nq); }
  else s->ik_success = 0;
  return ok;
}
void rcso_robot_reset(rcso_sim* s) { /* SimRobot.cpp:193-205 */
  for (int i = 0; i < s->rc.njoints; i++) {
    s->d->qpos[s->rc.joint_qposadr[i]] = s->rc.q_home[i];
    s->d->ctrl[s->rc.actuator_id[i]] = s->rc.q_home[i];
  }
}
void rcso_robot_state(rcso_sim* s, int* ik_success, int* collision, int* is_moving, int* is_arrived,
                      double* previous_angles, double* target_angles) {
  if (ik_success) *ik_success = s->ik_success;
  if (collision) *collision = s->collision;
  if (is_moving) *is_moving = s->is_moving;
  if (is_arrived) *is_arrived = s->is_arrived;
  if (previous_angles) memcpy(previous_angles, s->previous_angles, sizeof(double) * (size_t)s->rc.njoints);
  if (target_angles) memcpy(target_angles, s->target_angles, sizeof(double) * (size_t)s->rc.njoints);
}
int rcso_gripper_set_normalized_width(rcso_sim* s, double width, double force) { /* SimGripper.cpp:79-92 */
  if (width < 0 || width > 1 || force < 0) return -1;
  s->g_last_commanded_width = width;
  s->d->ctrl[s->gc.actuator_id] = width * (s->gc.max_actuator_width - s->gc.min_actuator_width) + s->gc.min_actuator_width;
  return 0;
}
double rcso_gripper_get_normalized_width(rcso_sim* s) { /* SimGripper.cpp:93-106 */
  double w = (s->d->qpos[s->gc.joint_qposadr] - s->gc.min_joint_width) / (s->gc.max_joint_width - s->gc.min_joint_width);
  if (w < 0) w = 0; else if (w > 1) w = 1;
  return w;
}
int rcso_gripper_is_grasped(rcso_sim* s) { /* SimGripper.cpp:132-141 */
  double w = rcso_gripper_get_normalized_width(s);
  return s->g_last_commanded_width - s->gc.epsilon_inner < w && w < s->g_last_commanded_width + s->gc.epsilon_outer;
}
void rcso_gripper_reset(rcso_sim* s) { /* SimGripper.cpp:158-163 */
  s->g_last_commanded_width = 0; s->g_is_moving = 0; s->g_last_width = 0; s->g_collision = 0;
  s->d->qpos[s->gc.joint_qposadr] = s->gc.max_joint_width;
  s->d->ctrl[s->gc.actuator_id] = s->gc.max_actuator_width;
}
void rcso_gripper_state(rcso_sim* s, double* lcw, int* is_moving, double* last_width, int* collision) {
  if (lcw) *lcw = s->g_last_commanded_width;
  if (is_moving) *is_moving = s->g_is_moving;
  if (last_width) *last_width = s->g_last_width;
  if (collision) *collision = s->g_collision;
}

/* ================================================================== env loop (CPU baseline workload) */
typedef struct {
  const rcso_model* m; const rcso_robot_cfg* rc; const rcso_gripper_cfg* gc;
  int env_begin, env_end, nsteps, episode_len, async_control;
  const double* actions; double max_mov; const double *low, *high;
  long long physics_steps; double* obs_out;
} worker_t;

#define ENV_OBS_STRIDE 28
/* obs[0:21] = tquat, joints, xyzrpy, gripper; obs[21] = info["gripper_width"]; obs[22:28] = info flags collision,
 * ik_success, is_sim_converged, is_grasped, robot collision, gripper collision (envs/sim.py:60-66,125-131) */
static void env_obs(rcso_sim* s, double gripper_obs, double* obs /* ENV_OBS_STRIDE */) { /* base.py:246-253,710-719 */
  double pose[7];
  rcso_robot_get_cartesian_position(s, pose);
  memcpy(obs, pose, 7 * sizeof(double));
  rcso_robot_get_joint_position(s, obs + 7);
  rcso_pose_xyzrpy(pose, obs + 14);
  obs[20] = gripper_obs;
  double gw = s->gc.enabled ? rcso_gripper_get_normalized_width(s) : 0;
  int rc = s->collision, gcol = s->gc.enabled ? s->g_collision : 0;
  obs[21] = gw;
  obs[22] = rc || gcol;
  obs[23] = s->ik_success;
  obs[24] = rcso_sim_is_converged(s);
  obs[25] = gw > 0.01 && gw < 0.99;
  obs[26] = rc;
  obs[27] = gcol;
}
static void env_reset(rcso_sim* s, long long* psteps) { /* base.py:703-708, envs/sim.py:68-76 */
  if (s->gc.enabled) rcso_gripper_reset(s);
  rcso_sim_reset(s);
  rcso_robot_reset(s);
  rcso_sim_step(s, 1);
  *psteps += 1;
}
static void* worker(void* arg) {
  worker_t* w = (worker_t*)arg;
  int nj = w->rc->njoints;
  for (int e = w->env_begin; e < w->env_end; e++) {
    rcso_sim* s = rcso_sim_new(w->m, w->rc, w->gc);
    rcso_sim_set_config(s, w->async_control, 30, 500);
    double prev[8]; int have_prev = 0; double grip_obs = 1;
    env_reset(s, &w->physics_steps);
    for (int t = 0; t < w->nsteps; t++) {
      if (w->episode_len > 0 && t > 0 && t % w->episode_len == 0) { env_reset(s, &w->physics_steps); grip_obs = 1; }
      const double* a = w->actions + ((size_t)e * w->nsteps + t) * 8;
      double origin[8], joints[8];
      rcso_robot_get_joint_position(s, origin); /* RelativeActionSpace.action, base.py:469-488 */
      for (int i = 0; i < nj; i++) {
        double lim = a[i] < -w->max_mov ? -w->max_mov : (a[i] > w->max_mov ? w->max_mov : a[i]);
        double v = origin[i] + lim;
        joints[i] = v < w->low[i] ? w->low[i] : (v > w->high[i] ? w->high[i] : v);
      }
      if (s->gc.enabled) { /* GripperWrapper.action binary, base.py:721-735 */
        double g = nearbyint(a[7]);
        g = g < 0 ? 0 : (g > 1 ? 1 : g);
        rcso_gripper_set_normalized_width(s, g == 0 ? 0.0 : 1.0, 0);
        grip_obs = g;
      }
      int changed = !have_prev; /* RobotEnv.step dedupe, base.py:268-272 */
      for (int i = 0; i < nj && !changed; i++) if (fabs(joints[i] - prev[i]) > 1e-3) changed = 1;
      if (changed) rcso_robot_set_joint_position(s, joints);
      memcpy(prev, joints, sizeof(double) * (size_t)nj); have_prev = 1;
      if (w->async_control) { /* envs/sim.py:52-53 */
        int k = (int)nearbyint(1.0 / 30 / w->m->timestep);
        rcso_sim_step(s, k);
        w->physics_steps += k;
      } else {
        rcso_sim_step_until_convergence(s);
        w->physics_steps += s->convergence_steps;
      }
      if (w->obs_out) env_obs(s, grip_obs, w->obs_out + ((size_t)e * w->nsteps + t) * ENV_OBS_STRIDE);
    }
    rcso_sim_free(s);
  }
  return NULL;
}
double rcso_bench_env_steps(const rcso_model* m, const rcso_robot_cfg* rc, const rcso_gripper_cfg* gc, int nenv,
                            int nthreads, int nsteps, int episode_len, int async_control, const double* actions,
                            double max_mov, const double* joint_low, const double* joint_high,
                            long long* physics_steps_out, double* obs_out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nenv) nthreads = nenv;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  worker_t* ws = (worker_t*)calloc((size_t)nthreads, sizeof(worker_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    ws[t] = (worker_t){m, rc, gc, (int)((long long)nenv * t / nthreads), (int)((long long)nenv * (t + 1) / nthreads),
                       nsteps, episode_len, async_control, actions, max_mov, joint_low, joint_high, 0, obs_out};
    pthread_create(&th[t], NULL, worker, &ws[t]);
  }
  long long total = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); total += ws[t].physics_steps; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (physics_steps_out) *physics_steps_out = total;
  free(th);
  free(ws);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
