// Batched damped-least-squares CLIK: the algorithm of rcs.common.Pin::inverse
// (/root/reference/src/rcs/Kinematics.cpp:28-68; eps 1e-4, IT_MAX 1000, DT 0.1, damp 1e-6 from
// include/rcs/Kinematics.h:32-35), which the reference runs on Pinocchio (forwardKinematics,
// computeFrameJacobian LOCAL, log6, Jlog6, 6x6 LDLT, integrate). One environment per thread: the
// loop is a ~100-iteration serial chain of tiny dense operations on a 6 x nq Jacobian, so the
// parallelism is across environments; the 6x6 solves are far too small for tensor-core tiles.

enum { IK_SCRATCH_REALS = 8, IK_MAXQ = 9 };

// NQ / NCH > 0 fix the number of IK joints and the length of the body chain at compile time: every loop unrolls and the
// Jacobian / chain arrays stay in registers instead of thread-local memory (dynamic indexing); 0 = read them at run time.
template <int NQ, int NCH>
RCSB_DEV void ik_site_fk(const RcsbModel& m, const real* q, int nqm_rt, real* R, real* p, real* J) {
  const int nqm = NQ > 0 ? NQ : nqm_rt;
  constexpr int CH = NCH > 0 ? NCH : RCSB_MAXB;
  // chain root -> site body, rotation-matrix form (as st_kinematics): local frame [R | t] of every body from its constant
  // offset (b_rot, b_pos) and its joint motion (Rodrigues), composed with the parent's world frame
  int chain[CH], n = 0;
  if (NCH > 0) {
    int b = m.rb_site_body;
#pragma unroll
    for (int i = 0; i < CH; i++) { chain[i] = b; b = m.b_parent[b >= 0 ? b : 0]; }
    n = CH;
  } else {
    for (int b = m.rb_site_body; b >= 0; b = m.b_parent[b]) chain[n++] = b;
  }
  real Rw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pw[3] = {0, 0, 0};
  real anchors[CH][3], axes[CH][3];
  int jtype[CH], jdof[CH], nj = 0;
#pragma unroll
  for (int cj = 0; cj < CH; cj++) {
    const int ci = n - 1 - cj;
    if (NCH == 0 && ci < 0) break;
    const int b = chain[NCH > 0 ? CH - 1 - cj : ci];
    const real* Rb = m.b_rot[b];
    real Rl[9], tl[3];
    if (m.b_jtype[b] == RCSB_JNT_FREE) {  // not part of an arm chain: treated as the fixed offset
      for (int i = 0; i < 9; i++) Rl[i] = Rb[i];
      copy3(tl, m.b_pos[b]);
    } else {
      const int qa = m.b_qadr[b];
      const real qj = (qa < nqm ? q[qa] : (real)0) - m.qpos0[qa];
      const real* u = m.b_jaxis[b];
      const real* jp = m.b_jpos[b];
      if (m.b_jtype[b] == RCSB_JNT_SLIDE) {
        real v[3];
        mulmat3(v, Rb, u);
        for (int i = 0; i < 9; i++) Rl[i] = Rb[i];
        tl[0] = m.b_pos[b][0] + v[0] * qj; tl[1] = m.b_pos[b][1] + v[1] * qj; tl[2] = m.b_pos[b][2] + v[2] * qj;
      } else {
        real sn, co;
        sincos(qj, &sn, &co);
        const real oc = 1 - co;
        real Rj[9] = {co + oc * u[0] * u[0], oc * u[0] * u[1] - sn * u[2], oc * u[0] * u[2] + sn * u[1],
                      oc * u[1] * u[0] + sn * u[2], co + oc * u[1] * u[1], oc * u[1] * u[2] - sn * u[0],
                      oc * u[2] * u[0] - sn * u[1], oc * u[2] * u[1] + sn * u[0], co + oc * u[2] * u[2]};
        for (int r = 0; r < 3; r++)
          for (int k = 0; k < 3; k++) Rl[3 * r + k] = Rb[3 * r] * Rj[k] + Rb[3 * r + 1] * Rj[3 + k] + Rb[3 * r + 2] * Rj[6 + k];
        real w[3], v[3];
        mulmat3(w, Rj, jp);
        w[0] = jp[0] - w[0]; w[1] = jp[1] - w[1]; w[2] = jp[2] - w[2];
        mulmat3(v, Rb, w);
        tl[0] = m.b_pos[b][0] + v[0]; tl[1] = m.b_pos[b][1] + v[1]; tl[2] = m.b_pos[b][2] + v[2];
      }
    }
    real Rn[9], v[3];
    mulmat3(v, Rw, tl);
    pw[0] += v[0]; pw[1] += v[1]; pw[2] += v[2];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) Rn[3 * r + k] = Rw[3 * r] * Rl[k] + Rw[3 * r + 1] * Rl[3 + k] + Rw[3 * r + 2] * Rl[6 + k];
    for (int i = 0; i < 9; i++) Rw[i] = Rn[i];
    if (m.b_jtype[b] != RCSB_JNT_FREE) {  // a rotation about the joint axis leaves axis and anchor in place
      mulmat3(axes[nj], Rw, m.b_jaxis[b]);
      mulmat3(v, Rw, m.b_jpos[b]);
      anchors[nj][0] = pw[0] + v[0]; anchors[nj][1] = pw[1] + v[1]; anchors[nj][2] = pw[2] + v[2];
      jtype[nj] = m.b_jtype[b];
      jdof[nj] = m.b_dadr[b];
      nj++;
    }
  }
  {
    real v[3];
    mulmat3(v, Rw, m.rb_site_pos);
    p[0] = pw[0] + v[0]; p[1] = pw[1] + v[1]; p[2] = pw[2] + v[2];
    const real* Rs = m.rb_site_rot;
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) R[3 * r + k] = Rw[3 * r] * Rs[k] + Rw[3 * r + 1] * Rs[3 + k] + Rw[3 * r + 2] * Rs[6 + k];
  }
  if (J) {
#pragma unroll
    for (int i = 0; i < 6 * (NQ > 0 ? NQ : IK_MAXQ); i++) if (i < 6 * nqm) J[i] = 0;
#pragma unroll
    for (int a = 0; a < CH; a++) {
      if (a >= nj || jdof[a] >= nqm) continue;
      real lin[3], ang[3] = {0, 0, 0}, r[3], l[3], w[3];
      if (jtype[a] == RCSB_JNT_HINGE) {
        r[0] = p[0] - anchors[a][0]; r[1] = p[1] - anchors[a][1]; r[2] = p[2] - anchors[a][2];
        cross3(lin, axes[a], r);
        copy3(ang, axes[a]);
      } else {
        copy3(lin, axes[a]);
      }
      mulmatT3(l, R, lin);
      mulmatT3(w, R, ang);
      // fixed-shape path: the dispatcher guarantees that the a-th joint of the chain drives dof a (static column index)
      const int col = NCH > 0 ? a : jdof[a];
      for (int k = 0; k < 3; k++) { J[k * nqm + col] = l[k]; J[(3 + k) * nqm + col] = w[k]; }
    }
  }
}
// log3 of a rotation together with sin / cos of its angle: one acos and one sincos per CLIK iteration serve the error
// (log6 of iMd) and the Jacobian of the logarithm (Jlog6 of iMd^-1, whose rotation vector is the negated one).
struct Log3 { real w[3], t, st, ct; };
RCSB_DEV void ik_log3(const real* R, Log3& L) {
  real ct = (real)0.5 * (R[0] + R[4] + R[8] - 1);
  ct = ct > 1 ? (real)1 : (ct < -1 ? (real)-1 : ct);
  const real t = acos(ct);
  L.t = t;
  sincos(t, &L.st, &L.ct);
  // Pinocchio 3.7 log3: near pi the axis comes from the diagonal, w_k^2 = theta^2 (R_kk - cos) / (1 - cos), signed by the
  // antisymmetric part (threshold pi - 1e-2); below eps^(1/4) the factor theta / sin(theta) is taken as 1
  if (t >= (real)3.14159265358979323846 - (real)1e-2) {
    const real beta = t * t / (1 - ct);
    const real a[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    for (int k = 0; k < 3; k++) {
      real v = (R[4 * k] - ct) * beta;
      L.w[k] = (a[k] > 0 ? (real)1 : (real)-1) * (v > 0 ? r_sqrt(v) : (real)0);
    }
    return;
  }
  const real f = (real)0.5 * (t > (real)1.220703125e-4 ? t / L.st : (real)1);
  L.w[0] = f * (R[7] - R[5]); L.w[1] = f * (R[2] - R[6]); L.w[2] = f * (R[3] - R[1]);
}
RCSB_DEV void ik_log6(const Log3& L, const real* p, real* out) {
  const real* w = L.w;
  const real t = L.t;
  real alpha, beta;
  real t2 = t * t;
  if (t < (real)1e-4) { alpha = 1 - t2 / 12 - t2 * t2 / 720; beta = (real)1 / 12 + t2 / 720; }
  else { real st = L.st, ct = L.ct; alpha = t * st / (2 * (1 - ct)); beta = 1 / t2 - st / (2 * t * (1 - ct)); }
  real wxp[3], wp = dot3(w, p);
  cross3(wxp, w, p);
  for (int k = 0; k < 3; k++) out[k] = alpha * p[k] - (real)0.5 * wxp[k] + beta * wp * w[k];
  copy3(out + 3, w);
}
RCSB_DEV void ik_skew(const real* v, real* S) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
// Jlog6 of the transform whose rotation vector is sign * L.w (sign = -1: the inverse rotation, same angle) and translation p
RCSB_DEV void ik_jlog6(const Log3& L, real sign, const real* p, real* J6) {
  const real w[3] = {sign * L.w[0], sign * L.w[1], sign * L.w[2]};
  const real t = L.t;
  real TL[9], alpha, diag, S[9];
  if (t < (real)1e-4) { alpha = (real)1 / 12 + t * t / 720; diag = (real)0.5 * (2 - t * t / 6); }
  else { real st = L.st, ct = L.ct; alpha = 1 / (t * t) - st / (2 * t * (1 - ct)); diag = (real)0.5 * (t * st / (1 - ct)); }
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) TL[3 * r + cc] = alpha * w[r] * w[cc];
  TL[0] += diag; TL[4] += diag; TL[8] += diag;
  ik_skew(w, S);
  for (int i = 0; i < 9; i++) TL[i] += (real)0.5 * S[i];
  real t2 = t * t, beta, bdot;
  if (t < (real)1e-4) { beta = (real)1 / 12 + t2 / 720; bdot = (real)1 / 360; }
  else {
    real st = L.st, ct = L.ct, tinv = 1 / t, t2inv = tinv * tinv, inv = 1 / (2 * (1 - ct));
    beta = t2inv - st * tinv * inv;
    bdot = -2 * t2inv * t2inv + (1 + st * tinv) * t2inv * inv;
  }
  real wTp = dot3(w, p), v3[3], B[9], TR[9];
  for (int k = 0; k < 3; k++) v3[k] = (bdot * wTp) * w[k] - (t2 * bdot + 2 * beta) * p[k];
  ik_skew(p, S);
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) B[3 * r + cc] = (real)0.5 * S[3 * r + cc] + v3[r] * w[cc] + beta * w[r] * p[cc];
  B[0] += wTp * beta; B[4] += wTp * beta; B[8] += wTp * beta;
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) TR[3 * r + cc] = B[3 * r] * TL[cc] + B[3 * r + 1] * TL[3 + cc] + B[3 * r + 2] * TL[6 + cc];
  for (int i = 0; i < 36; i++) J6[i] = 0;
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) {
      J6[6 * r + cc] = TL[3 * r + cc];
      J6[6 * r + 3 + cc] = TR[3 * r + cc];
      J6[6 * (3 + r) + 3 + cc] = TL[3 * r + cc];
    }
}
// 6 x 6 LDL^T solve (Eigen's ldlt().solve in Kinematics.cpp:57-59), fully unrolled: every index is static, so L and D
// live in registers (the rolled version kept them in thread-local memory and spent 40 % of the IK kernel there). The 21
// divisions by the pivots are 6 reciprocals and multiplications (an FP64 division is a ~20-instruction dependent chain):
// entries of L differ from the divided ones by at most one ulp.
RCSB_DEV void ik_ldlt6(const real* A, real* b) {
  real Lm[36], D[6], Dinv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    real s = A[6 * j + j];
#pragma unroll
    for (int k = 0; k < 6; k++) if (k < j) s -= Lm[6 * j + k] * Lm[6 * j + k] * D[k];
    D[j] = s;
    const real sinv = (real)1 / s;
    Dinv[j] = sinv;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (i <= j) continue;
      real t = A[6 * i + j];
#pragma unroll
      for (int k = 0; k < 6; k++) if (k < j) t -= Lm[6 * i + k] * Lm[6 * j + k] * D[k];
      Lm[6 * i + j] = t * sinv;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int k = 0; k < 6; k++) if (k < i) b[i] -= Lm[6 * i + k] * b[k];
#pragma unroll
  for (int i = 0; i < 6; i++) b[i] *= Dinv[i];
#pragma unroll
  for (int i = 5; i >= 0; i--)
#pragma unroll
    for (int k = 0; k < 6; k++) if (k > i) b[i] -= Lm[6 * k + i] * b[k];
}

// returns success; q_out[nqm]
template <int NQ, int NCH>
RCSB_DEV int ik_solve_t(const RcsbModel& m, const real* pose7, const real* q0, int nq0, real* q_out, int* iters_out) {
  constexpr int QM = NQ > 0 ? NQ : IK_MAXQ;
  const int nqm = NQ > 0 ? NQ : (m.rb_ik_nq < IK_MAXQ ? m.rb_ik_nq : IK_MAXQ);
  real inv_tcp[7], goal[7], Rd[9];
  pose_inverse(m.rb_tcp_offset, inv_tcp);
  pose_mul(pose7, inv_tcp, goal);
  Quat qg = {goal[3], goal[4], goal[5], goal[6]};
  q_to_mat(qg, Rd);
  real q[QM], J[6 * QM], Jn[6 * QM];
#pragma unroll
  for (int i = 0; i < QM; i++) if (i < nqm) q[i] = i < nq0 ? q0[i] : (real)0;
  int success = 0, it;
  for (it = 0;; it++) {
    real R[9], p[3], Ri[9], pi[3], err[6];
    ik_site_fk<NQ, NCH>(m, q, nqm, R, p, J);
    real dp[3] = {goal[0] - p[0], goal[1] - p[1], goal[2] - p[2]};
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) Ri[3 * r + cc] = R[r] * Rd[cc] + R[3 + r] * Rd[3 + cc] + R[6 + r] * Rd[6 + cc];
    mulmatT3(pi, R, dp);
    Log3 lg;
    ik_log3(Ri, lg);
    ik_log6(lg, pi, err);
    real en = 0;
    for (int k = 0; k < 6; k++) en += err[k] * err[k];
    if (r_sqrt(en) < (real)1e-4) { success = 1; break; }
    if (it >= 1000) break;
    real pinv[3], Jl[36], JJt[36], y[6];
    mulmatT3(pinv, Ri, pi);  // iMd^-1: rotation Ri^T (rotation vector -w, same angle), translation -Ri^T pi
    pinv[0] = -pinv[0]; pinv[1] = -pinv[1]; pinv[2] = -pinv[2];
    ik_jlog6(lg, (real)-1, pinv, Jl);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc < QM; cc++) {
        if (cc >= nqm) continue;
        real s = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) s += Jl[6 * r + k] * J[k * nqm + cc];
        Jn[r * nqm + cc] = -s;
      }
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc < 6; cc++) {
        real s = 0;
#pragma unroll
        for (int k = 0; k < QM; k++) if (k < nqm) s += Jn[r * nqm + k] * Jn[cc * nqm + k];
        JJt[6 * r + cc] = s;
      }
    for (int k = 0; k < 6; k++) { JJt[7 * k] += (real)1e-6; y[k] = err[k]; }
    ik_ldlt6(JJt, y);
#pragma unroll
    for (int cc = 0; cc < QM; cc++) {
      if (cc >= nqm) continue;
      real s = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += Jn[k * nqm + cc] * y[k];
      q[cc] += -s * (real)0.1;
    }
  }
  if (iters_out) *iters_out = it;
  if (success) {
#pragma unroll
    for (int k = 0; k < QM; k++) if (k < nqm) q_out[k] = q[k];
  }
  return success;
}
// dispatch on the shapes of the shipped arms (FR3 + fingers: 9 joints in the IK model, chain of 7 bodies; xArm7: 7 / 7)
RCSB_DEV int ik_solve(const RcsbModel& m, const real* pose7, const real* q0, int nq0, real* q_out, int* iters_out) {
  const int nqm = m.rb_ik_nq < IK_MAXQ ? m.rb_ik_nq : IK_MAXQ;
  int n = 0, canonical = 1;  // canonical: every chain body has a hinge / slide joint and the i-th one drives dof i
  for (int b = m.rb_site_body; b >= 0; b = m.b_parent[b]) n++;
  for (int b = m.rb_site_body, i = n - 1; b >= 0; b = m.b_parent[b], i--)
    if (m.b_jtype[b] == RCSB_JNT_FREE || m.b_dadr[b] != i) canonical = 0;
  if (!canonical) return ik_solve_t<0, 0>(m, pose7, q0, nq0, q_out, iters_out);
  if (nqm == 9 && n == 7) return ik_solve_t<9, 7>(m, pose7, q0, nq0, q_out, iters_out);
  if (nqm == 7 && n == 7) return ik_solve_t<7, 7>(m, pose7, q0, nq0, q_out, iters_out);
  return ik_solve_t<0, 0>(m, pose7, q0, nq0, q_out, iters_out);
}


// ------------------------------------------------------------------ 8 lanes per environment (small and medium batches)
// The thread-per-environment solver above leaves a B200 almost idle at a few thousand environments (one warp per SM, a
// serial chain of ~5 k FP64 instructions per CLIK iteration). Here 8 lanes share one environment, 4 environments per
// warp: lane g owns joint g of the chain (its local frame, its Jacobian column, its q); the world frames come from a
// 3-level inclusive scan of the affine transforms (width-8 shuffles), the 6 x 6 normal matrix J J^T from a butterfly
// reduction of the lanes' outer products, and the tiny error / Jlog6 / LDL^T parts are evaluated redundantly by the 8
// lanes (identical inputs, identical results), so nothing goes through memory. Canonical chains only (every chain body has
// a hinge / slide joint and the i-th one drives dof i, at most 8 of them): FR3 (7 of 9 IK dofs) and xArm7.
#ifndef RCSB_HOST_EMU
RCSB_DEV real g8_idx(real x, int src) { return __shfl_sync(0xffffffffu, x, src, 8); }
RCSB_DEV real g8_up(real x, int d) { return __shfl_up_sync(0xffffffffu, x, d, 8); }
RCSB_DEV real g8_xor(real x, int d) { return __shfl_xor_sync(0xffffffffu, x, d, 8); }
RCSB_DEV void ik_local_frame(const RcsbModel& m, int b, real q, real* Rl, real* tl) {  // as in ik_site_fk
  const real* Rb = m.b_rot[b];
  const int qa = m.b_qadr[b];
  const real qj = q - m.qpos0[qa];
  const real* u = m.b_jaxis[b];
  const real* jp = m.b_jpos[b];
  if (m.b_jtype[b] == RCSB_JNT_SLIDE) {
    real v[3];
    mulmat3(v, Rb, u);
    for (int i = 0; i < 9; i++) Rl[i] = Rb[i];
    tl[0] = m.b_pos[b][0] + v[0] * qj; tl[1] = m.b_pos[b][1] + v[1] * qj; tl[2] = m.b_pos[b][2] + v[2] * qj;
  } else {
    real sn, co;
    sincos(qj, &sn, &co);
    const real oc = 1 - co;
    real Rj[9] = {co + oc * u[0] * u[0], oc * u[0] * u[1] - sn * u[2], oc * u[0] * u[2] + sn * u[1],
                  oc * u[1] * u[0] + sn * u[2], co + oc * u[1] * u[1], oc * u[1] * u[2] - sn * u[0],
                  oc * u[2] * u[0] - sn * u[1], oc * u[2] * u[1] + sn * u[0], co + oc * u[2] * u[2]};
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) Rl[3 * r + k] = Rb[3 * r] * Rj[k] + Rb[3 * r + 1] * Rj[3 + k] + Rb[3 * r + 2] * Rj[6 + k];
    real w[3], v[3];
    mulmat3(w, Rj, jp);
    w[0] = jp[0] - w[0]; w[1] = jp[1] - w[1]; w[2] = jp[2] - w[2];
    mulmat3(v, Rb, w);
    tl[0] = m.b_pos[b][0] + v[0]; tl[1] = m.b_pos[b][1] + v[1]; tl[2] = m.b_pos[b][2] + v[2];
  }
}
// canonical chain length (0 when the chain is not canonical or longer than 8 bodies)
RCSB_DEV int ik_canonical_chain(const RcsbModel& m) {
  int n = 0;
  for (int b = m.rb_site_body; b >= 0; b = m.b_parent[b]) n++;
  if (n < 1 || n > 8) return 0;
  for (int b = m.rb_site_body, i = n - 1; b >= 0; b = m.b_parent[b], i--)
    if (m.b_jtype[b] == RCSB_JNT_FREE || m.b_dadr[b] != i || m.b_qadr[b] != i) return 0;
  return n;
}
// All 32 lanes of the warp call this together. g = lane & 7; q_g = this lane's joint value (start value in, solution out);
// valid = this group has an environment. Returns success (uniform within the group).
RCSB_DEV int ik_solve8(const RcsbModel& m, int nch, const real* pose7, real& q_g, int g, int valid, int* iters_out) {
  real inv_tcp[7], goal[7], Rd[9];
  pose_inverse(m.rb_tcp_offset, inv_tcp);
  pose_mul(pose7, inv_tcp, goal);
  Quat qg = {goal[3], goal[4], goal[5], goal[6]};
  q_to_mat(qg, Rd);
  int b = m.rb_site_body;  // lane g < nch owns the g-th body of the chain, counted from the root
  for (int i = nch - 1; i > g; i--) b = m.b_parent[b >= 0 ? b : 0];
  const bool link = g < nch;
  if (!link) b = m.rb_site_body;
  const int hinge = m.b_jtype[b] == RCSB_JNT_HINGE;
  int active = valid, success = 0, it = 0;
  for (;;) {
    // ---- local frame of every chain body, then the inclusive scan: lane g ends with the world frame of body g
    real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    if (link) ik_local_frame(m, b, q_g, R, t);
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      real Ru[9], tu[3];
#pragma unroll
      for (int i = 0; i < 9; i++) Ru[i] = g8_up(R[i], d);
#pragma unroll
      for (int i = 0; i < 3; i++) tu[i] = g8_up(t[i], d);
      if (g >= d) {
        real Rn[9], v[3];
        mulmat3(v, Ru, t);
        t[0] = tu[0] + v[0]; t[1] = tu[1] + v[1]; t[2] = tu[2] + v[2];
        for (int r = 0; r < 3; r++)
          for (int k = 0; k < 3; k++) Rn[3 * r + k] = Ru[3 * r] * R[k] + Ru[3 * r + 1] * R[3 + k] + Ru[3 * r + 2] * R[6 + k];
        for (int i = 0; i < 9; i++) R[i] = Rn[i];
      }
    }
    // joint axis and anchor in the world (a rotation about the joint axis leaves both in place)
    real axis[3], anchor[3], v[3];
    mulmat3(axis, R, m.b_jaxis[b]);
    mulmat3(v, R, m.b_jpos[b]);
    anchor[0] = t[0] + v[0]; anchor[1] = t[1] + v[1]; anchor[2] = t[2] + v[2];
    // site pose from the last chain lane, broadcast to the group
    real Rs[9], ps[3];
    {
      const real* Sl = m.rb_site_rot;
      mulmat3(v, R, m.rb_site_pos);
      ps[0] = t[0] + v[0]; ps[1] = t[1] + v[1]; ps[2] = t[2] + v[2];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) Rs[3 * r + k] = R[3 * r] * Sl[k] + R[3 * r + 1] * Sl[3 + k] + R[3 * r + 2] * Sl[6 + k];
#pragma unroll
      for (int i = 0; i < 9; i++) Rs[i] = g8_idx(Rs[i], nch - 1);
#pragma unroll
      for (int i = 0; i < 3; i++) ps[i] = g8_idx(ps[i], nch - 1);
    }
    // ---- error in the local frame (redundant in the 8 lanes)
    real Ri[9], pi[3], err[6], dp[3] = {goal[0] - ps[0], goal[1] - ps[1], goal[2] - ps[2]};
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) Ri[3 * r + cc] = Rs[r] * Rd[cc] + Rs[3 + r] * Rd[3 + cc] + Rs[6 + r] * Rd[6 + cc];
    mulmatT3(pi, Rs, dp);
    Log3 lg;
    ik_log3(Ri, lg);
    ik_log6(lg, pi, err);
    real en = 0;
    for (int k = 0; k < 6; k++) en += err[k] * err[k];
    if (active && r_sqrt(en) < (real)1e-4) { success = 1; active = 0; }
    if (active && it >= 1000) active = 0;
    if (!__any_sync(0xffffffffu, active)) break;
    // ---- this lane's Jacobian column in the LOCAL frame, mapped through -Jlog6(iMd^-1)
    real col[6] = {0, 0, 0, 0, 0, 0};
    if (link) {
      real lin[3], ang[3] = {0, 0, 0}, r3[3];
      if (hinge) {
        r3[0] = ps[0] - anchor[0]; r3[1] = ps[1] - anchor[1]; r3[2] = ps[2] - anchor[2];
        cross3(lin, axis, r3);
        copy3(ang, axis);
      } else {
        copy3(lin, axis);
      }
      mulmatT3(col, Rs, lin);
      mulmatT3(col + 3, Rs, ang);
    }
    real pinv[3], Jl[36], jn[6];
    mulmatT3(pinv, Ri, pi);  // iMd^-1: rotation Ri^T (rotation vector -w, same angle), translation -Ri^T pi
    pinv[0] = -pinv[0]; pinv[1] = -pinv[1]; pinv[2] = -pinv[2];
    ik_jlog6(lg, (real)-1, pinv, Jl);
#pragma unroll
    for (int r = 0; r < 6; r++) {
      real sacc = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) sacc += Jl[6 * r + k] * col[k];
      jn[r] = -sacc;
    }
    // ---- J J^T + damping: butterfly sum of the lanes' outer products (every lane ends with the same 21 sums)
    real JJt[36], y[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc <= r; cc++) {
        real p = jn[r] * jn[cc];
        p += g8_xor(p, 1);
        p += g8_xor(p, 2);
        p += g8_xor(p, 4);
        JJt[6 * r + cc] = p; JJt[6 * cc + r] = p;
      }
    for (int k = 0; k < 6; k++) { JJt[7 * k] += (real)1e-6; y[k] = err[k]; }
    ik_ldlt6(JJt, y);
    real dq = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) dq += jn[k] * y[k];
    if (active) { q_g += -dq * (real)0.1; it++; }
  }
  if (iters_out) *iters_out = it;
  return success;
}
#endif

// one environment per thread; apply != 0 restates SimRobot::set_cartesian_position (SimRobot.cpp:145-155)
RCSB_DEV void ik_env(const RcsbModel* sm, real* /*scratch*/, int /*lane*/, int env, const real* pose, const real* q0,
                     real* q_out, int* success, int* iters, int apply, real* sr, int* si) {
  const RcsbModel& m = *sm;
  const int nj = m.rb_njoints, nqm = m.rb_ik_nq < IK_MAXQ ? m.rb_ik_nq : IK_MAXQ;
  real q0l[RCSB_MAXJ], q[IK_MAXQ];
  real* row = sr ? sr + (size_t)env * m.lay.nsr : nullptr;
  for (int i = 0; i < nj; i++) q0l[i] = apply ? row[m.lay.o_q + m.rb_qadr[i]] : q0[(size_t)env * nj + i];
  int it = 0;
  int ok = ik_solve(m, pose + (size_t)env * 7, q0l, nj, q, &it);
  if (iters) iters[env] = it;
  if (success) success[env] = ok;
  if (q_out && ok) for (int i = 0; i < nqm; i++) q_out[(size_t)env * nqm + i] = q[i];
  if (apply) {
    int* irow = si + (size_t)env * RCSB_I_TAIL;
    irow[RCSB_I_IK_SUCCESS] = ok;
    if (ok) {  // set_joint_position(joint_vals)
      for (int i = 0; i < nj; i++) {
        row[m.lay.o_rcs + RCSB_S_TARGET + i] = q[i];
        row[m.lay.o_rcs + RCSB_S_PREV + i] = row[m.lay.o_q + m.rb_qadr[i]];
        row[m.lay.o_ctrl + m.rb_act[i]] = q[i];
      }
      irow[RCSB_I_MOVING] = 1;
      irow[RCSB_I_ARRIVED] = 0;
    }
  }
}

// ------------------------------------------------------------------ Cartesian actions of the Gym layer, one env per thread
// RelativeActionSpace.action (python/rcs/envs/base.py:490-578, RelativeTo.LAST_STEP) + RobotEnv.step's dedupe and
// dispatch (base.py:255-288) + SimRobot::set_cartesian_position (SimRobot.cpp:145-155). act is [N][6] xyzrpy
// (CARTESIAN_TRPY) or [N][7] xyz + quat xyzw (CARTESIAN_TQuat); relative != 0 applies the clipped offset to the current
// Cartesian position of the robot.
enum { RCSB_CART_TRPY = 0, RCSB_CART_TQUAT = 1 };
RCSB_DEV Quat q_from_rpy(real roll, real pitch, real yaw) {  // Rz(yaw) Ry(pitch) Rx(roll), include/rcs/Pose.h:37-43
  real cr = cos(roll / 2), sr = sin(roll / 2), cp = cos(pitch / 2), sp = sin(pitch / 2), cy = cos(yaw / 2), sy = sin(yaw / 2);
  Quat q = {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
  return q;
}
// Pose::limit_rotation_angle (Pose.cpp:184-192): Eigen angularDistance to identity and slerp from identity
RCSB_DEV Quat q_limit_angle(Quat q, real max_angle) {
  real vn = r_sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  real angle = 2 * atan2(vn, r_abs(q.w));
  if (!(angle > max_angle && max_angle >= 0)) return q;
  real t = max_angle / angle;
  // Eigen's slerp(t, other) from the identity: d = dot = q.w, shortest arc through |d|
  real d = q.w, ad = r_abs(d), s0, s1;
  if (ad >= (real)1 - (real)2.220446049250313e-16) { s0 = 1 - t; s1 = t; }
  else {
    real theta = acos(ad), st = sin(theta);
    s0 = sin((1 - t) * theta) / st;
    s1 = sin(t * theta) / st;
  }
  if (d < 0) s1 = -s1;
  Quat r = {s1 * q.x, s1 * q.y, s1 * q.z, s0 + s1 * q.w};
  return r;
}
RCSB_DEV void robot_cartesian_from_row(const RcsbModel& m, const real* row, real* pose7) {  // as robot_cartesian_position
  const real* sp = row + m.lay.o_rcs + RCSB_S_SITEPOS;
  real site[7], base[7], binv[7], t[7];
  Quat qs = q_norm(q_from_mat(sp + 3));
  site[0] = sp[0]; site[1] = sp[1]; site[2] = sp[2];
  site[3] = qs.x; site[4] = qs.y; site[5] = qs.z; site[6] = qs.w;
  Quat qb = {m.rb_base_quat[1], m.rb_base_quat[2], m.rb_base_quat[3], m.rb_base_quat[0]};
  qb = q_norm(qb);
  base[0] = m.rb_base_pos[0]; base[1] = m.rb_base_pos[1]; base[2] = m.rb_base_pos[2];
  base[3] = qb.x; base[4] = qb.y; base[5] = qb.z; base[6] = qb.w;
  pose_inverse(base, binv);
  pose_mul(binv, site, t);
  pose_mul(t, m.rb_tcp_offset, pose7);
}
// Optional per-environment state of RelativeTo.CONFIGURED_ORIGIN (base.py:443-467, 490-578): the origin pose set at
// reset and the last clipped offset. Device arrays owned by the host env object; null for the other modes.
struct CartOrigin { const real* origin; real* last; int* have_last; };
enum { RCSB_REL_ABS = 0, RCSB_REL_LAST_STEP = 1, RCSB_REL_CONFIGURED_ORIGIN = 2 };
RCSB_DEV void pose_limit(const real* in, real max_trans, real max_rot, real* out) {  // limit_translation_length . limit_rotation_angle
  const real tn = r_sqrt(in[0] * in[0] + in[1] * in[1] + in[2] * in[2]);
  const real sc = (tn > max_trans && max_trans >= 0) ? max_trans / tn : (real)1;
  Quat q = q_norm(q_limit_angle(Quat{in[3], in[4], in[5], in[6]}, max_rot));
  out[0] = in[0] * sc; out[1] = in[1] * sc; out[2] = in[2] * sc;
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}
// The Gym layer of one Cartesian action: RelativeActionSpace.action, RobotEnv.step's dedupe. Every calling lane computes
// the same values; only `writer` lanes commit state (previous action, CONFIGURED_ORIGIN bookkeeping), after `sync`
// (a warp barrier in the lanes-per-environment kernel) so that no lane still reads what another one overwrites.
// Returns 1 when a new command has to be executed; pose7 receives its absolute goal pose.
template <bool SYNC>
RCSB_DEV int cart_goal(const RcsbModel& m, int env, const real* act, int kind, int relative, real max_trans, real max_rot,
                       real* row, int* irow, CartOrigin co, int writer, real* pose7) {
  const int na = kind == RCSB_CART_TRPY ? 6 : 7;
  const real* a = act + (size_t)env * na;
  real absact[7], newlast[7];
  for (int i = 0; i < na; i++) absact[i] = a[i];
  if (relative != RCSB_REL_ABS) {
    real origin[7], off[7], in[7], prod[7];
    Quat qo = kind == RCSB_CART_TRPY ? q_from_rpy(a[3], a[4], a[5]) : q_norm(Quat{a[3], a[4], a[5], a[6]});
    in[0] = a[0]; in[1] = a[1]; in[2] = a[2]; in[3] = qo.x; in[4] = qo.y; in[5] = qo.z; in[6] = qo.w;
    if (relative == RCSB_REL_LAST_STEP) {
      robot_cartesian_from_row(m, row, origin);
      pose_limit(in, max_trans, max_rot, off);
    } else {  // CONFIGURED_ORIGIN: the offset itself may move by at most (max_trans, max_rot) per step
      for (int i = 0; i < 7; i++) origin[i] = co.origin[(size_t)env * 7 + i];
      if (!co.have_last[env]) pose_limit(in, max_trans, max_rot, off);
      else {
        real last[7], linv[7], diff[7], cd[7];
        for (int i = 0; i < 7; i++) last[i] = co.last[(size_t)env * 7 + i];
        pose_inverse(last, linv);
        pose_mul(in, linv, diff);
        pose_limit(diff, max_trans, max_rot, cd);
        pose_mul(cd, last, off);
      }
      for (int i = 0; i < 7; i++) newlast[i] = off[i];
    }
    pose_mul(off, origin, prod);  // only its rotation is used: the translations add (base.py:519-522)
    const real lo[3] = {(real)-0.855, (real)-0.855, (real)0}, hi[3] = {(real)0.855, (real)0.855, (real)1.188};
    for (int k = 0; k < 3; k++) {
      real v = origin[k] + off[k];
      absact[k] = v < lo[k] ? lo[k] : (v > hi[k] ? hi[k] : v);
    }
    if (kind == RCSB_CART_TRPY) {
      real x6[6];
      pose_xyzrpy(prod, x6);
      absact[3] = x6[3]; absact[4] = x6[4]; absact[5] = x6[5];
    } else {
      absact[3] = prod[3]; absact[4] = prod[4]; absact[5] = prod[5]; absact[6] = prod[6];
    }
  }
  // RobotEnv.step: skip the command when the action equals the previous one (atol 1e-3, rtol 0)
  real* prev = row + m.lay.o_rcs + RCSB_S_PREVACT;
  int changed = !irow[RCSB_I_HAVE_PREV_ACTION];
  for (int i = 0; i < na; i++)
    if (!(r_abs(absact[i] - prev[i]) <= (real)1e-3)) changed = 1;
#ifndef RCSB_HOST_EMU
  if (SYNC) __syncwarp();
#endif
  if (writer) {
    for (int i = 0; i < na; i++) prev[i] = absact[i];
    irow[RCSB_I_HAVE_PREV_ACTION] = 1;
    if (relative == RCSB_REL_CONFIGURED_ORIGIN) {
      for (int i = 0; i < 7; i++) co.last[(size_t)env * 7 + i] = newlast[i];
      co.have_last[env] = 1;
    }
  }
  Quat qt = kind == RCSB_CART_TRPY ? q_from_rpy(absact[3], absact[4], absact[5]) : Quat{absact[3], absact[4], absact[5], absact[6]};
  qt = q_norm(qt);
  pose7[0] = absact[0]; pose7[1] = absact[1]; pose7[2] = absact[2];
  pose7[3] = qt.x; pose7[4] = qt.y; pose7[5] = qt.z; pose7[6] = qt.w;
  return changed;
}
RCSB_DEV void cart_action_env(const RcsbModel* sm, int env, const real* act, int kind, int relative, real max_trans, real max_rot,
                              real* sr, int* si, CartOrigin co) {
  const RcsbModel& m = *sm;
  real* row = sr + (size_t)env * m.lay.nsr;
  int* irow = si + (size_t)env * RCSB_I_TAIL;
  real pose7[7];
  if (!cart_goal<false>(m, env, act, kind, relative, max_trans, max_rot, row, irow, co, 1, pose7)) return;
  ik_env(sm, nullptr, 0, 0, pose7, nullptr, nullptr, nullptr, nullptr, 1, row, irow);  // env 0 of the row pointers
}
#ifndef RCSB_HOST_EMU
// ---- the 8-lanes-per-environment entry points: Pin::inverse / set_cartesian_position and the Cartesian Gym action
RCSB_DEV void ik_commit8(const RcsbModel& m, int nch, int g, int valid, int ok, real q_g, real* row, int* irow) {
  if (!valid) return;
  if (g == 0) irow[RCSB_I_IK_SUCCESS] = ok;
  if (ok) {  // set_joint_position(joint_vals)
    if (g < m.rb_njoints) {
      const real qv = g < nch ? q_g : (real)0;
      row[m.lay.o_rcs + RCSB_S_TARGET + g] = qv;
      row[m.lay.o_rcs + RCSB_S_PREV + g] = row[m.lay.o_q + m.rb_qadr[g]];
      row[m.lay.o_ctrl + m.rb_act[g]] = qv;
    }
    if (g == 0) { irow[RCSB_I_MOVING] = 1; irow[RCSB_I_ARRIVED] = 0; }
  }
}
RCSB_DEV void ik_env8(const RcsbModel* sm, int nch, int env, int valid, int g, const real* pose, const real* q0, real* q_out,
                      int* success, int* iters, int apply, real* sr, int* si) {
  const RcsbModel& m = *sm;
  const int nj = m.rb_njoints, nqm = m.rb_ik_nq < IK_MAXQ ? m.rb_ik_nq : IK_MAXQ;
  real* row = sr ? sr + (size_t)env * m.lay.nsr : nullptr;
  real q_g = 0;
  if (g < nch && g < nj) q_g = apply ? row[m.lay.o_q + m.rb_qadr[g]] : q0[(size_t)env * nj + g];
  real p7[7];
  for (int i = 0; i < 7; i++) p7[i] = pose[(size_t)env * 7 + i];
  int it = 0;
  const int ok = ik_solve8(m, nch, p7, q_g, g, valid, &it);
  if (!valid) return;
  if (g == 0) {
    if (iters) iters[env] = it;
    if (success) success[env] = ok;
  }
  if (q_out && ok) {
    if (g < nch) q_out[(size_t)env * nqm + g] = q_g;
    if (g == 0) for (int i = nch; i < nqm; i++) q_out[(size_t)env * nqm + i] = 0;
  }
  if (apply) ik_commit8(m, nch, g, valid, ok, q_g, row, si + (size_t)env * RCSB_I_TAIL);
}
RCSB_DEV void cart_action_env8(const RcsbModel* sm, int nch, int env, int valid, int g, const real* act, int kind, int relative,
                               real max_trans, real max_rot, real* sr, int* si, CartOrigin co) {
  const RcsbModel& m = *sm;
  real* row = sr + (size_t)env * m.lay.nsr;
  int* irow = si + (size_t)env * RCSB_I_TAIL;
  real pose7[7];
  const int changed = cart_goal<true>(m, env, act, kind, relative, max_trans, max_rot, row, irow, co, valid && g == 0, pose7);
  const int run = valid && changed;
  real q_g = 0;
  if (g < nch && g < m.rb_njoints) q_g = row[m.lay.o_q + m.rb_qadr[g]];
  const int ok = ik_solve8(m, nch, pose7, q_g, g, run, nullptr);
  ik_commit8(m, nch, g, run, ok, q_g, row, irow);
}
#endif
