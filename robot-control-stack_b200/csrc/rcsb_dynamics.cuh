// Per-environment smooth dynamics and collision detection, one environment per warp.
//
// B200-native re-design of the arithmetic the reference reaches through mj_step1
// (/root/reference/src/sim/sim.cpp:110): forward kinematics, tree COM frames, composite rigid
// body mass matrix + Cholesky, velocity-stage bias forces (RNE with zero acceleration), passive
// damping and gravity compensation, broad + narrow phase collision with exact temporal coherence.
// Stages are either parallel-fors over independent work items (body | dof | matrix entry | geom pair)
// or short root-to-leaf / leaf-to-root passes along the dof chain with one lane per vector component
// (MuJoCo's "com frame" spatial vectors keep those passes to one fused multiply-add per lane and link).
// This file is a variant body: rcsb_variant.cuh includes it once per kernel shape (MD() / LAY macros).

// ------------------------------------------------------------------ dense Cholesky / triangular solves
// A (n x n, row-major, shared memory) -> strictly-lower part holds L, dinv[j] = 1 / L[j][j]; A's diagonal is left
// untouched. Device version: lane i owns row i in registers, pivots and multipliers travel by warp shuffles
// (right-looking, no shared-memory round trips, no barriers inside the factorisation). The host-emulation build
// keeps the plain column version (one lane).
#ifdef RCSB_HOST_EMU
RCSB_DEV_NOINLINE void chol_factor(const Ctx& c, real* A, real* dinv, int n, int blocks = 0) {
  for (int j = 0; j < n; j++) {
    real d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (d < RCSB_MINVAL) d = RCSB_MINVAL;
    real inv = (real)1 / r_sqrt(d);
    for (int i = j + 1; i < n; i++) {
      real t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t * inv;
    }
    dinv[j] = inv;
  }
}
// x <- (L L^T)^{-1} x ; y is scratch of length n
RCSB_DEV_NOINLINE void chol_solve(const Ctx& c, const real* L, const real* dinv, int n, real* x, real* y, int lo = 0, int hi = 1 << 30) {
  for (int k = 0; k < n; k++) {
    real xk = x[k] * dinv[k];
    for (int i = k + 1; i < n; i++) x[i] -= L[i * n + k] * xk;
    y[k] = xk;
  }
  for (int k = n - 1; k >= 0; k--) {
    real yk = y[k] * dinv[k];
    for (int i = 0; i < k; i++) y[i] -= L[k * n + i] * yk;
    x[k] = yk;
  }
}
// factor A in place and solve A x = b for x (in place); when A1 is given, factor it too (same size, no solve)
RCSB_DEV_NOINLINE void chol_factor_solve(const Ctx& c, real* A, real* dinv, int n, real* x, real* y, real* A1, real* dinv1,
                                         int blocks = 0) {
  chol_factor(c, A, dinv, n);
  chol_solve(c, A, dinv, n, x, y);
  if (A1) chol_factor(c, A1, dinv1, n);
}
// x <- (L L^T)^{-1} x for a factor that is block diagonal by kinematic tree (dense substitution handles it as is)
RCSB_DEV_NOINLINE void chol_solve_blocks(const Ctx& c, const real* L, const real* dinv, int n, real* x, real* y) {
  chol_solve(c, L, dinv, n, x, y);
}
// X[r] <- (L L^T)^{-1} X[r] for nrhs right-hand sides (rows of X, row stride n)
RCSB_DEV_NOINLINE void chol_solve_multi(const Ctx& c, const real* L, const real* dinv, int n, real* X, int nrhs, real* y) {
  for (int r = 0; r < nrhs; r++) chol_solve(c, L, dinv, n, X + r * n, y);
}
#else
// Half-warp formulation for N <= 16: lanes 0..15 hold system 0, lanes 16..31 system 1 (when present); lane li owns row
// li in registers, pivots and multipliers travel by width-16 shuffles (right-looking, no shared-memory round trips and
// no barriers inside the factorisation). With N a compile-time constant every loop unrolls.
RCSB_DEV real half_bcast(real x, int src) { return __shfl_sync(0xffffffffu, x, src, 16); }
template <int N>
RCSB_DEV void chol_rows_factor(int li, real (&a)[N]) {
#pragma unroll
  for (int j = 0; j < N; j++) {
    real d = half_bcast(a[j], j);
    if (d < RCSB_MINVAL) d = RCSB_MINVAL;
    real inv = rsqrt(d);
    real lij = a[j] * inv;  // lane > j: L[lane][j]; lane == j: L[j][j]
    a[j] = li == j ? inv : lij;
#pragma unroll
    for (int k = j + 1; k < N; k++) {
      real lkj = half_bcast(lij, k);
      if (li >= k) a[k] -= lij * lkj;
    }
  }
}
// x <- (L L^T)^{-1} x with row li of L in a[] (a[li] = 1 / L[li][li]) and column li of L in col[]
template <int N>
RCSB_DEV real chol_rows_solve(int li, const real (&a)[N], const real (&col)[N], real xi, int lo = 0, int hi = N) {
  // [lo, hi): a diagonal block of L outside of which the right-hand side is zero (dofs of one kinematic tree when L
  // factors M): the other steps would only move zeros around, so they are skipped (uniform branch)
  real di = (real)0;
#pragma unroll
  for (int k = 0; k < N; k++) if (k == li) di = a[k];
#pragma unroll
  for (int j = 0; j < N; j++) {  // forward: L z = x
    if (j < lo || j >= hi) continue;
    real zj = half_bcast(xi * di, j);
    if (li == j) xi = zj;
    else if (li > j) xi -= a[j] * zj;
  }
#pragma unroll
  for (int j = N - 1; j >= 0; j--) {  // backward: L^T w = z
    if (j < lo || j >= hi) continue;
    real wj = half_bcast(xi * di, j);
    if (li == j) xi = wj;
    else if (li < j) xi -= col[j] * wj;
  }
  return xi;
}
template <int N>
RCSB_DEV void chol_n(const Ctx& c, real* A0, real* dinv0, real* x, real* A1, real* dinv1) {
  const int li = c.lane & 15, half = c.lane >> 4;
  real* A = half ? A1 : A0;
  real* dinv = half ? dinv1 : dinv0;
  const bool on = A != nullptr && li < N;
  real a[N], col[N];
#pragma unroll
  for (int k = 0; k < N; k++) a[k] = (on && k <= li) ? A[li * N + k] : (k == li ? (real)1 : (real)0);
  chol_rows_factor<N>(li, a);
  if (on) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (k < li) A[li * N + k] = a[k];
      else if (k == li) dinv[li] = a[k];
    }
  }
  if (x == nullptr) return;
  RCSB_SYNC();
#pragma unroll
  for (int j = 0; j < N; j++) col[j] = (!half && li < N && j > li) ? A0[j * N + li] : (real)0;
  real xi = (!half && li < N) ? x[li] : (real)0;
  xi = chol_rows_solve<N>(li, a, col, xi);
  if (!half && li < N) x[li] = xi;
}
template <int N>
RCSB_DEV void chol_solve_n(const Ctx& c, const real* L, const real* dinv, real* x, int lo, int hi) {
  const int li = c.lane & 15, half = c.lane >> 4;
  const bool on = !half && li < N && li >= lo && li < hi;
  real a[N], col[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    a[k] = (on && k < li) ? L[li * N + k] : (real)0;
    col[k] = (on && k > li) ? L[k * N + li] : (real)0;
  }
  real di = on ? dinv[li] : (real)1;
#pragma unroll
  for (int k = 0; k < N; k++) if (k == li) a[k] = di;
  real xi = on ? x[li] : (real)0;
  xi = chol_rows_solve<N>(li, a, col, xi, lo, hi);
  if (on) x[li] = xi;
}
// ---- matrices that are block diagonal by kinematic tree (M always; the constraint Hessian while no constraint row
// couples two trees; the implicit integrator's matrix): two blocks are factored / solved at a time, one per half-warp,
// with the rows-in-registers scheme above. N bounds the block size at compile time; a block smaller than N is padded
// with identity rows. Off-block entries of the storage are left as they are (zeros): the result is the Cholesky factor
// of the whole matrix, so the dense triangular solves stay valid on it.
template <int N>
RCSB_DEV void chol_pair(const Ctx& c, real* A, real* dinv, int n, int lo0, int n0, int lo1, int n1, real* x) {
  const int li = c.lane & 15, half = c.lane >> 4;
  const int lo = half ? lo1 : lo0, nb = half ? n1 : n0;
  const bool on = li < nb;
  real a[N], col[N];
#pragma unroll
  for (int k = 0; k < N; k++) a[k] = (on && k <= li) ? A[(lo + li) * n + lo + k] : (k == li ? (real)1 : (real)0);
  chol_rows_factor<N>(li, a);
  if (on) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (k < li) A[(lo + li) * n + lo + k] = a[k];
      else if (k == li) dinv[lo + li] = a[k];
    }
  }
  if (x == nullptr) return;
  RCSB_SYNC();
#pragma unroll
  for (int j = 0; j < N; j++) col[j] = (on && j > li && j < nb) ? A[(lo + j) * n + lo + li] : (real)0;
  real xi = on ? x[lo + li] : (real)0;
  xi = chol_rows_solve<N>(li, a, col, xi);
  if (on) x[lo + li] = xi;
}
template <int N>
RCSB_DEV void chol_pair_solve(const Ctx& c, const real* L, const real* dinv, int n, int lo0, int n0, int lo1, int n1, real* x) {
  const int li = c.lane & 15, half = c.lane >> 4;
  const int lo = half ? lo1 : lo0, nb = half ? n1 : n0;
  const bool on = li < nb;
  real a[N], col[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    a[k] = (on && k < li) ? L[(lo + li) * n + lo + k] : (real)0;
    col[k] = (on && k > li && k < nb) ? L[(lo + k) * n + lo + li] : (real)0;
  }
  const real di = on ? dinv[lo + li] : (real)1;
#pragma unroll
  for (int k = 0; k < N; k++) if (k == li) a[k] = di;
  real xi = on ? x[lo + li] : (real)0;
  xi = chol_rows_solve<N>(li, a, col, xi);
  if (on) x[lo + li] = xi;
}
// tree blocks of the dof range [0, n): returns the count (0 when the trees' dofs are not contiguous blocks <= 9 wide)
RCSB_DEV int tree_blocks(const RcsbModel& m, int n, int* lo, int* sz) {
  int nt = 0;
  for (int j = 0; j < n;) {
    const int hi = m.d_tree_hi[j];
    if (m.d_tree_lo[j] != j || hi <= j || hi - j > 9 || nt >= RCSB_MAXROOT) return 0;
    lo[nt] = j; sz[nt] = hi - j; nt++;
    j = hi;
  }
  return nt;
}
// factor (and optionally solve x in place) block by block; returns 0 when the model has no usable block structure
RCSB_DEV int chol_blocks(const Ctx& c, real* A, real* dinv, int n, real* x) {
  const RcsbModel& m = CMODEL(c);
  if (MD(nroot) < 2) return 0;
  int lo[RCSB_MAXROOT], sz[RCSB_MAXROOT];
  const int nt = tree_blocks(m, n, lo, sz);
  if (nt < 2) return 0;
  RCSB_SYNC();
  for (int t = 0; t < nt; t += 2) {
    const int n0 = sz[t], n1 = t + 1 < nt ? sz[t + 1] : 0, l1 = t + 1 < nt ? lo[t + 1] : 0;
    const int mx = n0 > n1 ? n0 : n1;
    if (mx <= 7) chol_pair<7>(c, A, dinv, n, lo[t], n0, l1, n1, x);
    else chol_pair<9>(c, A, dinv, n, lo[t], n0, l1, n1, x);
  }
  RCSB_SYNC();
  return 1;
}
RCSB_DEV int chol_blocks_solve(const Ctx& c, const real* L, const real* dinv, int n, real* x) {
  const RcsbModel& m = CMODEL(c);
  if (MD(nroot) < 2) return 0;
  int lo[RCSB_MAXROOT], sz[RCSB_MAXROOT];
  const int nt = tree_blocks(m, n, lo, sz);
  if (nt < 2) return 0;
  RCSB_SYNC();
  for (int t = 0; t < nt; t += 2) {
    const int n0 = sz[t], n1 = t + 1 < nt ? sz[t + 1] : 0, l1 = t + 1 < nt ? lo[t + 1] : 0;
    const int mx = n0 > n1 ? n0 : n1;
    if (mx <= 7) chol_pair_solve<7>(c, L, dinv, n, lo[t], n0, l1, n1, x);
    else chol_pair_solve<9>(c, L, dinv, n, lo[t], n0, l1, n1, x);
  }
  RCSB_SYNC();
  return 1;
}
RCSB_DEV_NOINLINE void chol_factor(const Ctx& c, real* A, real* dinv, int n, int blocks = 0) {
  __builtin_assume(__isShared(A));
  __builtin_assume(__isShared(dinv));
  if (blocks && chol_blocks(c, A, dinv, n, nullptr)) return;
  RCSB_SYNC();
  switch (n) {  // dof counts of the supported scenes: xArm7 (7), FR3 + fingers (9), xArm7 + brick (13), FR3 + fingers + cube (15)
    case 7: chol_n<7>(c, A, dinv, nullptr, nullptr, nullptr); break;
    case 9: chol_n<9>(c, A, dinv, nullptr, nullptr, nullptr); break;
    case 13: chol_n<13>(c, A, dinv, nullptr, nullptr, nullptr); break;
    case 15: chol_n<15>(c, A, dinv, nullptr, nullptr, nullptr); break;
    default:  // any other size: column version on shared memory
      for (int j = 0; j < n; j++) {
        RCSB_SYNC();
        real d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (d < RCSB_MINVAL) d = RCSB_MINVAL;
        real inv = rsqrt(d);
        PFOR(ii, n - j - 1) {
          int i = j + 1 + ii;
          real t = A[i * n + j];
          for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
          A[i * n + j] = t * inv;
        }
        if (c.lane == 0) dinv[j] = inv;
      }
  }
  RCSB_SYNC();
}
// x <- (L L^T)^{-1} x ; y is unused scratch (kept for the common signature). [lo, hi): see chol_rows_solve.
RCSB_DEV_NOINLINE void chol_solve(const Ctx& c, const real* L, const real* dinv, int n, real* x, real* y, int lo = 0, int hi = 1 << 30) {
  const int lane = c.lane;
  if (hi > n) hi = n;
  __builtin_assume(__isShared(L));
  __builtin_assume(__isShared(dinv));
  __builtin_assume(__isShared(x));
  RCSB_SYNC();
  switch (n) {
    case 7: chol_solve_n<7>(c, L, dinv, x, lo, hi); break;
    case 9: chol_solve_n<9>(c, L, dinv, x, lo, hi); break;
    case 13: chol_solve_n<13>(c, L, dinv, x, lo, hi); break;
    case 15: chol_solve_n<15>(c, L, dinv, x, lo, hi); break;
    default: {
      real xi = lane < n ? x[lane] : (real)0;
      real di = lane < n ? dinv[lane] : (real)0;
      for (int j = 0; j < n; j++) {  // forward: L z = x
        real zj = warp_bcast(xi * di, j);
        if (lane == j) xi = zj;
        else if (lane > j && lane < n) xi -= L[lane * n + j] * zj;
      }
      for (int j = n - 1; j >= 0; j--) {  // backward: L^T w = z
        real wj = warp_bcast(xi * di, j);
        if (lane == j) xi = wj;
        else if (lane < j) xi -= L[j * n + lane] * wj;
      }
      if (lane < n) x[lane] = xi;
    }
  }
  RCSB_SYNC();
}
// factor A in place and solve A x = b for x (in place); when A1 is given, factor it too in the other half-warp
RCSB_DEV_NOINLINE void chol_factor_solve(const Ctx& c, real* A, real* dinv, int n, real* x, real* y, real* A1, real* dinv1,
                                         int blocks = 0) {
  __builtin_assume(__isShared(A));
  __builtin_assume(__isShared(dinv));
  __builtin_assume(__isShared(x));
  if (blocks && chol_blocks(c, A, dinv, n, x)) {  // A1 (the integrator's matrix) is block diagonal whenever A is
    if (A1) {
      __builtin_assume(__isShared(A1));
      __builtin_assume(__isShared(dinv1));
      chol_blocks(c, A1, dinv1, n, nullptr);
    }
    return;
  }
  RCSB_SYNC();
  switch (n) {
    case 7: chol_n<7>(c, A, dinv, x, A1, dinv1); break;
    case 9: chol_n<9>(c, A, dinv, x, A1, dinv1); break;
    case 13: chol_n<13>(c, A, dinv, x, A1, dinv1); break;
    case 15: chol_n<15>(c, A, dinv, x, A1, dinv1); break;
    default:
      chol_factor(c, A, dinv, n);
      chol_solve(c, A, dinv, n, x, y);
      if (A1) chol_factor(c, A1, dinv1, n);
      return;
  }
  RCSB_SYNC();
}
// x <- (L L^T)^{-1} x for a factor that is block diagonal by kinematic tree: both half-warps solve a block each
RCSB_DEV_NOINLINE void chol_solve_blocks(const Ctx& c, const real* L, const real* dinv, int n, real* x, real* y) {
  __builtin_assume(__isShared(L));
  __builtin_assume(__isShared(dinv));
  __builtin_assume(__isShared(x));
  if (chol_blocks_solve(c, L, dinv, n, x)) return;
  chol_solve(c, L, dinv, n, x, y);
}
// X[r] <- (L L^T)^{-1} X[r] for nrhs right-hand sides (rows of X, row stride n), one right-hand side per lane: serial
// substitution restricted to the kinematic trees the row touches (L is block diagonal by tree). The noslip pass needs
// M^-1 J^T for every friction row; solving them side by side replaces one warp-wide solve per row.
RCSB_DEV_NOINLINE void chol_solve_multi(const Ctx& c, const real* L, const real* dinv, int n, real* X, int nrhs, real* y) {
  const RcsbModel& m = CMODEL(c);
  __builtin_assume(__isShared(L));
  __builtin_assume(__isShared(dinv));
  __builtin_assume(__isShared(X));
  RCSB_SYNC();
  PFOR(r, nrhs) {
    real* x = X + r * n;
    int lo = n, hi = 0;
    for (int k = 0; k < n; k++)
      if (x[k] != 0) {
        lo = m.d_tree_lo[k] < lo ? m.d_tree_lo[k] : lo;
        hi = m.d_tree_hi[k] > hi ? m.d_tree_hi[k] : hi;
      }
    for (int j = lo; j < hi; j++) {  // forward: L z = x
      const real z = x[j] * dinv[j];
      x[j] = z;
      for (int i = j + 1; i < hi; i++) x[i] -= L[i * n + j] * z;
    }
    for (int j = hi - 1; j >= lo; j--) {  // backward: L^T w = z
      const real w = x[j] * dinv[j];
      x[j] = w;
      for (int i = lo; i < j; i++) x[i] -= L[j * n + i] * w;
    }
  }
  RCSB_SYNC();
}
#endif

// ------------------------------------------------------------------ forward kinematics
// Rotation-matrix form: every body's local transform (static offset composed with its joint motion) is built in
// parallel, one lane per body; the chain is then walked with 12 lanes per body (9 matrix entries + 3 position
// entries, each a 3-term dot product), so the serial depth per body is a handful of FMAs instead of a lane-serial
// quaternion chain.
RCSB_DEV void st_kinematics(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  real* q = WR(q);
  real* L4 = WR(bquat);  // scratch: every body's local frame [R | t] as a 3 x 4 row-major block (o_bquat holds 12*nb)
  PFOR(b, MD(nb)) {
    real R[9], t[3];
    if (m.b_jtype[b] == RCSB_JNT_FREE) {
      real* qq = q + m.b_qadr[b];
      quat_normalize(qq + 3);
      quat_to_mat(R, qq + 3);
      copy3(t, qq);
    } else {
      const real* Rb = m.b_rot[b];
      const real* u = m.b_jaxis[b];
      const real* jp = m.b_jpos[b];
      real qq = q[m.b_qadr[b]] - m.qpos0[m.b_qadr[b]];
      if (m.b_jtype[b] == RCSB_JNT_SLIDE) {
        real v[3];
        mulmat3(v, Rb, u);
        for (int i = 0; i < 9; i++) R[i] = Rb[i];
        t[0] = m.b_pos[b][0] + v[0] * qq; t[1] = m.b_pos[b][1] + v[1] * qq; t[2] = m.b_pos[b][2] + v[2] * qq;
      } else {
        real s, co;
        sincos(qq, &s, &co);
        const real oc = 1 - co;  // Rodrigues rotation about the joint axis
        real Rj[9] = {co + oc * u[0] * u[0], oc * u[0] * u[1] - s * u[2], oc * u[0] * u[2] + s * u[1],
                      oc * u[1] * u[0] + s * u[2], co + oc * u[1] * u[1], oc * u[1] * u[2] - s * u[0],
                      oc * u[2] * u[0] - s * u[1], oc * u[2] * u[1] + s * u[0], co + oc * u[2] * u[2]};
        for (int r = 0; r < 3; r++)
          for (int k = 0; k < 3; k++) R[3 * r + k] = Rb[3 * r] * Rj[k] + Rb[3 * r + 1] * Rj[3 + k] + Rb[3 * r + 2] * Rj[6 + k];
        // the body origin moves when the joint anchor is off-origin: t = b_pos + Rb*(jp - Rj*jp)
        real w[3], v[3];
        mulmat3(w, Rj, jp);
        w[0] = jp[0] - w[0]; w[1] = jp[1] - w[1]; w[2] = jp[2] - w[2];
        mulmat3(v, Rb, w);
        t[0] = m.b_pos[b][0] + v[0]; t[1] = m.b_pos[b][1] + v[1]; t[2] = m.b_pos[b][2] + v[2];
      }
    }
    real* o = L4 + 12 * b;
    for (int r = 0; r < 3; r++) { o[4 * r] = R[3 * r]; o[4 * r + 1] = R[3 * r + 1]; o[4 * r + 2] = R[3 * r + 2]; o[4 * r + 3] = t[r]; }
  }
  RCSB_SYNC();
  // chain walk, 12 lanes per body: entry (r, k) of [R | p]_world = [R_parent | p_parent] * [R | t]_local
  for (int b = 0; b < MD(nb); b++) {
    const int p = m.b_parent[b];
    PFOR1(e, 12) {
      const int r = e >> 2, k = e & 3;
      const real* Lc = L4 + 12 * b + k;
      real val;
      if (p < 0) val = Lc[4 * r];
      else {
        const real* Rp = WR(bmat) + 9 * p + 3 * r;
        val = k == 3 ? WR(bpos)[3 * p + r] : (real)0;
        val += Rp[0] * Lc[0];
        val += Rp[1] * Lc[4];
        val += Rp[2] * Lc[8];
      }
      real* dst = k == 3 ? WR(bpos) + 3 * b + r : WR(bmat) + 9 * b + 3 * r + k;
      *dst = val;
    }
    RCSB_SYNC();
  }
}

// world pose of collidable geom g
RCSB_DEV void geom_frame(const Ctx& c, int g, real* pos, real* mat) {
  const RcsbModel& m = CMODEL(c);
  int b = m.g_body[g];
  const real* Rl = m.g_rot[g];
  if (b < 0) {
    copy3(pos, m.g_pos[g]);
    for (int i = 0; i < 9; i++) mat[i] = Rl[i];
  } else {
    const real* R = WR(bmat) + 9 * b;
    const real* p = WR(bpos) + 3 * b;
    real v[3];
    mulmat3(v, R, m.g_pos[g]);
    pos[0] = p[0] + v[0]; pos[1] = p[1] + v[1]; pos[2] = p[2] + v[2];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) mat[3 * r + k] = R[3 * r] * Rl[k] + R[3 * r + 1] * Rl[3 + k] + R[3 * r + 2] * Rl[6 + k];
  }
}

// ------------------------------------------------------------------ COM frames, inertias, motion axes, geom centres
// world position of the body's centre of mass
RCSB_DEV void body_com(const Ctx& c, int b, real* o) {
  const RcsbModel& m = CMODEL(c);
  real v[3];
  const real* p = WR(bpos) + 3 * b;
  mulmat3(v, WR(bmat) + 9 * b, m.b_ipos[b]);
  o[0] = p[0] + v[0]; o[1] = p[1] + v[1]; o[2] = p[2] + v[2];
}
RCSB_DEV void st_com(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  real* mc = WR(crb);  // scratch (crb is written later): mass-weighted body COMs
  PFOR(b, MD(nb)) {
    real o[3];
    body_com(c, b, o);
    mc[3 * b] = m.b_mass[b] * o[0]; mc[3 * b + 1] = m.b_mass[b] * o[1]; mc[3 * b + 2] = m.b_mass[b] * o[2];
  }
  RCSB_SYNC();
  PFOR(e, 3 * MD(nroot)) {
    const int r = e / 3, k = e - 3 * r;
    real s = 0;
    for (int b = 0; b < MD(nb); b++)
      if (m.b_root[b] == r) s += mc[3 * b + k];
    WR(rootcom)[e] = s * m.r_invmass[r];
  }
  RCSB_SYNC();
  PFOR(b, MD(nb)) {  // inertia about the tree COM, world axes
    const real* R = WR(bmat) + 9 * b;
    const real* I = m.b_inertia[b];
    const real* rc = WR(rootcom) + 3 * m.b_root[b];
    real bc[3];
    body_com(c, b, bc);
    real mass = m.b_mass[b], d[3] = {bc[0] - rc[0], bc[1] - rc[1], bc[2] - rc[2]};
    real A[9];  // R * I
    for (int r = 0; r < 3; r++) {
      real r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
      A[3 * r] = r0 * I[0] + r1 * I[3] + r2 * I[4];
      A[3 * r + 1] = r0 * I[3] + r1 * I[1] + r2 * I[5];
      A[3 * r + 2] = r0 * I[4] + r1 * I[5] + r2 * I[2];
    }
    real dd = dot3(d, d);
    real* ci = WR(cinert) + 10 * b;
    ci[0] = A[0] * R[0] + A[1] * R[1] + A[2] * R[2] + mass * (dd - d[0] * d[0]);
    ci[1] = A[3] * R[3] + A[4] * R[4] + A[5] * R[5] + mass * (dd - d[1] * d[1]);
    ci[2] = A[6] * R[6] + A[7] * R[7] + A[8] * R[8] + mass * (dd - d[2] * d[2]);
    ci[3] = A[0] * R[3] + A[1] * R[4] + A[2] * R[5] - mass * d[0] * d[1];
    ci[4] = A[0] * R[6] + A[1] * R[7] + A[2] * R[8] - mass * d[0] * d[2];
    ci[5] = A[3] * R[6] + A[4] * R[7] + A[5] * R[8] - mass * d[1] * d[2];
    ci[6] = mass * d[0]; ci[7] = mass * d[1]; ci[8] = mass * d[2];
    ci[9] = mass;
  }
  PFOR(j, MD(nv)) {  // motion axis of dof j about the tree COM
    int b = m.d_body[j];
    const real* rc = WR(rootcom) + 3 * m.b_root[b];
    const real* R = WR(bmat) + 9 * b;
    const real* pos = WR(bpos) + 3 * b;
    real* cd = WR(cdof) + 6 * j;
    int jt = m.b_jtype[b];
    // a rotation about the joint axis leaves axis and anchor in place, so both come from the body frame
    real axis[3], an[3];
    if (jt == RCSB_JNT_FREE) copy3(an, pos);
    else {
      mulmat3(axis, R, m.b_jaxis[b]);
      mulmat3(an, R, m.b_jpos[b]);
      an[0] += pos[0]; an[1] += pos[1]; an[2] += pos[2];
    }
    real off[3] = {rc[0] - an[0], rc[1] - an[1], rc[2] - an[2]};
    if (jt == RCSB_JNT_FREE) {
      int a = j - m.b_dadr[b];
      if (a < 3) {
        cd[0] = cd[1] = cd[2] = 0;
        cd[3] = a == 0; cd[4] = a == 1; cd[5] = a == 2;
      } else {
        real ax[3] = {R[a - 3], R[3 + a - 3], R[6 + a - 3]};
        copy3(cd, ax);
        cross3(cd + 3, ax, off);
      }
    } else if (jt == RCSB_JNT_SLIDE) {
      cd[0] = cd[1] = cd[2] = 0;
      copy3(cd + 3, axis);
    } else {
      copy3(cd, axis);
      cross3(cd + 3, axis, off);
    }
  }
  PFOR1(e, 12) {  // attachment site pose (SimRobot::get_cartesian_position reads it after the step): xpos[3] | xmat[9]
    const int b = m.rb_site_body;
    real* sp = WR(rcs) + RCSB_S_SITEPOS;
    const real* Rl = m.rb_site_rot;
    if (e < 3) {
      real val = m.rb_site_pos[e];
      if (b >= 0) {
        const real* R = WR(bmat) + 9 * b + 3 * e;
        val = WR(bpos)[3 * b + e] + (R[0] * m.rb_site_pos[0] + R[1] * m.rb_site_pos[1] + R[2] * m.rb_site_pos[2]);
      }
      sp[e] = val;
    } else {
      const int r = (e - 3) / 3, k = (e - 3) - 3 * r;
      real val = Rl[3 * r + k];
      if (b >= 0) {
        const real* R = WR(bmat) + 9 * b + 3 * r;
        val = R[0] * Rl[k] + R[1] * Rl[3 + k] + R[2] * Rl[6 + k];
      }
      sp[e] = val;
    }
  }
  RCSB_SYNC();
}

// ------------------------------------------------------------------ composite inertia, mass matrix, factorisation
RCSB_DEV void st_crb(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  int nv = MD(nv);
  // composite inertias: every body sums the inertias of its subtree (descendant mask, leaves first) - one lane per
  // (body, component), no serial leaf-to-root chain through shared memory
  {
    const int nb = MD(nb), n = nb * 10;
#pragma unroll
    for (int base = 0; base < n; base += RCSB_NLANES) {
      const int e = base + c.lane, bmin = base / 10;
      if (e < n) {
        const int b = e / 10, k = e - 10 * b;
        const unsigned mask = (unsigned)m.b_descmask[b];
        real s = 0;
#pragma unroll
        for (int d = nb - 1; d >= bmin; d--)
          if ((mask >> d) & 1u) s += WR(cinert)[10 * d + k];
        WR(crb)[e] = s;
      }
    }
  }
  RCSB_SYNC();
  real* buf = WR(crbbuf);
  PFOR(i, nv) { mul_inert_vec(buf + 6 * i, WR(crb) + 10 * m.d_body[i], WR(cdof) + 6 * i); }
  RCSB_SYNC();
  PFOR(t, nv * (nv + 1) / 2) {
    const int i = m.tri_i[t], j = m.tri_j[t];
    {
      real val = 0;
      if ((m.d_ancmask[i] >> j) & 1u) {
        const real* cd = WR(cdof) + 6 * j;
        const real* bf = buf + 6 * i;
        val = cd[0] * bf[0] + cd[1] * bf[1] + cd[2] * bf[2] + cd[3] * bf[3] + cd[4] * bf[4] + cd[5] * bf[5];
        if (i == j) val += m.d_armature[i];
      }
      WR(M)[i * nv + j] = val;
      WR(M)[j * nv + i] = val;
    }
  }
  // the factorisation of M is deferred (ensure_chol_M copies M into the solver scratch): the all-equality path never needs it
  if (c.lane == 0) { CWI(c)[LAY.oi_misc + MI_HAVE_L] = 0; CWI(c)[LAY.oi_misc + MI_HAVE_H2] = 0; }
  RCSB_SYNC();
}

// ------------------------------------------------------------------ velocity stage: bias, passive, gravity compensation
RCSB_DEV void st_velocity(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv), nb = MD(nb);
  const real* v = WR(v);
  // pass A: spatial velocity accumulated up to and including every dof (a body's velocity is that of its last dof) -
  // one lane per (dof, component) sums over the dof's ancestor mask in root-to-leaf order
  {
    const int n = nv * 6;
#pragma unroll
    for (int base = 0; base < n; base += RCSB_NLANES) {
      const int e = base + c.lane;
      const int jmax = (base + RCSB_NLANES - 1) / 6 < nv - 1 ? (base + RCSB_NLANES - 1) / 6 : nv - 1;
      if (e < n) {
        const int j = e / 6, k = e - 6 * j;
        const unsigned mask = (unsigned)m.d_ancmask[j];
        real s = 0;
#pragma unroll
        for (int a = 0; a <= jmax; a++)
          if ((mask >> a) & 1u) s += WR(cdof)[6 * a + k] * v[a];
        WR(cvel)[e] = s;
      }
    }
  }
  RCSB_SYNC();
  // time derivative of every motion axis: (velocity accumulated before the dof) x axis
  PFOR(j, nv) {
    real* cdd = WR(cdofdot) + 6 * j;
    const int pre = m.d_pre[j];
    if (m.d_dotzero[j] || pre < 0) {
      for (int k = 0; k < 6; k++) cdd[k] = 0;
    } else {
      cross_motion(cdd, WR(cvel) + 6 * pre, WR(cdof) + 6 * j);
    }
  }
  RCSB_SYNC();
  // pass B: spatial acceleration with qacc = 0 and gravity folded into the root, same ancestor-mask sum
  {
    const int n = nv * 6;
#pragma unroll
    for (int base = 0; base < n; base += RCSB_NLANES) {
      const int e = base + c.lane;
      const int jmax = (base + RCSB_NLANES - 1) / 6 < nv - 1 ? (base + RCSB_NLANES - 1) / 6 : nv - 1;
      if (e < n) {
        const int j = e / 6, k = e - 6 * j;
        const unsigned mask = (unsigned)m.d_ancmask[j];
        real s = k >= 3 ? -m.gravity[k - 3] : (real)0;
#pragma unroll
        for (int a = 0; a <= jmax; a++)
          if ((mask >> a) & 1u) s += WR(cdofdot)[6 * a + k] * v[a];
        WR(cacc)[e] = s;
      }
    }
  }
  RCSB_SYNC();
  PFOR(b, nb) {
    const int jl = m.b_lastdof[b];
    real Ia[6], Iv[6], x[6];
    mul_inert_vec(Ia, WR(cinert) + 10 * b, WR(cacc) + 6 * jl);
    mul_inert_vec(Iv, WR(cinert) + 10 * b, WR(cvel) + 6 * jl);
    cross_force(x, WR(cvel) + 6 * jl, Iv);
    for (int k = 0; k < 6; k++) WR(cfrc)[6 * b + k] = Ia[k] + x[k];  // cdof_dot is dead: cfrc shares its slots
    // gravity-compensation wrench [torque; force] about the tree COM (force -gcmass*g at the compensated mass centre)
    real* gw = WR(gcw) + 6 * jl;  // the acceleration in this slot was consumed above, by this lane only
    real gm = m.b_gcmass[b];
    if (gm != 0) {
      const real* rc = WR(rootcom) + 3 * m.b_root[b];
      const real* p = WR(bpos) + 3 * b;
      real o[3], fg[3] = {-gm * m.gravity[0], -gm * m.gravity[1], -gm * m.gravity[2]};
      mulmat3(o, WR(bmat) + 9 * b, m.b_gcpos[b]);
      o[0] += p[0] - rc[0]; o[1] += p[1] - rc[1]; o[2] += p[2] - rc[2];
      cross3(gw, o, fg);
      copy3(gw + 3, fg);
    } else {
      for (int k = 0; k < 6; k++) gw[k] = 0;
    }
  }
  RCSB_SYNC();
  // leaves to root: every body collects the forces / compensation wrenches of its subtree (lanes 0-5 | 6-11)
  for (int b = nb - 1; b > 0; b--) {
    const int p = m.b_parent[b];
    if (p >= 0) PFOR1(k, 12) {
      if (k < 6) WR(cfrc)[6 * p + k] += WR(cfrc)[6 * b + k];
      else WR(gcw)[6 * m.b_lastdof[p] + k - 6] += WR(gcw)[6 * m.b_lastdof[b] + k - 6];
    }
    RCSB_SYNC();
  }
  PFOR(j, nv) {
    const int bj = m.d_body[j];
    const real* cd = WR(cdof) + 6 * j;
    const real* f = WR(cfrc) + 6 * bj;
    const real* g = WR(gcw) + 6 * m.b_lastdof[bj];
    WR(bias)[j] = cd[0] * f[0] + cd[1] * f[1] + cd[2] * f[2] + cd[3] * f[3] + cd[4] * f[4] + cd[5] * f[5];
    real gc = cd[0] * g[0] + cd[1] * g[1] + cd[2] * g[2] + cd[3] * g[3] + cd[4] * g[4] + cd[5] * g[5];
    WR(gravc)[j] = gc;
    WR(passive)[j] = -m.d_damping[j] * v[j] + (m.d_actgravcomp[j] ? (real)0 : gc);
  }
  RCSB_SYNC();
}

// ------------------------------------------------------------------ collision
// support mapping of geom g (world pose gp, gR) along world direction dir; warp-cooperative for meshes
RCSB_DEV void support(const Ctx& c, int g, const real* gp, const real* gR, const real* dir, real* out) {
  const RcsbModel& m = CMODEL(c);
  real dl[3], v[3] = {0, 0, 0};
  mulmatT3(dl, gR, dir);
  int type = m.g_type[g];
  if (type == RCSB_GEOM_MESH) {
    const real* verts = c.verts + 3 * m.g_vertadr[g];
    int n = m.g_vertnum[g], best = 0x7fffffff;
    real bd = (real)-1e300;
    PFOR(i, n) {
      real s = RCSB_LDG(verts + 3 * i) * dl[0] + RCSB_LDG(verts + 3 * i + 1) * dl[1] + RCSB_LDG(verts + 3 * i + 2) * dl[2];
#if defined(RCSB_HOST_EMU) && defined(RCSB_EMU_REVERSE)
      if (s > bd || (s == bd && i < best)) { bd = s; best = i; }
#else
      if (s > bd) { bd = s; best = i; }
#endif
    }
    warp_argmax(bd, best);
    v[0] = RCSB_LDG(verts + 3 * best); v[1] = RCSB_LDG(verts + 3 * best + 1); v[2] = RCSB_LDG(verts + 3 * best + 2);
  } else if (type == RCSB_GEOM_BOX) {
    for (int k = 0; k < 3; k++) v[k] = dl[k] >= 0 ? m.g_size[g][k] : -m.g_size[g][k];
  } else if (type == RCSB_GEOM_CAPSULE) {
    real n = norm3(dl);
    if (n > RCSB_MINVAL) for (int k = 0; k < 3; k++) v[k] = m.g_size[g][0] * dl[k] / n;
    v[2] += dl[2] >= 0 ? m.g_size[g][1] : -m.g_size[g][1];
  } else if (type == RCSB_GEOM_SPHERE) {
    real n = norm3(dl);
    if (n > RCSB_MINVAL) for (int k = 0; k < 3; k++) v[k] = m.g_size[g][0] * dl[k] / n;
  } else if (type == RCSB_GEOM_CYLINDER) {  // rim point in the xy direction, cap by the sign of z
    real n = r_sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (n > RCSB_MINVAL) { v[0] = dl[0] / n * m.g_size[g][0]; v[1] = dl[1] / n * m.g_size[g][0]; }
    v[2] = dl[2] > 0 ? m.g_size[g][1] : (dl[2] < 0 ? -m.g_size[g][1] : (real)0);
  }
  real o[3];
  mulmat3(o, gR, v);
  out[0] = o[0] + gp[0]; out[1] = o[1] + gp[1]; out[2] = o[2] + gp[2];
}

struct Sup { real v[3], v1[3], v2[3]; };
// world frames of the geom pair in the narrow phase: pointers into the warp's workspace (o_pairfr), filled by pair_frames
struct PairFrames { int g1, g2; const real *p1, *R1, *p2, *R2; };

// world frames of both geoms of the pair, 12 lanes per geom (3 position + 9 rotation entries), into o_pairfr
RCSB_DEV void pair_frames(const Ctx& c, PairFrames& pf) {
  const RcsbModel& m = CMODEL(c);
  real* fr = WR(pairfr);
  RCSB_SYNC();  // the previous pair's readers are done
  PFOR1(e, 24) {
    const int g = e < 12 ? pf.g1 : pf.g2, i = e < 12 ? e : e - 12, b = m.g_body[g];
    real val;
    if (i < 3) {
      val = m.g_pos[g][i];
      if (b >= 0) {
        const real* R = WR(bmat) + 9 * b + 3 * i;
        val = WR(bpos)[3 * b + i] + (R[0] * m.g_pos[g][0] + R[1] * m.g_pos[g][1] + R[2] * m.g_pos[g][2]);
      }
    } else {
      const int r = (i - 3) / 3, k = (i - 3) - 3 * r;
      const real* Rl = m.g_rot[g];
      val = Rl[3 * r + k];
      if (b >= 0) {
        const real* R = WR(bmat) + 9 * b + 3 * r;
        val = R[0] * Rl[k] + R[1] * Rl[3 + k] + R[2] * Rl[6 + k];
      }
    }
    fr[e] = val;
  }
  RCSB_SYNC();
  pf.p1 = fr; pf.R1 = fr + 3; pf.p2 = fr + 12; pf.R2 = fr + 15;
}
RCSB_DEV void mink_support(const Ctx& c, const PairFrames& pf, const real* dir, Sup& s) {
  real nd[3] = {-dir[0], -dir[1], -dir[2]};
  real a[3], b[3];
  support(c, pf.g1, pf.p1, pf.R1, dir, a);
  support(c, pf.g2, pf.p2, pf.R2, nd, b);
  RCSB_SYNC();  // s is shared by the warp: no lane may still be reading the record this call overwrites
  // every lane holds the same finished values in registers and stores them: writes only, nothing is read back here
  for (int k = 0; k < 3; k++) { s.v1[k] = a[k]; s.v2[k] = b[k]; s.v[k] = a[k] - b[k]; }
}
RCSB_DEV real origin_tri_dist(const real* A, const real* B, const real* C, real* w) {
  // closest point of triangle ABC to the origin (Ericson, Real-Time Collision Detection 5.1.5)
  real ab[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, ac[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
  real ap[3] = {-A[0], -A[1], -A[2]};
  real d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  bool done = false;
  if (d1 <= 0 && d2 <= 0) { copy3(w, A); done = true; }
  real bp[3] = {-B[0], -B[1], -B[2]};
  real d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (!done && d3 >= 0 && d4 <= d3) { copy3(w, B); done = true; }
  real vc = d1 * d4 - d3 * d2;
  if (!done && vc <= 0 && d1 >= 0 && d3 <= 0) {
    real t = d1 / (d1 - d3);
    for (int k = 0; k < 3; k++) w[k] = A[k] + t * ab[k];
    done = true;
  }
  real cp[3] = {-C[0], -C[1], -C[2]};
  real d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (!done && d6 >= 0 && d5 <= d6) { copy3(w, C); done = true; }
  real vb = d5 * d2 - d1 * d6;
  if (!done && vb <= 0 && d2 >= 0 && d6 <= 0) {
    real t = d2 / (d2 - d6);
    for (int k = 0; k < 3; k++) w[k] = A[k] + t * ac[k];
    done = true;
  }
  real va = d3 * d6 - d5 * d4;
  if (!done && va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    real t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) w[k] = B[k] + t * (C[k] - B[k]);
    done = true;
  }
  if (!done) {
    real den = (real)1 / (va + vb + vc), vv = vb * den, uu = vc * den;
    for (int k = 0; k < 3; k++) w[k] = A[k] + ab[k] * vv + ac[k] * uu;
  }
  return r_sqrt(dot3(w, w));
}
RCSB_DEV void portal_dir(const Sup* s, real* dir) {
  real a[3] = {s[2].v[0] - s[1].v[0], s[2].v[1] - s[1].v[1], s[2].v[2] - s[1].v[2]};
  real b[3] = {s[3].v[0] - s[1].v[0], s[3].v[1] - s[1].v[1], s[3].v[2] - s[1].v[2]};
  cross3(dir, a, b);
  normalize3(dir);
}
RCSB_DEV int portal_reach_tol(const Sup* s, const Sup& v4, const real* dir) {
  real dv4 = dot3(v4.v, dir);
  real a = dv4 - dot3(s[1].v, dir), b = dv4 - dot3(s[2].v, dir), cc = dv4 - dot3(s[3].v, dir);
  real mn = a < b ? a : b;
  mn = mn < cc ? mn : cc;
  return mn <= (real)1e-6;
}
RCSB_DEV void expand_portal(Sup* s, const Sup& v4) {
  real x[3];
  cross3(x, v4.v, s[0].v);
  const int which = dot3(s[1].v, x) > 0 ? (dot3(s[2].v, x) > 0 ? 1 : 3) : (dot3(s[3].v, x) > 0 ? 2 : 1);
  RCSB_SYNC();  // every lane has taken its decision before the shared portal changes
  s[which] = v4;
  RCSB_SYNC();
}
// Minkowski Portal Refinement penetration query (algorithm of libccd's ccdMPRPenetration, which
// MuJoCo 3.2.6 uses for convex pairs). Warp-uniform control flow; only support() is cooperative.
// When the query ends on a support check "the Minkowski difference does not reach past the origin along dir",
// dir is a separating direction; it is handed back through sep (sep[3] = 1) so the caller can re-validate it with
// a single support pair on the next substeps instead of repeating the whole query.
RCSB_DEV_NOINLINE int mpr_penetration(const Ctx& c, const PairFrames& pf, real* depth, real* dir_out, real* pos, real* sep) {
  sep[3] = 0;
  // the portal lives in the warp's workspace: every lane computes the same values, so one shared copy replaces 32
  // per-thread copies in local memory; warp barriers separate the reads of a record from its replacement
  const RcsbModel& m = CMODEL(c);
  (void)m;
  Sup* s = (Sup*)WR(sup);
  Sup& v4 = s[4];
  real dir[3], va[3], vb[3];
  {
    real v0[3] = {pf.p1[0] - pf.p2[0], pf.p1[1] - pf.p2[1], pf.p1[2] - pf.p2[2]};
    if (r_abs(v0[0]) < (real)1e-12 && r_abs(v0[1]) < (real)1e-12 && r_abs(v0[2]) < (real)1e-12) v0[0] += (real)1e-5;
    RCSB_SYNC();
    for (int k = 0; k < 3; k++) { s[0].v1[k] = pf.p1[k]; s[0].v2[k] = pf.p2[k]; s[0].v[k] = v0[k]; }  // plain stores only
  }
  dir[0] = -s[0].v[0]; dir[1] = -s[0].v[1]; dir[2] = -s[0].v[2];
  normalize3(dir);
  mink_support(c, pf, dir, s[1]);
  if (dot3(s[1].v, dir) <= 0) { copy3(sep, dir); sep[3] = 1; return 0; }
  cross3(dir, s[0].v, s[1].v);
  if (dot3(dir, dir) < (real)1e-24) {
    *depth = norm3(s[1].v);
    for (int k = 0; k < 3; k++) { dir_out[k] = s[1].v[k]; pos[k] = (real)0.5 * (s[1].v1[k] + s[1].v2[k]); }
    normalize3(dir_out);
    return 1;
  }
  normalize3(dir);
  mink_support(c, pf, dir, s[2]);
  if (dot3(s[2].v, dir) <= 0) { copy3(sep, dir); sep[3] = 1; return 0; }
  for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
  cross3(dir, va, vb);
  normalize3(dir);
  const int flip = dot3(dir, s[0].v) > 0;
  RCSB_SYNC();
  if (flip) {
    Sup t1 = s[1], t2 = s[2];
    RCSB_SYNC();
    s[1] = t2; s[2] = t1;
    dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2];
  }
  RCSB_SYNC();
  for (int it = 0;; it++) {
    if (it > 100) return 0;
    mink_support(c, pf, dir, s[3]);
    if (dot3(s[3].v, dir) <= 0) { copy3(sep, dir); sep[3] = 1; return 0; }
    int cont = 0;
    cross3(va, s[1].v, s[3].v);
    if (dot3(va, s[0].v) < (real)-1e-18) cont = 2;
    if (!cont) {
      cross3(va, s[3].v, s[2].v);
      if (dot3(va, s[0].v) < (real)-1e-18) cont = 1;
    }
    RCSB_SYNC();
    if (!cont) break;
    s[cont] = s[3];
    RCSB_SYNC();
    for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
    cross3(dir, va, vb);
    normalize3(dir);
  }
  for (int it = 0;; it++) {
    portal_dir(s, dir);
    if (dot3(s[1].v, dir) >= 0) break;
    mink_support(c, pf, dir, v4);
    if (dot3(v4.v, dir) < 0) { copy3(sep, dir); sep[3] = 1; return 0; }
    if (portal_reach_tol(s, v4, dir) || it >= 50) return 0;
    expand_portal(s, v4);
  }
  for (int it = 0;; it++) {
    portal_dir(s, dir);
    mink_support(c, pf, dir, v4);
    if (portal_reach_tol(s, v4, dir) || it >= 50) {
      real w[3];
      *depth = origin_tri_dist(s[1].v, s[2].v, s[3].v, w);
      if (*depth < (real)1e-12) copy3(dir_out, dir);
      else { copy3(dir_out, w); normalize3(dir_out); }
      // contact position: barycentric blend of the witness points (libccd findPos)
      real b[4], vec[3], sum;
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
      real p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
      for (int i = 0; i < 4; i++)
        for (int k = 0; k < 3; k++) { p1[k] += b[i] * s[i].v1[k]; p2[k] += b[i] * s[i].v2[k]; }
      for (int k = 0; k < 3; k++) pos[k] = (real)0.5 * (p1[k] + p2[k]) / sum;
      return 1;
    }
    expand_portal(s, v4);
  }
}

// append one contact (all lanes call with identical arguments; lane 0 writes)
// like_prev: the previous contact of the list belongs to the same geom pair and has the same normal (the other points of
// a plane-box, plane-capsule, plane-hull or box-box face contact): its frame and mixed solver parameters are copied instead
// of being derived again (identical values; the derivation is a serial chain with a square root and divisions in lane 0)
RCSB_DEV void add_contact(const Ctx& c, int& ncon, int g1, int g2, real dist, const real* pos, const real* normal,
                          real margin, real gap, int like_prev = 0) {
  const RcsbModel& m = CMODEL(c);
  if (ncon >= MD(maxcon)) {  // reduced layout: the full-capacity launch redoes this step; full layout: drop and count
    if (c.lane == 0) WI(misc)[MD(cap_reduced) ? MI_OVERFLOW : MI_WARN] += 1;
    return;
  }
  if (c.lane == 0) {
    real* cr = WR(con) + RCSB_C_REALS * ncon;
    int* ci = WI(con) + RCSB_CI_INTS * ncon;
    if (like_prev && ncon > 0) {
      const real* pr = cr - RCSB_C_REALS;
      const int* pi = ci - RCSB_CI_INTS;
      for (int k = 0; k < RCSB_C_REALS; k++) cr[k] = pr[k];
      ci[RCSB_CI_G0] = g1; ci[RCSB_CI_G1] = g2; ci[RCSB_CI_DIM] = pi[RCSB_CI_DIM]; ci[RCSB_CI_EFC] = -1;
      cr[RCSB_C_DIST] = dist;
      copy3(cr + RCSB_C_POS, pos);
      cr[RCSB_C_INCMARGIN] = margin - gap;
      cr[RCSB_C_MU] = 0;
      ncon++;
      return;
    }
    cr[RCSB_C_DIST] = dist;
    copy3(cr + RCSB_C_POS, pos);
    copy3(cr + RCSB_C_FRAME, normal);
    make_frame(cr + RCSB_C_FRAME);
    cr[RCSB_C_INCMARGIN] = margin - gap;
    ci[RCSB_CI_G0] = g1; ci[RCSB_CI_G1] = g2;
    int p1 = m.g_priority[g1], p2 = m.g_priority[g2];
    real mix, fr[3];
    int dim;
    if (p1 == p2) {
      real s1 = m.g_solmix[g1], s2 = m.g_solmix[g2];
      if (s1 >= RCSB_MINVAL && s2 >= RCSB_MINVAL) mix = s1 / (s1 + s2);
      else if (s1 < RCSB_MINVAL && s2 < RCSB_MINVAL) mix = (real)0.5;
      else mix = s1 < RCSB_MINVAL ? (real)0 : (real)1;
      for (int k = 0; k < 3; k++) fr[k] = m.g_friction[g1][k] > m.g_friction[g2][k] ? m.g_friction[g1][k] : m.g_friction[g2][k];
      dim = m.g_condim[g1] > m.g_condim[g2] ? m.g_condim[g1] : m.g_condim[g2];
    } else {
      int g = p1 > p2 ? g1 : g2;
      mix = p1 > p2 ? (real)1 : (real)0;
      for (int k = 0; k < 3; k++) fr[k] = m.g_friction[g][k];
      dim = m.g_condim[g];
    }
    ci[RCSB_CI_DIM] = dim;
    ci[RCSB_CI_EFC] = -1;
    const real *r1 = m.g_solref[g1], *r2 = m.g_solref[g2];
    if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) cr[RCSB_C_SOLREF + k] = mix * r1[k] + (1 - mix) * r2[k];
    else for (int k = 0; k < 2; k++) cr[RCSB_C_SOLREF + k] = r1[k] < r2[k] ? r1[k] : r2[k];
    for (int k = 0; k < 5; k++) cr[RCSB_C_SOLIMP + k] = mix * m.g_solimp[g1][k] + (1 - mix) * m.g_solimp[g2][k];
    for (int k = 0; k < 3; k++) cr[RCSB_C_FRIC + k] = fr[k] < (real)1e-5 ? (real)1e-5 : fr[k];  // slide, spin, roll
    cr[RCSB_C_MU] = 0;
  }
  ncon++;
}


// ---- box-box (mjc_BoxBox, restated as separating axes + face clipping; see oracle/mj_collision.c): up to 8 contacts.
// Every lane runs the same scalar code on identical data (add_contact lets lane 0 write); out of line. The clip polygons
// live in the warp's shared workspace (the tabletop scene runs this every step for the brick on the table).
RCSB_DEV_NOINLINE void box_box(const Ctx& c, int& ncon, const PairFrames& pf, real margin, real gap, real* clear_out) {
  const RcsbModel& m = CMODEL(c);
  const real *R1 = pf.R1, *R2 = pf.R2, *p1 = pf.p1, *p2 = pf.p2;
  const real *a = m.g_size[pf.g1], *b = m.g_size[pf.g2];
  real Rel[9], Q[9], t[3], dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  mulmatT3(t, R1, dp);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Rel[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
      Q[3 * i + j] = r_abs(Rel[3 * i + j]);
    }
  real best = (real)-1e300;
  int code = -1, flip = 0;
  for (int i = 0; i < 3; i++) {
    real s = r_abs(t[i]) - (a[i] + b[0] * Q[3 * i] + b[1] * Q[3 * i + 1] + b[2] * Q[3 * i + 2]);
    if (s > best) { best = s; code = i; flip = t[i] < 0; }
  }
  for (int j = 0; j < 3; j++) {
    real tj = t[0] * Rel[j] + t[1] * Rel[3 + j] + t[2] * Rel[6 + j];
    real s = r_abs(tj) - (a[0] * Q[j] + a[1] * Q[3 + j] + a[2] * Q[6 + j] + b[j]);
    if (s > best) { best = s; code = 3 + j; flip = tj < 0; }
  }
  // a face axis separates the boxes by more than the margin (or they touch exactly): no contact; the clearance feeds the
  // group's separation budget
  if (best > margin - (real)1e-12) { *clear_out = best > margin ? best - margin : (real)0; return; }
  real ebest = (real)-1e300, en[3] = {0, 0, 0};
  int ecode = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      real l2 = 1 - Rel[3 * i + j] * Rel[3 * i + j];
      if (l2 < (real)1e-10) continue;
      real l = r_sqrt(l2);
      real proj = t[i2] * Rel[3 * i1 + j] - t[i1] * Rel[3 * i2 + j];
      real ra = a[i1] * Q[3 * i2 + j] + a[i2] * Q[3 * i1 + j];
      real rb = b[j1] * Q[3 * i + j2] + b[j2] * Q[3 * i + j1];
      real s = (r_abs(proj) - (ra + rb)) / l;
      if (s > margin) return;  // separated along an edge-edge axis only: no usable clearance bound (budget stays 0)
      if (s > ebest) {
        ebest = s; ecode = 3 * i + j;
        real ax[3] = {0, 0, 0};
        ax[i1] = -Rel[3 * i2 + j] / l; ax[i2] = Rel[3 * i1 + j] / l;
        if (proj < 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
        en[0] = ax[0]; en[1] = ax[1]; en[2] = ax[2];
      }
    }
  if (ecode >= 0 && ebest * (real)1.05 > best) {  // edge-edge: closest points of the two edge lines
    const int i = ecode / 3, j = ecode % 3;
    real nw[3];
    mulmat3(nw, R1, en);
    real pa[3], pb[3], ua[3] = {R1[i], R1[3 + i], R1[6 + i]}, ub[3] = {R2[j], R2[3 + j], R2[6 + j]};
    copy3(pa, p1); copy3(pb, p2);
    for (int k = 0; k < 3; k++) {
      if (k != i) {
        real ak[3] = {R1[k], R1[3 + k], R1[6 + k]};
        real sg = dot3(nw, ak) > 0 ? a[k] : -a[k];
        for (int cc = 0; cc < 3; cc++) pa[cc] += sg * ak[cc];
      }
      if (k != j) {
        real bk[3] = {R2[k], R2[3 + k], R2[6 + k]};
        real sg = dot3(nw, bk) > 0 ? -b[k] : b[k];
        for (int cc = 0; cc < 3; cc++) pb[cc] += sg * bk[cc];
      }
    }
    real w[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    real uaub = dot3(ua, ub), q1 = dot3(ua, w), q2 = -dot3(ub, w), den = 1 - uaub * uaub;
    real alpha = 0, beta = 0;
    if (den > (real)1e-10) { alpha = (q1 + uaub * q2) / den; beta = (uaub * q1 + q2) / den; }
    real pos[3];
    for (int cc = 0; cc < 3; cc++) pos[cc] = (real)0.5 * ((pa[cc] + alpha * ua[cc]) + (pb[cc] + beta * ub[cc]));
    add_contact(c, ncon, pf.g1, pf.g2, ebest, pos, nw, margin, gap);
    return;
  }
  // face contact: reference face on box 1 (code < 3) or on box 2
  const real *Rr, *Ri, *pr, *pi, *sr, *si;
  const int ax = code < 3 ? code : code - 3;
  if (code < 3) { Rr = R1; Ri = R2; pr = p1; pi = p2; sr = a; si = b; }
  else { Rr = R2; Ri = R1; pr = p2; pi = p1; sr = b; si = a; }
  const real sgn = (code < 3) ? (flip ? (real)-1 : (real)1) : (flip ? (real)1 : (real)-1);
  const real nr[3] = {sgn * Rr[ax], sgn * Rr[3 + ax], sgn * Rr[6 + ax]};
  int iax = 0;
  real md = -1, isg = 1;
  for (int k = 0; k < 3; k++) {
    real dk = nr[0] * Ri[k] + nr[1] * Ri[3 + k] + nr[2] * Ri[6 + k];
    if (r_abs(dk) > md) { md = r_abs(dk); iax = k; isg = dk > 0 ? (real)-1 : (real)1; }
  }
  const int u1 = (iax + 1) % 3, u2 = (iax + 2) % 3, r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  // The clip polygons (a quadrilateral cut by four half-planes: 8 vertices at most) live in the warp's shared workspace, in
  // the slots of the Minkowski portal that this pair type does not use: as thread-local arrays they were indexed
  // dynamically, i.e. in local memory, and the serial walk over them was two thirds of the tabletop scene's collision
  // stage. Lane 0 stores, everybody reads after a warp barrier; the control flow stays uniform.
  real (*poly)[3] = (real (*)[3])WR(sup);
  real (*tmp)[3] = poly + 8;
  int np = 4;
  RCSB_SYNC();  // the previous user of the portal slots is done
  for (int v = 0; v < 4; v++) {
    const real c0 = (v == 0 || v == 3) ? (real)1 : (real)-1, c1 = v < 2 ? (real)1 : (real)-1;
    real e[3];
    for (int cc = 0; cc < 3; cc++)
      e[cc] = pi[cc] + isg * si[iax] * Ri[3 * cc + iax] + c0 * si[u1] * Ri[3 * cc + u1] + c1 * si[u2] * Ri[3 * cc + u2] - pr[cc];
    if (c.lane == 0) {
      poly[v][0] = e[0] * Rr[r1] + e[1] * Rr[3 + r1] + e[2] * Rr[6 + r1];
      poly[v][1] = e[0] * Rr[r2] + e[1] * Rr[3 + r2] + e[2] * Rr[6 + r2];
      poly[v][2] = dot3(e, nr) - sr[ax];
    }
  }
  RCSB_SYNC();
  for (int side = 0; side < 4 && np > 0; side++) {  // Sutherland-Hodgman against the reference rectangle
    const int cx = side >> 1;
    const real sg = (side & 1) ? (real)-1 : (real)1, lim = cx == 0 ? sr[r1] : sr[r2];
    int nq = 0;
    for (int v = 0; v < np; v++) {
      const real *A = poly[v], *B = poly[v + 1 == np ? 0 : v + 1];
      const real a3[3] = {A[0], A[1], A[2]}, b3[3] = {B[0], B[1], B[2]};
      real da = lim - sg * a3[cx], db = lim - sg * b3[cx];
      if (da >= 0) {
        if (c.lane == 0 && nq < 8) copy3(tmp[nq], a3);
        nq++;
      }
      if ((da >= 0) != (db >= 0)) {
        real f = da / (da - db);
        if (c.lane == 0 && nq < 8) for (int k = 0; k < 3; k++) tmp[nq][k] = a3[k] + f * (b3[k] - a3[k]);
        nq++;
      }
    }
    np = nq < 8 ? nq : 8;
    RCSB_SYNC();  // the clipped polygon is complete: it is the next side's input (the two buffers swap roles)
    real (*sw)[3] = poly; poly = tmp; tmp = sw;
  }
  real nw[3] = {nr[0], nr[1], nr[2]};
  if (code >= 3) { nw[0] = -nw[0]; nw[1] = -nw[1]; nw[2] = -nw[2]; }  // contact normal runs from geom 1 to geom 2
  int cnt = 0;
  for (int v = 0; v < np && cnt < 8; v++) {
    real dist = poly[v][2];
    if (dist > margin) continue;
    real pos[3];
    for (int cc = 0; cc < 3; cc++)
      pos[cc] = pr[cc] + poly[v][0] * Rr[3 * cc + r1] + poly[v][1] * Rr[3 * cc + r2] + (sr[ax] + (real)0.5 * dist) * nr[cc];
    add_contact(c, ncon, pf.g1, pf.g2, dist, pos, nw, margin, gap, cnt > 0);
    cnt++;
  }
}
// ---- plane-mesh (mjc_PlaneConvex): the deepest hull vertex plus neighbours of it in the hull's edge graph that are within
// the margin and at least 0.3 * rbound away from the points already taken, 3 contacts at most
RCSB_DEV_NOINLINE void plane_mesh(const Ctx& c, int& ncon, const PairFrames& pf, const real* n, real margin, real gap, real* clear_out) {
  const RcsbModel& m = CMODEL(c);
  const int g = pf.g2;
  const real* verts = c.verts + 3 * m.g_vertadr[g];
  real nd[3] = {-n[0], -n[1], -n[2]}, dl[3];
  mulmatT3(dl, pf.R2, nd);
  int best = 0x7fffffff;
  real bd = (real)-1e300;
  PFOR(i, m.g_vertnum[g]) {
    real s = RCSB_LDG(verts + 3 * i) * dl[0] + RCSB_LDG(verts + 3 * i + 1) * dl[1] + RCSB_LDG(verts + 3 * i + 2) * dl[2];
#if defined(RCSB_HOST_EMU) && defined(RCSB_EMU_REVERSE)
    if (s > bd || (s == bd && i < best)) { bd = s; best = i; }
#else
    if (s > bd) { bd = s; best = i; }
#endif
  }
  warp_argmax(bd, best);
  real taken[3][3];
  int count = 0;
  const real thr = (real)0.3 * CMODEL_G(c).g_rbound0[g], thr2 = thr * thr;  // tolplanemesh
  const int nb0 = c.vgraph ? RCSB_LDG(c.vgraph + m.g_vertadr[g] + best) : 0;
  const int nb1 = c.vgraph ? RCSB_LDG(c.vgraph + m.g_vertadr[g] + best + 1) : 0;
  const int* nbr = c.vgraph + m.nmeshvert + 1;
  for (int k = -1; k < nb1 - nb0 && count < 3; k++) {
    const int vi = k < 0 ? best : RCSB_LDG(nbr + nb0 + k);
    real vl[3] = {RCSB_LDG(verts + 3 * vi), RCSB_LDG(verts + 3 * vi + 1), RCSB_LDG(verts + 3 * vi + 2)}, v[3], pos[3];
    mulmat3(v, pf.R2, vl);
    v[0] += pf.p2[0]; v[1] += pf.p2[1]; v[2] += pf.p2[2];
    real dist = (v[0] - pf.p1[0]) * n[0] + (v[1] - pf.p1[1]) * n[1] + (v[2] - pf.p1[2]) * n[2];
    if (dist > margin) {
      if (k < 0) { *clear_out = dist - margin; return; }
      continue;
    }
    int close = 0;
    for (int t = 0; t < 3; t++)
      if (t < count) {
        real e[3] = {v[0] - taken[t][0], v[1] - taken[t][1], v[2] - taken[t][2]};
        if (dot3(e, e) < thr2) close = 1;
      }
    if (close) continue;
    for (int t = 0; t < 3; t++) if (t == count) copy3(taken[t], v);
    count++;
    for (int kk = 0; kk < 3; kk++) pos[kk] = v[kk] - (real)0.5 * dist * n[kk];
    add_contact(c, ncon, pf.g1, pf.g2, dist, pos, n, margin, gap, count > 1);
  }
}

// append `value` to `list` for every lane with `hit`, preserving lane order; returns the new count (uniform)
RCSB_DEV int compact_append(const Ctx& c, int hit, int value, int count, uint16_t* list) {
#ifdef RCSB_HOST_EMU
  if (hit) { if (count < RCSB_MAXCAND) list[count] = (uint16_t)value; count++; }
  return count;
#else
  unsigned mask = warp_ballot(hit);
  if (hit) {
    int slot = count + __popc(mask & ((1u << c.lane) - 1u));
    if (slot < RCSB_MAXCAND) list[slot] = (uint16_t)value;
  }
  return count + __popc(mask);
#endif
}
// slot of this lane's item in a list that holds `count` entries, lanes with `hit` in lane order; advances count
RCSB_DEV int compact_slot(const Ctx& c, int hit, int& count) {
#ifdef RCSB_HOST_EMU
  int slot = count;
  count += hit ? 1 : 0;
  return slot;
#else
  unsigned mask = warp_ballot(hit);
  int slot = count + __popc(mask & ((1u << c.lane) - 1u));
  count += __popc(mask);
  return slot;
#endif
}
// separating-axis test of two oriented boxes (rotation A/B row-major, centres ca/cb, half sizes ha/hb); 1 = separated.
// *gap receives a lower bound of the distance between the boxes beyond the margin: the widest clearance over the six
// face axes (unit vectors), or 0 when only an edge-edge axis separates them.
RCSB_DEV int obb_separated(const real* A, const real* ca, const real* ha, const real* B, const real* cb, const real* hb,
                           real margin, real* gap) {
  real R[9], AR[9], t[3], dv[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
  mulmatT3(t, A, dv);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
      AR[3 * i + j] = r_abs(R[3 * i + j]) + (real)1e-12;
    }
  real best = 0;
  for (int i = 0; i < 3; i++) {
    real g = r_abs(t[i]) - (ha[i] + hb[0] * AR[3 * i] + hb[1] * AR[3 * i + 1] + hb[2] * AR[3 * i + 2] + margin);
    best = g > best ? g : best;
  }
  for (int j = 0; j < 3; j++) {
    real g = r_abs(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) - (ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j] + hb[j] + margin);
    best = g > best ? g : best;
  }
  *gap = best;
  if (best > 0) return 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      real ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j];
      real rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
      if (r_abs(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb + margin) return 1;
    }
  return 0;
}

// ---- separation budgets (temporal coherence, exact): every collision group (all geom pairs between two bodies) keeps
// a lower bound of the distance between its geoms beyond the contact margin, measured the last time the group went
// through the phases below and reduced after every step by an upper bound of the relative displacement the joints on
// the tree path between the two bodies can have caused (st_integrate). While the budget is positive no pair of the
// group can be in contact, so the group is skipped; the contact set is the one a full pass would produce.
#define RCSB_BUDGET_BIG 1e30f
RCSB_DEV void budget_min(const Ctx& c, float* bud, int g, real gap) {
  float f = (float)(gap * (real)0.999999);  // never above the measured gap
  if (!(f > 0)) f = 0;
#ifdef RCSB_HOST_EMU
  if (f < bud[g]) bud[g] = f;
#else
  atomicMin((unsigned*)(bud + g), __float_as_uint(f));  // non-negative floats order like their bit patterns
#endif
}
RCSB_DEV float reach_decode(uint16_t h) {  // upper half of a float32 bit pattern (rcsb_layout.h: rcsb_reach_encode)
#ifdef RCSB_HOST_EMU
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
#else
  return __uint_as_float((uint32_t)h << 16);
#endif
}
RCSB_DEV void budget_reset(const Ctx& c) {  // positions changed by something other than a step: every group is due
  const RcsbModel& m = CMODEL(c);
  float* bud = (float*)WR(cbud);
  PFOR(g, MD(ngrp)) { bud[g] = 0; }
}
// after the velocities of the step are final: positions move by h * qvel
RCSB_DEV void budget_advance(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  float* bud = (float*)WR(cbud);
  const real h = m.timestep;
  PFOR(g, MD(ngrp)) {
    const uint16_t* reach = m.grp_reach[g];  // zero off the tree path between the group's bodies
    real s = 0;
    for (int j = 0; j < MD(nv); j++) s += r_abs(WR(v)[j]) * (real)reach_decode(reach[j]);
    bud[g] -= (float)(h * s * (real)1.000001) + 1e-6f;
  }
}
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
#define RCSB_COL_PROBE(i)                                                                                               \
  do {                                                                                                                  \
    if (blockIdx.x == 0 && c.lane == 0 && rcsb_trace_step[threadIdx.x >> 5] < RCSB_TRACE_STEPS)                         \
      rcsb_trace_col[rcsb_trace_step[threadIdx.x >> 5]][i][threadIdx.x >> 5] = (unsigned)(clock64() - t_col_);          \
  } while (0)
#else
#define RCSB_COL_PROBE(i) ((void)0)
#endif
RCSB_DEV void st_collision(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  float* bud = (float*)WR(cbud);
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
  const long long t_col_ = clock64();
#endif
  // ---- groups whose separation budget is used up go through the phases; the others cannot be in contact
  uint32_t act[(RCSB_MAXGRP + 31) / 32] = {0, 0};
#ifdef RCSB_HOST_EMU
  for (int g = 0; g < MD(ngrp); g++) if (!(bud[g] > 0)) act[g >> 5] |= 1u << (g & 31);
#else
  for (int w = 0; w * 32 < MD(ngrp); w++) {
    const int g = w * 32 + c.lane;
    act[w] = warp_ballot(g < MD(ngrp) && !(bud[g] > 0));
  }
#endif
  if (c.lane == 0) { WI(misc)[MI_OVERFLOW] = 0; WI(misc)[MI_NCON] = 0; }
  if (!(act[0] | act[1])) {
    RCSB_SYNC();
    return;
  }
  PFOR(g, MD(ngrp)) { if ((act[g >> 5] >> (g & 31)) & 1u) bud[g] = RCSB_BUDGET_BIG; }
  // ---- broad phase: bounding spheres about the local AABB centres (plane: signed distance), one pair per lane,
  //      survivors compacted in pair order
  PFOR(g, MD(ng)) {  // bounding-volume centres for the broad phase
    int b = m.g_body[g];
    real* o = WR(gpos) + 3 * g;
    if (b < 0) copy3(o, m.g_bpos[g]);
    else {
      real v[3];
      mulmat3(v, WR(bmat) + 9 * b, m.g_bpos[g]);
      const real* p = WR(bpos) + 3 * b;
      o[0] = p[0] + v[0]; o[1] = p[1] + v[1]; o[2] = p[2] + v[2];
    }
  }
  RCSB_SYNC();
  RCSB_COL_PROBE(0);
  uint16_t* candA = (uint16_t*)WR(cand);  // candidate lists live in the stage-local union next to the geom centres
  uint16_t* candB = candA + RCSB_MAXCAND;
  int ncandA = 0;
  for (int base = 0; base < MD(npair); base += RCSB_NLANES) {
    // an event has one or two due groups: most 32-pair blocks hold none of their pairs and are skipped as a whole
    if (!((act[0] & m.pair_blk_grps[base >> 5][0]) | (act[1] & m.pair_blk_grps[base >> 5][1]))) continue;
    int p = base + c.lane, hit = 0;
    if (p < MD(npair)) {
      const int grp = m.pair_grp[p];
      if ((act[grp >> 5] >> (grp & 31)) & 1u) {
        int g1 = m.pair[p][0], g2 = m.pair[p][1];
        real margin = m.g_margin[g1] > m.g_margin[g2] ? m.g_margin[g1] : m.g_margin[g2];
        const real *a = WR(gpos) + 3 * g1, *b = WR(gpos) + 3 * g2;
        real d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, gap;
        if (m.g_type[g1] == RCSB_GEOM_PLANE) {
          const real* R = m.g_rot[g1];  // planes are static in all supported scenes
          real n[3] = {R[2], R[5], R[8]};
          gap = dot3(d, n) - (m.g_rbound[g2] + margin);
          hit = !(dot3(d, n) > m.g_rbound[g2] + margin);
        } else {
          real bound = m.g_rbound[g1] + m.g_rbound[g2] + margin;
          hit = !(dot3(d, d) > bound * bound);
          // a single-precision root, shortened by more than its rounding error: the gap only has to be a lower bound
          gap = hit ? (real)0 : (real)(sqrtf((float)dot3(d, d)) * 0.9999997f) - bound;
        }
        if (!hit) budget_min(c, bud, grp, gap);
      }
    }
    ncandA = compact_append(c, hit, p, ncandA, candA);
  }
  if (ncandA > RCSB_MAXCAND) {  // dropped candidates: their groups stay due
    if (c.lane == 0) WI(misc)[MI_WARN] += 1;
    ncandA = RCSB_MAXCAND;
    RCSB_SYNC();
    PFOR(g, MD(ngrp)) { if ((act[g >> 5] >> (g & 31)) & 1u) bud[g] = 0; }
  }
  RCSB_SYNC();
  RCSB_COL_PROBE(1);
  // ---- mid phase: oriented boxes (local AABBs) by separating axes, one candidate per lane. Conservative: a
  //      rejected pair cannot intersect (the hull lies inside its box), so the contact set is unchanged.
  int ncand = 0;
  for (int base = 0; base < ncandA; base += RCSB_NLANES) {
    int ic = base + c.lane, hit = 0, p = 0;
    if (ic < ncandA) {
      p = candA[ic];
      int g1 = m.pair[p][0], g2 = m.pair[p][1];
      real margin = m.g_margin[g1] > m.g_margin[g2] ? m.g_margin[g1] : m.g_margin[g2];
      real p2[3], R2[9], gap = 0;
      geom_frame(c, g2, p2, R2);
      const real* c2 = WR(gpos) + 3 * g2;
      const real* hb = m.g_aabb[g2] + 3;
      if (m.g_type[g1] == RCSB_GEOM_PLANE) {
        const real* R1 = m.g_rot[g1];
        real n[3] = {R1[2], R1[5], R1[8]}, r = 0;
        for (int j = 0; j < 3; j++) r += hb[j] * r_abs(n[0] * R2[j] + n[1] * R2[3 + j] + n[2] * R2[6 + j]);
        real dist = (c2[0] - m.g_pos[g1][0]) * n[0] + (c2[1] - m.g_pos[g1][1]) * n[1] + (c2[2] - m.g_pos[g1][2]) * n[2];
        hit = !(dist - r > margin);
        gap = dist - r - margin;
      } else {
        real p1[3], R1[9];
        geom_frame(c, g1, p1, R1);
        hit = !obb_separated(R1, WR(gpos) + 3 * g1, m.g_aabb[g1] + 3, R2, c2, hb, margin, &gap);
      }
      if (!hit) budget_min(c, bud, m.pair_grp[p], gap);
    }
    ncand = compact_append(c, hit, p, ncand, candB);
  }
  RCSB_SYNC();
  RCSB_COL_PROBE(2);
  // ---- narrow phase, candidates in pair order
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
  const long long t_narrow_ = clock64();
  if (blockIdx.x == 0 && c.lane == 0 && rcsb_trace_step[threadIdx.x >> 5] < RCSB_TRACE_STEPS)  // due groups | broad | mid survivors
    rcsb_trace_aux[rcsb_trace_step[threadIdx.x >> 5]][threadIdx.x >> 5] = (unsigned)(__popc(act[0]) + __popc(act[1])) | ((unsigned)ncandA << 8) | ((unsigned)ncand << 16);
#endif
  int ncon = 0;
  for (int ic = 0; ic < ncand; ic++) {
    int p = candB[ic];
    PairFrames pf;
    pf.g1 = m.pair[p][0]; pf.g2 = m.pair[p][1];
    real margin = m.g_margin[pf.g1] > m.g_margin[pf.g2] ? m.g_margin[pf.g1] : m.g_margin[pf.g2];
    real gap = m.g_gap[pf.g1] > m.g_gap[pf.g2] ? m.g_gap[pf.g1] : m.g_gap[pf.g2];
    pair_frames(c, pf);
    real clear = 0;  // lower bound of the pair's distance beyond the margin when it ends up without a contact
    int t1 = m.g_type[pf.g1], t2 = m.g_type[pf.g2];
    if (t1 == RCSB_GEOM_PLANE) {
      real n[3] = {pf.R1[2], pf.R1[5], pf.R1[8]};
      if (t2 == RCSB_GEOM_MESH) {
        plane_mesh(c, ncon, pf, n, margin, gap, &clear);
      } else if (t2 == RCSB_GEOM_BOX) {
        int cnt = 0;
        real mind = (real)1e30;
        for (int i = 0; i < 8 && cnt < 4; i++) {
          const real* sz = m.g_size[pf.g2];
          real loc[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]}, cw[3], pos[3];
          mulmat3(cw, pf.R2, loc);
          cw[0] += pf.p2[0]; cw[1] += pf.p2[1]; cw[2] += pf.p2[2];
          real dist = (cw[0] - pf.p1[0]) * n[0] + (cw[1] - pf.p1[1]) * n[1] + (cw[2] - pf.p1[2]) * n[2];
          mind = dist < mind ? dist : mind;
          if (dist > margin) continue;
          for (int k = 0; k < 3; k++) pos[k] = cw[k] - (real)0.5 * dist * n[k];
          add_contact(c, ncon, pf.g1, pf.g2, dist, pos, n, margin, gap, cnt > 0);
          cnt++;
        }
        if (cnt == 0) clear = mind - margin;
      } else if (t2 == RCSB_GEOM_CAPSULE) {
        real r = m.g_size[pf.g2][0], hl = m.g_size[pf.g2][1];
        real ax[3] = {pf.R2[2] * hl, pf.R2[5] * hl, pf.R2[8] * hl};
        real mind = (real)1e30;
        int cnt = 0;
        for (int s = 0; s < 2; s++) {
          real cw[3], pos[3];
          for (int k = 0; k < 3; k++) cw[k] = pf.p2[k] + (s == 0 ? ax[k] : -ax[k]);
          real dist = (cw[0] - pf.p1[0]) * n[0] + (cw[1] - pf.p1[1]) * n[1] + (cw[2] - pf.p1[2]) * n[2] - r;
          mind = dist < mind ? dist : mind;
          if (dist > margin) continue;
          for (int k = 0; k < 3; k++) pos[k] = cw[k] - n[k] * (r + (real)0.5 * dist);
          add_contact(c, ncon, pf.g1, pf.g2, dist, pos, n, margin, gap, cnt > 0);
          cnt++;
        }
        if (cnt == 0) clear = mind - margin;
      }
    } else if (t1 == RCSB_GEOM_BOX && t2 == RCSB_GEOM_BOX) {
      box_box(c, ncon, pf, margin, gap, &clear);
    } else {
      real depth, dir[3], pos[3];
      // depth < 1e-12: exactly touching pair (finger pads at qpos0), not a constraint; see oracle/mj_collision.c
      // Separating-direction cache (RCSB_SEPSLOTS direct-mapped slots in the persistent state row): a direction that
      // separated the pair on an earlier step is re-checked with one support pair; disjoint pairs stay disjoint for many
      // steps (link5 / link7 overlap in their boxes in every pose), so the full query only runs when a cached direction
      // stops separating.
      real* sc = WR(sepcache) + 4 * (p & (RCSB_SEPSLOTS - 1));
      int skip = 0;
      if (sc[0] == (real)(p + 1)) {
        Sup& s = ((Sup*)WR(sup))[4];
        mink_support(c, pf, sc + 1, s);
        skip = dot3(s.v, sc + 1) <= 0;
        // the Minkowski difference reaches at most dot(s.v, dir) <= 0 along the unit direction: the geoms are at least
        // that far apart
        if (skip) clear = -dot3(s.v, sc + 1) - margin;
      }
      if (!skip) {
        real sep[4];
        int hit = mpr_penetration(c, pf, &depth, dir, pos, sep);
        if (hit && depth >= (real)1e-12) add_contact(c, ncon, pf.g1, pf.g2, -depth, pos, dir, margin, gap);
        RCSB_SYNC();
        if (c.lane == 0) {
          if (!hit && sep[3] != 0) { sc[0] = (real)(p + 1); sc[1] = sep[0]; sc[2] = sep[1]; sc[3] = sep[2]; }
          else if (sc[0] == (real)(p + 1)) sc[0] = 0;
        }
        RCSB_SYNC();
      }
    }
    if (c.lane == 0) budget_min(c, bud, m.pair_grp[p], clear);
  }
  if (c.lane == 0) WI(misc)[MI_NCON] = ncon;
  RCSB_SYNC();
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
  if (blockIdx.x == 0 && c.lane == 0 && rcsb_trace_step[threadIdx.x >> 5] < RCSB_TRACE_STEPS)
    rcsb_trace[rcsb_trace_step[threadIdx.x >> 5]][9][threadIdx.x >> 5] = (unsigned)(clock64() - t_narrow_);
#endif
}
