// Constraint construction, Newton solver with elliptic cones, noslip pass, actuation, implicitfast
// integration and the RCS callback layer -- the arithmetic the reference reaches through mj_step1's
// mj_makeConstraint and mj_step2 (/root/reference/src/sim/sim.cpp:110-112) plus the callbacks RCS runs
// between and after them (/root/reference/src/sim/sim.cpp:14-61, SimRobot.cpp:156-191,
// SimGripper.cpp:108-151). One environment per warp; see rcsb_dynamics.cuh.

#define EFC(arr) (WR(efc) + (arr)*MD(maxefc))
#define EFCI(arr) (WI(efc) + (arr)*MD(maxefc))

RCSB_DEV void get_impedance(const real* solimp, real pos, real margin, real* imp) {
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = dmin < (real)0.0001 ? (real)0.0001 : (dmin > (real)0.9999 ? (real)0.9999 : dmin);
  dmax = dmax < (real)0.0001 ? (real)0.0001 : (dmax > (real)0.9999 ? (real)0.9999 : dmax);
  if (width < RCSB_MINVAL) width = RCSB_MINVAL;
  mid = mid < (real)0.0001 ? (real)0.0001 : (mid > (real)0.9999 ? (real)0.9999 : mid);
  if (power < 1) power = 1;
  real x = r_abs(pos - margin) / width, y;
  if (x >= 1) { *imp = dmax; return; }
  if (x <= 0) { *imp = dmin; return; }
  if (power == 1) y = x;
  else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);  // the default; avoids pow()
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  *imp = dmin + y * (dmax - dmin);
}

// frame-row . translational Jacobian column k of a world point attached to moving body `body`
RCSB_DEV real jac_dot(const Ctx& c, int body, int k, const real* pos, const real* dirv) {
  const RcsbModel& m = CMODEL(c);
  if (body < 0 || !((m.b_dofmask[body] >> k) & 1u)) return 0;
  const real* cd = WR(cdof) + 6 * k;
  const real* rc = WR(rootcom) + 3 * m.b_root[body];
  real off[3] = {pos[0] - rc[0], pos[1] - rc[1], pos[2] - rc[2]}, t[3];
  cross3(t, cd, off);
  return dirv[0] * (cd[3] + t[0]) + dirv[1] * (cd[4] + t[1]) + dirv[2] * (cd[5] + t[2]);
}

RCSB_DEV void st_make_constraint(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv), maxefc = MD(maxefc);
  const real* q = WR(q);
  int* etype = EFCI(RCSB_EI_TYPE);
  int* eid = EFCI(RCSB_EI_ID);
  int ncon = WI(misc)[MI_NCON];
  int nefc = 0, ne = 0, nf = 0, nl = 0;
  // ---- row table (equality -> friction loss -> limits -> contacts), identical in every lane
  for (int e = 0; e < MD(neq); e++)
    if (m.e_active[e]) {
      if (c.lane == 0) { etype[nefc] = RCSB_EQ; eid[nefc] = e; }
      nefc++; ne++;
    }
  // friction-loss rows (one per dof with frictionloss > 0) and joint-limit rows (dof j, side s -> item 2j+s), one item
  // per lane, compacted in item order
  for (int base = 0; base < nv; base += RCSB_NLANES) {
    int j = base + c.lane;
    int hit = j < nv && m.d_frictionloss[j] > 0;
    int slot = compact_slot(c, hit, nefc);
    if (hit) { etype[slot] = RCSB_FRICTION_DOF; eid[slot] = j; }
  }
  nf = nefc - ne;
  for (int base = 0; base < 2 * nv; base += RCSB_NLANES) {
    int item = base + c.lane, j = item >> 1, side = item & 1, hit = 0;
    if (item < 2 * nv && m.d_limited[j]) {
      real qj = q[m.d_qadr[j]];
      real dist = side ? (m.d_range[j][1] - qj) : (qj - m.d_range[j][0]);
      hit = dist < m.d_margin[j];
    }
    int slot = compact_slot(c, hit, nefc);
    if (hit) {
      if (slot < maxefc) { etype[slot] = RCSB_LIMIT; eid[slot] = item; }
      else WI(misc)[MD(cap_reduced) ? MI_OVERFLOW : MI_WARN] = 1;
    }
  }
  if (nefc > maxefc) nefc = maxefc;
  nl = nefc - ne - nf;
  for (int ci = 0; ci < ncon; ci++) {
    const real* cr = WR(con) + RCSB_C_REALS * ci;
    int* cii = WI(con) + RCSB_CI_INTS * ci;
    int dim = cii[RCSB_CI_DIM];
    int rows = MD(cone_elliptic) ? dim : (dim == 1 ? 1 : 2 * (dim - 1));
    int addr = -1;
    if (cr[RCSB_C_DIST] < cr[RCSB_C_INCMARGIN]) {
      if (nefc + rows <= maxefc) {
        addr = nefc;
        if (c.lane == 0)
          for (int r = 0; r < rows; r++) { etype[nefc + r] = MD(cone_elliptic) ? RCSB_CONTACT_ELL : RCSB_CONTACT_PYR; eid[nefc + r] = ci; }
        nefc += rows;
      } else if (c.lane == 0) {
        WI(misc)[MD(cap_reduced) ? MI_OVERFLOW : MI_WARN] += 1;
      }
    }
    if (c.lane == 0) cii[RCSB_CI_EFC] = addr;
  }
  // does any constraint row couple two kinematic trees? (while none does, the constraint Hessian stays block diagonal by
  // tree like M itself, and the solver factors the blocks side by side)
  int coupled = 0;
  if (MD(nroot) > 1) {
    for (int e = 0; e < MD(neq); e++)
      if (m.e_active[e] && m.e_dof2[e] >= 0 && m.d_tree_lo[m.e_dof1[e]] != m.d_tree_lo[m.e_dof2[e]]) coupled = 1;
    for (int ci = 0; ci < ncon; ci++) {
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      const int b1 = m.g_body[cii[RCSB_CI_G0]], b2 = m.g_body[cii[RCSB_CI_G1]];
      if (b1 >= 0 && b2 >= 0 && m.b_root[b1] != m.b_root[b2]) coupled = 1;
    }
  }
  // kinematic trees touched by the rows the noslip pass works on (bits 8.. of the same slot; all of them when there are
  // friction-loss rows)
  unsigned nsroots = 0;
  if (MD(noslip_iterations) > 0 && MD(nroot) > 1) {
    if (nf > 0) nsroots = 0xffu;
    for (int ci = 0; ci < ncon; ci++) {
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      if (cii[RCSB_CI_EFC] < 0 || cii[RCSB_CI_DIM] < 3) continue;
      const int b1 = m.g_body[cii[RCSB_CI_G0]], b2 = m.g_body[cii[RCSB_CI_G1]];
      if (b1 >= 0) nsroots |= 1u << m.b_root[b1];
      if (b2 >= 0) nsroots |= 1u << m.b_root[b2];
    }
  }
  if (c.lane == 0) {
    WI(misc)[MI_NEFC] = nefc; WI(misc)[MI_NE] = ne; WI(misc)[MI_NF] = nf; WI(misc)[MI_NL] = nl;
    WI(misc)[MI_COUPLED] = coupled | (int)(nsroots << 8);
  }
  RCSB_SYNC();
  // ---- Jacobian
  PFOR(e, nefc * nv) {
    int r = e / nv, k = e - r * nv, type = etype[r], id = eid[r];
    real val = 0;
    if (type == RCSB_EQ) {
      if (k == m.e_dof1[id]) val = 1;
      else if (k == m.e_dof2[id]) {
        int qa = m.d_qadr[k];
        real p2 = q[qa] - m.qpos0[qa];
        const real* pc = m.e_poly[id];
        val = -(pc[1] + p2 * (2 * pc[2] + p2 * (3 * pc[3] + p2 * 4 * pc[4])));
      }
    } else if (type == RCSB_FRICTION_DOF) {
      val = k == id;
    } else if (type == RCSB_LIMIT) {
      if (k == (id >> 1)) val = (id & 1) ? (real)-1 : (real)1;
    } else {
      const real* cr = WR(con) + RCSB_C_REALS * id;
      const int* cii = WI(con) + RCSB_CI_INTS * id;
      int b1 = m.g_body[cii[RCSB_CI_G0]], b2 = m.g_body[cii[RCSB_CI_G1]];
      int idx = r - cii[RCSB_CI_EFC];
      if (type == RCSB_CONTACT_ELL) {
        const real* fr = cr + RCSB_C_FRAME + 3 * idx;
        val = jac_dot(c, b2, k, cr + RCSB_C_POS, fr) - jac_dot(c, b1, k, cr + RCSB_C_POS, fr);
      } else {
        const real* fn = cr + RCSB_C_FRAME;
        const real* ft = cr + RCSB_C_FRAME + 3 * (1 + idx / 2);
        real jn = jac_dot(c, b2, k, cr + RCSB_C_POS, fn) - jac_dot(c, b1, k, cr + RCSB_C_POS, fn);
        real jt = jac_dot(c, b2, k, cr + RCSB_C_POS, ft) - jac_dot(c, b1, k, cr + RCSB_C_POS, ft);
        val = jn + ((idx & 1) ? (real)-1 : (real)1) * cr[RCSB_C_FRIC] * jt;
      }
    }
    WR(J)[e] = val;
  }
  // ---- per-row position, impedance, regulariser
  PFOR(r, nefc) {
    int type = etype[r], id = eid[r];
    real pos = 0, margin = 0, floss = 0, diag, K, B, imp;
    const real *solref, *solimp;
    int friction_row = 0;
    if (type == RCSB_EQ) {
      int d1 = m.e_dof1[id], d2 = m.e_dof2[id];
      real p1 = q[m.d_qadr[d1]] - m.qpos0[m.d_qadr[d1]];
      const real* pc = m.e_poly[id];
      if (d2 >= 0) {
        real p2 = q[m.d_qadr[d2]] - m.qpos0[m.d_qadr[d2]];
        pos = p1 - (pc[0] + p2 * (pc[1] + p2 * (pc[2] + p2 * (pc[3] + p2 * pc[4]))));
        diag = m.d_invweight0[d1] + m.d_invweight0[d2];
      } else {
        pos = p1 - pc[0];
        diag = m.d_invweight0[d1];
      }
      solref = m.e_solref[id]; solimp = m.e_solimp[id];
    } else if (type == RCSB_FRICTION_DOF) {
      floss = m.d_frictionloss[id];
      diag = m.d_invweight0[id];
      solref = m.d_solref[id]; solimp = m.d_solimp[id];  // host fills defaults for friction rows
      friction_row = 1;
    } else if (type == RCSB_LIMIT) {
      int j = id >> 1;
      real qj = q[m.d_qadr[j]];
      pos = (id & 1) ? (m.d_range[j][1] - qj) : (qj - m.d_range[j][0]);
      margin = m.d_margin[j];
      diag = m.d_invweight0[j];
      solref = m.d_solref[j]; solimp = m.d_solimp[j];
    } else {
      const real* cr = WR(con) + RCSB_C_REALS * id;
      const int* cii = WI(con) + RCSB_CI_INTS * id;
      int first = (r == cii[RCSB_CI_EFC]);
      diag = m.g_invweight[cii[RCSB_CI_G0]] + m.g_invweight[cii[RCSB_CI_G1]];
      solref = cr + RCSB_C_SOLREF; solimp = cr + RCSB_C_SOLIMP;
      if (type == RCSB_CONTACT_ELL) {
        if (first) { pos = cr[RCSB_C_DIST]; margin = cr[RCSB_C_INCMARGIN]; } else friction_row = 1;
      } else {
        pos = cr[RCSB_C_DIST]; margin = cr[RCSB_C_INCMARGIN];
        diag *= 1 + cr[RCSB_C_FRIC] * cr[RCSB_C_FRIC];
      }
    }
    if (diag < RCSB_MINVAL) diag = RCSB_MINVAL;
    get_impedance(solimp, pos, margin, &imp);
    real dmax = solimp[1];
    dmax = dmax < (real)0.0001 ? (real)0.0001 : (dmax > (real)0.9999 ? (real)0.9999 : dmax);
    if (solref[0] > 0) {
      real tc = solref[0], dr = solref[1];
      if (tc < 2 * m.timestep) tc = 2 * m.timestep;
      K = 1 / (dmax * dmax * tc * tc * dr * dr);
      B = 2 / (dmax * tc);
    } else {
      K = -solref[0] / (dmax * dmax);
      B = -solref[1] / dmax;
    }
    if (friction_row) K = 0;
    real R = (1 - imp) / imp * diag;
    if (R < RCSB_MINVAL) R = RCSB_MINVAL;
    EFC(RCSB_E_POS)[r] = pos; EFC(RCSB_E_MARGIN)[r] = margin; EFC(RCSB_E_FLOSS)[r] = floss;
    EFC(RCSB_E_R)[r] = R;
    EFC(RCSB_E_K)[r] = B;                             // damping coefficient, consumed below
    EFC(RCSB_E_AREF)[r] = K * imp * (pos - margin);   // stiffness term, consumed below
  }
  RCSB_SYNC();
  if (MD(cone_elliptic)) {
    PFOR(ci, ncon) {  // friction rows of elliptic cones: R tied to the normal row through impratio
      real* cr = WR(con) + RCSB_C_REALS * ci;
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      int a = cii[RCSB_CI_EFC], dim = cii[RCSB_CI_DIM];
      if (a >= 0 && dim >= 3) {
        real ir = m.impratio < RCSB_MINVAL ? RCSB_MINVAL : m.impratio;
        real* R = EFC(RCSB_E_R);
        R[a + 1] = R[a] / ir;
        cr[RCSB_C_MU] = cr[RCSB_C_FRIC] * r_sqrt(R[a + 1] / R[a]);
        R[a + 2] = R[a + 1];  // condim 3: both tangential friction coefficients are equal
      }
    }
    RCSB_SYNC();
  } else {
    PFOR(ci, ncon) {  // pyramidal cones: every edge gets Rpy = 2 mu^2 R_first, mu = friction / sqrt(impratio)
      real* cr = WR(con) + RCSB_C_REALS * ci;
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      int a = cii[RCSB_CI_EFC], dim = cii[RCSB_CI_DIM];
      if (a >= 0 && dim >= 3) {
        real ir = m.impratio < RCSB_MINVAL ? RCSB_MINVAL : m.impratio;
        real mu = cr[RCSB_C_FRIC] / r_sqrt(ir);
        cr[RCSB_C_MU] = mu;
        real Rpy = 2 * mu * mu * EFC(RCSB_E_R)[a];
        if (Rpy < RCSB_MINVAL) Rpy = RCSB_MINVAL;
        for (int j = 0; j < 2 * (dim - 1); j++) EFC(RCSB_E_R)[a + j] = Rpy;
      }
    }
    RCSB_SYNC();
  }
  PFOR(r, nefc) {
    real vel = 0;
    for (int k = 0; k < nv; k++) vel += WR(J)[r * nv + k] * WR(v)[k];
    EFC(RCSB_E_D)[r] = 1 / EFC(RCSB_E_R)[r];
    EFC(RCSB_E_AREF)[r] = -EFC(RCSB_E_K)[r] * vel - EFC(RCSB_E_AREF)[r];
  }
  RCSB_SYNC();
}

// ------------------------------------------------------------------ actuation and smooth acceleration
RCSB_DEV void st_actuation(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  PFOR(a, MD(nu)) {
    real len, vel;
    if (m.a_trntype[a] == RCSB_TRN_JOINT) {
      int d = m.a_trnid[a];
      len = m.a_gear[a] * WR(q)[m.d_qadr[d]];
      vel = m.a_gear[a] * WR(v)[d];
    } else {
      const real* co = m.t_coef[m.a_trnid[a]];
      len = 0; vel = 0;
      for (int k = 0; k < nv; k++)
        if (co[k] != 0) { len += co[k] * WR(q)[m.d_qadr[k]]; vel += co[k] * WR(v)[k]; }
      len *= m.a_gear[a]; vel *= m.a_gear[a];
    }
    real ctrl = WR(ctrl)[a];
    if (m.a_ctrllimited[a]) ctrl = ctrl < m.a_ctrlrange[a][0] ? m.a_ctrlrange[a][0] : (ctrl > m.a_ctrlrange[a][1] ? m.a_ctrlrange[a][1] : ctrl);
    real f = m.a_gain[a] * ctrl + m.a_bias[a][0] + m.a_bias[a][1] * len + m.a_bias[a][2] * vel;
    if (m.a_forcelimited[a]) f = f < m.a_forcerange[a][0] ? m.a_forcerange[a][0] : (f > m.a_forcerange[a][1] ? m.a_forcerange[a][1] : f);
    WR(aforce)[a] = f;
  }
  RCSB_SYNC();
  PFOR(k, nv) {
    real s = 0;
    for (int a = 0; a < MD(nu); a++) s += m.a_moment[a][k] * WR(aforce)[a];
    if (m.d_actgravcomp[k]) s += WR(gravc)[k];
    if (m.d_actfrclimited[k]) s = s < m.d_actfrcrange[k][0] ? m.d_actfrcrange[k][0] : (s > m.d_actfrcrange[k][1] ? m.d_actfrcrange[k][1] : s);
    WR(actfrc)[k] = s;
    real sm = WR(passive)[k] - WR(bias)[k] + s;
    WR(smooth)[k] = sm;
  }
  RCSB_SYNC();
}
// Cholesky factor of M on demand (o_L holds a copy of M until then)
RCSB_DEV void ensure_chol_M(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  if (!WI(misc)[MI_HAVE_L]) {
    if (c.lane == 0) WI(misc)[MI_HAVE_H2] = 0;  // o_L is about to hold M's factor
    PFOR(e, MD(nv) * MD(nv)) { WR(L)[e] = WR(M)[e]; }
    chol_factor(c, WR(L), WR(L) + MD(nv) * MD(nv), MD(nv), 1);  // M is block diagonal by kinematic tree
    if (c.lane == 0) WI(misc)[MI_HAVE_L] = 1;
    RCSB_SYNC();
  }
}
// qacc_smooth = M^-1 qfrc_smooth (mj_fwdAcceleration); only the general solver path and nefc == 0 need it
RCSB_DEV void compute_qacc_smooth(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  ensure_chol_M(c);
  PFOR(k, MD(nv)) { WR(qacc_smooth)[k] = WR(smooth)[k]; }
  chol_solve_blocks(c, WR(L), WR(L) + MD(nv) * MD(nv), MD(nv), WR(qacc_smooth), WR(tmp));
}

// ------------------------------------------------------------------ constraint cost, forces, states
struct CostOut { real cost, gauss; };

RCSB_DEV real constraint_update(const Ctx& c, int nefc, int ncon, int want_hess) {
  const RcsbModel& m = CMODEL(c);
  const real* jar = EFC(RCSB_E_JAR);
  real* force = EFC(RCSB_E_FORCE);
  int* state = EFCI(RCSB_EI_STATE);
  const int* etype = EFCI(RCSB_EI_TYPE);
  real cost = 0;
  PFOR(r, nefc) {
    int type = etype[r];
    real D = EFC(RCSB_E_D)[r], x = jar[r];
    if (type == RCSB_EQ) {
      force[r] = -D * x; state[r] = RCSB_QUADRATIC; cost += (real)0.5 * D * x * x;
    } else if (type == RCSB_FRICTION_DOF) {
      real f = EFC(RCSB_E_FLOSS)[r], R = EFC(RCSB_E_R)[r];
      if (x <= -R * f) { force[r] = f; state[r] = RCSB_LINEARNEG; cost += (real)-0.5 * R * f * f - f * x; }
      else if (x >= R * f) { force[r] = -f; state[r] = RCSB_LINEARPOS; cost += (real)-0.5 * R * f * f + f * x; }
      else { force[r] = -D * x; state[r] = RCSB_QUADRATIC; cost += (real)0.5 * D * x * x; }
    } else if (type != RCSB_CONTACT_ELL) {
      if (x < 0) { force[r] = -D * x; state[r] = RCSB_QUADRATIC; cost += (real)0.5 * D * x * x; }
      else { force[r] = 0; state[r] = RCSB_SATISFIED; }
    }
  }
  if (MD(cone_elliptic)) {
    PFOR(ci, ncon) {
      const real* cr = WR(con) + RCSB_C_REALS * ci;
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      int i = cii[RCSB_CI_EFC], dim = cii[RCSB_CI_DIM];
      if (i < 0) continue;
      real mu = cr[RCSB_C_MU], fr = cr[RCSB_C_FRIC], U[3] = {jar[i] * mu, 0, 0}, T2 = 0;
      for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * fr; T2 += U[j] * U[j]; }
      real N = U[0], T = r_sqrt(T2);
      real* H = WR(conehess) + 9 * ci;
      if (want_hess) for (int k = 0; k < 9; k++) H[k] = 0;
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        for (int j = 0; j < dim; j++) { force[i + j] = 0; state[i + j] = RCSB_SATISFIED; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          real Dj = EFC(RCSB_E_D)[i + j], xj = jar[i + j];
          force[i + j] = -Dj * xj; state[i + j] = RCSB_QUADRATIC; cost += (real)0.5 * Dj * xj * xj;
        }
      } else {
        real Dm = EFC(RCSB_E_D)[i] / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += (real)0.5 * Dm * NmT * NmT;
        real f0 = -Dm * NmT * mu;
        force[i] = f0;
        for (int j = 1; j < dim; j++) force[i + j] = -f0 / T * U[j] * fr;
        for (int j = 0; j < dim; j++) state[i + j] = RCSB_CONE;
        if (want_hess) {
          real g[3] = {mu, 0, 0};
          for (int j = 1; j < dim; j++) g[j] = -mu * fr * U[j] / T;
          for (int a = 0; a < dim; a++)
            for (int b = 0; b < dim; b++) H[3 * a + b] = Dm * g[a] * g[b];
          for (int a = 1; a < dim; a++)
            for (int b = 1; b < dim; b++) {
              real d2T = fr * fr * ((a == b ? (real)1 : (real)0) / T - U[a] * U[b] / (T * T * T));
              H[3 * a + b] += Dm * NmT * (-mu) * d2T;
            }
        }
      }
    }
  }
  RCSB_SYNC();
  return warp_sum(cost);
}

// cost at acceleration vector `acc` (shared memory): fills Ma, jar, force, state; returns total cost
RCSB_DEV_NOINLINE real total_cost(const Ctx& c, const real* acc, int nefc, int ncon, int want_hess, real* gauss_out) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  real g = 0;
#ifndef RCSB_HOST_EMU
  __builtin_assume(__isShared(acc));
#endif
  PFOR(i, nv) {
    real s = 0;
    for (int j = 0; j < nv; j++) s += WR(M)[i * nv + j] * acc[j];
    WR(Ma)[i] = s;
    g += (real)0.5 * (s - WR(smooth)[i]) * (acc[i] - WR(qacc_smooth)[i]);
  }
  PFOR(r, nefc) {
    real s = 0;
    for (int k = 0; k < nv; k++) s += WR(J)[r * nv + k] * acc[k];
    EFC(RCSB_E_JAR)[r] = s - EFC(RCSB_E_AREF)[r];
  }
  RCSB_SYNC();
  g = warp_sum(g);
  real cost = constraint_update(c, nefc, ncon, want_hess);
  if (gauss_out) *gauss_out = g;
  return cost + g;
}

// value, first and second derivative of the cost along the search direction at step alpha
RCSB_DEV_NOINLINE void line_eval(const Ctx& c, int nefc, int ncon, real alpha, real qG0, real qG1, real qG2, real* val, real* d1,
                        real* d2) {
  const RcsbModel& m = CMODEL(c);
  const real* jar = EFC(RCSB_E_JAR);
  const real* Jv = EFC(RCSB_E_JV);
  const int* etype = EFCI(RCSB_EI_TYPE);
  real v = 0, g1 = 0, g2 = 0;
  PFOR(r, nefc) {
    int type = etype[r];
    real D = EFC(RCSB_E_D)[r], x = jar[r] + alpha * Jv[r], jv = Jv[r];
    if (type == RCSB_EQ) {
      v += (real)0.5 * D * x * x; g1 += D * x * jv; g2 += D * jv * jv;
    } else if (type == RCSB_FRICTION_DOF) {
      real f = EFC(RCSB_E_FLOSS)[r], R = EFC(RCSB_E_R)[r];
      if (x <= -R * f) { v += (real)-0.5 * R * f * f - f * x; g1 += -f * jv; }
      else if (x >= R * f) { v += (real)-0.5 * R * f * f + f * x; g1 += f * jv; }
      else { v += (real)0.5 * D * x * x; g1 += D * x * jv; g2 += D * jv * jv; }
    } else if (type != RCSB_CONTACT_ELL) {
      if (x < 0) { v += (real)0.5 * D * x * x; g1 += D * x * jv; g2 += D * jv * jv; }
    }
  }
  if (MD(cone_elliptic)) {
    PFOR(ci, ncon) {
      const real* cr = WR(con) + RCSB_C_REALS * ci;
      const int* cii = WI(con) + RCSB_CI_INTS * ci;
      int i = cii[RCSB_CI_EFC], dim = cii[RCSB_CI_DIM];
      if (i < 0) continue;
      real mu = cr[RCSB_C_MU], fr = cr[RCSB_C_FRIC], U[3], V[3], T2 = 0;
      U[0] = (jar[i] + alpha * Jv[i]) * mu; V[0] = Jv[i] * mu;
      for (int j = 1; j < dim; j++) {
        U[j] = (jar[i + j] + alpha * Jv[i + j]) * fr;
        V[j] = Jv[i + j] * fr;
        T2 += U[j] * U[j];
      }
      real N = U[0], T = r_sqrt(T2);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          real xj = jar[i + j] + alpha * Jv[i + j], Dj = EFC(RCSB_E_D)[i + j], jv = Jv[i + j];
          v += (real)0.5 * Dj * xj * xj; g1 += Dj * xj * jv; g2 += Dj * jv * jv;
        }
      } else {
        real Dm = EFC(RCSB_E_D)[i] / (mu * mu * (1 + mu * mu)), NmT = N - mu * T, UV = 0, VV = 0;
        for (int j = 1; j < dim; j++) { UV += U[j] * V[j]; VV += V[j] * V[j]; }
        real T1 = UV / T, T2d = VV / T - UV * UV / (T * T * T), N1 = V[0];
        v += (real)0.5 * Dm * NmT * NmT;
        g1 += Dm * NmT * (N1 - mu * T1);
        g2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2d));
      }
    }
  }
  v = warp_sum(v); g1 = warp_sum(g1); g2 = warp_sum(g2);
  *val = v + alpha * alpha * qG2 + alpha * qG1 + qG0;
  *d1 = g1 + 2 * alpha * qG2 + qG1;
  *d2 = g2 + 2 * qG2;
}

RCSB_DEV real line_search(const Ctx& c, int nefc, int ncon, real qG0, real qG1, real qG2, real gtol, int maxiter) {
  real v0, d10, d20, v, d1, d2;
  line_eval(c, nefc, ncon, 0, qG0, qG1, qG2, &v0, &d10, &d20);
  if (d10 >= 0 || d20 <= 0) return 0;
  real lo = 0, hi = -1, alpha = -d10 / d20;
  for (int it = 0; it < maxiter; it++) {
    line_eval(c, nefc, ncon, alpha, qG0, qG1, qG2, &v, &d1, &d2);
    if (r_abs(d1) < gtol) break;
    if (d1 < 0) lo = alpha; else hi = alpha;
    real next = (d2 > 0) ? alpha - d1 / d2 : (real)-1;
    if (hi > 0) {
      if (!(next > lo && next < hi)) next = (real)0.5 * (lo + hi);
      if (hi - lo < (real)1e-15 * (1 + r_abs(hi))) { alpha = next; break; }
    } else if (!(next > lo)) {
      next = 2 * alpha + (real)1e-12;
    }
    alpha = next;
  }
  line_eval(c, nefc, ncon, alpha, qG0, qG1, qG2, &v, &d1, &d2);
  if (v > v0) return 0;
  return alpha;
}

RCSB_DEV void compute_qfc(const Ctx& c, int nefc) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  PFOR(k, nv) {
    real s = 0;
    for (int r = 0; r < nefc; r++) s += WR(J)[r * nv + k] * EFC(RCSB_E_FORCE)[r];
    WR(qfc)[k] = s;
    WR(grad)[k] = WR(Ma)[k] - WR(smooth)[k] - s;
  }
  RCSB_SYNC();
}

// smallest range of dofs, made of whole kinematic trees, that holds every non-zero of a constraint Jacobian row: M^-1
// applied to the row stays inside it (M is block diagonal by tree), so the triangular solves can skip the rest
RCSB_DEV void row_dof_range(const Ctx& c, const real* row, int* lo_out, int* hi_out) {
  const RcsbModel& m = CMODEL(c);
  int lo = MD(nv), hi = 0;
  PFOR(k, MD(nv)) {
    if (row[k] != 0) {
      lo = m.d_tree_lo[k] < lo ? m.d_tree_lo[k] : lo;
      hi = m.d_tree_hi[k] > hi ? m.d_tree_hi[k] : hi;
    }
  }
#ifndef RCSB_HOST_EMU
  for (int o = 16; o > 0; o >>= 1) {
    int l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
#endif
  if (hi <= lo) { lo = 0; hi = MD(nv); }
  *lo_out = lo; *hi_out = hi;
}
// ------------------------------------------------------------------ noslip post-pass
// roots == 0: o_L holds M's factor and qacc_smooth is up to date. roots != 0 (bit r = kinematic tree r): o_L holds the
// integrator matrix's factor, whose blocks of the trees in `roots` are M's own; every friction row lies inside those trees,
// so only their dofs are read and written (the other trees keep the acceleration the solver gave them: their constraint
// forces do not change).
RCSB_DEV_NOINLINE void solve_noslip(const Ctx& c, int nefc, int ncon, unsigned roots) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  int ne = WI(misc)[MI_NE], nf = WI(misc)[MI_NF];
  int any = nf > 0;
  for (int ci = 0; ci < ncon; ci++)
    if (WI(con)[RCSB_CI_INTS * ci + RCSB_CI_EFC] >= 0 && WI(con)[RCSB_CI_INTS * ci + RCSB_CI_DIM] >= 3) any = 1;
  if (!any) return;
  if (roots) {  // qacc_smooth of the touched trees (the entries of the other trees are not M^-1 qfrc_smooth: never read)
    PFOR(k, nv) { WR(qacc_smooth)[k] = WR(smooth)[k]; }
    chol_solve_blocks(c, WR(L), WR(L) + nv * nv, nv, WR(qacc_smooth), WR(tmp));
  }
  // unregularised residual offsets b = J*qacc_smooth - aref
  PFOR(r, nefc) {
    real s = 0;
    for (int k = 0; k < nv; k++) s += WR(J)[r * nv + k] * WR(qacc_smooth)[k];
    EFC(RCSB_E_B)[r] = s - EFC(RCSB_E_AREF)[r];
  }
  // M^-1 J^T of every friction row: [friction-loss rows | 2 rows per contact]
  real* MinvJ = WR(noslip);
  real* Ablk = MinvJ + (nf + 2 * MD(maxcon)) * nv;  // 1 value per dof-friction row, 4 per contact
  PFOR(e, nf * nv) { MinvJ[e] = WR(J)[ne * nv + e]; }
  for (int ci = 0; ci < ncon; ci++) {
    const int a = WI(con)[RCSB_CI_INTS * ci + RCSB_CI_EFC];
    const int live = a >= 0 && WI(con)[RCSB_CI_INTS * ci + RCSB_CI_DIM] >= 3;
    PFOR(e, 2 * nv) { MinvJ[(nf + 2 * ci) * nv + e] = live ? WR(J)[(a + 1) * nv + e] : (real)0; }
  }
  chol_solve_multi(c, WR(L), WR(L) + nv * nv, nv, MinvJ, nf + 2 * ncon, WR(tmp));  // all friction rows side by side
  RCSB_SYNC();
  PFOR(i, nf) {
    real s = 0;
    for (int k = 0; k < nv; k++) s += WR(J)[(ne + i) * nv + k] * MinvJ[i * nv + k];
    Ablk[i] = s;
  }
  PFOR(e, ncon * 4) {
    int ci = e >> 2, r = (e >> 1) & 1, s2 = e & 1;
    int a = WI(con)[RCSB_CI_INTS * ci + RCSB_CI_EFC];
    real s = 0;
    if (a >= 0 && WI(con)[RCSB_CI_INTS * ci + RCSB_CI_DIM] >= 3)
      for (int k = 0; k < nv; k++) s += WR(J)[(a + 1 + r) * nv + k] * MinvJ[(nf + 2 * ci + s2) * nv + k];
    Ablk[nf + e] = s;
  }
  compute_qfc(c, nefc);  // qfc = J^T f for the current forces
  real scale = (real)1 / (m.meaninertia * (nv > 1 ? nv : 1));
  real* f = EFC(RCSB_E_FORCE);
  for (int iter = 0; iter < MD(noslip_iterations); iter++) {
    real improvement = 0;
    for (int i = 0; i < nf; i++) {
      int r = ne + i;
      real res = EFC(RCSB_E_B)[r];
      for (int k = 0; k < nv; k++) res += MinvJ[i * nv + k] * WR(qfc)[k];
      real Aii = Ablk[i];
      if (Aii < RCSB_MINVAL) continue;
      real old = f[r], fn = old - res / Aii, fl = EFC(RCSB_E_FLOSS)[r];
      fn = fn < -fl ? -fl : (fn > fl ? fl : fn);
      real df = fn - old;
      improvement -= (real)0.5 * df * df * Aii + df * res;
      RCSB_SYNC();
      if (c.lane == 0) f[r] = fn;
      PFOR(k, nv) { WR(qfc)[k] += WR(J)[r * nv + k] * df; }
      RCSB_SYNC();
    }
    for (int ci = 0; ci < ncon; ci++) {
      int a = WI(con)[RCSB_CI_INTS * ci + RCSB_CI_EFC];
      if (a < 0 || WI(con)[RCSB_CI_INTS * ci + RCSB_CI_DIM] < 3 || !MD(cone_elliptic)) continue;
      const real* cr = WR(con) + RCSB_C_REALS * ci;
      real fn = f[a], res[2], old[2] = {f[a + 1], f[a + 2]}, bc[2], v[2] = {0, 0};
      const real* Ac = Ablk + nf + 4 * ci;
      for (int j = 0; j < 2; j++) {
        res[j] = EFC(RCSB_E_B)[a + 1 + j];
        for (int k = 0; k < nv; k++) res[j] += MinvJ[(nf + 2 * ci + j) * nv + k] * WR(qfc)[k];
      }
      if (fn >= RCSB_MINVAL) {
        for (int r = 0; r < 2; r++) bc[r] = res[r] - Ac[2 * r] * old[0] - Ac[2 * r + 1] * old[1];
        real det = Ac[0] * Ac[3] - Ac[1] * Ac[2];
        if (det > (real)1e-10) {
          v[0] = -(Ac[3] * bc[0] - Ac[1] * bc[1]) / det;
          v[1] = -(-Ac[2] * bc[0] + Ac[0] * bc[1]) / det;
        }
        real fr = cr[RCSB_C_FRIC];
        real e = (v[0] * v[0] + v[1] * v[1]) / (fr * fr);
        if (det <= (real)1e-10 || e > fn * fn) {
          // QCQP in two variables: min 0.5 x'Ax + x'b  s.t. |x/fr| <= fn, Newton on the multiplier
          real As[4] = {Ac[0] * fr * fr, Ac[1] * fr * fr, Ac[2] * fr * fr, Ac[3] * fr * fr}, bs[2] = {bc[0] * fr, bc[1] * fr};
          real la = 0, v0 = 0, v1 = 0;
          int ok = 1;
          for (int it = 0; it < 20; it++) {
            real a00 = As[0] + la, a11 = As[3] + la, dt = a00 * a11 - As[1] * As[2];
            if (dt < (real)1e-10) { ok = 0; break; }
            real P00 = a11 / dt, P01 = -As[1] / dt, P10 = -As[2] / dt, P11 = a00 / dt;
            v0 = -P00 * bs[0] - P01 * bs[1];
            v1 = -P10 * bs[0] - P11 * bs[1];
            real val = v0 * v0 + v1 * v1 - fn * fn;
            if (val < (real)1e-10) break;
            real pv0 = P00 * v0 + P01 * v1, pv1 = P10 * v0 + P11 * v1;
            real deriv = -2 * (v0 * pv0 + v1 * pv1);
            real delta = -val / deriv;
            if (delta < (real)1e-10) break;
            la += delta;
          }
          v[0] = ok ? v0 * fr : (real)0;
          v[1] = ok ? v1 * fr : (real)0;
        }
      }
      real df[2] = {v[0] - old[0], v[1] - old[1]};
      improvement -= (real)0.5 * (df[0] * (Ac[0] * df[0] + Ac[1] * df[1]) + df[1] * (Ac[2] * df[0] + Ac[3] * df[1])) + df[0] * res[0] + df[1] * res[1];
      RCSB_SYNC();
      if (c.lane == 0) { f[a + 1] = v[0]; f[a + 2] = v[1]; }
      PFOR(k, nv) { WR(qfc)[k] += WR(J)[(a + 1) * nv + k] * df[0] + WR(J)[(a + 2) * nv + k] * df[1]; }
      RCSB_SYNC();
    }
    if (improvement * scale < m.noslip_tolerance) break;
  }
  if (roots) {
    PFOR(k, nv) { WR(search)[k] = WR(qfc)[k]; }
    chol_solve_blocks(c, WR(L), WR(L) + nv * nv, nv, WR(search), WR(tmp));
    PFOR(k, nv) { if ((roots >> m.b_root[m.d_body[k]]) & 1u) WR(qacc)[k] = WR(search)[k] + WR(qacc_smooth)[k]; }
    RCSB_SYNC();
    return;
  }
  PFOR(k, nv) { WR(qacc)[k] = WR(qfc)[k]; }
  chol_solve_blocks(c, WR(L), WR(L) + nv * nv, nv, WR(qacc), WR(tmp));
  PFOR(k, nv) { WR(qacc)[k] += WR(qacc_smooth)[k]; }
  RCSB_SYNC();
}

// General path: MuJoCo's Newton solver (warm start choice, Hessian with cone blocks, exact line search) and the noslip
// post-pass. Out of line: the direct solve in st_constraint_solve handles almost every step, so this code stays out of
// the hot instruction stream.
RCSB_DEV void build_integrator_matrix(const Ctx& c, real* dst);
RCSB_DEV_NOINLINE void newton_solve(const Ctx& c, int nefc, int ncon, int have_direct) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  compute_qacc_smooth(c);
  // warm start: keep qacc_warmstart only if its cost beats qacc_smooth's. have_direct: o_qacc holds the last direct
  // active-set solution (the exact minimiser for zones that turned out not to be the final ones); when it costs less than
  // both it is the starting point - the minimiser the iteration converges to is the same, it is just reached sooner.
  real gauss;
  real cost_smooth = total_cost(c, WR(qacc_smooth), nefc, ncon, 0, nullptr);
  real cost_direct = have_direct ? total_cost(c, WR(qacc), nefc, ncon, 0, nullptr) : (real)1e300;
  real cost = total_cost(c, WR(warm), nefc, ncon, 1, &gauss);
  const int use_warm = cost < cost_smooth;
  const int use_direct = cost_direct < (use_warm ? cost : cost_smooth);
  if (!use_direct) {
    const real* start = use_warm ? WR(warm) : WR(qacc_smooth);
    PFOR(k, nv) { WR(qacc)[k] = start[k]; }
  }
  RCSB_SYNC();
  real scale = (real)1 / (m.meaninertia * (nv > 1 ? nv : 1));
  if (use_direct || !use_warm) cost = total_cost(c, WR(qacc), nefc, ncon, 1, &gauss);
  int iter = 0;
  const int* state = EFCI(RCSB_EI_STATE);
  while (iter < m.iterations) {
    compute_qfc(c, nefc);
    // Hessian H = M + J^T diag(D_active) J + cone blocks
    PFOR(t, nv * (nv + 1) / 2) {
      const int a = m.tri_i[t], b = m.tri_j[t];
      real h = WR(M)[a * nv + b];
      for (int r = 0; r < nefc; r++)
        if (state[r] == RCSB_QUADRATIC) h += EFC(RCSB_E_D)[r] * WR(J)[r * nv + a] * WR(J)[r * nv + b];
      for (int ci = 0; ci < ncon; ci++) {
        int i = WI(con)[RCSB_CI_INTS * ci + RCSB_CI_EFC];
        if (i < 0 || state[i] != RCSB_CONE) continue;
        int dim = WI(con)[RCSB_CI_INTS * ci + RCSB_CI_DIM];
        const real* Hc = WR(conehess) + 9 * ci;
        for (int r = 0; r < dim; r++)
          for (int s = 0; s < dim; s++) h += Hc[3 * r + s] * WR(J)[(i + r) * nv + a] * WR(J)[(i + s) * nv + b];
      }
      WR(H)[a * nv + b] = h;
      WR(H)[b * nv + a] = h;
    }
    PFOR(k, nv) { WR(search)[k] = WR(grad)[k]; }
    chol_factor_solve(c, WR(H), WR(H) + nv * nv, nv, WR(search), WR(tmp), nullptr, nullptr, !(WI(misc)[MI_COUPLED] & 1));
    PFOR(k, nv) { WR(search)[k] = -WR(search)[k]; }
    RCSB_SYNC();
    real qG1 = 0, qG2 = 0, sn = 0;
    PFOR(i, nv) {
      real s = 0;
      for (int j = 0; j < nv; j++) s += WR(M)[i * nv + j] * WR(search)[j];
      WR(Mv)[i] = s;
      qG1 += WR(search)[i] * (WR(Ma)[i] - WR(smooth)[i]);
      qG2 += (real)0.5 * WR(search)[i] * s;
      sn += WR(search)[i] * WR(search)[i];
    }
    PFOR(r, nefc) {
      real s = 0;
      for (int k = 0; k < nv; k++) s += WR(J)[r * nv + k] * WR(search)[k];
      EFC(RCSB_E_JV)[r] = s;
    }
    RCSB_SYNC();
    qG1 = warp_sum(qG1); qG2 = warp_sum(qG2); sn = r_sqrt(warp_sum(sn));
    if (sn < RCSB_MINVAL) break;
    real gtol = m.tolerance * m.ls_tolerance * sn / scale;
    real alpha = line_search(c, nefc, ncon, gauss, qG1, qG2, gtol, m.ls_iterations);
    if (alpha == 0) break;
    PFOR(k, nv) { WR(qacc)[k] += alpha * WR(search)[k]; }
    RCSB_SYNC();
    real oldcost = cost;
    cost = total_cost(c, WR(qacc), nefc, ncon, 1, &gauss);
    iter++;
    real gn = 0;
    PFOR(k, nv) {
      real s = 0;
      for (int r = 0; r < nefc; r++) s += WR(J)[r * nv + k] * EFC(RCSB_E_FORCE)[r];
      real g = WR(Ma)[k] - WR(smooth)[k] - s;
      gn += g * g;
    }
    gn = warp_sum(gn);
    real improvement = scale * (oldcost - cost), gradient = scale * r_sqrt(gn);
    if (improvement < m.tolerance || gradient < m.tolerance) break;
  }
  if (c.lane == 0) WI(misc)[MI_SOLVER_ITER] = iter;
  compute_qfc(c, nefc);
  PFOR(k, nv) { WR(warm)[k] = WR(qacc)[k]; }  // mj_fwdConstraint saves the warm start before the noslip post-pass
  RCSB_SYNC();
  if (MD(noslip_iterations) > 0) solve_noslip(c, nefc, ncon, 0);
}

// ------------------------------------------------------------------ constrained acceleration (Newton)
RCSB_DEV void build_integrator_matrix(const Ctx& c, real* dst);
RCSB_DEV void st_constraint_solve(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  int nefc = WI(misc)[MI_NEFC], ncon = WI(misc)[MI_NCON];
  if (nefc == 0) {
    compute_qacc_smooth(c);
    PFOR(k, nv) { WR(qacc)[k] = WR(qacc_smooth)[k]; WR(warm)[k] = WR(qacc_smooth)[k]; WR(qfc)[k] = 0; }
    if (c.lane == 0) WI(misc)[MI_SOLVER_ITER] = 0;
    RCSB_SYNC();
    return;
  }
  {
    // Direct active-set solve. Rows: equality / friction-loss / joint-limit, pyramidal contact rows (plain inequality
    // rows) and elliptic contacts in their top (separating: inactive) or bottom (force strictly inside the friction
    // cone, i.e. sticking: every row of the contact a plain quadratic) zone; a contact that lands on the cone surface
    // hands the problem to the Newton solver. The cost is strictly convex and piecewise
    // quadratic in qacc: with the zone of every row fixed (quadratic, linear with constant force, or inactive) the
    // stationarity condition is the linear system
    //   (M + sum_quadratic D J^T J) qacc = qfrc_smooth + sum_quadratic D aref J^T + sum_linear force J^T,
    // and a solution whose rows land in the zones that were assumed satisfies the optimality conditions of the whole
    // problem, i.e. it IS the minimiser the Newton iteration converges to. Zones are guessed (limits active, friction
    // from the warm start), solved, checked, and re-guessed from the solution (at most twice; once with pyramidal cones); all-equality problems
    // verify on the first pass. Only if that fails does the general Newton path below run.
    int* state = EFCI(RCSB_EI_STATE);
    const int* etype = EFCI(RCSB_EI_TYPE);
    PFOR(r, nefc) {  // zones of the previous step's solution (qacc_warmstart) are the first guess
      int st = RCSB_QUADRATIC;
      const int type = etype[r];
      if (type == RCSB_FRICTION_DOF || type == RCSB_LIMIT || type == RCSB_CONTACT_PYR) {
        real x = -EFC(RCSB_E_AREF)[r];
        for (int k = 0; k < nv; k++) x += WR(J)[r * nv + k] * WR(warm)[k];
        if (type == RCSB_FRICTION_DOF) {
          real f = EFC(RCSB_E_FLOSS)[r], R = EFC(RCSB_E_R)[r];
          st = x <= -R * f ? RCSB_LINEARNEG : (x >= R * f ? RCSB_LINEARPOS : RCSB_QUADRATIC);
        } else {
          st = x < 0 ? RCSB_QUADRATIC : RCSB_SATISFIED;
        }
      }
      state[r] = st;
    }
    RCSB_SYNC();
    // noslip needs M's own factor in o_L afterwards, so the integrator matrix is not factored alongside then
    const int need_noslip = MD(noslip_iterations) > 0 && (WI(misc)[MI_NF] > 0 || ncon > 0);
    // The noslip rows only touch kinematic trees whose block of the integrator's matrix IS M's block (a free object: no
    // damping, no actuator) and no row couples two trees: the noslip pass then works on those blocks of the integrator's
    // factor and M is never factored on its own (one factorisation less per step; the resting-cube steps of pick-up).
    const unsigned nsroots = (unsigned)WI(misc)[MI_COUPLED] >> 8;
    const int ns_plain = need_noslip && MD(nroot) > 1 && !(WI(misc)[MI_COUPLED] & 1) && nsroots != 0 && (nsroots & ~(unsigned)m.root_plain) == 0;
    int verified = 0, attempt = 0;
    // Pyramidal cones put 2 (dim - 1) plain inequality rows on every contact. A sliding or rocking body whose zones do not
    // settle on the second pass almost never settles on a third (0.1 % of the tabletop steps) but oscillates between
    // active sets, so those environments go to the Newton solver one pass earlier (six passes instead: -11 %).
    const int max_attempts = MD(cone_elliptic) ? 3 : 2;
    for (; attempt < max_attempts && !verified; attempt++) {
      PFOR(t, nv * (nv + 1) / 2) {
        const int a = m.tri_i[t], b = m.tri_j[t];
        real h = WR(M)[a * nv + b];
        for (int r = 0; r < nefc; r++)
          if (state[r] == RCSB_QUADRATIC) h += EFC(RCSB_E_D)[r] * WR(J)[r * nv + a] * WR(J)[r * nv + b];
        WR(H)[a * nv + b] = h;
        WR(H)[b * nv + a] = h;
      }
      PFOR(k, nv) {
        real s = WR(smooth)[k];
        for (int r = 0; r < nefc; r++) {
          int st = state[r];
          if (st == RCSB_QUADRATIC) s += EFC(RCSB_E_D)[r] * EFC(RCSB_E_AREF)[r] * WR(J)[r * nv + k];
          else if (st == RCSB_LINEARNEG) s += EFC(RCSB_E_FLOSS)[r] * WR(J)[r * nv + k];
          else if (st == RCSB_LINEARPOS) s -= EFC(RCSB_E_FLOSS)[r] * WR(J)[r * nv + k];
        }
        WR(qacc)[k] = s;
      }
      // the integrator's matrix does not depend on the constraint forces: on the first pass it is factored in the
      // idle half of the warp (o_L is free here: M's own factor is only needed by the Newton / noslip path)
      const int dual = attempt == 0 && nv <= 16 && !WI(misc)[MI_HAVE_L] && (!need_noslip || ns_plain);
      if (dual) build_integrator_matrix(c, WR(L));
      chol_factor_solve(c, WR(H), WR(H) + nv * nv, nv, WR(qacc), WR(tmp), dual ? WR(L) : nullptr, dual ? WR(L) + nv * nv : nullptr,
                        !(WI(misc)[MI_COUPLED] & 1));
      if (dual && c.lane == 0) WI(misc)[MI_HAVE_H2] = 1;
      int changed = 0, on_cone = 0;
      PFOR(r, nefc) {
        real s = 0;
        for (int k = 0; k < nv; k++) s += WR(J)[r * nv + k] * WR(qacc)[k];
        real jar = s - EFC(RCSB_E_AREF)[r], D = EFC(RCSB_E_D)[r], force;
        int type = etype[r], st;
        EFC(RCSB_E_JAR)[r] = jar;
        if (type == RCSB_CONTACT_ELL) continue;  // zone of the whole contact: below
        if (type == RCSB_EQ) { st = RCSB_QUADRATIC; force = -D * jar; }
        else if (type == RCSB_FRICTION_DOF) {
          real f = EFC(RCSB_E_FLOSS)[r], R = EFC(RCSB_E_R)[r];
          if (jar <= -R * f) { st = RCSB_LINEARNEG; force = f; }
          else if (jar >= R * f) { st = RCSB_LINEARPOS; force = -f; }
          else { st = RCSB_QUADRATIC; force = -D * jar; }
        } else {
          if (jar < 0) { st = RCSB_QUADRATIC; force = -D * jar; } else { st = RCSB_SATISFIED; force = 0; }
        }
        if (st != state[r]) changed = 1;
        state[r] = st;
        EFC(RCSB_E_FORCE)[r] = force;
      }
      if (MD(cone_elliptic) && ncon > 0) {
        RCSB_SYNC();
        const real* jar = EFC(RCSB_E_JAR);
        PFOR(ci, ncon) {
          const real* cr = WR(con) + RCSB_C_REALS * ci;
          const int* cii = WI(con) + RCSB_CI_INTS * ci;
          const int i = cii[RCSB_CI_EFC], dim = cii[RCSB_CI_DIM];
          if (i < 0) continue;
          real mu = cr[RCSB_C_MU], fr = cr[RCSB_C_FRIC], N = jar[i] * mu, T2 = 0;
          for (int j = 1; j < dim; j++) { real u = jar[i + j] * fr; T2 += u * u; }
          real T = r_sqrt(T2);
          int st;
          if (N >= mu * T || (T <= 0 && N >= 0)) st = RCSB_SATISFIED;
          else if (mu * N + T <= 0 || (T <= 0 && N < 0)) st = RCSB_QUADRATIC;
          else { st = RCSB_CONE; on_cone = 1; }
          for (int j = 0; j < dim; j++) {
            if (st != state[i + j]) changed = 1;
            state[i + j] = st;
            EFC(RCSB_E_FORCE)[i + j] = st == RCSB_QUADRATIC ? -EFC(RCSB_E_D)[i + j] * jar[i + j] : (real)0;
          }
        }
      }
      verified = !warp_any(changed);
      on_cone = warp_any(on_cone);
      RCSB_SYNC();
      if (on_cone) { verified = 0; break; }
    }
    if (verified) {
      PFOR(k, nv) {
        real s = 0;
        for (int r = 0; r < nefc; r++) s += WR(J)[r * nv + k] * EFC(RCSB_E_FORCE)[r];
        WR(qfc)[k] = s;
      }
      if (c.lane == 0) WI(misc)[MI_SOLVER_ITER] = attempt;
      PFOR(k, nv) { WR(warm)[k] = WR(qacc)[k]; }  // saved before noslip, as mj_fwdConstraint does
      RCSB_SYNC();
      if (need_noslip) {  // the post-pass works on M's own factor and the unconstrained acceleration
        if (ns_plain && WI(misc)[MI_HAVE_H2]) {
          solve_noslip(c, nefc, ncon, nsroots);
        } else {
          compute_qacc_smooth(c);
          solve_noslip(c, nefc, ncon, 0);
        }
      }
      return;
    }
  }
  newton_solve(c, nefc, ncon, 1);
}

// ------------------------------------------------------------------ implicitfast / Euler integration + mj_advance
// M - h * d(passive + actuator force)/d(qvel): the implicitfast system matrix (symmetric; plain M for Euler)
RCSB_DEV void build_integrator_matrix(const Ctx& c, real* dst) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  const real h = m.timestep;
  PFOR(t, nv * (nv + 1) / 2) {
    const int i = m.tri_i[t], j = m.tri_j[t], e = i * nv + j;
    real dv = (i == j) ? -m.d_damping[i] + (MD(implicitfast) ? m.d_kvdiag[i] : (real)0) : (real)0;
    if (MD(implicitfast)) {
      for (int sa = 0; sa < m.n_special; sa++) {
        int a = m.a_special[sa];
        // a joint transmission has one non-zero moment entry: it only reaches the diagonal entry of its own dof
        if (m.a_trntype[a] == RCSB_TRN_JOINT && (i != j || m.a_trnid[a] != i)) continue;
        real bv = m.a_bias[a][2];
        if (bv == 0) continue;
        real fa = WR(aforce)[a];
        if (m.a_forcelimited[a] && (fa <= m.a_forcerange[a][0] || fa >= m.a_forcerange[a][1])) continue;
        dv += bv * m.a_moment[a][i] * m.a_moment[a][j];
      }
    }
    real val = WR(M)[e] - h * dv;
    dst[i * nv + j] = val;
    dst[j * nv + i] = val;
  }
}
RCSB_DEV void st_integrate(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  const int nv = MD(nv);
  const real h = m.timestep;
  PFOR(k, nv) { WR(search)[k] = WR(smooth)[k] + WR(qfc)[k]; }
  if (WI(misc)[MI_HAVE_H2]) {  // factored next to the constraint Hessian (st_constraint_solve)
    chol_solve_blocks(c, WR(L), WR(L) + nv * nv, nv, WR(search), WR(tmp));
  } else {
    build_integrator_matrix(c, WR(H));
    chol_factor_solve(c, WR(H), WR(H) + nv * nv, nv, WR(search), WR(tmp), nullptr, nullptr, 1);
  }
  PFOR(k, nv) { WR(v)[k] += h * WR(search)[k]; }  // qacc_warmstart was saved by the constraint solve, before noslip
  RCSB_SYNC();
  budget_advance(c);  // collision groups: the positions are about to move by h * qvel
  PFOR(b, MD(nb)) {
    int qa = m.b_qadr[b], da = m.b_dadr[b];
    real* q = WR(q);
    const real* v = WR(v);
    if (m.b_jtype[b] == RCSB_JNT_FREE) {
      for (int k = 0; k < 3; k++) q[qa + k] += h * v[da + k];
      real wv[3] = {v[da + 3], v[da + 4], v[da + 5]};
      real ang = norm3(wv) * h;
      if (ang > 0) {
        normalize3(wv);
        real s = sin((real)0.5 * ang), dq[4] = {cos((real)0.5 * ang), wv[0] * s, wv[1] * s, wv[2] * s}, qn[4];
        quat_mul(qn, q + qa + 3, dq);
        quat_normalize(qn);
        q[qa + 3] = qn[0]; q[qa + 4] = qn[1]; q[qa + 5] = qn[2]; q[qa + 6] = qn[3];
      }
    } else {
      q[qa] += h * v[da];
    }
  }
  RCSB_SYNC();
  PFOR(i, MD(nq)) { WR(cbq)[i] = WR(q)[i]; }  // the qpos the separation budgets now refer to
  RCSB_SYNC();
}

// ------------------------------------------------------------------ RCS device layer: callbacks
// Persistent per-env RCS state lives in the workspace region o_rcs (reals: RCSB_S_* tail) and
// oi_rcs (ints: RCSB_I_*). Semantics: sim.cpp:14-61, SimRobot.cpp:156-191, SimGripper.cpp:93-151.
#define RS(i) (WR(rcs)[(i)])
#define RI(i) (CWI(c)[LAY.oi_misc + MI_COUNT + (i)])

RCSB_DEV real gripper_width(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  real w = (WR(q)[m.gr_qadr] - m.gr_min_joint) / (m.gr_max_joint - m.gr_min_joint);
  return w < 0 ? (real)0 : (w > 1 ? (real)1 : w);
}
// all lanes evaluate callbacks redundantly on identical data; lane 0 commits state
RCSB_DEV int run_callback(const Ctx& c, int kind) {
  const RcsbModel& m = CMODEL(c);
  int ret = 0;
  if (kind == RCSB_CB_ARRIVED) {
    real mx = 0;
    for (int i = 0; i < MD(rb_njoints); i++) {
      real e = r_abs(WR(q)[m.rb_qadr[i]] - RS(RCSB_S_TARGET + i));
      mx = e > mx ? e : mx;
    }
    RCSB_SYNC();
    if (c.lane == 0) RI(RCSB_I_ARRIVED) = mx < m.rb_joint_tol;
  } else if (kind == RCSB_CB_MOVING) {
    real mx = 0;
    for (int i = 0; i < MD(rb_njoints); i++) {
      real e = r_abs(WR(q)[m.rb_qadr[i]] - RS(RCSB_S_PREV + i));
      mx = e > mx ? e : mx;
    }
    RCSB_SYNC();
    if (c.lane == 0) {
      for (int i = 0; i < MD(rb_njoints); i++) RS(RCSB_S_PREV + i) = WR(q)[m.rb_qadr[i]];
      RI(RCSB_I_MOVING) = mx > (real)0.0001;
    }
  } else if (kind == RCSB_CB_ROBOT_CONV) {
    ret = !RI(RCSB_I_IK_SUCCESS) ? 1 : (RI(RCSB_I_ARRIVED) && !RI(RCSB_I_MOVING));
  } else if (kind == RCSB_CB_ROBOT_COLL) {
    int ncon = WI(misc)[MI_NCON];
    for (int i = 0; i < ncon && !ret; i++) {
      const int* ci = WI(con) + RCSB_CI_INTS * i;
      if ((m.g_role[ci[RCSB_CI_G0]] | m.g_role[ci[RCSB_CI_G1]]) & RCSB_ROLE_ARM) ret = 1;
    }
    RCSB_SYNC();
    if (c.lane == 0) RI(RCSB_I_COLLISION) = ret;
  } else if (kind == RCSB_CB_GRIP_CONV) {
    real w = gripper_width(c);
    int moving = r_abs(RS(RCSB_S_GLW) - w) > (real)0.001 * (m.gr_max_act - m.gr_min_act);
    RCSB_SYNC();
    if (c.lane == 0) { RI(RCSB_I_G_MOVING) = moving; RS(RCSB_S_GLW) = w; }
    ret = !moving;
  } else if (kind == RCSB_CB_GRIP_COLL) {
    int ncon = WI(misc)[MI_NCON];
    for (int i = 0; i < ncon && !ret; i++) {
      const int* ci = WI(con) + RCSB_CI_INTS * i;
      int r0 = m.g_role[ci[RCSB_CI_G0]], r1 = m.g_role[ci[RCSB_CI_G1]];
      if ((r0 & RCSB_ROLE_FINGER) && (r1 & RCSB_ROLE_FINGER)) continue;
      if (((r0 | r1) & RCSB_ROLE_GRIPPER) && !(r1 & RCSB_ROLE_IGNORED)) ret = 1;
    }
    RCSB_SYNC();
    if (c.lane == 0) RI(RCSB_I_G_COLLISION) = ret;
  }
  RCSB_SYNC();
  return ret;
}
RCSB_DEV int cb_registered(const RcsbModel& m, int kind) {
  if (kind <= RCSB_CB_ROBOT_CONV) return m.rb_register_convergence;
  if (kind == RCSB_CB_ROBOT_COLL) return 1;
  return MD(gr_enabled);
}
RCSB_DEV real cb_period(const RcsbModel& m, int kind) { return kind <= RCSB_CB_ROBOT_COLL ? m.rb_cb_period : m.gr_cb_period; }

// plain callbacks between the two halves of the step (sim.cpp:38-47); clocks compare in double
RCSB_DEV void invoke_callbacks(const Ctx& c, double time) {
  const RcsbModel& m = CMODEL(c);
  for (int kind = RCSB_CB_ARRIVED; kind <= RCSB_CB_MOVING; kind++) {
    if (!cb_registered(m, kind)) continue;
    double dt = time - CCLK(c)[RCSB_D_CBLAST + kind];
    if (dt > (double)cb_period(m, kind)) {
      run_callback(c, kind);
      if (c.lane == 0) CCLK(c)[RCSB_D_CBLAST + kind] = time;
      RCSB_SYNC();
    }
  }
}
// condition callbacks after the step (sim.cpp:49-61): any-list first, then all-list
RCSB_DEV int invoke_condition_callbacks(const Ctx& c, double time) {
  const RcsbModel& m = CMODEL(c);
  const int any_list[2] = {RCSB_CB_ROBOT_COLL, RCSB_CB_GRIP_COLL};
  const int all_list[2] = {RCSB_CB_ROBOT_CONV, RCSB_CB_GRIP_CONV};
  for (int pass = 0; pass < 2; pass++)
    for (int i = 0; i < 2; i++) {
      int kind = pass == 0 ? any_list[i] : all_list[i];
      if (!cb_registered(m, kind)) continue;
      double dt = time - CCLK(c)[RCSB_D_CBLAST + kind];
      if (dt > (double)cb_period(m, kind)) {
        int r = run_callback(c, kind);
        if (c.lane == 0) { RI(RCSB_I_CBRET + kind) = r; CCLK(c)[RCSB_D_CBLAST + kind] = time; }
        RCSB_SYNC();
      }
    }
  int any = 0, all = 1;
  for (int i = 0; i < 2; i++) {
    if (cb_registered(m, any_list[i]) && RI(RCSB_I_CBRET + any_list[i])) any = 1;
    if (cb_registered(m, all_list[i]) && !RI(RCSB_I_CBRET + all_list[i])) all = 0;
  }
  return any || all;
}

// ------------------------------------------------------------------ one physics step (step1; callbacks; step2)
RCSB_DEV int state_is_bad(const Ctx& c) {
  const RcsbModel& m = CMODEL(c);
  int bad = 0;
  PFOR(i, MD(nq)) { real x = WR(q)[i]; if (!(x == x) || x > (real)1e10 || x < (real)-1e10) bad = 1; }
  PFOR(i, MD(nv)) { real x = WR(v)[i]; if (!(x == x) || x > (real)1e10 || x < (real)-1e10) bad = 1; }
  return warp_any(bad);
}
RCSB_DEV void reset_data(const Ctx& c, double* time) {  // mj_resetData
  const RcsbModel& m = CMODEL(c);
  PFOR(i, MD(nq)) { WR(q)[i] = m.qpos0[i]; }
  PFOR(i, MD(nv)) { WR(v)[i] = 0; WR(warm)[i] = 0; }
  PFOR(i, MD(nu)) { WR(ctrl)[i] = 0; }
  if (c.lane == 0) *time = 0;
  budget_reset(c);
  RCSB_SYNC();
}
// RCSB_STAGE: CTA barrier (lockstep launches only) + the stage; the profiling build (-DRCSB_STAGE_TIMING) also
// accumulates clock64() per stage for warp 0 of CTA 0 into rcsb_stage_cycles[].
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
#define RCSB_STAGE(idx, call)                                                                        \
  do {                                                                                               \
    RCSB_STAGE_SYNC(idx);                                                                            \
    long long t0_ = clock64();                                                                       \
    call;                                                                                            \
    long long dt_ = clock64() - t0_;                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) rcsb_stage_cycles[idx] += (unsigned long long)dt_;      \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && rcsb_trace_step[threadIdx.x >> 5] < RCSB_TRACE_STEPS) \
      rcsb_trace[rcsb_trace_step[threadIdx.x >> 5]][idx][threadIdx.x >> 5] = (unsigned)dt_;          \
  } while (0)
#else
#define RCSB_STAGE(idx, call) do { RCSB_STAGE_SYNC(idx); call; } while (0)
#endif
// returns 1 when the step did not fit the reduced workspace layout: nothing persistent was changed, the caller hands
// the environment to the full-capacity launch
RCSB_DEV int physics_step(const Ctx& c, double* time) {
  const RcsbModel& m = CMODEL(c);
  if (state_is_bad(c)) {
    if (c.lane == 0) WI(misc)[MI_WARN] += 1;
    reset_data(c, time);
  }
  // Lockstep (fixed-substep launches): the step is far more straight-line code than the instruction cache holds, so
  // the warps of a CTA are re-aligned by CTA barriers (c.lockstep: bit i = before stage i, bit 9 = end of the step);
  // warps that run the same code together share each fetched line instead of streaming the whole program once per
  // warp. Measured best: ONE barrier per step, right after the collision stage -- the one whose duration varies per
  // environment (collision groups that are due) -- i.e. before the velocity stage (mask 0x010); a barrier before every
  // stage costs 10 %, one before the collision stage 24 %, none at all 45 %.
  // ---- mj_step1
  RCSB_STAGE(0, st_kinematics(c));
  RCSB_STAGE(1, st_com(c));
  RCSB_STAGE(2, st_crb(c));
  RCSB_STAGE(3, st_collision(c));
  RCSB_STAGE(4, st_velocity(c));
  RCSB_STAGE(5, st_make_constraint(c));
  if (MD(cap_reduced) && WI(misc)[MI_OVERFLOW]) {
    for (int i = 6; i < 9; i++) RCSB_STAGE_SYNC(i);  // the barriers of the stages this warp skips
    RCSB_STEP_SYNC();
    return 1;
  }
  // ---- RCS plain callbacks see pre-integration time and qpos
  invoke_callbacks(c, *time);
  // ---- mj_step2
  RCSB_STAGE(6, st_actuation(c));
  RCSB_STAGE(7, st_constraint_solve(c));
  RCSB_STAGE(8, st_integrate(c));
  RCSB_STEP_SYNC();
#if defined(RCSB_STAGE_TIMING) && !defined(RCSB_HOST_EMU)
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) rcsb_trace_step[threadIdx.x >> 5] += 1;
#endif
  if (c.lane == 0) {  // the clock lives in shared memory: one writer
    *time += (double)m.timestep;
    RI(RCSB_I_TOTAL_STEPS) += 1;
  }
  RCSB_SYNC();
  return 0;
}
