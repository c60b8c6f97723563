// Per-environment "programs": what one launch does to one environment, as a sequence of ops that
// restate the reference's device and Gym layers on the warp that owns the environment:
//   Sim::step / step_until_convergence / reset      /root/reference/src/sim/sim.cpp:84-138
//   SimRobot set/get/reset                          /root/reference/src/sim/SimRobot.cpp:114-205
//   SimGripper set/get/reset                        /root/reference/src/sim/SimGripper.cpp:79-165
//   RelativeActionSpace / GripperWrapper / RobotEnv /root/reference/python/rcs/envs/base.py:246-288,469-488,710-735
//   RobotSimWrapper / GripperWrapperSim             /root/reference/python/rcs/envs/sim.py:49-76,125-131

// ------------------------------------------------------------------ Pose math (xyz + quat xyzw), Eigen semantics
struct Quat { real x, y, z, w; };
RCSB_DEV Quat q_norm(Quat q) {
  real n = r_sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  if (n > 0) { q.x /= n; q.y /= n; q.z /= n; q.w /= n; }
  return q;
}
RCSB_DEV Quat q_mul(Quat a, Quat b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
RCSB_DEV void q_rot(Quat q, const real* v, real* out) {
  real qv[3] = {q.x, q.y, q.z}, uv[3], t[3];
  cross3(uv, qv, v);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(t, qv, uv);
  out[0] = v[0] + q.w * uv[0] + t[0]; out[1] = v[1] + q.w * uv[1] + t[1]; out[2] = v[2] + q.w * uv[2] + t[2];
}
RCSB_DEV void q_to_mat(Quat q, real* R) {
  real tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  real twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
       tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
RCSB_DEV Quat q_from_mat(const real* mm) {
  Quat q;
  real t = mm[0] + mm[4] + mm[8];
  if (t > 0) {
    t = r_sqrt(t + 1);
    q.w = (real)0.5 * t;
    t = (real)0.5 / t;
    q.x = (mm[7] - mm[5]) * t; q.y = (mm[2] - mm[6]) * t; q.z = (mm[3] - mm[1]) * t;
  } else {
    int i = 0;
    if (mm[4] > mm[0]) i = 1;
    if (mm[8] > mm[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    real qq[3];
    t = r_sqrt(mm[4 * i] - mm[4 * j] - mm[4 * k] + 1);
    qq[i] = (real)0.5 * t;
    t = (real)0.5 / t;
    q.w = (mm[3 * k + j] - mm[3 * j + k]) * t;
    qq[j] = (mm[3 * j + i] + mm[3 * i + j]) * t;
    qq[k] = (mm[3 * k + i] + mm[3 * i + k]) * t;
    q.x = qq[0]; q.y = qq[1]; q.z = qq[2];
  }
  return q;
}
RCSB_DEV void pose_mul(const real* a, const real* b, real* out) {
  Quat qa = {a[3], a[4], a[5], a[6]}, qb = {b[3], b[4], b[5], b[6]};
  real t[3];
  q_rot(qa, b, t);
  Quat q = q_norm(q_mul(qa, qb));
  out[0] = t[0] + a[0]; out[1] = t[1] + a[1]; out[2] = t[2] + a[2];
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}
RCSB_DEV void pose_inverse(const real* a, real* out) {
  Quat cq = {-a[3], -a[4], -a[5], a[6]};
  real t[3];
  q_rot(cq, a, t);
  Quat q = q_norm(cq);
  out[0] = -t[0]; out[1] = -t[1]; out[2] = -t[2];
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}
RCSB_DEV void pose_xyzrpy(const real* a, real* out6) {  // Eigen eulerAngles(2,1,0) range convention
  Quat q = {a[3], a[4], a[5], a[6]};
  real mm[9];
  q_to_mat(q, mm);
  real r0 = atan2(mm[3], mm[0]);
  real c2 = r_sqrt(mm[8] * mm[8] + mm[7] * mm[7]);
  real r1;
  if (r0 < 0) { r0 += (real)3.14159265358979323846; r1 = atan2(-mm[6], -c2); }
  else r1 = atan2(-mm[6], c2);
  real s1 = sin(r0), c1 = cos(r0);
  real r2 = atan2(s1 * mm[2] - c1 * mm[5], c1 * mm[4] - s1 * mm[1]);
  out6[0] = a[0]; out6[1] = a[1]; out6[2] = a[2];
  out6[3] = r2; out6[4] = r1; out6[5] = r0;
}
// SimRobot::get_cartesian_position: base^-1 * Pose(site_xmat, site_xpos) * tcp_offset
RCSB_DEV void robot_cartesian_position(const Ctx& c, real* pose7) {
  const RcsbModel& m = CMODEL(c);
  const real* sp = WR(rcs) + RCSB_S_SITEPOS;
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

// ------------------------------------------------------------------ state row <-> workspace
RCSB_DEV void load_env(const Ctx& c, const real* sr, const double* sd, const int* si) {
  const RcsbModel& m = CMODEL(c);
  PFOR(i, LAY.nsr) { CW(c)[i] = sr[i]; }
  PFOR(i, RCSB_D_TAIL) { CCLK(c)[i] = sd[i]; }
  PFOR(i, RCSB_I_TAIL) { CWI(c)[LAY.oi_misc + MI_COUNT + i] = si[i]; }
  PFOR(i, MI_COUNT) { CWI(c)[LAY.oi_misc + i] = 0; }
  {  // separation budgets are valid only for the qpos they were advanced to
    int same = 1;
    PFOR(i, MD(nq)) { if (!(WR(q)[i] == WR(cbq)[i])) same = 0; }
    if (!warp_all(same)) budget_reset(c);
  }
  RCSB_SYNC();
}
RCSB_DEV void store_env(const Ctx& c, real* sr, double* sd, int* si) {
  const RcsbModel& m = CMODEL(c);
  RCSB_SYNC();
  if (c.lane == 0) {
    RI(RCSB_I_NCON) = WI(misc)[MI_NCON];
    RI(RCSB_I_NEFC) = WI(misc)[MI_NEFC];
    RI(RCSB_I_SOLVER_ITER) = WI(misc)[MI_SOLVER_ITER];
    RI(RCSB_I_WARN) += WI(misc)[MI_WARN];
  }
  RCSB_SYNC();
  PFOR(i, LAY.nsr) { sr[i] = CW(c)[i]; }
  PFOR(i, RCSB_D_TAIL) { sd[i] = CCLK(c)[i]; }
  PFOR(i, RCSB_I_TAIL) { si[i] = CWI(c)[LAY.oi_misc + MI_COUNT + i]; }
}

// ------------------------------------------------------------------ device-layer ops
RCSB_DEV void op_set_joint_position(const Ctx& c, const real* qd) {  // SimRobot.cpp:123-131
  const RcsbModel& m = CMODEL(c);
  PFOR(i, MD(rb_njoints)) {
    RS(RCSB_S_TARGET + i) = qd[i];
    RS(RCSB_S_PREV + i) = WR(q)[m.rb_qadr[i]];
    WR(ctrl)[m.rb_act[i]] = qd[i];
  }
  if (c.lane == 0) { RI(RCSB_I_MOVING) = 1; RI(RCSB_I_ARRIVED) = 0; }
  RCSB_SYNC();
}
RCSB_DEV void op_set_gripper(const Ctx& c, real width) {  // SimGripper.cpp:79-92 (argument validated on the host)
  const RcsbModel& m = CMODEL(c);
  if (c.lane == 0) {
    RS(RCSB_S_GLCW) = width;
    WR(ctrl)[m.gr_act] = width * (m.gr_max_act - m.gr_min_act) + m.gr_min_act;
  }
  RCSB_SYNC();
}

// the whole per-launch program for the environment loaded in the workspace
// GripperWrapper.action (base.py:721-735): binary (round, clip, grasp() / open()) or continuous (clip, set_normalized_width)
RCSB_DEV void op_gripper_action(const Ctx& c, const RcsbLaunch& L, int env) {
  const RcsbModel& m = CMODEL(c);
  if (!(L.ops & (RCSB_OP_ACT_GRIPPER_BIN | RCSB_OP_ACT_GRIPPER_CONT)) || !MD(gr_enabled)) return;
  real g = L.act_gripper[(size_t)env * L.act_gstride];
  if (L.ops & RCSB_OP_ACT_GRIPPER_BIN) g = rint(g);
  g = g < 0 ? (real)0 : (g > 1 ? (real)1 : g);
  op_set_gripper(c, (L.ops & RCSB_OP_ACT_GRIPPER_BIN) ? (g == 0 ? (real)0 : (real)1) : g);
  if (c.lane == 0) RS(RCSB_S_GCMD) = g;
  RCSB_SYNC();
}
RCSB_DEV void run_env_pre_ops(const Ctx& c, const RcsbLaunch& L, int env) {
  const RcsbModel& m = CMODEL(c);
  const unsigned ops = L.ops;
  // ops that write qpos directly invalidate the collision groups' separation budgets
  if (ops & (RCSB_OP_GRIPPER_RESET | RCSB_OP_SIM_RESET | RCSB_OP_ROBOT_RESET | RCSB_OP_SET_JOINTS_HARD)) budget_reset(c);
  if ((ops & RCSB_OP_GRIPPER_RESET) && MD(gr_enabled)) {  // SimGripper.cpp:158-163
    if (c.lane == 0) {
      RS(RCSB_S_GLCW) = 0; RS(RCSB_S_GLW) = 0; RI(RCSB_I_G_MOVING) = 0; RI(RCSB_I_G_COLLISION) = 0;
      WR(q)[m.gr_qadr] = m.gr_max_joint;
      WR(ctrl)[m.gr_act] = m.gr_max_act;
    }
    RCSB_SYNC();
  }
  if (ops & RCSB_OP_ENV_RESET_FLAGS) {
    if (c.lane == 0) RS(RCSB_S_GCMD) = -1;
    RCSB_SYNC();
  }
  if (ops & RCSB_OP_SIM_RESET) {  // sim.cpp:117-138
    reset_data(c, &CCLK(c)[RCSB_D_TIME]);
    PFOR(i, RCSB_NCB) { CCLK(c)[RCSB_D_CBLAST + i] = 0; }
    RCSB_SYNC();
  }
  if (ops & RCSB_OP_ROBOT_RESET) {  // SimRobot.cpp:193-205
    PFOR(i, MD(rb_njoints)) { WR(q)[m.rb_qadr[i]] = m.rb_q_home[i]; WR(ctrl)[m.rb_act[i]] = m.rb_q_home[i]; }
    RCSB_SYNC();
  }
  if (ops & RCSB_OP_SET_JOINTS_HARD) {
    PFOR(i, MD(rb_njoints)) {
      real v = L.act_joints[(size_t)env * L.act_jstride + i];
      WR(q)[m.rb_qadr[i]] = v; WR(ctrl)[m.rb_act[i]] = v;
    }
    RCSB_SYNC();
  }
  if (ops & (RCSB_OP_ACT_JOINTS_REL | RCSB_OP_ACT_JOINTS_ABS)) {
    real* jt = WR(tmp);
    PFOR(i, MD(rb_njoints)) {
      real a = L.act_joints[(size_t)env * L.act_jstride + i];
      if (ops & RCSB_OP_ACT_JOINTS_REL) {  // base.py:475-488
        real lim = a < -L.max_mov ? -L.max_mov : (a > L.max_mov ? L.max_mov : a);
        real v = WR(q)[m.rb_qadr[i]] + lim;
        a = v < L.jlow[i] ? L.jlow[i] : (v > L.jhigh[i] ? L.jhigh[i] : v);
      }
      jt[i] = a;
    }
    RCSB_SYNC();
    op_gripper_action(c, L, env);
    int changed = !RI(RCSB_I_HAVE_PREV_ACTION);  // base.py:268-272: not allclose(a, prev, atol=1e-3, rtol=0)
    for (int i = 0; i < MD(rb_njoints); i++)
      if (!(r_abs(jt[i] - RS(RCSB_S_PREVACT + i)) <= (real)1e-3)) changed = 1;
    RCSB_SYNC();
    if (changed) op_set_joint_position(c, jt);
    PFOR(i, MD(rb_njoints)) { RS(RCSB_S_PREVACT + i) = jt[i]; }
    if (c.lane == 0) RI(RCSB_I_HAVE_PREV_ACTION) = 1;
    RCSB_SYNC();
  } else {
    op_gripper_action(c, L, env);
  }
  if (ops & RCSB_OP_SET_JOINTS) op_set_joint_position(c, L.act_joints + (size_t)env * L.act_jstride);
  if ((ops & RCSB_OP_SET_GRIPPER) && MD(gr_enabled)) op_set_gripper(c, L.act_gripper[(size_t)env * L.act_gstride]);
}
// CTA barriers a lockstep warp owes for `nsteps` physics steps it does not run
RCSB_DEV void skip_step_barriers(const Ctx& c, int nsteps) {
#ifndef RCSB_HOST_EMU
  int n = nsteps * __popc((unsigned)c.lockstep);
  for (int i = 0; i < n; i++) RCSB_GROUP_BARRIER();
#endif
}
RCSB_DEV void run_env_program(const Ctx& c, const RcsbLaunch& L, int env) {
  const RcsbModel& m = CMODEL(c);
  const unsigned ops = L.ops;
  const int resuming = L.phase == 1;
  int overflow = 0;
  if (!resuming) run_env_pre_ops(c, L, env);
  if (ops & RCSB_OP_STEP_K) {  // sim.cpp:108-115
    const int k = resuming ? RI(RCSB_I_RESUME) : L.k;
    for (int i = 0; i < k; i++) {
      if (physics_step(c, &CCLK(c)[RCSB_D_TIME])) {
        overflow = k - i;
        skip_step_barriers(c, k - i - 1);
        break;
      }
    }
  }
  if (ops & RCSB_OP_STEP_CONV) {  // sim.cpp:84-106
    int steps = 0, converged = 0;
    RCSB_SYNC();
    if (resuming) steps = RI(RCSB_I_CONV_STEPS);
    else PFOR(i, RCSB_NCB) { RI(RCSB_I_CBRET + i) = 0; }
    RCSB_SYNC();
    for (;;) {
      const int alive = !overflow && !converged && (L.max_convergence_steps == -1 || steps < L.max_convergence_steps);
#ifndef RCSB_HOST_EMU
      // lockstep launches: one CTA-wide vote per step keeps the warps of the CTA on the same code (instruction cache)
      // until the last of their environments has converged; warps that are done keep voting
      if (c.conv_vote) { if (!rcsb_cta_vote(alive)) break; }
      else
#endif
      if (!alive) break;
      if (alive) {
        if (physics_step(c, &CCLK(c)[RCSB_D_TIME])) { overflow = 1; continue; }
        steps++;
        converged = invoke_condition_callbacks(c, CCLK(c)[RCSB_D_TIME]);
      }
    }
    RCSB_SYNC();
    if (c.lane == 0) { RI(RCSB_I_CONVERGED) = converged; RI(RCSB_I_CONV_STEPS) = steps; }
    RCSB_SYNC();
  }
  RCSB_SYNC();
  if (c.lane == 0) {
    RI(RCSB_I_RESUME) = overflow;
#ifndef RCSB_HOST_EMU
    if (overflow) L.overflow_list[atomicAdd(L.overflow_count, 1)] = env;
#else
    if (overflow) L.overflow_list[(*L.overflow_count)++] = env;
#endif
  }
  RCSB_SYNC();
  if (overflow) return;  // the full-capacity launch finishes the steps and packs the observation
  if (L.con_n && (ops & (RCSB_OP_STEP_K | RCSB_OP_STEP_CONV))) {  // contact list of the last step1, as the callbacks saw it
    const int ncon = WI(misc)[MI_NCON];
    if (c.lane == 0) L.con_n[env] = ncon;
    PFOR(i, L.con_cap) {
      int* g = L.con_geom + ((size_t)env * L.con_cap + i) * 2;
      real* o = L.con_real ? L.con_real + ((size_t)env * L.con_cap + i) * RCSB_CON_EXPORT_REALS : nullptr;
      if (i < ncon) {
        const int* ci = WI(con) + RCSB_CI_INTS * i;
        const real* cr = WR(con) + RCSB_C_REALS * i;
        g[0] = m.g_origid[ci[RCSB_CI_G0]]; g[1] = m.g_origid[ci[RCSB_CI_G1]];
        if (o) {
          o[0] = cr[RCSB_C_DIST];
          for (int k = 0; k < 3; k++) { o[1 + k] = cr[RCSB_C_POS + k]; o[4 + k] = cr[RCSB_C_FRAME + k]; }
        }
      } else {
        g[0] = g[1] = -1;
        if (o) for (int k = 0; k < RCSB_CON_EXPORT_REALS; k++) o[k] = 0;
      }
    }
  }
  if ((ops & RCSB_OP_FRAMES) && L.frames) {
    // kinematics of the CURRENT qpos (as mj_forward would give it to a renderer). The free-joint quaternions are
    // re-normalised inside st_kinematics: qpos is put back afterwards so that observing never changes the state.
    real* keep = WR(M);  // rebuilt by every step, and outside the regions the kinematics stage writes
    PFOR(i, MD(nq)) { keep[i] = WR(q)[i]; }
    RCSB_SYNC();
    st_kinematics(c);
    PFOR(i, MD(nq)) { WR(q)[i] = keep[i]; }
    real* out = L.frames + (size_t)env * MD(nb) * 12;
    PFOR(e, MD(nb) * 12) {
      const int b = e / 12, k = e - 12 * b;
      out[e] = k < 3 ? WR(bpos)[3 * b + k] : WR(bmat)[9 * b + k - 3];
    }
    RCSB_SYNC();
  }
  if ((ops & RCSB_OP_OBS) && c.lane == 0) {
    real pose[7];
    robot_cartesian_position(c, pose);
    const real gw = MD(gr_enabled) ? gripper_width(c) : (real)0;
    const int rc = RI(RCSB_I_COLLISION), gc = MD(gr_enabled) ? RI(RCSB_I_G_COLLISION) : 0;
    int f[RCSB_INFO_DIM];
    f[0] = rc || gc;                       // envs/sim.py:61,127-128
    f[1] = RI(RCSB_I_IK_SUCCESS);
    f[2] = RI(RCSB_I_CONVERGED);
    f[3] = gw > (real)0.01 && gw < (real)0.99;  // envs/sim.py:130
    f[4] = rc || !RI(RCSB_I_IK_SUCCESS);   // truncated, envs/sim.py:66
    f[5] = rc; f[6] = gc;
    f[7] = RI(RCSB_I_CONV_STEPS);
    if (L.obs) {
      real* o = L.obs + (size_t)env * RCSB_OBS_DIM;
      for (int i = 0; i < 7; i++) o[i] = pose[i];
      for (int i = 0; i < 7; i++) o[7 + i] = i < MD(rb_njoints) ? WR(q)[m.rb_qadr[i]] : (real)0;
      pose_xyzrpy(pose, o + 14);
      o[20] = RS(RCSB_S_GCMD) < 0 ? (real)1 : RS(RCSB_S_GCMD);
      o[21] = gw;
      for (int i = 0; i < RCSB_INFO_DIM; i++) o[22 + i] = (real)f[i];
    }
    if (L.info) {
      int* fo = L.info + (size_t)env * RCSB_INFO_DIM;
      for (int i = 0; i < RCSB_INFO_DIM; i++) fo[i] = f[i];
    }
  }
}
