// Device-side model and per-environment state layout of the batched rigid-body backend.
//
// The model is the "fused" form of the compiled scene (rcs_b200/devmodel.py): bodies without joints
// are folded into their nearest moving ancestor (or into the world), so the kinematic tree holds one
// body per joint; only collidable geoms are kept. It restates the mjModel subset that
// /root/reference/src/sim/{sim,SimRobot,SimGripper}.cpp reach through mjModel*/mjData*.
#pragma once
#include <stdint.h>

#ifndef RCSB_REAL
#define RCSB_REAL double  // reference arithmetic is float64 (mjtNum, Eigen double)
#endif
typedef RCSB_REAL real;

enum {
  RCSB_MAXB = 12,     // moving bodies
  RCSB_MAXV = 16,     // dofs
  RCSB_MAXQ = 18,     // qpos
  RCSB_MAXU = 8,      // actuators
  RCSB_MAXG = 32,     // collidable geoms
  RCSB_MAXPAIR = 320, // candidate geom pairs
  RCSB_MAXT = 2,      // fixed tendons
  RCSB_MAXEQ = 2,     // joint equalities
  RCSB_MAXROOT = 4,   // kinematic trees
  RCSB_MAXJ = 8,      // robot arm joints
  RCSB_MAXCAND = 48,  // broad-phase survivors per step (28 at the FR3 home pose)
  RCSB_SEPSLOTS = 4,  // direct-mapped slots of the separating-direction cache (convex narrow phase)
  RCSB_MAXGRP = 64,   // collision groups (pairs of bodies that own at least one candidate geom pair)
};

enum { RCSB_JNT_FREE = 0, RCSB_JNT_SLIDE = 2, RCSB_JNT_HINGE = 3 };
enum { RCSB_GEOM_PLANE = 0, RCSB_GEOM_SPHERE = 2, RCSB_GEOM_CAPSULE = 3, RCSB_GEOM_CYLINDER = 5, RCSB_GEOM_BOX = 6, RCSB_GEOM_MESH = 7 };
enum { RCSB_TRN_JOINT = 0, RCSB_TRN_TENDON = 3 };
enum { RCSB_EQ = 0, RCSB_FRICTION_DOF = 1, RCSB_LIMIT = 3, RCSB_CONTACT_PYR = 6, RCSB_CONTACT_ELL = 7 };
enum { RCSB_SATISFIED = 0, RCSB_QUADRATIC = 1, RCSB_LINEARNEG = 2, RCSB_LINEARPOS = 3, RCSB_CONE = 4 };
// geom role bits for the RCS collision callbacks (SimRobot.cpp:172-182, SimGripper.cpp:108-130)
enum { RCSB_ROLE_ARM = 1, RCSB_ROLE_GRIPPER = 2, RCSB_ROLE_FINGER = 4, RCSB_ROLE_IGNORED = 8 };
// callback kinds, in the registration order of SimRobot / SimGripper constructors
enum { RCSB_CB_ARRIVED = 0, RCSB_CB_MOVING = 1, RCSB_CB_ROBOT_CONV = 2, RCSB_CB_ROBOT_COLL = 3, RCSB_CB_GRIP_CONV = 4,
       RCSB_CB_GRIP_COLL = 5, RCSB_NCB = 6 };

// per-contact record in the workspace (reals)
enum { RCSB_C_DIST = 0, RCSB_C_POS = 1, RCSB_C_FRAME = 4, RCSB_C_FRIC = 13, RCSB_C_SOLREF = 16, RCSB_C_SOLIMP = 18,
       RCSB_C_MU = 23, RCSB_C_INCMARGIN = 24, RCSB_C_REALS = 25 };
enum { RCSB_CI_G0 = 0, RCSB_CI_G1 = 1, RCSB_CI_DIM = 2, RCSB_CI_EFC = 3, RCSB_CI_INTS = 4 };
// per-constraint-row scalars (reals), stored as arrays of length maxefc each
enum { RCSB_E_POS = 0, RCSB_E_MARGIN, RCSB_E_FLOSS, RCSB_E_D, RCSB_E_R, RCSB_E_AREF, RCSB_E_FORCE, RCSB_E_JAR, RCSB_E_JV,
       RCSB_E_B, RCSB_E_K, RCSB_E_NARR };
enum { RCSB_EI_TYPE = 0, RCSB_EI_ID, RCSB_EI_STATE, RCSB_EI_NARR };

// ---- per-environment persistent state in HBM: struct-of-arrays by field group, env-major rows
//   sr[N][nsr]  reals  : qpos[nq] qvel[nv] ctrl[nu] qacc_warmstart[nv] | RCS tail (RCSB_S_*)
//   sd[N][RCSB_D_TAIL] doubles : simulation time and callback clocks (always double: the callback
//                        cadence depends on float64 accumulation of time, sim.cpp:14-23)
//   si[N][RCSB_I_TAIL] ints   : flags and counters
enum {
  RCSB_S_PREV = 0,                          // previous_angles[MAXJ]   (SimRobotState)
  RCSB_S_TARGET = RCSB_S_PREV + RCSB_MAXJ,  // target_angles[MAXJ]
  RCSB_S_GLCW = RCSB_S_TARGET + RCSB_MAXJ,  // gripper last_commanded_width
  RCSB_S_GLW,                               // gripper last_width
  RCSB_S_GCMD,                              // GripperWrapper._last_gripper_cmd (-1 = None), base.py:684-735
  RCSB_S_PREVACT,                           // RobotEnv.prev_action joints[MAXJ], base.py:268-287
  RCSB_S_SITEPOS = RCSB_S_PREVACT + RCSB_MAXJ,  // attachment site xpos[3] from the last step1
  RCSB_S_SITEMAT = RCSB_S_SITEPOS + 3,      // attachment site xmat[9]
  RCSB_S_TAIL = RCSB_S_SITEMAT + 9,
};
enum { RCSB_D_TIME = 0, RCSB_D_CBLAST = 1, RCSB_D_TAIL = 1 + RCSB_NCB };
enum {
  RCSB_I_IK_SUCCESS = 0, RCSB_I_COLLISION, RCSB_I_MOVING, RCSB_I_ARRIVED, RCSB_I_G_MOVING, RCSB_I_G_COLLISION,
  RCSB_I_CONVERGED, RCSB_I_CONV_STEPS, RCSB_I_CBRET,  // RCSB_NCB last_return_value flags follow
  RCSB_I_NCON = RCSB_I_CBRET + RCSB_NCB, RCSB_I_NEFC, RCSB_I_SOLVER_ITER, RCSB_I_WARN, RCSB_I_TOTAL_STEPS,
  RCSB_I_HAVE_PREV_ACTION,  // RobotEnv.prev_action is not None (base.py:268-272)
  RCSB_I_RESUME,  // substeps this launch still owes the environment (it outgrew the reduced workspace layout)
  RCSB_I_TAIL
};

// Per-warp workspace layout: offsets in reals (o_*) and ints (oi_*). A pure function of the model's shape
// (rcsb_make_layout), so kernels specialised for a fixed shape fold every offset into an instruction immediate.
struct RcsbLayout {
  int ws_reals, ws_ints, nsr;  // nsr = reals per env in HBM (dynamic state + RCS tail)
  int o_q, o_v, o_ctrl, o_warm, o_bpos, o_bquat, o_bmat, o_rootcom, o_cinert, o_crb, o_crbbuf,
      o_cdof, o_cdofdot, o_cvel, o_cacc, o_cfrc, o_gcw, o_M, o_L, o_H, o_bias, o_passive, o_gravc, o_actfrc, o_smooth,
      o_qacc_smooth, o_qacc, o_qfc, o_grad, o_search, o_Ma, o_Mv, o_tmp, o_aforce, o_gpos, o_cand, o_pairfr, o_sup, o_con,
      o_J, o_efc, o_conehess, o_noslip, o_site, o_rcs, o_sepcache, o_cbud, o_cbq;
  int ws_doubles;  // RCSB_D_TAIL doubles per warp follow the reals
  int oi_con, oi_efc, oi_misc;
};
// The part of a model that fixes code shape: loop bounds, workspace layout, enabled features.
struct RcsbShape {
  int nq, nv, nu, nb, ng, npair, nt, neq, nroot, maxcon, maxefc, rb_njoints, cone_elliptic, implicitfast,
      noslip_iterations, cap_reduced, gr_enabled, ngrp, nfl;  // nfl: dofs with friction loss (rows of the noslip workspace)
};

struct RcsbModel {
  // ---- sizes / options
  int nq, nv, nu, nb, ng, npair, nt, neq, nroot, nmeshvert;
  int nfl;  // dofs with frictionloss > 0, derived in rcsb_model_finalize_layout
  int root_plain;  // bit r: the implicit integrator's matrix block of kinematic tree r is M's own block (no damping, no velocity-dependent actuator on its dofs); derived
  int cone_elliptic, implicitfast, iterations, ls_iterations, noslip_iterations;
  int maxcon, maxefc;  // per-env capacities of the contact / constraint workspaces
  // Two workspace layouts share one kernel: the full-capacity one above and a reduced one (fast_maxcon / fast_maxefc,
  // 0 = none) that is small enough for twice the warps per SM. A launch runs every environment in the reduced layout
  // first; an environment whose contact / constraint count does not fit (cap_reduced != 0 in the model copy it ran
  // with) stops before the step that overflowed and is finished by a second launch in the full layout.
  int fast_maxcon, fast_maxefc, cap_reduced;
  real timestep, gravity[3], impratio, tolerance, ls_tolerance, noslip_tolerance, meaninertia;
  // ---- moving bodies (one joint each)
  int b_parent[RCSB_MAXB], b_jtype[RCSB_MAXB], b_qadr[RCSB_MAXB], b_dadr[RCSB_MAXB], b_ndof[RCSB_MAXB], b_root[RCSB_MAXB];
  uint32_t b_ancmask[RCSB_MAXB];   // moving-body ancestors incl. self
  uint32_t b_descmask[RCSB_MAXB];  // descendants incl. self
  uint32_t b_dofmask[RCSB_MAXB];   // dofs of self and all ancestors
  real b_pos[RCSB_MAXB][3];                        // frame in parent moving body (or world) at qpos0 (b_quat: cold tail)
  real b_rot[RCSB_MAXB][9];                        // rotation matrix of b_quat, derived in rcsb_model_finalize_layout
  real b_jpos[RCSB_MAXB][3], b_jaxis[RCSB_MAXB][3];
  real b_mass[RCSB_MAXB], b_ipos[RCSB_MAXB][3], b_inertia[RCSB_MAXB][6];  // xx yy zz xy xz yz about COM, body axes
  real b_gcmass[RCSB_MAXB], b_gcpos[RCSB_MAXB][3];                        // sum(gravcomp*mass) and its centre
  // ---- dofs
  int d_body[RCSB_MAXV], d_qadr[RCSB_MAXV], d_limited[RCSB_MAXV], d_actfrclimited[RCSB_MAXV], d_actgravcomp[RCSB_MAXV],
      d_dotzero[RCSB_MAXV];
  uint32_t d_premask[RCSB_MAXV];  // dofs whose velocity enters cdof_dot of this dof
  // dof-chain form of the tree, derived in rcsb_model_finalize_layout: the dof whose accumulated spatial velocity
  // this dof adds to (-1: tree root), the dof whose accumulated velocity enters this dof's cdof_dot (-1: none), and
  // the last dof of every body (its accumulated velocity / acceleration is the body's)
  int d_parent[RCSB_MAXV], d_pre[RCSB_MAXV], b_lastdof[RCSB_MAXB];
  // dof range [lo, hi) of the kinematic tree every dof belongs to (M and its Cholesky factor are block diagonal by tree);
  // the whole range [0, nv) when the trees' dofs are not contiguous
  int d_tree_lo[RCSB_MAXV], d_tree_hi[RCSB_MAXV];
  uint32_t d_ancmask[RCSB_MAXV];  // ancestor dofs incl. self (sparsity of M)
  real d_armature[RCSB_MAXV], d_damping[RCSB_MAXV], d_frictionloss[RCSB_MAXV], d_invweight0[RCSB_MAXV];
  real d_range[RCSB_MAXV][2], d_margin[RCSB_MAXV], d_solref[RCSB_MAXV][2], d_solimp[RCSB_MAXV][5],
      d_actfrcrange[RCSB_MAXV][2];
  real qpos0[RCSB_MAXQ];
  // ---- kinematic trees
  real r_invmass[RCSB_MAXROOT];
  // ---- collidable geoms
  int g_body[RCSB_MAXG], g_type[RCSB_MAXG], g_vertadr[RCSB_MAXG], g_vertnum[RCSB_MAXG], g_origid[RCSB_MAXG],
      g_role[RCSB_MAXG], g_condim[RCSB_MAXG], g_priority[RCSB_MAXG];
  real g_pos[RCSB_MAXG][3];                        // in the moving body frame (world frame if g_body < 0; g_quat: cold tail)
  real g_rot[RCSB_MAXG][9];                        // rotation matrix of g_quat, derived
  real g_bpos[RCSB_MAXG][3];  // bounding-volume centre (local AABB centre) in the moving body frame; g_rbound is about it
  real g_size[RCSB_MAXG][3], g_rbound[RCSB_MAXG], g_aabb[RCSB_MAXG][6], g_friction[RCSB_MAXG][3], g_solref[RCSB_MAXG][2],
      g_solimp[RCSB_MAXG][5], g_solmix[RCSB_MAXG], g_margin[RCSB_MAXG], g_gap[RCSB_MAXG], g_invweight[RCSB_MAXG];
  uint8_t pair[RCSB_MAXPAIR][2];  // collidable-geom indices, lower geom type first
  // Collision groups, derived in rcsb_model_finalize_layout: all geom pairs between the same two bodies form a group
  // with one separation budget (rcsb_dynamics.cuh: st_collision). grp_mask = dofs on the tree path between the two
  // bodies; sum over them of |dq_j| * grp_reach[g][j] (cold tail) bounds the relative displacement of the group's geoms.
  int ngrp;
  uint8_t pair_grp[RCSB_MAXPAIR];
  uint32_t pair_blk_grps[(RCSB_MAXPAIR + 31) / 32][2];  // bit mask of the collision groups that own a pair of the 32-pair block; derived
  uint32_t grp_mask[RCSB_MAXGRP];
  // ---- tendons, equalities, actuators
  real t_coef[RCSB_MAXT][RCSB_MAXV];
  int e_dof1[RCSB_MAXEQ], e_dof2[RCSB_MAXEQ], e_active[RCSB_MAXEQ];
  real e_poly[RCSB_MAXEQ][5], e_solref[RCSB_MAXEQ][2], e_solimp[RCSB_MAXEQ][5];
  int a_trntype[RCSB_MAXU], a_trnid[RCSB_MAXU], a_ctrllimited[RCSB_MAXU], a_forcelimited[RCSB_MAXU];
  // implicitfast derivative: joint actuators that can never be force-clamped fold into a per-dof constant
  int n_special, a_special[RCSB_MAXU];  // actuators that need the per-step treatment (tendon transmission or forcerange)
  real d_kvdiag[RCSB_MAXV];             // sum of bias2*gear^2 of the folded joint actuators
  real a_moment[RCSB_MAXU][RCSB_MAXV];  // d(actuator length)/d(qpos of dof), derived in rcsb_model_finalize_layout
  uint8_t tri_i[RCSB_MAXV * (RCSB_MAXV + 1) / 2], tri_j[RCSB_MAXV * (RCSB_MAXV + 1) / 2];  // lower-triangle entry list (i >= j), derived
  real a_gear[RCSB_MAXU], a_gain[RCSB_MAXU], a_bias[RCSB_MAXU][3], a_ctrlrange[RCSB_MAXU][2], a_forcerange[RCSB_MAXU][2];
  // ---- RCS device layer: SimRobot / SimGripper configuration
  int rb_njoints, rb_qadr[RCSB_MAXJ], rb_act[RCSB_MAXJ], rb_site_body, rb_register_convergence, rb_ik_nq;
  real rb_site_rot[9];  // rotation matrix of rb_site_quat, derived
  real rb_site_pos[3], rb_site_quat[4], rb_base_pos[3], rb_base_quat[4] /* wxyz */, rb_tcp_offset[7] /* xyz+xyzw */;
  real rb_q_home[RCSB_MAXJ], rb_joint_tol, rb_cb_period;
  int gr_enabled, gr_act, gr_qadr;
  real gr_eps_inner, gr_eps_outer, gr_cb_period, gr_max_act, gr_min_act, gr_max_joint, gr_min_joint;
  // per group and dof on its tree path: upper bound, over all poses, of the distance between the joint anchor and any
  // collidable point of the group's body that the dof moves (1 for translational dofs; 0 off the path). Read by every
  // group on every step (budget_advance), so it is staged with the hot part.
  // Stored as the upper half of the float32 bit pattern, rounded up (rcsb_reach_encode / rcsb_reach_decode): 2 KB.
  uint16_t grp_reach[RCSB_MAXGRP][RCSB_MAXV];
  // ---- per-warp workspace layout (offsets in reals / ints), filled by rcsb_model_finalize
  RcsbLayout lay;
  // ---- cold tail: NOT staged into shared memory (RCSB_MODEL_HOT_BYTES ends here); device code reaches it through the
  //      global-memory copy of the model (CMODEL_G), where the few hot rows stay in L1
  int cold_begin;
  int8_t grp_body[RCSB_MAXGRP][2];              // the two bodies of every collision group (-1 = world)
  real b_quat[RCSB_MAXB][4], g_quat[RCSB_MAXG][4];  // host-side inputs of b_rot / g_rot (rcsb_model_finalize_layout)
  real g_rbound0[RCSB_MAXG];                    // mjModel geom_rbound (about the geom frame origin): plane-mesh point spacing
};


// op bits, executed in this order within one launch
enum {
  RCSB_OP_GRIPPER_RESET = 1 << 0,   // SimGripper::reset
  RCSB_OP_SIM_RESET = 1 << 1,       // Sim::reset (mj_resetData + callback clocks)
  RCSB_OP_ROBOT_RESET = 1 << 2,     // SimRobot::reset (set_joints_hard(q_home))
  RCSB_OP_ENV_RESET_FLAGS = 1 << 3, // GripperWrapper.reset: _last_gripper_cmd = None
  RCSB_OP_ACT_JOINTS_REL = 1 << 4,  // RelativeActionSpace (LAST_STEP) + RobotEnv.step dedupe + set_joint_position
  RCSB_OP_ACT_JOINTS_ABS = 1 << 5,  // RobotEnv.step dedupe + set_joint_position
  RCSB_OP_ACT_GRIPPER_BIN = 1 << 6, // GripperWrapper.action, binary
  RCSB_OP_SET_JOINTS = 1 << 7,      // SimRobot::set_joint_position (direct API, no dedupe)
  RCSB_OP_SET_GRIPPER = 1 << 8,     // SimGripper::set_normalized_width
  RCSB_OP_SET_JOINTS_HARD = 1 << 9, // SimRobot::set_joints_hard
  RCSB_OP_STEP_K = 1 << 10,         // Sim::step(k)
  RCSB_OP_STEP_CONV = 1 << 11,      // Sim::step_until_convergence
  RCSB_OP_OBS = 1 << 12,            // RobotEnv.get_obs + wrappers' observation/info
  RCSB_OP_ACT_GRIPPER_CONT = 1 << 13,  // GripperWrapper.action, continuous width
  RCSB_OP_FRAMES = 1 << 14,         // export the world frames [p | R] of every moving body (camera ray-caster input)
};
enum { RCSB_OBS_DIM = 30, RCSB_INFO_DIM = 8 };
// obs row: tquat[7] joints[7] xyzrpy[6] gripper[1] gripper_width[1] | the info row again, as reals (one packed block per
//          environment: what the multi-GPU exchange gathers and what the host-buffer path copies back)
// info row: collision, ik_success, is_sim_converged, is_grasped, truncated, robot_collision, gripper_collision, conv_steps

struct RcsbLaunch {
  int N, env_offset;
  unsigned ops;
  int k, max_convergence_steps;
  int conv_vote;   // step_until_convergence: 1 = static env -> warp map with a CTA-wide vote per step, 0 = dynamic scheduling
  int bar_groups;  // fixed-substep launches: number of separately aligned warp groups per CTA (named barriers, <= 15)
  int lockstep;  // fixed-substep launches: mask of CTA barriers (bit i: before stage i of the step, bit 9: at its end)
  int phase;     // 0: every environment, reduced or full layout; 1: full layout, only the environments in overflow_list
  int* overflow_list;   // [N] environments the reduced layout could not finish (phase 0 appends, phase 1 consumes)
  int* overflow_count;
  const real* act_joints;   // [N][act_jstride], njoints used
  const real* act_gripper;  // [N][act_gstride], first used
  int act_jstride, act_gstride;  // row strides in reals (njoints / 1 for separate arrays; a packed [N][njoints + 1] block: both njoints + 1)
  const unsigned char* mask;  // optional [N]: 0 = leave this env untouched
  real max_mov, jlow[RCSB_MAXJ], jhigh[RCSB_MAXJ];
  real* obs;   // [N][RCSB_OBS_DIM] or null
  int* info;   // [N][RCSB_INFO_DIM] or null
  // optional export of mjData.contact after the launch's last step (SimRobot.cpp:172-182, SimGripper.cpp:108-130 read
  // contact[i].geom[0/1]): count, geom ids of the compiled scene (mjModel numbering) in contact order, and
  // dist | pos[3] | normal[3] per contact; con_cap contacts per environment, unused slots -1 / 0
  int* con_n;      // [N] or null
  int* con_geom;   // [N][con_cap][2]
  real* con_real;  // [N][con_cap][RCSB_CON_EXPORT_REALS] or null
  int con_cap;
  real* frames;  // [N][nb][12] world frame of every moving body (position, then the row-major rotation), RCSB_OP_FRAMES
};
enum { RCSB_CON_EXPORT_REALS = 7 };

