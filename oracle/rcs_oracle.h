/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this. The product (robot-control-stack_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED against libmujoco 3.2.6 / Pinocchio 3.7.0: neither library nor any golden
 * mj_step / Pin::inverse vector exists in /root/reference or in this image. What IS
 * pinned: Pose arithmetic against /root/reference/python/tests/test_common.py, and the behavioural
 * envelopes of /root/reference/python/tests/test_sim_envs.py. Everything marked [3P] restates the
 * published MuJoCo 3.2.6 / Pinocchio 3.7.0 algorithms from their documentation.
 *
 * Double precision, single environment per rcso_data, no dependencies beyond libc/libm.
 */
#ifndef RCS_ORACLE_H
#define RCS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rcso_model rcso_model;
typedef struct rcso_data rcso_data;
typedef struct rcso_sim rcso_sim;

/* ---- model: restates the mjModel subset; filled field-by-field from the compiled scene ---- */
rcso_model* rcso_model_new(void);
void rcso_model_free(rcso_model* m);
int rcso_model_set_int(rcso_model* m, const char* field, const int* v, int n);     /* 0 ok, -1 unknown */
int rcso_model_set_real(rcso_model* m, const char* field, const double* v, int n); /* 0 ok, -1 unknown */
int rcso_model_finalize(rcso_model* m);                                            /* 0 ok */

/* ---- data: restates the mjData subset ---- */
rcso_data* rcso_data_new(const rcso_model* m);
void rcso_data_free(rcso_data* d);
double* rcso_data_real(rcso_data* d, const char* field, int* n); /* borrowed pointer into d */
int* rcso_data_int(rcso_data* d, const char* field, int* n);

/* mj_resetData, mj_step1, mj_step2, mj_step, mj_forward [3P]; call sites
 * /root/reference/src/sim/sim.cpp:110,112,118 */
void rcso_reset_data(const rcso_model* m, rcso_data* d);
void rcso_step1(const rcso_model* m, rcso_data* d);
void rcso_step2(const rcso_model* m, rcso_data* d);
void rcso_step(const rcso_model* m, rcso_data* d);
void rcso_forward(const rcso_model* m, rcso_data* d);

/* ---- RCS glue: Sim + SimRobot + SimGripper, /root/reference/src/sim/{sim,SimRobot,SimGripper}.cpp ---- */
typedef struct {
  int njoints;              /* 7 */
  int joint_qposadr[8];     /* qpos address of each arm joint */
  int actuator_id[8];
  int attachment_site;      /* site id */
  int base_body;            /* body id */
  int ncgeom;
  int cgeom[16];            /* arm_collision_geoms ids */
  double q_home[8];
  double joint_rotational_tolerance; /* SimRobot.h:15-16 */
  double seconds_between_callbacks;  /* SimRobot.h:17 */
  double tcp_offset[7];              /* xyz + quat(xyzw) */
  int register_convergence_callback;
  /* IK model (Pin): frame = attachment site */
  int ik_nq;                         /* model.nq of the kinematic model (9 for MJCF FR3) */
} rcso_robot_cfg;

typedef struct {
  int enabled;
  int actuator_id;
  int joint_qposadr;
  int ncgeom, cgeom[8];
  int ncfgeom, cfgeom[4];
  int nignored, ignored[8];
  double epsilon_inner, epsilon_outer, seconds_between_callbacks;
  double max_actuator_width, min_actuator_width, max_joint_width, min_joint_width;
} rcso_gripper_cfg;

rcso_sim* rcso_sim_new(const rcso_model* m, const rcso_robot_cfg* rc, const rcso_gripper_cfg* gc);
void rcso_sim_free(rcso_sim* s);
rcso_data* rcso_sim_data(rcso_sim* s);
void rcso_sim_set_config(rcso_sim* s, int async_control, int frequency, int max_convergence_steps);
void rcso_sim_step(rcso_sim* s, int k);              /* sim.cpp:108-115 */
void rcso_sim_step_until_convergence(rcso_sim* s);   /* sim.cpp:84-106 */
int rcso_sim_is_converged(rcso_sim* s);
int rcso_sim_convergence_steps(rcso_sim* s);
void rcso_sim_reset(rcso_sim* s);                    /* sim.cpp:117-138 */
/* SimRobot */
void rcso_robot_set_joint_position(rcso_sim* s, const double* q);        /* SimRobot.cpp:123-131 */
void rcso_robot_get_joint_position(rcso_sim* s, double* q);              /* :133-139 */
void rcso_robot_get_cartesian_position(rcso_sim* s, double* pose7);      /* :114-121 xyz+quat xyzw */
int rcso_robot_set_cartesian_position(rcso_sim* s, const double* pose7); /* :145-155 returns ik ok */
void rcso_robot_reset(rcso_sim* s);                                      /* :193-205 */
void rcso_robot_state(rcso_sim* s, int* ik_success, int* collision, int* is_moving, int* is_arrived,
                      double* previous_angles, double* target_angles);
/* SimGripper */
int rcso_gripper_set_normalized_width(rcso_sim* s, double width, double force); /* -1 invalid arg */
double rcso_gripper_get_normalized_width(rcso_sim* s);
int rcso_gripper_is_grasped(rcso_sim* s);
void rcso_gripper_reset(rcso_sim* s);
void rcso_gripper_state(rcso_sim* s, double* last_commanded_width, int* is_moving, double* last_width,
                        int* collision);

/* ---- Pin IK / FK, /root/reference/src/rcs/Kinematics.cpp:28-81 ---- */
int rcso_ik_inverse(const rcso_model* m, int site, int nq_model, const double* pose7, const double* q0, int nq0,
                    const double* tcp_offset7, double* q_out, int* iters);
void rcso_ik_forward(const rcso_model* m, int site, int nq_model, const double* q0, int nq0,
                     const double* tcp_offset7, double* pose7);

/* pinocchio::log3 as used by the IK (test hook: closed-form check near theta = pi) */
void rcso_log3(const double* R9_rowmajor, double* w3);

/* ---- Pose math, /root/reference/src/rcs/Pose.cpp ---- (pose7 = xyz + quat xyzw) */
void rcso_pose_mul(const double* a, const double* b, double* out);
void rcso_pose_inverse(const double* a, double* out);
void rcso_pose_from_rpy(const double* xyz, const double* rpy, double* out);
void rcso_pose_from_matrix(const double* R9_rowmajor, const double* xyz, double* out);
void rcso_pose_xyzrpy(const double* a, double* out6);
void rcso_pose_rotation_m(const double* a, double* R9_rowmajor);
double rcso_pose_total_angle(const double* a);
void rcso_pose_limit_rotation_angle(const double* a, double max_angle, double* out);
void rcso_pose_limit_translation_length(const double* a, double max_len, double* out);
void rcso_pose_interpolate(const double* a, const double* b, double progress, double* out);
int rcso_pose_is_close(const double* a, const double* b, double eps_r, double eps_t);

/* ---- Gym-level env loop of the benchmark workload (python/rcs/envs/{base,sim}.py semantics), used
 * only as the CPU baseline: runs `nsteps` env.step() calls with JOINTS relative control + binary
 * gripper over `nthreads` independent envs; returns wall seconds. actions: [nenv][nsteps][8]. ---- */
double rcso_bench_env_steps(const rcso_model* m, const rcso_robot_cfg* rc, const rcso_gripper_cfg* gc,
                            int nenv, int nthreads, int nsteps, int episode_len, int async_control,
                            const double* actions, double max_mov, const double* joint_low,
                            const double* joint_high, long long* physics_steps_out, double* obs_out);

#ifdef __cplusplus
}
#endif
#endif
