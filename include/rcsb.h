/* rcsb -- C ABI of the B200 batched rigid-body backend for Robot Control Stack.
 *
 * The reference has no C-ABI plugin table for its simulator; its boundary is the pybind11 module
 * rcs._core (/root/reference/src/pybind/rcs.cpp:420-527) whose C++ side reaches libmujoco through
 * mjModel* / mjData* (/root/reference/src/sim/sim.cpp, SimRobot.cpp, SimGripper.cpp). Every entry
 * point below names the reference interface it replaces; INTEGRATION.md shows the pybind/ctypes stub
 * a maintainer would add. Plain pointers and sizes only; device pointers are raw CUDA addresses.
 *
 * All functions return 0 on success and a negative code on failure unless stated otherwise;
 * rcsb_last_error() returns a description. There is no CPU execution path: without a CUDA device
 * rcsb_model_upload / rcsb_batch_* fail with RCSB_ERR_CUDA.
 */
#ifndef RCSB_H
#define RCSB_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rcsb_model rcsb_model;
typedef struct rcsb_batch rcsb_batch;

enum { RCSB_OK = 0, RCSB_ERR_FIELD = -1, RCSB_ERR_SIZE = -2, RCSB_ERR_MODEL = -3, RCSB_ERR_CUDA = -4, RCSB_ERR_ARG = -5 };

/* op bits of rcsb_batch_run, executed in this order for every selected environment */
enum {
  RCSB_RUN_GRIPPER_RESET = 1 << 0,   /* SimGripper::reset              SimGripper.cpp:158-165 */
  RCSB_RUN_SIM_RESET = 1 << 1,       /* Sim::reset                     sim.cpp:117-138 */
  RCSB_RUN_ROBOT_RESET = 1 << 2,     /* SimRobot::reset                SimRobot.cpp:193-216 */
  RCSB_RUN_ENV_RESET_FLAGS = 1 << 3, /* GripperWrapper.reset           python/rcs/envs/base.py:703-708 */
  RCSB_RUN_ACT_JOINTS_REL = 1 << 4,  /* RelativeActionSpace.action + RobotEnv.step   base.py:469-488, 255-288 */
  RCSB_RUN_ACT_JOINTS_ABS = 1 << 5,  /* RobotEnv.step (JOINTS)         base.py:255-288 */
  RCSB_RUN_ACT_GRIPPER_BIN = 1 << 6, /* GripperWrapper.action (binary) base.py:721-735 */
  RCSB_RUN_SET_JOINTS = 1 << 7,      /* SimRobot::set_joint_position   SimRobot.cpp:123-131 */
  RCSB_RUN_SET_GRIPPER = 1 << 8,     /* SimGripper::set_normalized_width  SimGripper.cpp:79-92 */
  RCSB_RUN_SET_JOINTS_HARD = 1 << 9, /* SimRobot::set_joints_hard      SimRobot.cpp:197-205 */
  RCSB_RUN_STEP_K = 1 << 10,         /* Sim::step(k)                   sim.cpp:108-115 */
  RCSB_RUN_STEP_CONV = 1 << 11,      /* Sim::step_until_convergence    sim.cpp:84-106 */
  RCSB_RUN_OBS = 1 << 12,            /* RobotEnv.get_obs + info        base.py:246-253, envs/sim.py:60-66,125-131 */
  RCSB_RUN_ACT_GRIPPER_CONT = 1 << 13, /* GripperWrapper.action (continuous width) base.py:721-735 */
  RCSB_RUN_FRAMES = 1 << 14           /* internal: body frames for rcsb_camera_depth (mjv_updateScene, camera.cpp:100-118) */
};

const char* rcsb_last_error(void);
int rcsb_version(void);
int rcsb_real_bytes(void); /* sizeof(real) the library was built with (8: float64, as the reference) */

/* ---- model: replaces mjModel (python/rcs/sim/sim.py:47-55 hands mjModel/mjData addresses to Sim) ---- */
rcsb_model* rcsb_model_new(void);
void rcsb_model_free(rcsb_model* m);
int rcsb_model_set_int(rcsb_model* m, const char* field, const int* v, int n);
int rcsb_model_set_real(rcsb_model* m, const char* field, const double* v, int n);
int rcsb_model_set_mesh_vertices(rcsb_model* m, const double* xyz, int nvert);
/* edge graph of the convex hulls (mjModel mesh_graph, used by mjc_PlaneConvex for multi-point plane-mesh contacts):
 * neighbours of pooled vertex v are nbr[adr[v] .. adr[v+1]) as vertex ids local to the geom's hull; optional */
int rcsb_model_set_mesh_graph(rcsb_model* m, const int* adr, int nadr, const int* nbr, int nnbr);
/* supporting planes of the convex hulls (n . x + d <= 0 inside, geom frame), pooled, with each collidable geom's range:
 * input of the depth ray-caster (rcsb_camera_depth); optional */
int rcsb_model_set_mesh_faces(rcsb_model* m, const double* planes, int nface, const int* geom_faceadr, const int* geom_facenum, int ng);
int rcsb_model_finalize(rcsb_model* m);         /* validates sizes, computes the workspace layout (host only) */
int rcsb_model_upload(rcsb_model* m, int device); /* copies constants and hull vertices to the CUDA device */
/* sizes of one environment's rows: reals, doubles, ints, obs reals, info ints */
int rcsb_model_dims(const rcsb_model* m, int* nsr, int* nsd, int* nsi, int* obs_dim, int* info_dim);
/* column offsets inside the real row: qpos, qvel, ctrl, qacc_warmstart, rcs tail */
int rcsb_model_offsets(const rcsb_model* m, int* o_qpos, int* o_qvel, int* o_ctrl, int* o_warm, int* o_tail);

/* shared-memory footprint that decides the warps per SM: bytes of one warp's workspace in the reduced layout (0 when the
 * model has none) and in the full layout, and the bytes the staged model takes per CTA (host only, no device needed) */
int rcsb_model_workspace_bytes(const rcsb_model* m, int* reduced_bytes, int* full_bytes, int* smem_header_bytes);

/* ---- batch: replaces N x (mjData + Sim + SimRobot + SimGripper) ----
 * sr/sd/si are caller-owned DEVICE arrays [n_envs][nsr] reals, [n_envs][nsd] doubles, [n_envs][nsi] ints
 * (allocated by the host language, e.g. torch tensors). stream is a cudaStream_t (may be 0). */
rcsb_batch* rcsb_batch_new(rcsb_model* m, int n_envs, void* sr, void* sd, void* si, void* stream);
void rcsb_batch_free(rcsb_batch* b);
int rcsb_batch_init_state(rcsb_batch* b); /* SimRobotState/SimGripperState defaults + mj_resetData on every env */

/* Contact list export, replaces the reads of mjData.contact[i].geom[0/1] / ncon (SimRobot.cpp:172-182,
 * SimGripper.cpp:108-130): after every following launch that steps, each environment's contact list of the last
 * mj_step1 is written to ncon_dev [n_envs], geom_dev [n_envs][cap][2] (geom ids of the compiled scene in mjModel
 * numbering, contact order, -1 beyond ncon) and, when not NULL, real_dev [n_envs][cap][7] reals (dist, pos[3],
 * normal[3]). ncon_dev == NULL switches the export off. Device pointers owned by the caller. */
int rcsb_batch_set_contact_export(rcsb_batch* b, int* ncon_dev, int* geom_dev, void* real_dev, int cap);

/* One launch: for every env (mask_dev == NULL) or every env with mask_dev[env] != 0, run the ops in
 * `ops`. act_joints_dev [n_envs][njoints] reals, act_gripper_dev [n_envs] reals, obs_dev
 * [n_envs][obs_dim] reals, info_dev [n_envs][info_dim] ints: device pointers, NULL where unused.
 * jlow/jhigh: host arrays [njoints] (joint limits of robots_meta_config, Robot.h:24-95). */
int rcsb_batch_run(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const void* act_joints_dev,
                   const void* act_gripper_dev, const unsigned char* mask_dev, double max_mov, const double* jlow,
                   const double* jhigh, void* obs_dev, int* info_dev);

/* Host-buffer variant of one env.step(): copies actions host->device, runs `ops`, copies obs/info
 * device->host on the batch stream and synchronises. Host buffers should be pinned. */
int rcsb_batch_run_host(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const double* act_joints_host,
                        const double* act_gripper_host, double max_mov, const double* jlow, const double* jhigh,
                        double* obs_host, int* info_host);

/* One env.step() of every environment through ONE packed host block each way: act_host [n_envs][njoints + 1] (joint
 * action, then the gripper action; pinned) -> H2D, the fused launch (`ops` as for rcsb_batch_run, RCSB_RUN_OBS implied),
 * D2H of obs_host [n_envs][obs_dim] -- the observation row carries the info flags as reals in its last 8 columns --
 * and a stream synchronise. Replaces the host round trip of SimEnvCreator's env.step (python/rcs/envs/sim.py:49-66).
 * A page-locked action block (cudaHostAlloc / cudaHostRegister / torch pin_memory) is not copied: the kernel reads the
 * action rows through the block's device-mapped pointer, so that transfer overlaps the physics; a pageable block takes a
 * staged copy. The observation block always leaves through one DMA copy. Same results either way.
 * RCSB_HOST_ZEROCOPY=0 in the environment forces the staged copy. */
int rcsb_env_step_host(rcsb_batch* b, unsigned ops, int k, int max_convergence_steps, const double* act_host, double max_mov,
                       const double* jlow, const double* jhigh, double* obs_host);

/* thin aliases with the reference's method names (all envs) */
int rcsb_sim_step(rcsb_batch* b, int k);                          /* Sim::step                 sim.cpp:108 */
int rcsb_sim_step_until_convergence(rcsb_batch* b, int max_steps); /* Sim::step_until_convergence sim.cpp:84 */
int rcsb_sim_reset(rcsb_batch* b);                                 /* Sim::reset                sim.cpp:117 */
int rcsb_robot_set_joint_position(rcsb_batch* b, const void* q_dev); /* SimRobot.cpp:123 */
int rcsb_robot_set_joints_hard(rcsb_batch* b, const void* q_dev);    /* SimRobot.cpp:197 */
int rcsb_robot_reset(rcsb_batch* b);                                 /* SimRobot.cpp:216 */
int rcsb_gripper_set_normalized_width(rcsb_batch* b, const void* width_dev); /* SimGripper.cpp:79 */
int rcsb_gripper_reset(rcsb_batch* b);                                        /* SimGripper.cpp:165 */
int rcsb_env_get_obs(rcsb_batch* b, void* obs_dev, int* info_dev);            /* base.py:246-253 */

/* Pin::inverse for every env (Kinematics.cpp:28-68): pose_dev [n][7] xyz+quat(xyzw) in the robot base
 * frame, q0_dev [n][njoints]; writes q_out_dev [n][ik_nq] and success_dev [n]. */
int rcsb_ik_inverse(rcsb_batch* b, const void* pose_dev, const void* q0_dev, void* q_out_dev, int* success_dev,
                    int* iters_dev);
/* SimRobot::set_cartesian_position for every env (SimRobot.cpp:145-155) */
int rcsb_robot_set_cartesian_position(rcsb_batch* b, const void* pose_dev);

/* Host-pointer variants for bindings that own no device memory (the compiled rcs_b200._core module):
 * synchronous, temporary device buffers per call. pose7_host [n][7], q0_host [n][njoints], q_out_host [n][ik_nq]. */
int rcsb_batch_info(rcsb_batch* b, int* n_envs, int* njoints, int* ik_nq, int* nq, int* nv, int* nu);
int rcsb_batch_read_row(rcsb_batch* b, int env, double* sr_row, double* sd_row, int* si_row); /* one environment's state rows */
int rcsb_robot_set_cartesian_position_host(rcsb_batch* b, const double* pose7_host);
int rcsb_ik_inverse_host(rcsb_batch* b, const double* pose7_host, const double* q0_host, double* q_out_host, int* success_host,
                         int* iters_host);

/* Cartesian Gym action for every env: RelativeActionSpace.action (python/rcs/envs/base.py:490-578, RelativeTo.LAST_STEP
 * when relative != 0: offset clipped to max_trans [m] / max_rot [rad], applied to the current Cartesian position, xyz
 * clipped to the workspace box) + RobotEnv.step's dedupe against the previous action and dispatch (base.py:255-288) +
 * SimRobot::set_cartesian_position (SimRobot.cpp:145-155). kind 0: act_dev [n][6] xyzrpy (CARTESIAN_TRPY); kind 1:
 * act_dev [n][7] xyz + quat xyzw (CARTESIAN_TQuat). */
int rcsb_env_cartesian_action(rcsb_batch* b, const void* act_dev, int kind, int relative, double max_trans, double max_rot);
/* The same with relative == 2 for RelativeTo.CONFIGURED_ORIGIN (base.py:443-467, 490-578): origin_dev [n][7] is the pose
 * RelativeActionSpace.reset() stored (xyz + quat xyzw), last_dev [n][7] / have_last_dev [n] the last clipped offset and
 * whether there is one (both updated by the call); all three are device arrays owned by the caller. relative 0 / 1 ignore
 * them. */
int rcsb_env_cartesian_action_origin(rcsb_batch* b, const void* act_dev, int kind, int relative, double max_trans, double max_rot,
                                     const void* origin_dev, void* last_dev, int* have_last_dev);

/* SimCameraSet depth frame of every environment (src/sim/camera.cpp:100-140 + python/rcs/camera/sim.py:45-95): the
 * camera (MuJoCo convention: looks along -z, +y up) sits at cam_pos / cam_rot (row-major) in the frame of moving body
 * cam_body (-1: world), vertical field of view fovy_deg, width x height pixels; out_dev [n_envs][height][width] uint16,
 * row 0 = top, = (uint16)(1000 * eye-space depth in metres clipped to [znear, zfar]) with physical_units, else
 * (uint16)(1000 * OpenGL window-space depth in [0, 1]). Rays are cast against the collidable geoms (convex hulls for
 * meshes), not rasterised visual meshes. cam_frames_dev (optional) [n_envs][12] reals receives the camera's world frame
 * of every environment (position, then the row-major rotation: mjData.cam_xpos / cam_xmat, for the extrinsics).
 * Rays are traced in float (frames and tile culling in double); RCSB_DEPTH_F64=1 in the environment selects double rays. */
int rcsb_camera_depth(rcsb_batch* b, int cam_body, const double* cam_pos, const double* cam_rot, double fovy_deg, int width, int height,
                      double znear, double zfar, int physical_units, void* out_dev, void* cam_frames_dev);
/* world frames of the moving bodies at the current qpos (mjData.xpos / xmat of the bodies that carry a joint):
 * frames_dev [n_envs][nb][12] reals, position then the row-major rotation */
int rcsb_body_frames(rcsb_batch* b, void* frames_dev);

/* evidence counters */
long long rcsb_launch_count(void); /* kernels launched by this library since load */
int rcsb_kernel_occupancy(rcsb_batch* b, int* warps_per_cta, int* smem_bytes, int* grid);
/* name of the kernel variant a batch launches: phase 0 = every environment (reduced workspace layout when the model has
   one), phase 1 = full-capacity pass for the environments that outgrew it. "generic" reads the model shape at run time;
   other names are kernels compiled for one fixed shape (csrc/rcsb_k_*.cu). Env RCSB_VARIANT=generic forces the former. */
const char* rcsb_kernel_variant(rcsb_batch* b, int phase);
/* profiling build only (-DRCSB_STAGE_TIMING): accumulated clock64() cycles per physics stage of warp 0 / CTA 0, then reset */
int rcsb_debug_stage_cycles(unsigned long long* out16);
/* Profiling build only: cycles of stage i (0..8) of the first max_steps (<= 256) physics steps every warp of CTA 0 ran since
 * the last call, out[max_steps][10][32] (step, stage, warp; stage 9 = the narrow phase inside the collision stage), followed
 * by [max_steps][32] collision counts (due groups | broad-phase survivors << 8 | mid-phase survivors << 16) and by [max_steps][3][32]
 * cycles from the start of the collision stage to the end of the geom-centre pass | the broad phase | the mid phase; the call
 * also clears the trace. */
int rcsb_debug_stage_trace(unsigned* out, int max_steps);

#ifdef __cplusplus
}
#endif
#endif
