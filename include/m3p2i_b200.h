/*
 * m3p2i_b200.h — C ABI of libm3p2i_b200.so, the B200-native rollout + update hot path of
 * the multi-modal MPPI planner (M3P2I).
 *
 * This is the drop-in boundary: every entry point replaces a piece of the reference's Python
 * hot path (paths relative to the reference checkout, src/m3p2i_aip/...):
 *
 *   m3p2i_create / m3p2i_destroy   <- MPPI.__init__ (planners/motion_planner/mppi.py:82-203) +
 *                                     IsaacGymWrapper.__init__/start_sim (utils/isaacgym_utils/
 *                                     isaacgym_wrapper.py:40-90): K rollout environments on one GPU
 *   m3p2i_set_scene_*              <- actor_utils.load_env_cfgs + IsaacGymWrapper.creat_env
 *                                     (actor_utils.py:94-101, isaacgym_wrapper.py:242-352)
 *   m3p2i_set_state                <- REACTIVE_TAMP.run_tamp: _dof_state[:] = ..; _root_state[:] = ..;
 *                                     set_dof_state_tensor / set_actor_root_state_tensor
 *                                     (scripts/reactive_tamp.py:45-48)
 *   m3p2i_set_objective            <- Objective.update_objective (cost_functions.py:15-17) and
 *                                     M3P2I.update_gripper_command (m3p2i.py:10-14)
 *   m3p2i_set_noise_table          <- the `delta` attribute of MPPI (mppi.py:100,386-392)
 *   m3p2i_set_noise_halton_spline  <- MPPI.get_samples (mppi.py:458-478) built on the device
 *   m3p2i_command                  <- MPPI.command (mppi.py:211-264): shift, sample, T-step rollout
 *                                     (mppi.py:275-332 -> reactive_tamp.py:63-73 -> isaacgym_wrapper.py:
 *                                     354-360 + cost_functions.py:19-36), softmin weights
 *                                     (mppi.py:430-456 / m3p2i.py:24-64), mean update
 *                                     (mppi.py:485-518 / m3p2i.py:66-92), Savitzky-Golay (mppi.py:257-263)
 *   m3p2i_update_only              <- MPPI._update_distribution / M3P2I._update_multi_modal_distribution
 *                                     on caller-supplied cost_horizon/actions (generic callback path)
 *   m3p2i_top_trajs                <- torch.topk(weights, 20) + index_select (mppi.py:248-254)
 *   m3p2i_get_buffer               <- attributes weights / cost_total / states / actions read after command()
 *   m3p2i_sim_*                    <- the IsaacGymWrapper facade on K persistent envs: step()
 *                                     (isaacgym_wrapper.py:354-360), set_dof_velocity_target_tensor (:196),
 *                                     apply_rigid_body_force_tensors (:202), tensor views (:98-112)
 *   m3p2i_peer_export / _attach    <- (nothing: the reference is single-GPU) K sharded over ranks; the all-gather
 *                                     of the discounted costs and the all-reduce of the weighted action sums are
 *                                     stores into peer HBM over NVLink issued by the rollout / weighted-sum kernels
 *   m3p2i_comm_init                <- (nothing) the same exchange as two NCCL collectives (fallback)
 *   m3p2i_phase_*                  <- (nothing) the same three phases host-staged, for any transport
 *
 * Conventions: plain C, caller owns every host array (C-contiguous fp32 / int32), the library owns all
 * device memory. Return 0 on success, negative code on failure; the message is in m3p2i_last_error().
 * A handle is not thread-safe; one CUDA stream per handle; calls are synchronous on return unless
 * stated otherwise. No exception ever crosses this boundary.
 */
#ifndef M3P2I_B200_H
#define M3P2I_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3P2I_MAX_NU 9
#define M3P2I_MAX_STATIC 8
#define M3P2I_NX 4          /* state row stored per step: reactive_tamp.py:66-69 */
#define M3P2I_TOP_N 20      /* mppi.py:248 */
#define M3P2I_MAX_HORIZON 64

#define M3P2I_ENV_POINT 0
#define M3P2I_ENV_PANDA 1

/* task ids (cost_functions.py:19-36) */
#define M3P2I_TASK_NAVIGATION 0
#define M3P2I_TASK_PUSH 1
#define M3P2I_TASK_PULL 2
#define M3P2I_TASK_PUSH_PULL 3
#define M3P2I_TASK_REACH 4
#define M3P2I_TASK_PICK 5
#define M3P2I_TASK_PLACE 6

/* gripper command (m3p2i.py:10-14, mppi.py:412-416) */
#define M3P2I_GRIPPER_NONE 0
#define M3P2I_GRIPPER_OPEN 1
#define M3P2I_GRIPPER_CLOSE 2

/* noise modes */
#define M3P2I_NOISE_TABLE 0   /* caller-supplied delta[K,T,nu] (reference semantics: sampled once, reused) */
#define M3P2I_NOISE_PHILOX 1  /* Philox4x32-10 counter RNG evaluated inside the rollout kernel */
#define M3P2I_NOISE_PHILOX_SPLINE 2 /* the same generator drawing max(T/4, 2) + 2 control points per (sample, dimension),
                                       blended over the horizon by a uniform quadratic B-spline rescaled to unit variance:
                                       smooth in time like the reference's Halton splines (mppi_utils.py:80-104,
                                       skill_utils.py:9-22), no table */

/* error codes */
#define M3P2I_OK 0
#define M3P2I_ERR_ARG -1
#define M3P2I_ERR_CUDA -2
#define M3P2I_ERR_STATE -3
#define M3P2I_ERR_NCCL -4
#define M3P2I_ERR_NO_DEVICE -5

/* buffers readable through m3p2i_get_buffer (device pointers, library-owned, layout in the comment) */
#define M3P2I_BUF_ACTIONS 0      /* [T][nu][K]  fp32, sample index fastest (coalesced) */
#define M3P2I_BUF_STATES 1       /* [T][4][K]   fp32 */
#define M3P2I_BUF_COST_HORIZON 2 /* [T][K]      fp32 */
#define M3P2I_BUF_COST_DISC 3    /* [K_global]  fp32, discounted cost J (all ranks after all-gather) */
#define M3P2I_BUF_COST_SUM 4     /* [K]         fp32, undiscounted sum over t */
#define M3P2I_BUF_WEIGHTS 5      /* [3][K_global] fp32: all / mode-1 / mode-2 weights */
#define M3P2I_BUF_NOISE 6        /* [K][T][nu]  fp32 noise table (table mode) */

typedef struct M3P2IConfig {
  int32_t env_type;            /* M3P2I_ENV_* (cfg.env_type) */
  int32_t num_samples;         /* K owned by this handle (= cfg.mppi.num_samples on one GPU) */
  int32_t horizon;             /* T (cfg.mppi.horizon) */
  int32_t nu;                  /* 2 (point) or 9 (panda) */
  int32_t multi_modal;         /* cfg.multi_modal */
  int32_t sample_null_action;  /* cfg.mppi.sample_null_action (mppi.py:300-302) */
  int32_t filter_u;            /* cfg.mppi.filter_u: Savitzky-Golay(9,2) on the returned action */
  int32_t noise_mode;          /* M3P2I_NOISE_* */
  int32_t num_samples_global;  /* K over all ranks (== num_samples when single GPU) */
  int32_t sample_offset;       /* global id of local sample 0 */
  int32_t substeps;            /* IsaacGymConfig.substeps (isaacgym_wrapper.py:10) */
  int32_t solver_passes;       /* contact solver sweeps per substep (our integrator; default 2) */
  int32_t lanes_per_sample;    /* rollout kernel shape: 0 = library chooses, 1 = one thread per sample,
                                  8 / 16 = lane-cooperative team of that many lanes per sample (panda_env) */
  int32_t update_cov;          /* cfg.mppi.update_cov (mppi.py:43,508-516): adapt the per-dimension noise variance from the
                                  weighted second moment of the samples (single-mode only, as in the reference) */
  int32_t reserved_i[2];
  float dt;                    /* cfg.isaacgym.dt */
  float gamma;                 /* cfg.mppi.rollout_var_discount (mppi.py:181) */
  float step_size_mean;        /* 0.98 (mppi.py:178) */
  float u_scale;               /* cfg.mppi.u_scale */
  float kp_suction;            /* cfg.kp_suction (skill_utils.py:86-90) */
  float pre_height_diff;       /* cfg.pre_height_diff (cost_functions.py:12) */
  float tilt_cos_theta;        /* 0.5 (cost_functions.py:13) */
  float reserved_f;
  float u_min[M3P2I_MAX_NU];   /* cfg.mppi.u_min */
  float u_max[M3P2I_MAX_NU];   /* cfg.mppi.u_max */
  float sigma[M3P2I_MAX_NU];   /* sqrt(diag(cfg.mppi.noise_sigma)) (mppi.py:175-176,394) */
  uint64_t seed;               /* Philox key */
} M3P2IConfig;

/* An oriented box (all boxes of the reference scenes are gym.create_box assets, actor_utils.py:69-75) */
typedef struct M3P2IBox {
  float pos[3];
  float half[3];
  float quat[4];   /* x, y, z, w */
  float mu;        /* friction of the shape (actor_utils.py:28) */
  int32_t actor;   /* row in root_state (pose re-read at set_state when movable), -1 = none */
} M3P2IBox;

/* A movable box body (planar in the point env, full 3-D cube in the panda env) */
typedef struct M3P2IBody {
  float half[3];
  float mass;
  float inertia;   /* about z (point env) / isotropic (panda cubes) */
  float mu;
  float r_eff;     /* mean contact-patch radius for torsional ground friction (point env) */
  int32_t actor;   /* row in root_state */
} M3P2IBody;

/* point_env (config/point_env/<n>_<name>.yaml, assets/urdf/pointRobot.urdf) */
typedef struct M3P2IPointScene {
  float robot_radius;    /* pointRobot.urdf:17 */
  float robot_mass;      /* pointRobot.urdf:11 */
  float robot_mu;        /* 0_point_robot.yaml:6 */
  float drive_damping;   /* isaacgym_wrapper.py:344 */
  float drive_effort;    /* pointRobot.urdf:36 */
  float gravity;         /* isaacgym_wrapper.py:25 (magnitude) */
  float ground_mu;       /* isaacgym_wrapper.py:466 */
  float contact_margin;  /* physx.contact_offset isaacgym_wrapper.py:30 */
  float baumgarte;       /* positional-error feedback per substep (our integrator) */
  float slop;            /* allowed penetration (our integrator) */
  float max_corr_vel;    /* cap on the positional-correction velocity (our integrator) */
  int32_t n_static;
  int32_t n_actors;      /* rows of root_state */
  int32_t reserved;
  M3P2IBody box;         /* the pushed/pulled block, 7_box.yaml */
  M3P2IBody dyn_obs;     /* 6_dyn_obs.yaml; its xy contact force is the collision cost */
  M3P2IBox statics[M3P2I_MAX_STATIC];
} M3P2IPointScene;

/* panda_env (config/panda_env/<n>_<name>.yaml, franka_panda.urdf) */
typedef struct M3P2IPandaScene {
  float base_pos[3];     /* panda_env/panda.yaml:7 */
  float gravity;
  float q_lower[M3P2I_MAX_NU];  /* franka_panda.urdf <limit lower> */
  float q_upper[M3P2I_MAX_NU];
  float qd_limit[M3P2I_MAX_NU]; /* <limit velocity> */
  float effort[M3P2I_MAX_NU];   /* <limit effort> */
  float drive_damping;   /* 600, isaacgym_wrapper.py:344 */
  float arm_inertia;     /* reflected inertia of an arm joint whose joint_inertia entry is <= 0 */
  float finger_mass;     /* mass of one finger (prismatic DoF) */
  float robot_mu;        /* friction of the panda shapes */
  float finger_half[3];  /* finger collision box (AABB of meshes/collision/finger.obj), centred at finger_center */
  float finger_center[3];/* in the left-finger frame; the right finger is the mirror image (urdf:220) */
  float hand_half[3];    /* hand collision box */
  float hand_center[3];
  float contact_margin;
  float baumgarte;
  float slop;
  float max_corr_vel;
  float penalty_stiffness; /* kinematic link vs static box: force = k * depth (reported as contact force only) */
  float joint_inertia[M3P2I_MAX_NU]; /* inertia the velocity drive of arm joint j works against (entries 0..6; <= 0:
                            arm_inertia). The URDF has no <inertial> blocks, so IsaacGym derives the link masses from
                            the collision meshes at its default density; tools/panda_inertia.py does the same and
                            reflects them onto the joint axes at the initial pose */
  float warm_start;      /* fraction of a contact's accumulated impulse re-applied in the next sub-step of a step() */
  float sleep_lin;       /* a supported, untouched cube slower than this (m/s; and sleep_ang rad/s) is asleep: it is */
  float sleep_ang;       /*   neither moved nor solved (PhysX sleeps resting bodies too); sleep_lin <= 0: never */
  float sleep_gap;       /* a corner counts as supporting when it is within this distance of the fixed box (m) */
  int32_t n_static;
  int32_t n_actors;
  int32_t idx_table;     /* index into statics[] of the bodies named by get_motion_cost (cost_functions.py:161-164) */
  int32_t idx_shelf;
  int32_t link_sweeps;   /* Gauss-Seidel sweeps over the finger/hand-cube contacts inside each solver pass (<= 0: 4):
                            the finger - cube - finger chain of a grasp needs them to settle within a sub-step */
  int32_t report_cube_contacts; /* 1: the contact force reported for the table / shelf stand also contains the cubes
                                   resting or sliding on them; 0 (default): only the robot's links count. The
                                   collision cost (cost_functions.py:158-169) is about the ROBOT hitting the furniture,
                                   and with 1 every rollout that slides cubeA on the table pays it */
  M3P2IBody cube_a;      /* 5_cubeA.yaml */
  M3P2IBody cube_b;      /* 6_cubeB.yaml */
  M3P2IBox statics[M3P2I_MAX_STATIC];
} M3P2IPandaScene;

/* Planner state that persists between command() calls (mppi.py:148-153,184-187) */
typedef struct M3P2IPlannerState {
  float mean_action[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  float mean_action_1[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  float mean_action_2[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  float best_traj[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  float best_traj_1[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  float best_traj_2[M3P2I_MAX_HORIZON * M3P2I_MAX_NU];
  double beta;                /* single-mode inverse temperature (adapted across calls for panda_env) */
  float cov_action[M3P2I_MAX_NU]; /* per-dimension noise variance (mppi.py:175; changes only with update_cov); the
                                  noise scale of the next command is its square root (scale_tril, mppi.py:176,516) */
  float reserved;
} M3P2IPlannerState;

/* Scalars produced by one command() */
typedef struct M3P2ICommandInfo {
  float eta[3];       /* normaliser of the all / mode-1 / mode-2 weights (mppi.py:441, m3p2i.py:33) */
  float beta[3];      /* beta used for each weight set */
  float min_cost[3];  /* min_k J_k per weight set */
  int32_t best_idx[3];/* global argmax of each weight set */
  float weight_push;  /* sum of weights[:K/2]  (m3p2i.py:18) */
  float weight_pull;  /* sum of weights[K/2:]  (m3p2i.py:19) */
  float mean_cost_sum;/* mean_k sum_t c_kt, the term aliased into cost_total (mppi.py:325) */
  float kernel_ms;    /* device time of the command's kernels (CUDA events on the handle's stream) */
  int32_t launches;   /* kernels launched by this command */
  int32_t beta_iters; /* iterations of the on-the-fly beta search (m3p2i.py:30-43), summed over sets */
  float rollout_ms;   /* device time of the fused rollout kernel alone (the roofline figure is computed from it) */
  int32_t rollout_lanes; /* lanes per sample of the rollout kernel this command used (1, 8 or 16) */
  float peer_wait_ms[2]; /* sharded over peer memory: device time this rank spent waiting for [0] the discounted costs of
                            all ranks (start of the softmin kernel: rollout skew between ranks + the NVLink stores) and
                            [1] the partial sums of all ranks (end of the weighted-sum kernel); 0 when unsharded */
  int32_t near_samples;  /* panda_env: samples of this shard the far-field kernel left to the full rollout kernel (the others
                            never came near a cube, the table or the shelf and were finished by it); -1: it did not run */
} M3P2ICommandInfo;

typedef struct M3P2IHandle_* m3p2i_handle;

const char* m3p2i_last_error(void);
int m3p2i_version(void);
int m3p2i_device_count(void);
/* sizeof() of an interface struct by its C name ("M3P2IConfig", ...), -1 if unknown: lets a foreign-language
 * binding verify its own struct layout at load time */
int m3p2i_abi_sizeof(const char* struct_name);

int m3p2i_create(const M3P2IConfig* cfg, int device, m3p2i_handle* out);
void m3p2i_destroy(m3p2i_handle h);

int m3p2i_set_scene_point(m3p2i_handle h, const M3P2IPointScene* scene);
int m3p2i_set_scene_panda(m3p2i_handle h, const M3P2IPandaScene* scene);

/* dof_state: [2*ndof] (pos, vel interleaved, isaacgym_wrapper.py:98-100); root_state: [n_actors*13]
 * (pos3, quat4 xyzw, linvel3, angvel3). One real state, broadcast to all K rollouts on the device. */
int m3p2i_set_state(m3p2i_handle h, const float* dof_state, const float* root_state);

int m3p2i_set_objective(m3p2i_handle h, int task_id, const float* goal, int goal_len, int gripper_cmd);

/* delta: [K_local, T, nu] (rows of this rank's shard) or NULL to drop the table. */
int m3p2i_set_noise_table(m3p2i_handle h, const float* delta);
/* Noise row [T,nu] of GLOBAL sample 0, for shards that do not own it (table mode, single-mode panda reach: every
 * sample's cost reads sample 0's cube position, cost_functions.py:98). NULL drops it. */
int m3p2i_set_noise_row0(m3p2i_handle h, const float* row0);
/* Builds the reference's once-sampled halton-spline table ON THE DEVICE for this shard's global samples (replaces
 * MPPI.get_samples, mppi.py:458-478, the K * nu scipy calls of skill_utils.bspline, skill_utils.py:9-22, and
 * generate_gaussian_halton_samples, mppi_utils.py:80-104): n_knots = T / knot_scale points per (sample, dimension) of
 * the Halton sequence in n_knots * nu dimensions -> sqrt(2) erfinv(2 u - 1) -> FITPACK smoothing spline (degree,
 * smoothing = 0.5 in the reference) sampled at T points. perms = NULL gives the plain Halton sequence (the reference's
 * use_ghalton=False branch, mppi_utils.py:82-87); perms = uint16 [n_knots * nu][perm_stride], row d a permutation of
 * 0 .. prime(d)-1, gives the generalised (digit-scrambled) sequence of ghalton.GeneralizedHalton(perms)
 * (mppi_utils.py:88-95: pass ghalton.EA_PERMS[:ndims], padded to perm_stride). Also fills the row of global sample 0
 * when this shard does not own it (m3p2i_set_noise_row0). Needs noise_mode = M3P2I_NOISE_TABLE. */
int m3p2i_set_noise_halton_spline(m3p2i_handle h, int knot_scale, int degree, float smoothing, const uint16_t* perms,
                                  int perm_stride);
/* Writes the noise the kernel uses (table or Philox) as [K_local, T, nu]; before delta[-1]=0 is applied. */
int m3p2i_get_noise(m3p2i_handle h, float* out_delta);

int m3p2i_get_planner_state(m3p2i_handle h, M3P2IPlannerState* out);
int m3p2i_set_planner_state(m3p2i_handle h, const M3P2IPlannerState* in);

/* Savitzky-Golay smoothing matrix [T,T] (row-major, out = S @ action); uploaded once by the host. */
int m3p2i_set_filter_matrix(m3p2i_handle h, const float* S);

/* One planner tick. out_action: [T,nu]; out_cost_total: [K_local] or NULL (mppi.py:325 quirk included);
 * info may be NULL: the command then records no timing events on the stream (kernel_ms / rollout_ms are only measured
 * for callers that ask for the scalars). */
int m3p2i_command(m3p2i_handle h, float* out_action, float* out_cost_total, M3P2ICommandInfo* info);

/* Same tick without any host<->device copy of inputs or results (state, means stay resident);
 * used to time the device-resident path. Results are fetched later with m3p2i_fetch_result. */
int m3p2i_command_resident(m3p2i_handle h, M3P2ICommandInfo* info);
int m3p2i_fetch_result(m3p2i_handle h, float* out_action, float* out_cost_total);

/* Open-loop rollout of caller-supplied actions [K_local,T,nu]; writes states [K,T,4], cost_horizon [K,T]
 * (either may be NULL). No planner state is touched. Used by the parity tests. */
int m3p2i_rollout_actions(m3p2i_handle h, const float* actions, float* out_states, float* out_cost_horizon);

/* Generic callback path, first half (mppi.py:237-242,381-416): applies the one-step shift to the stored sequences and
 * writes this tick's perturbed action sequences out_actions [K_local,T,nu] (before u_scale), exactly the rows the
 * fused command would roll out. The caller then runs its own dynamics / cost callbacks T times and finishes the
 * tick with m3p2i_update_only. */
int m3p2i_sample_actions(m3p2i_handle h, float* out_actions);

/* Update on caller-supplied arrays: cost_horizon [K_local,T], actions [K_local,T,nu] (generic callback path,
 * mppi.py:327-331). out_action [T,nu] is the new mean_action (unfiltered). */
int m3p2i_update_only(m3p2i_handle h, const float* cost_horizon, const float* actions,
                      float* out_mean_action, M3P2ICommandInfo* info);

/* top-n weights and their state trajectories: out_idx [n] (global ids), out_weights [n],
 * out_trajs [n,T,2] = states[idx][:, :, (0,2)] (mppi.py:248-254). Trajectories of samples owned by other
 * ranks are zero-filled. */
int m3p2i_top_trajs(m3p2i_handle h, int n, int32_t* out_idx, float* out_weights, float* out_trajs);

/* Run all work of this handle on `cuda_stream` (a cudaStream_t; NULL = the handle's own stream), so that a caller
 * can bracket commands with events of a stream it owns. */
int m3p2i_set_stream(m3p2i_handle h, void* cuda_stream);

int m3p2i_get_buffer(m3p2i_handle h, int which, void** dev_ptr, size_t* bytes);
/* Copies a buffer to the host in the reference's layout ([K,T,nu], [K,T,4], [K,T], [K], [3,K]). */
int m3p2i_read_buffer(m3p2i_handle h, int which, float* out, size_t count);

/* ---- persistent K-env simulation facade (IsaacGymWrapper surface) ---- */
int m3p2i_sim_reset(m3p2i_handle h);  /* copy the state given by set_state into all K envs */
int m3p2i_sim_set_velocity_target(m3p2i_handle h, const float* u /* [K,nu] */);
int m3p2i_sim_apply_forces(m3p2i_handle h, const float* f_robot /* [K,2] or NULL */,
                           const float* f_box /* [K,2] or NULL */);
int m3p2i_sim_step(m3p2i_handle h);
/* Per-env state upload (inverse of m3p2i_sim_read): dof_state [K,2*ndof], root_state [K,n_actors,13]; only the
 * rows of movable bodies are read from root_state. Either pointer may be NULL (keeps that part). */
int m3p2i_sim_write(m3p2i_handle h, const float* dof_state, const float* root_state);
/* Task cost of every persistent env in its current state, out_cost [K] (Objective.compute_cost,
 * cost_functions.py:19-36). Like the reference's pull cost it also arms the suction forces for the next step. */
int m3p2i_sim_cost(m3p2i_handle h, float* out_cost);
/* dof_state [K,2*ndof], root_state [K,n_actors,13], link_state [K,n_links,13] (panda: leftfinger,
 * rightfinger, hand; point: robot body), contact_force [K,n_contact,3] (point: dyn-obs; panda: table,
 * shelf_stand, cubeB). Any pointer may be NULL. */
int m3p2i_sim_read(m3p2i_handle h, float* dof_state, float* root_state, float* link_state,
                   float* contact_force);

/* ---- multi-GPU: K sharded over ranks, one process per GPU ---- */
/* Number of floats in the packed partial-sum buffer exchanged by the second collective:
 * [sum w a | sum w1 a | sum w2 a | best row x3] (6 T nu), sum of the undiscounted costs (1), sum w a^2 (T nu). */
int m3p2i_partials_len(m3p2i_handle h);
/* Host-staged three-phase tick (any transport, e.g. torch.distributed gloo); the NCCL path of
 * m3p2i_command runs the same three phases with the two collectives enqueued on the handle's stream.
 *   phase_rollout : shift + sample + rollout of the local shard, returns its discounted costs J [K_local]
 *   phase_partials: given J of ALL samples [K_global] (all-gather), computes the weights and returns this
 *                   rank's packed partial sums [partials_len]
 *   phase_finish  : given the element-wise SUM over ranks of the partials (all-reduce), updates the means */
int m3p2i_phase_rollout(m3p2i_handle h, float* out_cost_disc_local);
int m3p2i_phase_partials(m3p2i_handle h, const float* cost_disc_global, float* out_partials);
int m3p2i_phase_finish(m3p2i_handle h, const float* partials_sum, float* out_action,
                       float* out_cost_total, M3P2ICommandInfo* info);
int m3p2i_comm_unique_id(void* out_id128 /* 128 bytes */);
int m3p2i_comm_init(m3p2i_handle h, int rank, int nranks, const void* id128);

/* Exchange over NVLink peer memory (replaces m3p2i_comm_init's two NCCL collectives per command): each rank's
 * rollout kernel stores its discounted costs into every peer's HBM and the weighted-sum kernel does the same with the
 * packed partial sums, so a sharded command stays at 3 kernel launches with no collective call.
 * Every rank: m3p2i_peer_export -> exchange the descriptors through the host (torch.distributed all_gather_object,
 * MPI, ...) -> m3p2i_peer_attach(all descriptors, indexed by rank) -> host barrier -> commands. Ranks may be processes
 * (cudaIpc) or handles of one process. All ranks must issue the same sequence of commands. At most 8 ranks (one
 * NVSwitch domain). A peer that stops delivering makes the next fetch fail with M3P2I_ERR_STATE instead of hanging:
 * waits inside the kernels give up after M3P2I_PEER_TIMEOUT_MS (environment, read at attach; default 30000). */
typedef struct M3P2IPeerHandle {
  unsigned char ipc[64];  /* cudaIpcMemHandle_t of the rank's mailbox */
  int64_t pid;            /* exporting process */
  uint64_t ptr;           /* device pointer, valid inside the exporting process */
  uint64_t bytes;         /* mailbox size (must agree on all ranks) */
  int32_t device;
  int32_t reserved;
} M3P2IPeerHandle;
int m3p2i_peer_export(m3p2i_handle h, M3P2IPeerHandle* out);
int m3p2i_peer_attach(m3p2i_handle h, int rank, int nranks, const M3P2IPeerHandle* all);

#ifdef __cplusplus
}
#endif
#endif /* M3P2I_B200_H */
