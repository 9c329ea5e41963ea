/* b2s.h -- C-ABI of libb2s.so, the B200-native PushEnv substep path.
 *
 * Drop-in boundary: this library replaces what the
 * reference reaches through `robovat.simulation.physics.BulletPhysics`
 * (robovat/simulation/physics/bullet_physics.py:28-1304 -> pybullet 2.6.5),
 * `BulletCamera._frames` (robovat/simulation/camera/bullet_camera.py:188-235),
 * the per-substep Python in `Simulator.step` / `ControllableBody.update`
 * (robovat/simulation/simulator.py:94-103, controllable_body.py:387-413),
 * the PushEnv phase machine (robovat/envs/push/push_env.py:631-937) and
 * `push_reward.get_reward_fn` (robovat/reward_fns/push_reward.py:272-374),
 * for B independent environments at once.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a raw host or device address.
 *   - every entry point returns 0 on success or a negative B2S_E_* code; the
 *     message is available from b2s_last_error() (thread-local).  Nothing
 *     throws across the ABI.  There is NO CPU fallback: without a usable CUDA
 *     device b2s_create fails with B2S_E_CUDA.
 *   - all device work is enqueued on the caller's stream (cudaStream_t passed
 *     as void*); no hidden synchronisation except in the *_sync / host-copy
 *     helpers that say so.
 *   - quaternions are [x,y,z,w]; poses are 7 floats (pos3, quat4); metres,
 *     radians, seconds (same as the reference, bullet_physics.py:122-127).
 *   - one world per GPU per process; one host thread per world.
 */
#ifndef B2S_H_
#define B2S_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100            /* 0.1.0 */
#define B2S_NUM_JOINTS 7           /* Sawyer limb: right_j0 .. right_j6 */
#define B2S_MAX_LINKS 12           /* arm collision links (base + 7 + hand/finger) */
#define B2S_MAX_TILES 32

/* error codes */
#define B2S_OK 0
#define B2S_E_INVALID (-1)         /* bad argument (reference: ValueError) */
#define B2S_E_CUDA (-2)            /* CUDA runtime error / no device */
#define B2S_E_STATE (-3)           /* call order violated (e.g. step before load_scene) */
#define B2S_E_CAPACITY (-4)        /* a per-env capacity (pairs, manifolds) overflowed */
#define B2S_E_UNSUPPORTED (-5)     /* reference: NotImplementedError */

/* phases of PushEnv._execute_action (push_env.py:121-127) + idle/settle */
enum {
  B2S_PHASE_INITIAL = 0, B2S_PHASE_PRE = 1, B2S_PHASE_START = 2, B2S_PHASE_MOTION = 3,
  B2S_PHASE_POST = 4, B2S_PHASE_OFFSTAGE = 5, B2S_PHASE_DONE = 6,
  B2S_PHASE_SETTLE = 7,            /* Simulator.wait_until_stable after 'done' (push_env.py:723) */
  B2S_PHASE_IDLE = 8,              /* no action in flight: env is frozen by b2s_env_substeps */
  /* between two episodes of a device-side rollout (b2s_rollout_*): the freshly sampled scene drops
   * (wait_until_stable with the loose thresholds, push_env.py:443-447), then the final wait (:456-458) */
  B2S_PHASE_RESET_DROP = 9, B2S_PHASE_RESET_WAIT = 10
};

/* reward tasks (push_reward.py:283-296) */
enum { B2S_TASK_NONE = 0, B2S_TASK_CLEARING = 1, B2S_TASK_INSERTION = 2, B2S_TASK_CROSSING = 3 };

/* static-body flags */
#define B2S_STATIC_ON_TABLE 1u     /* z follows the per-env table height offset */
#define B2S_STATIC_IS_TABLE 2u     /* counts for check_contact(arm, table) (push_env.py:850) */
#define B2S_STATIC_NO_COLLIDE 4u   /* visual only (rendered, never collides) */
#define B2S_STATIC_IS_TILE 8u      /* loaded after the movables (push_env.py:343-357): body uid = index + num_movables */

/* Solver / world parameters.  Every Bullet default named in SURVEY.md 3.4 is a
 * field here so nothing is hard-wired; b2s_default_params() fills them. */
typedef struct B2SParams {
  int32_t num_envs;                /* B on this rank */
  int32_t env_id_offset;           /* global id of local env 0 (RNG key, sharding) */
  int32_t max_movables;            /* Nmax = config MAX_MOVABLE_BODIES */
  int32_t max_pairs;               /* broad-phase pair capacity per env */
  int32_t max_manifolds;           /* persistent (non-empty) manifolds per env */
  int32_t solver_iterations;       /* 50  (pybullet default numSolverIterations) */
  int32_t friction_dirs;           /* 2 = pyramid on plane-space tangents; 1 = velocity aligned */
  int32_t gjk_max_iters;           /* 32 */
  int32_t epa_max_iters;           /* 32 */
  int32_t ik_max_iters;            /* 20  (pybullet default maxNumIterations) */
  int32_t ik_interval;             /* 10  STEPS_TO_UPDATE_IK  controllable_body.py:25 */
  int32_t check_done_interval;     /* 100 STEPS_TO_CHECK_DONE controllable_body.py:22 */
  int32_t steps_check;             /* SIM.STEPS_CHECK         push_env.py:662 */
  int32_t max_phase_steps;         /* SIM.MAX_PHASE_STEPS     push_env.py:682 */
  int32_t max_motion_steps;        /* SIM.MAX_MOTION_STEPS    push_env.py:676 */
  int32_t max_offstage_steps;      /* SIM.MAX_OFFSTAGE_STEPS  push_env.py:679 */
  int32_t stable_check_after;      /* 100  simulator.py:329 */
  int32_t stable_min_steps;        /* 100  simulator.py:330 */
  int32_t stable_max_steps;        /* 2000 simulator.py:331 */
  int32_t clamp_joint_velocity;    /* 1: enforce speed*max_velocity (sawyer_sim.py:203-206 intends it;
                                      pybullet ignores it, bullet_physics.py:1092) ; 0: reference-literal */
  int32_t cam_height, cam_width;   /* depth / segmentation image */
  int32_t num_points;              /* OBS.NUM_POINTS per body */
  int32_t task;                    /* B2S_TASK_* */
  int32_t max_contacts;            /* contact points per env handed to the solver */
  int32_t max_colliders;           /* convex hulls per env (statics + arm links + movable hulls) */
  int32_t warps_per_block;         /* warps per block of the substep kernel; 0 = what the library was built for */
  int32_t envs_per_block;          /* environment slots per block of the substep kernel; 0 = derived from num_envs and the SM count */
  int32_t export_debug;            /* 1: every substep also writes the inspection arrays (B2S_ARR_PAIR_KEYS, LINK_POSES, LINK_VEL);
                                      0 (product default): they are only refreshed by the calls that need them */
  int32_t num_goal_steps;          /* NUM_GOAL_STEPS (push_env.py:259-262, 803-806): 0 = None: one (start, motion) pair per action;
                                      G >= 1: an action is [G][4] and the arm goes post -> pre G times before it leaves */
  int32_t use_crop;                /* 1: SegmentedPointCloudObs keeps only points inside [crop_min, crop_max] (OBS.CROP_MIN/MAX,
                                      camera_obs.py:187-193) */

  double time_step;                /* dt; reference default 1e-3 (simulator.py:26) */
  float gravity[3];                /* (0,0,-9.8) simulator.py:27 */
  float erp2;                      /* 0.08 contact ERP (pybullet) */
  float linear_slop;               /* 1e-5 */
  float warmstart;                 /* 0.85 */
  float residual_threshold;        /* 1e-7 least-squares residual early exit */
  float linear_damping;            /* 0.04 */
  float angular_damping;           /* 0.04 */
  float breaking_factor;           /* 0.02: breaking threshold = factor * min(bounding radius) */
  float ik_damping;                /* 0.1 DLS lambda */
  float ik_residual;               /* 1e-4 */
  float ik_max_step;               /* pi/4: max |dq| per DLS iteration */
  float position_gain;             /* 0.05 controllable_body.py:17 */
  float velocity_gain;             /* 1.0  controllable_body.py:18 */
  float joint_pos_threshold;       /* LIMB_POSITION_THRESHOLD, default 0.008726640 (:15) */
  float joint_vel_threshold;       /* 0.05 (:16) */
  float limb_timeout;              /* LIMB_TIMEOUT, default 15.0 (:14) */
  float limb_velocity_ratio;       /* LIMB_MAX_VELOCITY_RATIO */
  float stable_lin_threshold;      /* 0.005 simulator.py:327 */
  float stable_ang_threshold;      /* 0.005 simulator.py:328 */
  /* PushEnv action -> waypoints (push_env.py:71-81, 752-786) */
  float cspace_low[3], cspace_high[3];
  float translation_x, translation_y;
  float finger_tip_offset;         /* ARM.FINGER_TIP_OFFSET */
  float gripper_safe_height;       /* ARM.GRIPPER_SAFE_HEIGHT */
  float offstage_positions[B2S_NUM_JOINTS];
  float min_delta_position, min_delta_angle;      /* push_env.py:917-918 */
  float table_workspace_low[2], table_workspace_high[2];   /* push_env.py:85-90 */
  /* camera (bullet_camera.py:17-19) */
  float cam_near, cam_far;
  float crop_min[3], crop_max[3];  /* world frame, metres */
  /* torsional friction of the movables (tools/templates/urdf_template.xml:12-14: rolling 0.001, spinning 0.001 in every
   * generated URDF; Body.set_dynamics forwards `rolling_friction` as the spinning value too, body.py:225-230).  Per contact
   * point of a movable: one spinning row about the normal and two rolling rows about the tangents, angular only, limited
   * to mu_c * min(normal impulse, 1) with mu_c = roll_A * friction_B + roll_B * friction_A (statics and arm links: 0).
   * rolling_friction == 0 switches the three rows off (as Bullet does).  Needs friction_dirs == 2. */
  float rolling_friction, spinning_friction;
} B2SParams;

/* Scene description: host pointers, copied to the device by b2s_load_scene.
 * Produced by the host URDF/OBJ loader (robovat_b200/assets.py), which stands
 * for pybullet.loadURDF (bullet_physics.py:173-181). */
typedef struct B2SSceneDesc {
  /* convex-hull library (V-HACD output: <= 64 vertices per hull) */
  int32_t num_verts;
  const float* verts;              /* [num_verts][3] asset-local, centred on the asset's COM */
  int32_t num_hulls;
  const int32_t* hull_vert_off;    /* [num_hulls] */
  const int32_t* hull_vert_cnt;    /* [num_hulls] 1..64 (1 = sphere, 2 = capsule with margin) */
  const float* hull_margin;        /* [num_hulls] collision margin (radius for spheres/capsules) */
  /* outward face planes of each hull (n.x <= d), used by the depth/segmentation raster only */
  int32_t num_planes;
  const float* planes;             /* [num_planes][4] nx ny nz d, asset-local */
  const int32_t* hull_plane_off;   /* [num_hulls] */
  const int32_t* hull_plane_cnt;   /* [num_hulls] */
  /* assets = compounds of hulls */
  int32_t num_assets;
  const int32_t* asset_hull_off;   /* [num_assets] */
  const int32_t* asset_hull_cnt;   /* [num_assets] */
  /* static bodies in reference uid order: ground, table, [wall], tiles */
  int32_t num_statics;
  const int32_t* static_asset;     /* [num_statics] */
  const float* static_pose;        /* [num_statics][7] */
  const float* static_friction;    /* [num_statics] */
  const uint32_t* static_flags;    /* [num_statics] B2S_STATIC_* */
  /* movable asset pool (MOVABLE.<NAME>.PATHS / TARGET_PATHS, push_env.py:99-113) */
  int32_t num_movable_assets;
  const int32_t* movable_assets;   /* asset ids */
  int32_t num_target_assets;
  const int32_t* target_assets;    /* asset ids used for body 0 when the layout has a target */
  /* arm: serial chain of B2S_NUM_JOINTS revolute joints + fixed end-effector frame */
  float arm_base_pose[7];
  float joint_origin[B2S_NUM_JOINTS][7];   /* joint frame in parent link frame */
  float joint_axis[B2S_NUM_JOINTS][3];
  float joint_lower[B2S_NUM_JOINTS], joint_upper[B2S_NUM_JOINTS], joint_max_velocity[B2S_NUM_JOINTS];
  float ee_pose[7];                /* END_EFFCTOR_NAME frame in the last link's frame */
  int32_t num_links;               /* collision links: link k is rigidly attached after joint link_joint[k] */
  int32_t link_joint[B2S_MAX_LINKS];       /* -1 = fixed to the arm base */
  int32_t link_asset[B2S_MAX_LINKS];
  float link_pose[B2S_MAX_LINKS][7];       /* collision frame in the joint's link frame */
  float arm_friction;
  /* reward / reset layout (robovat/envs/push/layouts.py) in tile units */
  float tile_size; float tile_offset[2];
  int32_t num_region; float region[B2S_MAX_TILES][2];
  int32_t num_goal;   float goal[B2S_MAX_TILES][2];
  int32_t num_target; float target[B2S_MAX_TILES][2];
  int32_t num_obstacle; float obstacle[B2S_MAX_TILES][2];
  /* movable sampling ranges (MOVABLE.<NAME>.{SCALE,MASS,FRICTION,MARGIN,POSE}) */
  float scale_range[2], mass_range[2], friction_range[2];
  float pose_x[2], pose_y[2], pose_z[2], pose_roll[2], pose_pitch[2], pose_yaw[2];
  float placement_margin;
  int32_t min_movables;            /* MIN_MOVABLE_BODIES (max = params.max_movables) */
  float table_height_range[2];     /* TABLE.HEIGHT_RANGE arm_env.py:81-82 */
  float safe_drop_height;          /* 0.2 push_env.py:529 */
} B2SSceneDesc;

/* Caller-owned device buffers (torch tensors; tensor.data_ptr()).  The world
 * borrows them; they must outlive it or be re-bound.  NULL = not used. */
typedef struct B2SBuffers {
  float* body_state;       /* [13][B][Nmax]  px py pz qx qy qz qw vx vy vz wx wy wz of the movables */
  float* joint_state;      /* [2][7][B]      q, qdot of the limb */
  float* action;           /* [B][G][4]      PushEnv action in [-1,1] (G = max(1, params.num_goal_steps)) */
  float* obs_position;     /* [B][Nmax][3]   PoseObs 'position' (pose_obs.py:53-73), zero padded */
  int32_t* num_movables;   /* [B] */
  uint8_t* body_mask;      /* [B][Nmax]      movable_body_mask (push_env.py:363-365) */
  float* depth;            /* [B][H][W]      linear depth, metres */
  uint8_t* segmask;        /* [B][H][W]      body uid, 255 = background (bullet_camera.py:212) */
  float* point_cloud;      /* [B][Nmax][P][3] SegmentedPointCloudObs (camera_obs.py:182-212) */
  float* reward;           /* [B] */
  uint8_t* termination;    /* [B] */
  uint8_t* is_safe;        /* [B] attributes['is_safe']      (push_env.py:715) */
  uint8_t* is_effective;   /* [B] attributes['is_effective'] (push_env.py:727) */
  float* episode_return;   /* [B] running sum of rewards */
} B2SBuffers;

/* names for b2s_array(): world-owned device arrays exposed for inspection */
enum {
  B2S_ARR_MANIFOLD_KEYS = 0,   /* int32 [B][max_manifolds]  (colliderA << 16) | colliderB, -1 = empty */
  B2S_ARR_MANIFOLD_NPTS = 1,   /* int32 [B][max_manifolds] */
  B2S_ARR_MANIFOLD_PTS = 2,    /* float [B][max_manifolds][4][B2S_CP_FLOATS] */
  B2S_ARR_NUM_MANIFOLDS = 3,   /* int32 [B] */
  B2S_ARR_PAIR_KEYS = 4,       /* int32 [B][max_pairs] broad-phase pairs of the last substep (written when params.export_debug) */
  B2S_ARR_NUM_PAIRS = 5,       /* int32 [B] */
  B2S_ARR_PHASE = 6,           /* int32 [B] */
  B2S_ARR_NUM_STEPS = 7,       /* int32 [B] substeps executed since reset (Simulator.num_steps) */
  B2S_ARR_CTRL = 8,            /* float [B][B2S_CTRL_FLOATS] controller targets */
  B2S_ARR_CTRL_FLAGS = 9,      /* int32 [B][4] link_target_active, joint_target_active, qd_target_is_none, interrupt */
  B2S_ARR_LINK_POSES = 10,     /* float [B][num_links+1][7] world collision frames, last = end effector (refreshed by
                                  b2s_forward_kinematics / b2s_render / b2s_reset; every substep when params.export_debug) */
  B2S_ARR_MOV_PARAMS = 11,     /* float [4][B][Nmax] asset(as int bits), scale, mass, friction */
  B2S_ARR_TABLE_DZ = 12,       /* float [B] */
  B2S_ARR_ERROR_FLAGS = 13,    /* int32 [B] bit0 pair overflow, bit1 manifold overflow, bit2 non-finite state,
                                  bit3 contact overflow, bit4 colour overflow, bit5 collider overflow, bit6 solver invariant,
                                  bit7 reset found no placement with the MARGIN clearance (re-sample the env),
                                  bit8 a rollout's reset found no valid scene in max_reset_retries re-samples (the env stops) */
  B2S_ARR_WAYPOINTS = 14,      /* float [B][G][2][7] start / end gripper poses of every goal step */
  B2S_ARR_STATUS = 15,         /* float [B][2][Nmax][4] start/end status: pos3 + yaw (push_env.py:925-937) */
  B2S_ARR_CONTACT_FLAGS = 16,  /* int32 [B] bit0 arm-table, bit1 arm-movable, per last substep */
  B2S_ARR_PHASE_STATE = 17,    /* int32 [B][8] max_phase_steps, num_waypoints, settle_steps, stable_steps, ... */
  B2S_ARR_SOLVER_STATS = 18,   /* int32 [B][4] rows, colours, iterations used, contacts of the last substep */
  B2S_ARR_CTRL_TIME = 19,      /* double [B][5] link start/stop, joint start/stop, gripper-ready time */
  B2S_ARR_LINK_VEL = 20,       /* float [B][num_links][6] linear + angular velocity of the collision frames (as LINK_POSES) */
  B2S_ARR_NUM_COLLIDERS = 21,  /* int32 [B] */
  B2S_ARR_COL_SLOT = 22,       /* int32 [B][max_colliders] body slot of each collider */
  B2S_ARR_COL_HULL = 23,       /* int32 [B][max_colliders] hull id of each collider */
  B2S_ARR_PROF = 24,           /* uint64 [8] stage timing of the substep kernel (ns; tuning builds only) */
  B2S_ARR_NUM_EPISODES = 25,   /* int32 [B] episodes finished by b2s_rollout_* (RobotEnv.num_episodes, robot_env.py:262) */
  B2S_ARR_ROLLOUT_STATE = 26,  /* int32 [B][4] rollout: steps of the current episode, episodes of this rollout, re-samples, spare */
  B2S_ARR_RAY_SCENE = 27,      /* bytes: the raster's per-env camera-space scene and per-tile hull lists (valid after b2s_render;
                                  layout in b2s_obs.cu; inspection / tuning only) */
  B2S_ARR_COUNT = 28
};
#define B2S_CP_FLOATS 16   /* localA3 localB3 normalB3 dist lambda_n lambda_t1 lambda_t2, then 3 spare words: in point 0 of a
                              manifold they hold the GJK simplex of the pair's last call (int n, (ia | ib << 8) x 4) */
#define B2S_CTRL_FLOATS 40

typedef struct B2SWorld B2SWorld;

int b2s_version(void);
const char* b2s_last_error(void);
/* fills every field with the defaults quoted above (sizes are left 0) */
int b2s_default_params(B2SParams* params);

/* replaces BulletPhysics.__init__/reset/start/set_gravity (bullet_physics.py:31-137) */
int b2s_create(const B2SParams* params, int device, B2SWorld** out);
int b2s_destroy(B2SWorld* world);
/* replaces the pybullet.loadURDF calls of BulletPhysics.add_body (:143-186) for the whole scene */
int b2s_load_scene(B2SWorld* world, const B2SSceneDesc* scene);
int b2s_bind_buffers(B2SWorld* world, const B2SBuffers* buffers);
int b2s_get_params(const B2SWorld* world, B2SParams* out);

/* RobotEnv.reset scene part (robot_env.py:202-235, push_env.py:331-471): samples the table
 * height, movable count/assets/scale/mass/friction and drop poses with Philox keyed by
 * (seed, global env id), clears contacts, puts the arm at OFFSTAGE_POSITIONS.
 * env_mask: device uint8 [B] or NULL = all.  Follow with b2s_settle(). */
int b2s_reset(B2SWorld* world, const uint8_t* env_mask_dev, uint64_t seed, void* stream);
/* Simulator.wait_until_stable(movables, lin, ang, max_steps) for every env (simulator.py:325-376) */
int b2s_settle(B2SWorld* world, float lin_threshold, float ang_threshold, int max_steps, void* stream);
/* the same for the envs of env_mask_dev only (device uint8 [B], NULL = all): a partial reset must not step the
 * physics of the other envs (the reference has one world per env, so its wait never touches another env) */
int b2s_settle_masked(B2SWorld* world, const uint8_t* env_mask_dev, float lin_threshold, float ang_threshold, int max_steps,
                      void* stream);
/* end of RobotEnv.reset (robot_env.py:224-235): the settled scene becomes the episode's first observation, i.e. the
 * `prev_obs_data` of the first PushReward.get_reward (push_reward.py:396-405).  Stores the current movable xy as the
 * reward's previous state for the envs of env_mask_dev (NULL = all). */
int b2s_begin_episode(B2SWorld* world, const uint8_t* env_mask_dev, void* stream);

/* Simulator.step x n for every env, phase machine untouched
 * (= ControllableBody.update + pybullet.stepSimulation; simulator.py:94-103).  The raw substep call of the parity
 * tests; bench.py steps through b2s_rollout_run (`value`) and b2s_env_async_step_free (`e2e`). */
int b2s_step(B2SWorld* world, int n_substeps, void* stream);
/* the same substeps as n launches of one substep each (launch-granular profiling; results identical to b2s_step) */
int b2s_step_staged(B2SWorld* world, int n_substeps, void* stream);

/* PushEnv._execute_action (push_env.py:631-733): b2s_set_action computes the waypoints from
 * buffers.action and starts the phase machine ('initial'); b2s_env_substeps advances every env whose
 * action is still in flight by up to n substeps (finished envs are frozen) and writes the number of
 * unfinished envs to *unfinished_host (after synchronising the stream) when it is not NULL. */
int b2s_set_action(B2SWorld* world, void* stream);
int b2s_env_substeps(B2SWorld* world, int n_substeps, int* unfinished_host, void* stream);
/* the same as a free-running launch (see B2SRollout.free_running): n_substeps x (envs in flight) substeps in total, at
 * most 4 n_substeps for one env.  For callers that loop until every action has finished (PushEnv._execute_action does):
 * which env gets how many substeps per call does not matter to them, and no SM waits for the slowest block. */
int b2s_env_substeps_free(B2SWorld* world, int n_substeps, int* unfinished_host, void* stream);
/* convenience: set_action + env_substeps until every env finished (or max_substeps) */
int b2s_env_step(B2SWorld* world, int chunk, int max_substeps, void* stream);

/* Episodes on the device: the loops of robovat/io/episode_generation.py:41-61 (`action = policy.action(obs);
 * obs, reward, done, _ = env.step(action)` until done or num_steps) and :88-112 (episode after episode, `env.reset()`
 * in between) with HeuristicPushPolicy (robovat/policies/push_policy.py:33-52 over heuristic_push_sampler.py:66-123),
 * for every env independently and without returning to the host: an env that finishes an action computes its reward
 * (PushReward), records the transition, draws its next action and starts it inside the same kernel launch; an env
 * whose episode is over samples its next scene, lets it drop and settle (re-sampling scenes whose bodies fell off the
 * table) and starts the next episode.  Nobody waits for the slowest env of the batch -- this is what
 * tools/parallel_run.py's independent worker processes amount to.  All pointers are device pointers owned by the
 * caller; output pointers may be NULL.  Records of episode ep of env e start at index (e * num_episodes + ep). */
enum { B2S_POLICY_HEURISTIC = 0, B2S_POLICY_AIMED = 1 };
typedef struct B2SRollout {
  int32_t num_actions;         /* A: steps per episode at most (config MAX_STEPS) */
  int32_t max_attempts;        /* HEURISTICS.MAX_ATTEMPS of the rejection sampler (1..65535) */
  int32_t num_episodes;        /* EP: episodes per env in this rollout; an env goes idle after its last one */
  int32_t max_reset_retries;   /* re-samples of a scene that came out invalid before the env gives up (error bit8) */
  uint64_t seed;               /* policy seed: Philox key; counter = (stream 1, step, global env id, RobotEnv.num_episodes, attempt) */
  uint64_t reset_seed;         /* scene seed of the resets between episodes (as b2s_reset's seed) */
  float drop_lin_threshold, drop_ang_threshold;   /* 0.1, 0.1  push_env.py:443-447 */
  int32_t drop_max_steps;      /* 500 */
  int32_t policy_kind;         /* B2S_POLICY_HEURISTIC (HeuristicPushSampler), B2S_POLICY_AIMED (synthetic workloads: start
                                  8 cm behind a random body in a random direction and push through it; always a contact) */
  int32_t free_running;        /* 0: every launch advances every env by exactly `chunk` substeps.  1: a launch ends when it has
                                  executed chunk x (running envs) substeps IN TOTAL, all thread blocks stopping together -- envs
                                  that are cheap to step get ahead of expensive ones and no SM waits for the slowest block.
                                  What an env computes is unchanged (its episodes do not depend on the schedule); only how far
                                  each env has got when a call returns is. */
  int32_t reserved;
  const float* first_action;   /* [B][4] action of step 0 of episode 0, or NULL: drawn by the device policy like the others */
  float* actions;              /* out [B][EP][A][4] */
  float* rewards;              /* out [B][EP][A] */
  float* positions;            /* out [B][EP][A+1][Nmax][3] PoseObs 'position' before step 0 and after every step */
  uint8_t* flags;              /* out [B][EP][A] bit0 is_safe, bit1 is_effective, bit2 termination (reward fn), bit3 unsafe at 'done' */
  int32_t* substeps;           /* out [B][EP][A] Simulator.num_steps after the step */
  int32_t* lengths;            /* out [B][EP] steps taken (the episode stops at termination / unsafe-at-done / A) */
  float* returns;              /* out [B][EP] RobotEnv.episode_reward at the end of the episode */
} B2SRollout;
/* starts episode 0 in every env from its current (reset and settled) state */
int b2s_rollout_begin(B2SWorld* world, const B2SRollout* rollout, void* stream);
/* advances the rollout by launches of `chunk` substeps queued back to back until every env has finished its last
 * episode or max_substeps per env were launched (the count of unfinished envs is read back asynchronously two launches
 * behind, so the queue never drains).  *unfinished_host (may be NULL) receives the final count after synchronising the
 * stream.  May be called repeatedly: a rollout keeps its state between calls. */
int b2s_rollout_run(B2SWorld* world, int chunk, int max_substeps, int* unfinished_host, void* stream);

/* Asynchronous stepping with the policy on the host (the other way to never wait for the slowest env: what
 * tools/parallel_run.py's workers do, each at its own pace, with `policy.action(obs)` / `env.step(action)` /
 * `env.reset()` of episode_generation.py:41-61 issued per env).  command_dev uint8 [B] (NULL = all 0), looked at only
 * for envs that are ready (idle): 1 = start the action in buffers.action[e] (PushEnv._execute_action), 2 = re-sample
 * the env's scene with reset_seed, let it drop and settle on the device (push_env.py:331-471; invalid scenes are
 * re-sampled up to 8 times).  Then every busy env advances by up to n_substeps.  An env whose action finishes computes
 * its reward (buffers.reward / termination / episode_return / is_safe / is_effective, PoseObs row) and waits; an env
 * whose reset finishes publishes its first observation and waits.  status_dev uint8 [B] (may be NULL): bit0 ready
 * for a command, bit1 an action finished during this call, bit2 a reset finished during this call, bit3 the env ended
 * its last action unsafe at 'done' (RobotEnv._done, push_env.py:719-721). */
int b2s_env_async_step(B2SWorld* world, const uint8_t* command_dev, int n_substeps, uint64_t reset_seed, uint8_t* status_dev,
                       void* stream);
/* the same with a free-running launch (see B2SRollout.free_running): n_substeps x (busy envs) substeps in total, at most
 * 4 n_substeps for one env */
int b2s_env_async_step_free(B2SWorld* world, const uint8_t* command_dev, int n_substeps, uint64_t reset_seed, uint8_t* status_dev,
                            void* stream);

/* robot commands outside the phase machine (sawyer_sim.py:186-308); poses/q are device pointers */
int b2s_arm_move_to_gripper_pose(B2SWorld* world, const float* pose_dev /*[B][7]*/, const uint8_t* env_mask_dev, void* stream);
int b2s_arm_move_to_joint_positions(B2SWorld* world, const float* q_dev /*[B][7]*/, const uint8_t* env_mask_dev, void* stream);
int b2s_arm_reset_targets(B2SWorld* world, const uint8_t* env_mask_dev, void* stream);
/* BulletPhysics.position_control_array (bullet_physics.py:1061-1104): latch the POSITION_CONTROL motor
 * targets of the 7 limb joints (q_dev, qd_dev: device [B][7]; qd_dev NULL = zero target velocity).  Used
 * when the reference's own Python ControllableBody drives the arm substep by substep. */
int b2s_set_motor_targets(B2SWorld* world, const float* q_dev, const float* qd_dev, const uint8_t* env_mask_dev, void* stream);
/* re-derive the collider list after the host edited num_movables / the movable assets
 * (BulletPhysics.add_body / remove_body of a movable, bullet_physics.py:143-195) */
int b2s_rebuild_colliders(B2SWorld* world, void* stream);
/* SawyerSim.is_limb_ready (sawyer_sim.py:394-400) -> uint8 [B] */
int b2s_arm_is_ready(B2SWorld* world, uint8_t* out_dev, void* stream);
/* BulletPhysics.compute_inverse_kinematics (bullet_physics.py:1203-1262), one solve from q_start */
int b2s_inverse_kinematics(B2SWorld* world, const float* pose_dev /*[B][7]*/, const float* q_start_dev /*[7][B]*/,
                           float* q_out_dev /*[7][B]*/, void* stream);
/* forward kinematics of the bound joint state -> B2S_ARR_LINK_POSES */
int b2s_forward_kinematics(B2SWorld* world, void* stream);

/* Simulator.check_contact(arm, table) / (arm, movables) (simulator.py:246-287): uint8 [B] each */
int b2s_query_contacts(B2SWorld* world, uint8_t* arm_table_dev, uint8_t* arm_movable_dev, void* stream);

/* observations: PoseObs + attribute obs -> bound buffers */
int b2s_observe(B2SWorld* world, void* stream);
/* BulletCamera._frames + set_calibration (bullet_camera.py:188-258): K row-major [9], R [9], t [3]
 * (x_cam = R x_world + t), optionally per env (per_env != 0 -> arrays are [B][..]) on the HOST. */
int b2s_set_camera(B2SWorld* world, const float* K, const float* R, const float* t, int per_env);
int b2s_render(B2SWorld* world, void* stream);
/* SegmentedPointCloudObs.get_observation (camera_obs.py:182-212) from the last render */
int b2s_point_cloud(B2SWorld* world, uint64_t seed, void* stream);

/* PushReward.get_reward (push_reward.py:396-405): prev_xy / next_xy device [B][Nmax][2];
 * NULL next = current body positions, NULL prev = positions stored by the last call. */
int b2s_reward(B2SWorld* world, const float* prev_xy_dev, const float* next_xy_dev, void* stream);

/* replaces tools/parallel_run.py: one all-gather of episode returns across ranks.
 * nccl_comm is an ncclComm_t; out_dev is [world_size * B] floats. */
int b2s_allgather_returns(B2SWorld* world, void* nccl_comm, float* out_dev, void* stream);

/* inspection: device pointer + byte size of a world-owned array */
int b2s_array(B2SWorld* world, int which, void** dev_ptr, int64_t* bytes);
/* sizeof(B2SParams / B2SSceneDesc / B2SBuffers / B2SRollout) for which = 0 / 1 / 2 / 3: lets a binding check its layout */
int b2s_sizeof(int which);
/* number of kernels this library has launched on this world since creation */
int64_t b2s_launch_count(const B2SWorld* world);
/* sum over envs of substeps executed since creation (device counter, synchronises) */
int64_t b2s_substeps_executed(B2SWorld* world, void* stream);

/* device-side SE(3) used by the kernels, exposed for parity against robovat.math /
 * third_party.transformations: arrays are device pointers of n items */
int b2s_se3_quat_from_euler(const float* euler_dev, float* quat_dev, int n, void* stream);
int b2s_se3_euler_from_quat(const float* quat_dev, float* euler_dev, int n, void* stream);
int b2s_se3_matrix_from_quat(const float* quat_dev, float* m_dev, int n, void* stream);
int b2s_se3_quat_multiply(const float* a_dev, const float* b_dev, float* out_dev, int n, void* stream);
int b2s_se3_pose_inverse(const float* pose_dev, float* out_dev, int n, void* stream);
int b2s_se3_pose_transform(const float* a_dev, const float* b_dev, float* out_dev, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* B2S_H_ */
