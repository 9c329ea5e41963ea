// b2s_dev.cuh -- device-side world description shared by the kernels and the host API.
//
// Data layout in HBM (all SoA over environments, env index fastest where one
// thread handles one env, env-major blocks where one warp handles one env):
//   body_state  float [13][B][Nmax]   px py pz qx qy qz qw vx vy vz wx wy wz  (torch-owned)
//   joint_state float [2][7][B]       q, qdot                                   (torch-owned)
//   manifolds   ping-pong  keys int32 [2][B][M], npts int32 [2][B][M],
//               pts float [2][B][M][4][16]; parity int32 [B]
//   scene       verts float4 [V], hull table, asset table, statics: read-only
// One environment is stepped by one warp; all per-substep intermediates (body
// table, collider table, AABBs, pair list, contact rows) live in that warp's
// slice of shared memory and never touch HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2s.h"
#include "../../include/b2s_geom.h"
#include "../../include/b2s_math.h"

typedef b2s_v3 V3;
typedef b2s_q4 Q4;
typedef b2s_m3 M3;

#define B2S_TYPE_STATIC 0
#define B2S_TYPE_KINEMATIC 1
#define B2S_TYPE_DYNAMIC 2

// narrow-phase unit: the lanes that work on one candidate pair.  B2S_HALF = log2(units per warp): 0 = the whole warp
// on one pair, 1 = two pairs of 16 lanes, 2 = four pairs of 8 lanes, 3 = eight pairs of 4 lanes.  Hulls have 8..64
// vertices and most of GJK is uniform simplex arithmetic, so narrow units waste fewer lanes; the units diverge only
// where their pairs differ.  A call takes about as long with one pair as with eight (it is a latency chain), and the
// stage ends with its slowest warp: with eight pairs per grab the ~92 pairs of a block's round are one grab per warp
// (measured in free-running rollouts: 16 lanes 21.7, 8 lanes 23.4, 4 lanes 24.5 M substeps/s).
#ifndef B2S_HALF
#define B2S_HALF 3
#endif
#if B2S_HALF
#define UW (32 >> B2S_HALF)
#define UL (lane & (UW - 1))
#define UB (lane & ~(UW - 1))
#define UM ((0xffffffffu >> (32 - UW)) << UB)
#define UH (lane / UW)
#define UNITS_PER_WARP (1 << B2S_HALF)
#else
#define UW 32
#define UL lane
#define UB 0
#define UM FULL
#define UH 0
#define UNITS_PER_WARP 1
#endif
#define EPA_MAXV 32
#define EPA_MAXF 96
#define FULL 0xffffffffu

struct DHull {
  int voff, vcnt;
  float margin, rad;
  float lc[3], lh[3];
  int poff, pcnt;
};

struct DAsset {
  int hoff, hcnt;
  float half[3];
  float pad;
};

// scalar scene data small enough to sit in the kernel parameter block / constant bank
struct DArm {
  float base[7];
  float joint_origin[B2S_NUM_JOINTS][7];
  float joint_axis[B2S_NUM_JOINTS][3];
  float lower[B2S_NUM_JOINTS], upper[B2S_NUM_JOINTS], max_vel[B2S_NUM_JOINTS];
  float ee[7];
  int num_links;
  int link_joint[B2S_MAX_LINKS];
  int link_asset[B2S_MAX_LINKS];
  float link_pose[B2S_MAX_LINKS][7];
  float friction;
};

struct DLayout {
  float tile_size, tile_offset[2];
  int num_region, num_goal, num_target, num_obstacle;
  float region[B2S_MAX_TILES][2], goal[B2S_MAX_TILES][2], target[B2S_MAX_TILES][2], obstacle[B2S_MAX_TILES][2];
  float scale_range[2], mass_range[2], friction_range[2];
  float pose_x[2], pose_y[2], pose_z[2], pose_roll[2], pose_pitch[2], pose_yaw[2];
  float placement_margin;
  int min_movables;
  float table_height_range[2], safe_drop_height;
  int num_movable_assets, num_target_assets;
};

// device-side episode driver (b2s_rollout_*, b2s_env_async_step; b2s_rollout.cuh)
enum { RO_OFF = 0, RO_EPISODES = 1, RO_ASYNC = 2 };    // DRollout.enabled
struct DRollout {
  int enabled, num_actions, max_attempts, num_episodes, max_reset_retries, drop_max_steps, policy_kind, free_running;
  float drop_lin, drop_ang;
  unsigned long long seed, reset_seed;
  float* actions; float* rewards; float* positions; uint8_t* flags; int32_t* substeps; int32_t* lengths; float* returns;
};

// Shared memory of a block = E per-environment regions, one scratch region per warp, E meta records.
// A warp may pick up any environment of its block in any stage, so everything that has to survive from
// one stage to the next (body table, colliders, pair list, contact list) is per environment; GJK, manifold
// staging, FK scratch and the B-side row data are per warp.  `con` aliases the narrow-phase scratch
// (stage .. simplex): the solve-stage tables are only alive in the solve stage.  The EPA polytope (rare path, 4 KB per
// warp) lives in global memory so that the block leaves more of the SM's 256 KB to the L1 cache.
struct SmemLayout {
  int body, col, pairs, cmk, pstage, ps_cap, words_env;        // per environment (pstage: staging records of the first ps_cap candidate pairs)
  int con, stage, fk, simplex, words_warp;                     // per warp: scratch
};
#define META_ACTIVE 0
#define META_NP 1
#define META_COST 2           // solver work of the environment's previous substep: stage C hands the expensive ones out first
#define META_NEWN 3
#define META_SETTLE_STEPS 4
#define META_SETTLE_STABLE 5
#define META_FINISHED 6
#define META_ENV 7            // environment index of the slot in this launch
#define META_WORDS 8

#define BODY_STRIDE 35   // pos3 R9 vel3 ang3 invm1 invI9 fric1 type1 quat4 = 34 (+1 pad, odd stride)
#define COL_STRIDE 13    // hull slot type|flags scale margin rad amin3 amax3 = 12 (+1)
#define FK_WORDS 96      // frames 7x7, axes 7x3, origins 7x3 = 91

struct DWorld {
  B2SParams P;
  int B, Nmax, Ns, L, NB, Hmax;
  int G;                 // goal steps per action: max(1, P.num_goal_steps)
  // scene (device, read-only)
  const float4* verts;
  const DHull* hulls;
  const DAsset* assets;
  const float4* planes;
  const int* static_asset;
  const float* static_pose;      // [Ns][7]
  const float* static_friction;
  const uint32_t* static_flags;
  const int* movable_assets;
  const int* target_assets;
  const DArm* arm;
  const DLayout* layout;
  // caller-owned buffers
  B2SBuffers buf;
  // world-owned arrays
  int32_t* man_keys;   // [2][B][M]
  int32_t* man_npts;   // [2][B][M]
  float* man_pts;      // [2][B][M][4][16]
  int32_t* man_parity; // [B]
  int32_t* num_manifolds;
  int32_t* pair_keys;
  int32_t* num_pairs;
  int32_t* phase;
  int32_t* num_steps;
  float* ctrl;
  int32_t* ctrl_flags;
  double* ctrl_time;
  float* link_poses;
  float* link_vel;
  float* mov_params;
  float* table_dz;
  int32_t* error_flags;
  float* waypoints;
  float* status;
  int32_t* contact_flags;
  int32_t* phase_state;
  int32_t* solver_stats;
  int32_t* ncol;
  int32_t* col_slot;
  int32_t* col_hull;
  int32_t* reset_count;
  float* prev_xy;
  float* cam;            // [B][21]
  unsigned long long* substeps;   // device counter: total env-substeps executed
  int32_t* unfinished;            // device counter used by env_substeps
  float* epa_scratch;             // [blocks][warps][EP_WORDS] EPA polytope (rare path: lives in L2, not in shared memory)
  int32_t* env_map;               // [blocks][E] environment stepped in a block slot (-1 none), re-dealt before every launch
  unsigned long long* prof;       // [8] stage timing counters (only written by -DB2S_PROF builds)
  float* pair_stage;              // [blocks][E][max_pairs][68] narrow-phase result of every candidate pair of the substep
  float* row_scratch;             // [blocks][warps][max_contacts][124] contact records of substep_post_big (large scenes only; L2 resident)
  int32_t* ro_state;              // [B][4] rollout: step of the episode, episode index, re-samples of the current reset, spare
  int32_t* num_episodes;          // [B] episodes finished by the device-side driver
  unsigned long long* free_target;   // [2] value of *substeps at which a free-running launch stops; start of the env window of the next one
  float* work_ema;                // [B] running mean of an environment's solver work (colours x iterations): what the deal of a free-running launch ranks by
  int32_t* async_events;          // [B] b2s_env_async_step: what happened to the env since the last call (B2S_ASYNC_*)
  DRollout ro;
  SmemLayout sm;
  unsigned char* ray_scratch;   // raster: per-environment camera-space scene + per-tile hull lists (allocated at the first render)
  int max_ray_planes;    // raster: upper bound of hull face planes / hulls in one environment
  int max_ray_cols;
  int envs_per_block;    // E: environment SLOTS of a block (capacity; the deal may leave some empty)
  int num_blocks;        // blocks of the substep kernel
  int reg_rows;          // 1: contacts fit one per lane and body slots one per lane (max_contacts <= 32, NB <= 32): small solve path
};

enum { MODE_RAW = 0, MODE_ENV = 1, MODE_SETTLE = 2 };

// The opt-in to more than 48 KB of dynamic shared memory is an attribute of (function, device): remember what was
// configured per device, not per process (a second world on another GPU would launch unconfigured otherwise).
#define B2S_MAX_DEVICES 64
template <class K>
static inline void b2s_opt_in_smem(K kernel, size_t smem, size_t* configured /* [B2S_MAX_DEVICES], zero initialised */) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= B2S_MAX_DEVICES) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); return; }
  if (smem > configured[dev]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = smem;
  }
}

// host launchers (defined next to their kernels)
// free_chunk > 0 (MODE_ENV only): a free-running launch -- the blocks stop together once the launch has executed
// free_chunk substeps per stepping environment IN TOTAL (n then only caps what one environment may take).
// run_ahead (free-running launches): a long solve does not hold its block, its environment sits out the block's next
// round(s) instead (b2s_step.cu).  Right for rollouts and asynchronous stepping, where only the total counts; wrong for
// the lock-step PushEnv.step, whose batch waits for exactly those environments.
void b2s_launch_substeps(const DWorld& W, int n, int mode, float lin, float ang, int max_steps, const uint8_t* env_mask, cudaStream_t s,
                         int free_chunk = 0, bool run_ahead = true);
void b2s_launch_begin_episode(const DWorld& W, const uint8_t* mask, cudaStream_t s);
size_t b2s_render_scratch_bytes(const DWorld& W);
void b2s_launch_rollout_begin(const DWorld& W, const float* first_action, cudaStream_t s);
void b2s_launch_async_commands(const DWorld& W, const uint8_t* command, cudaStream_t s);
void b2s_launch_async_status(const DWorld& W, uint8_t* status, cudaStream_t s);
// returns the blocks to launch; resident_blocks: how many blocks of the substep kernel the device holds at once
int b2s_launch_assign_envs(const DWorld& W, int mode, cudaStream_t s, int free_chunk = 0, int resident_blocks = 1 << 30);
void b2s_launch_count_running(const DWorld& W, cudaStream_t s);
void b2s_launch_staged(const DWorld& W, int n, cudaStream_t s, int64_t* launches);
void b2s_launch_reset(const DWorld& W, const uint8_t* mask, uint64_t seed, cudaStream_t s);
void b2s_launch_set_action(const DWorld& W, cudaStream_t s);
void b2s_launch_observe(const DWorld& W, cudaStream_t s);
void b2s_launch_reward(const DWorld& W, const float* prev_xy, const float* next_xy, cudaStream_t s);
void b2s_launch_arm_cmd(const DWorld& W, int cmd, const float* data, const uint8_t* mask, uint8_t* out, cudaStream_t s);
void b2s_launch_rebuild_colliders(const DWorld& W, cudaStream_t s);
void b2s_launch_ik(const DWorld& W, const float* pose, const float* q_start, float* q_out, cudaStream_t s);
void b2s_launch_fk(const DWorld& W, cudaStream_t s);
void b2s_launch_query_contacts(const DWorld& W, uint8_t* arm_table, uint8_t* arm_movable, cudaStream_t s);
void b2s_launch_export_manifolds(const DWorld& W, int32_t* keys, int32_t* npts, float* pts, cudaStream_t s);
void b2s_launch_render(const DWorld& W, cudaStream_t s);
void b2s_launch_point_cloud(const DWorld& W, uint64_t seed, cudaStream_t s);
void b2s_launch_se3(int op, const float* a, const float* b, float* out, int n, cudaStream_t s);
size_t b2s_smem_bytes(const DWorld& W);
