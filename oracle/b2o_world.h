/* b2o_world.h -- data model of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call it.
 *
 * PARITY UNPINNED (physics): the reference's arithmetic for this path lives in
 * the third-party wheel pybullet==2.6.5 (requirements.txt:9, setup.py:43),
 * which is neither vendored in /root/reference nor installable here, and the
 * reference has no tests or golden vectors for poses, contacts, IK or images
 * (SURVEY.md section 4).  This oracle therefore restates (a) robovat's own
 * control flow and constants exactly, citing file:line, and (b) the published
 * Bullet algorithms (GJK/EPA + persistent 4-point manifolds + sequential-impulse
 * PGS) from recall with every constant a named parameter in B2SParams.  The
 * parts that ARE pinned against the importable reference (robovat.math,
 * third_party.transformations, push_reward, layouts, heuristic_push_sampler,
 * Camera) are checked through tests/golden/ fixtures.
 *
 * The oracle is scalar and sequential: one environment at a time, plain loops.
 */
#ifndef B2O_WORLD_H_
#define B2O_WORLD_H_

#include <stdint.h>
#include <string>
#include <vector>

#include "../include/b2s.h"
#include "../include/b2s_geom.h"
#include "../include/b2s_math.h"

#ifndef B2S_F64
typedef float abi_float;     /* the float of the C-ABI structures (only differs from `float` in the double build, b2o_f64.h) */
#endif

namespace b2o {

typedef b2s_v3 V3;
typedef b2s_q4 Q4;
typedef b2s_m3 M3;

enum { TYPE_STATIC = 0, TYPE_KINEMATIC = 1, TYPE_DYNAMIC = 2 };
enum { EPA_MAXV = 32, EPA_MAXF = 96, MAX_COLOURS = 64 };

struct Hull {
  int voff, vcnt;
  float margin;
  V3 lc, lh;      /* local AABB centre / half extents of the vertices */
  float rad;      /* max |v| */
  int poff, pcnt; /* face planes (raster) */
};

struct Asset {
  int hoff, hcnt;
  V3 half;        /* half extents of the AABB of all hull vertices (inertia box) */
};

struct Scene {
  std::vector<V3> verts;
  std::vector<Hull> hulls;
  std::vector<Asset> assets;
  std::vector<float> planes; /* [n][4] */
  B2SSceneDesc d;            /* scalar fields + fixed-size arrays; pointer members are NOT valid */
  std::vector<int> static_asset;
  std::vector<float> static_pose, static_friction;
  std::vector<uint32_t> static_flags;
  std::vector<int> movable_assets, target_assets;
};

/* world transform of one collider for the current substep */
struct ColX {
  V3 pos; M3 R; float scale, margin, rad;
  int hull, slot, type; uint32_t flags;
  V3 amin, amax;
};

struct BodyX {       /* per body slot, rebuilt each substep */
  V3 pos; Q4 quat; V3 vel, ang;
  float inv_mass; M3 inv_inertia; /* world */
  float friction; int type;
};

struct ContactRow {
  V3 dir, angA, angB, iangA, iangB;
  V3 dirMA, dirMB;                 /* dir / mass of A, of B: what a unit impulse adds to the linear velocities */
  float inv_d, d, bias, lambda;
};
struct Contact {
  int slotA, slotB; float mu;
  ContactRow row[3];
  /* torsional rows (spinning about row[0].dir, rolling about row[1].dir and row[2].dir; angular only, no warm start):
     I^-1 axis of both bodies, 1/d, impulse; mu_t[0] = combined spinning, mu_t[1] = combined rolling coefficient */
  V3 tiA[3], tiB[3]; float tinv_d[3], tl[3]; float mu_t[2]; int tors;
  int manifold, point; int colour;
};

struct World {
  B2SParams P;
  Scene S;
  int B, Nmax, Ns, L, NB, Hmax;
  /* user-facing state (same layouts as B2SBuffers) */
  std::vector<float> body_state;   /* [13][B][Nmax] */
  std::vector<float> joint_state;  /* [2][7][B] */
  std::vector<float> action;       /* [B][4] */
  std::vector<float> obs_position; /* [B][Nmax][3] */
  std::vector<int32_t> num_movables;
  std::vector<uint8_t> body_mask;
  std::vector<float> depth; std::vector<uint8_t> segmask; std::vector<float> point_cloud;
  std::vector<float> reward; std::vector<uint8_t> termination, is_safe, is_effective;
  std::vector<float> episode_return;
  /* world-owned arrays (same layouts as the B2S_ARR_* arrays) */
  std::vector<int32_t> man_keys, man_npts, num_manifolds;
  std::vector<float> man_pts;
  std::vector<int32_t> pair_keys, num_pairs;
  std::vector<int32_t> phase, num_steps;
  std::vector<float> ctrl; std::vector<int32_t> ctrl_flags;
  std::vector<double> ctrl_time;   /* [B][5] link start/stop, joint start/stop, gripper ready */
  std::vector<float> link_poses;   /* [B][L+1][7] */
  std::vector<float> link_vel;     /* [B][L][6] */
  std::vector<float> mov_params;   /* [4][B][Nmax] */
  std::vector<float> table_dz;
  std::vector<int32_t> error_flags;
  std::vector<float> waypoints; std::vector<float> status;
  std::vector<int32_t> contact_flags, phase_state, solver_stats;
  std::vector<int32_t> ncol, col_slot, col_hull;  /* [B], [B][Hmax] */
  std::vector<int32_t> reset_count;
  /* device-side episode driver (b2s_rollout_*): settings, per-env state [B][4] (step, episode, re-samples, spare),
   * RobotEnv.num_episodes [B]; record pointers are caller-owned host arrays */
  struct Rollout {
    int enabled, num_actions, max_attempts, num_episodes, max_reset_retries, drop_max_steps, policy_kind;
    float drop_lin, drop_ang;
    uint64_t seed, reset_seed;
    float* actions; float* rewards; float* positions; uint8_t* flags; int32_t* substeps; int32_t* lengths; float* returns;
  } ro;
  std::vector<int32_t> ro_state, num_episodes, async_events;
  std::vector<float> prev_xy;      /* [B][Nmax][2] */
  std::vector<float> cam;          /* [B][21] K9 R9 t3 */
  int cam_per_env;
  int64_t substeps_executed;
  std::vector<int64_t> substeps_executed_env;   /* [B] substeps of each env since creation (num_steps restarts at a reset) */
  std::string err;
};

/* ---- physics (b2o_physics.cpp) ---- */
void derive_scene(Scene& S);
void build_colliders(World& w, int e);
void substep(World& w, int e);
int support(const World& w, const ColX& c, V3 d, V3* p);
/* narrow phase of one pair; returns 1 and fills the contact when distance < threshold */
int collide_pair(const World& w, const ColX& A, const ColX& B, float threshold,
                 V3* pA, V3* pB, V3* normal, float* distance, int32_t* cache);

/* ---- arm (b2o_arm.cpp) ---- */
void arm_fk(const World& w, const float* q, const float* qd, float* link_poses /*[L+1][7]*/,
            float* link_vel /*[L][6] or NULL*/);
void arm_ik(const World& w, const float* target_pose, const float* q_start, float* q_out);
void arm_update(World& w, int e);          /* ControllableBody.update + motor */
void arm_reset_targets(World& w, int e);
void arm_set_link_target(World& w, int e, const float* pose);
void arm_set_joint_target(World& w, int e, const float* q);
int arm_is_ready(World& w, int e);

/* ---- env (b2o_env.cpp) ---- */
void reset_env(World& w, int e, uint64_t seed);
void set_action(World& w, int e);
void env_substep(World& w, int e);         /* one substep + phase logic for an in-flight action */
void observe(World& w, int e);
void reward(World& w, int e, const float* prev_xy, const float* next_xy);
void policy_sample(World& w, int e, uint64_t seed, int action_index, int num_episodes, int max_attempts, float out[4]);
void policy_aimed(World& w, int e, uint64_t seed, int action_index, int num_episodes, float out[4]);
void rollout_begin(World& w, int e, const float* first_action);
void rollout_substep(World& w, int e);     /* env_substep + what follows an action / an episode / a reset */
void async_command(World& w, int e, int cmd);
void async_substep(World& w, int e);
void render(World& w, int e);
void point_cloud(World& w, int e, uint64_t seed);

}  // namespace b2o
#endif
