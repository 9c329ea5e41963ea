/* b2o_env.cpp -- CPU oracle: scene reset, PushEnv action execution (phase
 * machine), PoseObs and the PushEnv reward.
 *
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h).
 *
 * Restates, for one environment:
 *   PushEnv._reset_scene / _load_movable_bodies / _sample_body_poses*   push_env.py:331-597
 *   PushEnv._execute_action and helpers                                 push_env.py:631-937
 *   Simulator.wait_until_stable / check_stable                          simulator.py:289-376
 *   PoseObs.get_observation                                             pose_obs.py:53-73
 *   push_reward.get_reward_fn(...).reward_fn (is_planning=False)         push_reward.py:272-374
 * Deliberate deviations (documented in DESIGN.md): all movables are dropped at
 * once instead of one by one; the placement loop accepts the first valid
 * arrangement (the reference's `if i == num_attemps` exit test, push_env.py:520,
 * is a quirk); random draws are Philox keyed by (seed, global env id).
 */
#include <math.h>
#include <string.h>

#include "b2o_world.h"

namespace b2o {

struct Rng {
  uint32_t k0, k1, c1, c2, c3, blk;
  b2s_u4 buf; int have;
  Rng(uint64_t seed, uint32_t stream, uint32_t env, uint32_t attempt)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c1(stream), c2(env), c3(attempt), blk(0), have(0) {}
  uint32_t next() {
    if (!have) { buf = b2s_philox(k0, k1, blk++, c1, c2, c3); have = 4; }
    uint32_t r = (have == 4) ? buf.x : (have == 3) ? buf.y : (have == 2) ? buf.z : buf.w;
    --have;
    return r;
  }
  float uni(float lo, float hi) { return lo + (hi - lo) * b2s_u01(next()); }
  int below(int n) { int k = (int)(b2s_u01(next()) * (float)n); return k < n ? k : n - 1; }
};

static inline float& bs(World& w, int c, int e, int i) { return w.body_state[((size_t)c * w.B + e) * w.Nmax + i]; }
static inline float& mp_(World& w, int c, int e, int i) { return w.mov_params[((size_t)c * w.B + e) * w.Nmax + i]; }

void reset_env(World& w, int e, uint64_t seed) {
  const B2SSceneDesc& d = w.S.d;
  const B2SParams& P = w.P;
  const int Nmax = w.Nmax;
  Rng rng(seed, 0u, (uint32_t)(P.env_id_offset + e), (uint32_t)w.reset_count[e]);
  w.reset_count[e] += 1;
  float dz = rng.uni(d.table_height_range[0], d.table_height_range[1]);
  w.table_dz[e] = dz;
  int span = Nmax - d.min_movables + 1;
  int n = d.min_movables + rng.below(span > 0 ? span : 1);
  if (n > Nmax) n = Nmax;
  w.num_movables[e] = n;
  /* table top = z of the static flagged IS_TABLE (URDF origin at the top surface) */
  float table_z = 0.0f;
  for (int s = 0; s < w.Ns; ++s) if (w.S.static_flags[s] & B2S_STATIC_IS_TABLE) table_z = w.S.static_pose[s * 7 + 2] + dz;
  float px[64], py[64], pz[64], er[64], ep[64], ey[64];
  bool placed = false;
  for (int round = 0; round < 64; ++round) {
    bool all_ok = true;
    for (int i = 0; i < n; ++i) {
      bool ok = false;
      for (int att = 0; att <= 32 && !ok; ++att) {
        float x, y, z, ro, pi, ya;
        const bool use_target = (i == 0 && d.num_target > 0);
        const int nt = use_target ? d.num_target : d.num_obstacle;
        if (nt > 0) {
          const abi_float(*tiles)[2] = use_target ? d.target : d.obstacle;
          int t = rng.below(nt);
          x = rng.uni(d.tile_offset[0] + (tiles[t][0] - 0.5f) * d.tile_size, d.tile_offset[0] + (tiles[t][0] + 0.5f) * d.tile_size);
          y = rng.uni(d.tile_offset[1] + (tiles[t][1] - 0.5f) * d.tile_size, d.tile_offset[1] + (tiles[t][1] + 0.5f) * d.tile_size);
          z = table_z + d.safe_drop_height;
          ro = rng.uni(-B2S_PI, B2S_PI); pi = rng.uni(-B2S_HALF_PI, B2S_HALF_PI); ya = rng.uni(-B2S_PI, B2S_PI);
        } else {
          x = rng.uni(d.pose_x[0], d.pose_x[1]); y = rng.uni(d.pose_y[0], d.pose_y[1]);
          z = rng.uni(d.pose_z[0], d.pose_z[1]) + dz;
          ro = rng.uni(d.pose_roll[0], d.pose_roll[1]); pi = rng.uni(d.pose_pitch[0], d.pose_pitch[1]);
          ya = rng.uni(d.pose_yaw[0], d.pose_yaw[1]);
        }
        ok = true;
        for (int k = 0; k < i; ++k) {
          float dx = x - px[k], dy = y - py[k];
          if (sqrtf(dx * dx + dy * dy) < d.placement_margin) { ok = false; break; }
        }
        px[i] = x; py[i] = y; pz[i] = z; er[i] = ro; ep[i] = pi; ey[i] = ya;
      }
      if (!ok) { all_ok = false; break; }
    }
    if (all_ok) { placed = true; break; }
  }
  for (int i = 0; i < Nmax; ++i) {
    for (int c = 0; c < 13; ++c) bs(w, c, e, i) = 0.0f;
    bs(w, 6, e, i) = 1.0f;
    int32_t asset = 0; float scale = 1.0f, mass = 1.0f, fric = 0.0f;
    if (i < n) {
      if (i == 0 && d.num_target > 0 && d.num_target_assets > 0) asset = w.S.target_assets[rng.below(d.num_target_assets)];
      else asset = w.S.movable_assets[rng.below(d.num_movable_assets)];
      scale = rng.uni(d.scale_range[0], d.scale_range[1]);
      mass = rng.uni(d.mass_range[0], d.mass_range[1]);
      fric = rng.uni(d.friction_range[0], d.friction_range[1]);
      Q4 q = q_from_euler(er[i], ep[i], ey[i]);
      bs(w, 0, e, i) = px[i]; bs(w, 1, e, i) = py[i]; bs(w, 2, e, i) = pz[i];
      bs(w, 3, e, i) = q.x; bs(w, 4, e, i) = q.y; bs(w, 5, e, i) = q.z; bs(w, 6, e, i) = q.w;
    }
    float af; memcpy(&af, &asset, 4);
    mp_(w, 0, e, i) = af; mp_(w, 1, e, i) = scale; mp_(w, 2, e, i) = mass; mp_(w, 3, e, i) = fric;
    w.body_mask[(size_t)e * Nmax + i] = (i < n) ? 1 : 0;
  }
  for (int j = 0; j < 7; ++j) { w.joint_state[(0 * 7 + j) * w.B + e] = P.offstage_positions[j]; w.joint_state[(1 * 7 + j) * w.B + e] = 0.0f; }
  w.num_steps[e] = 0;
  w.phase[e] = B2S_PHASE_IDLE;
  w.num_manifolds[e] = 0;
  for (int k = 0; k < P.max_manifolds; ++k) { w.man_keys[(size_t)e * P.max_manifolds + k] = -1; w.man_npts[(size_t)e * P.max_manifolds + k] = 0; }
  memset(&w.man_pts[(size_t)e * P.max_manifolds * 4 * B2S_CP_FLOATS], 0, sizeof(float) * P.max_manifolds * 4 * B2S_CP_FLOATS);
  w.num_pairs[e] = 0;
  w.error_flags[e] = placed ? 0 : 128;      /* no arrangement with MARGIN clearance in 64 rounds */
  w.contact_flags[e] = 0;
  memset(&w.ctrl[(size_t)e * B2S_CTRL_FLOATS], 0, sizeof(float) * B2S_CTRL_FLOATS);
  memset(&w.ctrl_flags[(size_t)e * 4], 0, sizeof(int32_t) * 4);
  memset(&w.ctrl_time[(size_t)e * 5], 0, sizeof(double) * 5);
  w.ctrl_time[(size_t)e * 5 + 4] = 0.5;       /* grip(0) at reboot: ready 0.5 s later (sawyer_sim.py:392) */
  int32_t* ps = &w.phase_state[(size_t)e * 8];
  ps[1] = 0; ps[2] = 0; ps[3] = 0; ps[4] = 0; ps[5] = 0;   /* ps[0] (max_phase_steps) persists, push_env.py:133 */
  w.is_safe[e] = 1; w.is_effective[e] = 1;
  w.episode_return[e] = 0.0f; w.reward[e] = 0.0f; w.termination[e] = 0;
  build_colliders(w, e);
  /* ArmEnv._reset_robot: move_to_joint_positions(OFFSTAGE_POSITIONS) (arm_env.py:101-107) */
  float q[7], qd[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 7; ++j) q[j] = P.offstage_positions[j];
  arm_reset_targets(w, e);
  arm_set_joint_target(w, e, q);
  arm_fk(w, q, qd, &w.link_poses[(size_t)e * (w.L + 1) * 7], &w.link_vel[(size_t)e * w.L * 6]);
  observe(w, e);
  for (int i = 0; i < Nmax; ++i) { w.prev_xy[((size_t)e * Nmax + i) * 2] = w.obs_position[((size_t)e * Nmax + i) * 3]; w.prev_xy[((size_t)e * Nmax + i) * 2 + 1] = w.obs_position[((size_t)e * Nmax + i) * 3 + 1]; }
}

void observe(World& w, int e) {
  const int n = w.num_movables[e];
  for (int i = 0; i < w.Nmax; ++i) {
    float* o = &w.obs_position[((size_t)e * w.Nmax + i) * 3];
    if (i < n) { o[0] = bs(w, 0, e, i); o[1] = bs(w, 1, e, i); o[2] = bs(w, 2, e, i); }
    else { o[0] = o[1] = o[2] = 0.0f; }
  }
}

void reward_env(World& w, int e, const float* s0, const float* s1);

static void movable_status(World& w, int e, int which) {
  /* PushEnv._get_movable_status: positions + yaw (push_env.py:925-937) */
  for (int i = 0; i < w.Nmax; ++i) {
    float* s = &w.status[(((size_t)e * 2 + which) * w.Nmax + i) * 4];
    if (i < w.num_movables[e]) {
      s[0] = bs(w, 0, e, i); s[1] = bs(w, 1, e, i); s[2] = bs(w, 2, e, i);
      s[3] = yaw_from_q(q4(bs(w, 3, e, i), bs(w, 4, e, i), bs(w, 5, e, i), bs(w, 6, e, i)));
    } else { s[0] = s[1] = s[2] = s[3] = 0.0f; }
  }
}

void set_action(World& w, int e) {
  const B2SParams& P = w.P;
  const int G = P.num_goal_steps > 0 ? P.num_goal_steps : 1;
  float off[3], rng[3];
  for (int k = 0; k < 3; ++k) { off[k] = 0.5f * (P.cspace_high[k] + P.cspace_low[k]); rng[k] = 0.5f * (P.cspace_high[k] - P.cspace_low[k]); }
  const Q4 down = q_from_euler(B2S_PI, 0.0f, 0.0f);
  for (int g = 0; g < G; ++g) {
    /* PushEnv._compute_all_waypoints / _compute_waypoints (push_env.py:735-786) */
    const float* a = &w.action[((size_t)e * G + g) * 4];
    float x = a[0] * rng[0] + off[0], y = a[1] * rng[1] + off[1];
    float z = P.finger_tip_offset + off[2];
    float x2 = fminf(P.cspace_high[0], fmaxf(P.cspace_low[0], x + a[2] * P.translation_x));
    float y2 = fminf(P.cspace_high[1], fmaxf(P.cspace_low[1], y + a[3] * P.translation_y));
    float* wp = &w.waypoints[((size_t)e * G + g) * 14];
    wp[0] = x; wp[1] = y; wp[2] = z; wp[3] = down.x; wp[4] = down.y; wp[5] = down.z; wp[6] = down.w;
    wp[7] = x2; wp[8] = y2; wp[9] = z; wp[10] = down.x; wp[11] = down.y; wp[12] = down.z; wp[13] = down.w;
  }
  w.is_safe[e] = 1; w.is_effective[e] = 1;
  w.phase[e] = B2S_PHASE_INITIAL;
  int32_t* ps = &w.phase_state[(size_t)e * 8];
  ps[1] = 0; ps[2] = 0; ps[3] = 0; ps[4] = 0; ps[5] = 0; ps[6] += 1;
  movable_status(w, e, 0);
}

static void phase_logic(World& w, int e) {
  const B2SParams& P = w.P;
  int32_t* ps = &w.phase_state[(size_t)e * 8];
  int ph = w.phase[e];
  const int nsteps = w.num_steps[e];
  bool interrupt = ps[5] != 0;
  /* _is_phase_ready (push_env.py:812-837) */
  bool ready;
  if (interrupt) ready = true;
  else if (arm_is_ready(w, e) && (P.time_step * (double)nsteps >= w.ctrl_time[(size_t)e * 5 + 4])) { arm_reset_targets(w, e); ready = true; }
  else if (ps[0] < 0) ready = true;
  else if (nsteps >= ps[0]) { arm_reset_targets(w, e); ready = true; }
  else ready = false;
  float ee[7];
  {
    float q[7], lp[(B2S_MAX_LINKS + 1) * 7];
    for (int j = 0; j < 7; ++j) q[j] = w.joint_state[(0 * 7 + j) * w.B + e];
    arm_fk(w, q, q, lp, NULL);
    memcpy(ee, lp + w.L * 7, sizeof(float) * 7);
  }
  if (ready) {
    /* _get_next_phase (push_env.py:788-810) */
    if (interrupt && ph != B2S_PHASE_POST && ph != B2S_PHASE_OFFSTAGE && ph != B2S_PHASE_DONE) ph = B2S_PHASE_POST;
    else if (ph == B2S_PHASE_POST && P.num_goal_steps > 0 && ps[1] < P.num_goal_steps) ph = B2S_PHASE_PRE;
    else ph = ph + 1;
    ps[0] = nsteps + (ph == B2S_PHASE_MOTION ? P.max_motion_steps : ph == B2S_PHASE_OFFSTAGE ? P.max_offstage_steps : P.max_phase_steps);
    const int G = P.num_goal_steps > 0 ? P.num_goal_steps : 1;
    const float* wp = &w.waypoints[((size_t)e * G + (ps[1] < G - 1 ? ps[1] : G - 1)) * 14];    /* waypoints[num_waypoints] */
    float pose[7];
    if (ph == B2S_PHASE_PRE) { memcpy(pose, wp, sizeof(float) * 7); pose[2] = P.gripper_safe_height; arm_reset_targets(w, e); arm_set_link_target(w, e, pose); }
    else if (ph == B2S_PHASE_START) { arm_reset_targets(w, e); arm_set_link_target(w, e, wp); }
    else if (ph == B2S_PHASE_MOTION) { arm_reset_targets(w, e); arm_set_link_target(w, e, wp + 7); }
    else if (ph == B2S_PHASE_POST) { ps[1] += 1; memcpy(pose, ee, sizeof(float) * 7); pose[2] = P.gripper_safe_height; arm_reset_targets(w, e); arm_set_link_target(w, e, pose); }
    else if (ph == B2S_PHASE_OFFSTAGE) {
      float qo[7];
      for (int j = 0; j < 7; ++j) qo[j] = P.offstage_positions[j];
      arm_reset_targets(w, e); arm_set_joint_target(w, e, qo);
    }
  }
  interrupt = false;
  const int cf = w.contact_flags[e];
  if (ph == B2S_PHASE_MOTION && (cf & 1)) interrupt = true;        /* _check_singularity :839-855 */
  bool safe = true;                                                  /* _check_safety :857-898 */
  if (ph == B2S_PHASE_PRE) { if (cf & 2) safe = false; }
  else if (ph == B2S_PHASE_START) {
    if (cf & 2) {
      float start_z = P.finger_tip_offset + 0.5f * (P.cspace_high[2] + P.cspace_low[2]);
      float dist = ee[2] - start_z;
      if (!(fabsf(dist) <= 0.01f)) safe = false;
    }
  } else if (ph == B2S_PHASE_DONE) {
    if (cf & 2) safe = false;
    else
      for (int i = 0; i < w.num_movables[e]; ++i) {
        float x = bs(w, 0, e, i), y = bs(w, 1, e, i);
        if (x < P.table_workspace_low[0] || x > P.table_workspace_high[0] || y < P.table_workspace_low[1] || y > P.table_workspace_high[1]) { safe = false; break; }
      }
  }
  if (!safe) { interrupt = true; w.is_safe[e] = 0; }
  if (interrupt && ph == B2S_PHASE_DONE) ps[4] = 1;
  ps[5] = interrupt ? 1 : 0;
  w.phase[e] = ph;
}

static bool all_stable(World& w, int e, float lin, float ang) {
  for (int i = 0; i < w.num_movables[e]; ++i) {
    float lv = len(v3(bs(w, 7, e, i), bs(w, 8, e, i), bs(w, 9, e, i)));
    float av = len(v3(bs(w, 10, e, i), bs(w, 11, e, i), bs(w, 12, e, i)));
    if (lv >= lin || av >= ang) return false;
  }
  return true;
}

static void finish_action(World& w, int e) {
  const B2SParams& P = w.P;
  movable_status(w, e, 1);
  /* _check_effectiveness (push_env.py:900-923) */
  float dp = 0.0f, da = 0.0f;
  for (int i = 0; i < w.num_movables[e]; ++i) {
    const float* s0 = &w.status[(((size_t)e * 2 + 0) * w.Nmax + i) * 4];
    const float* s1 = &w.status[(((size_t)e * 2 + 1) * w.Nmax + i) * 4];
    dp = dp + len(v3(s1[0] - s0[0], s1[1] - s0[1], s1[2] - s0[2]));
    da = da + fabsf(b2s_wrap_pi(s1[3] - s0[3]));
  }
  w.is_effective[e] = (dp <= P.min_delta_position && da <= P.min_delta_angle) ? 0 : 1;
  w.phase[e] = B2S_PHASE_IDLE;
}

void env_substep(World& w, int e) {
  const B2SParams& P = w.P;
  int ph = w.phase[e];
  if (ph == B2S_PHASE_IDLE) return;
  int32_t* ps = &w.phase_state[(size_t)e * 8];
  if (ph < B2S_PHASE_DONE) {
    substep(w, e);
    if (w.num_steps[e] % P.steps_check != 0) return;
    phase_logic(w, e);
    if (w.phase[e] == B2S_PHASE_DONE) { w.phase[e] = B2S_PHASE_SETTLE; ps[2] = 0; ps[3] = 0; }
    return;
  }
  /* Simulator.wait_until_stable(movables) (simulator.py:325-376): after 'done' (SETTLE), or inside a rollout's reset
   * (RESET_DROP with the loose thresholds of push_env.py:443-447, then RESET_WAIT) */
  const bool drop = (ph == B2S_PHASE_RESET_DROP);
  const float slin = drop ? w.ro.drop_lin : P.stable_lin_threshold, sang = drop ? w.ro.drop_ang : P.stable_ang_threshold;
  const int smax = drop ? w.ro.drop_max_steps : P.stable_max_steps;
  substep(w, e);
  ps[2] += 1;
  if (ps[2] < P.stable_check_after) return;
  if (all_stable(w, e, slin, sang)) ps[3] += 1;
  if (!(ps[3] >= P.stable_min_steps || ps[2] >= smax)) return;
  if (ph == B2S_PHASE_SETTLE) finish_action(w, e);
  else if (drop) { w.phase[e] = B2S_PHASE_RESET_WAIT; ps[2] = 0; ps[3] = 0; }
  else w.phase[e] = B2S_PHASE_IDLE;          /* RESET_WAIT over: rollout_substep decides what follows */
}

/* ------------------------------------------------------------- reward ---- */
static bool on_tiles(float x, float y, const abi_float (*tiles)[2], int nt, float size, const abi_float* off, float max_dist) {
  /* check_on_tiles (push_reward.py:57-67) */
  bool any = false;
  for (int t = 0; t < nt; ++t) {
    float tx = off[0] + tiles[t][0] * size, ty = off[1] + tiles[t][1] * size;
    if (fabsf(x - tx) <= 0.5f * max_dist && fabsf(y - ty) <= 0.5f * max_dist) any = true;
  }
  return any;
}
static float tile_dist(float x, float y, const abi_float (*tiles)[2], int nt, float size, const abi_float* off) {
  /* get_tile_dists (push_reward.py:70-75) */
  float best = 3e38f;
  for (int t = 0; t < nt; ++t) {
    float dx = x - (off[0] + tiles[t][0] * size), dy = y - (off[1] + tiles[t][1] * size);
    float dd = sqrtf(dx * dx + dy * dy);
    if (dd < best) best = dd;
  }
  return best;
}
static float clearing_score(const float* xy, int n) {
  /* push_reward.py:101-107 (the second minimum overwrites the first) */
  float d1 = 0, d3 = 0;
  for (int i = 0; i < n; ++i) { d1 = d1 + fabsf(xy[i * 2] - 0.7f); d3 = d3 + fabsf(xy[i * 2 + 1] + 0.9f); }
  d1 = d1 / (float)n; d3 = d3 / (float)n;
  return -fminf(d1, d3);
}

void reward(World& w, int e, const float* prev_xy, const float* next_xy) {
  const int N = w.Nmax;
  const float* s0 = prev_xy + (size_t)e * N * 2;
  const float* s1 = next_xy + (size_t)e * N * 2;
  reward_env(w, e, s0, s1);
}

/* s0 / s1: xy of the Nmax bodies of env e before / after the action */
void reward_env(World& w, int e, const float* s0, const float* s1) {
  const B2SSceneDesc& d = w.S.d;
  const int N = w.Nmax;
  float r = 0.0f;
  bool term = false;
  const int task = w.P.task;
  if (task == B2S_TASK_NONE) { w.reward[e] = 1.0f; w.termination[e] = 0; w.episode_return[e] += 1.0f; return; }
  bool goal = false;
  float sc0 = 0.0f, sc1 = 0.0f;
  if (task == B2S_TASK_CROSSING) {
    term = !on_tiles(s1[0], s1[1], d.region, d.num_region, d.tile_size, d.tile_offset, d.tile_size * 1.5f);
    goal = on_tiles(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset, d.tile_size);
    sc0 = -tile_dist(s0[0], s0[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
    sc1 = -tile_dist(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
  } else if (task == B2S_TASK_INSERTION) {
    term = false;
    goal = on_tiles(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset, d.tile_size);
    sc0 = -tile_dist(s0[0], s0[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
    sc1 = -tile_dist(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
  } else {  /* clearing */
    term = false;
    goal = true;
    for (int i = 0; i < N; ++i)
      if (on_tiles(s1[i * 2], s1[i * 2 + 1], d.region, d.num_region, d.tile_size * 1.25f, d.tile_offset, d.tile_size * 1.25f)) goal = false;
    sc0 = clearing_score(s0, N);
    sc1 = clearing_score(s1, N);
  }
  bool goal_reached = goal && !term;
  bool penalty = term && !goal_reached;
  r = r + 100.0f * (goal_reached ? 1.0f : 0.0f);
  r = r + (-100.0f) * (penalty ? 1.0f : 0.0f);
  r = r + fabsf(sc1 - sc0) * 1.0f;
  r = r + (-1.0f);
  w.reward[e] = r;
  w.termination[e] = (term || goal_reached) ? 1 : 0;
  w.episode_return[e] += r;
}

/* ----------------------------------------------- episodes without the host ---- */
/* HeuristicPushSampler._sample (heuristic_push_sampler.py:66-123) with Philox draws: attempt k uses the counter
 * (stream 1 | action index << 8, global env id, k | num_episodes << 16), blocks 0 and 1. */
void policy_sample(World& w, int e, uint64_t seed, int action_index, int num_episodes, int max_attempts, float out[4]) {
  const B2SParams& P = w.P;
  int nb = w.num_movables[e]; if (nb < 1) nb = 1;
  const int target = num_episodes % nb;
  const float a = (float)(num_episodes * 42);                       /* SEED, heuristic_push_sampler.py:13 */
  const float base_angle = a - floorf(a / (2.0f * B2S_PI)) * (2.0f * B2S_PI);
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  for (int att = 0; att < max_attempts; ++att) {
    const uint32_t c1 = 1u | ((uint32_t)action_index << 8), c2 = (uint32_t)(P.env_id_offset + e);
    const uint32_t c3 = (uint32_t)att | ((uint32_t)num_episodes << 16);
    const b2s_u4 r0 = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), 0u, c1, c2, c3);
    const b2s_u4 r1 = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), 1u, c1, c2, c3);
    const float sx = -1.0f + 2.0f * b2s_u01(r0.x), sy = -1.0f + 2.0f * b2s_u01(r0.y);
    const float angle = base_angle + (-0.25f * B2S_PI + (0.5f * B2S_PI) * b2s_u01(r0.z));
    float sn, cs;
    b2s_sincos(angle, &sn, &cs);
    const float mx = fminf(1.0f, fmaxf(-1.0f, cs + (-0.3f + 0.6f * b2s_u01(r0.w))));
    const float my = fminf(1.0f, fmaxf(-1.0f, sn + (-0.3f + 0.6f * b2s_u01(r1.x))));
    const float offx = 0.5f * (P.cspace_high[0] + P.cspace_low[0]), offy = 0.5f * (P.cspace_high[1] + P.cspace_low[1]);
    const float rngx = 0.5f * (P.cspace_high[0] - P.cspace_low[0]), rngy = 0.5f * (P.cspace_high[1] - P.cspace_low[1]);
    const float x0 = sx * rngx + offx, y0 = sy * rngy + offy;
    const float x1 = fminf(P.cspace_high[0], fmaxf(P.cspace_low[0], x0 + mx * P.translation_x));
    const float y1 = fminf(P.cspace_high[1], fmaxf(P.cspace_low[1], y0 + my * P.translation_y));
    bool clear = true;                                               /* start_margin 0.05 */
    for (int i = 0; i < nb; ++i) {
      const float dx = bs(w, 0, e, i) - x0, dy = bs(w, 1, e, i) - y0;
      if (!(sqrtf(dx * dx + dy * dy) > 0.05f)) clear = false;
    }
    const float tx = bs(w, 0, e, target), ty = bs(w, 1, e, target);
    const float d0 = sqrtf((tx - x0) * (tx - x0) + (ty - y0) * (ty - y0));
    const float d1 = sqrtf((tx - x1) * (tx - x1) + (ty - y1) * (ty - y1));
    const bool touches = !(d0 >= 0.01f && d1 >= 0.01f);             /* motion_margin 0.01 */
    out[0] = sx; out[1] = sy; out[2] = mx; out[3] = my;
    if (clear && touches) return;
  }
}

/* B2S_POLICY_AIMED: a random body, a random direction, start 8 cm behind it, push through (synthetic workloads) */
void policy_aimed(World& w, int e, uint64_t seed, int action_index, int num_episodes, float out[4]) {
  const B2SParams& P = w.P;
  int nb = w.num_movables[e]; if (nb < 1) nb = 1;
  const b2s_u4 r = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), 0u, 2u | ((uint32_t)action_index << 8), (uint32_t)(P.env_id_offset + e), (uint32_t)num_episodes);
  int body = (int)(b2s_u01(r.x) * (float)nb); if (body > nb - 1) body = nb - 1;
  const float angle = -B2S_PI + (2.0f * B2S_PI) * b2s_u01(r.y);
  float sn, cs;
  b2s_sincos(angle, &sn, &cs);
  const float offx = 0.5f * (P.cspace_high[0] + P.cspace_low[0]), offy = 0.5f * (P.cspace_high[1] + P.cspace_low[1]);
  const float rngx = 0.5f * (P.cspace_high[0] - P.cspace_low[0]), rngy = 0.5f * (P.cspace_high[1] - P.cspace_low[1]);
  const float tx = bs(w, 0, e, body) - 0.08f * cs, ty = bs(w, 1, e, body) - 0.08f * sn;
  out[0] = fminf(1.0f, fmaxf(-1.0f, (tx - offx) / rngx));
  out[1] = fminf(1.0f, fmaxf(-1.0f, (ty - offy) / rngy));
  out[2] = cs; out[3] = sn;
}

static void rollout_policy(World& w, int e, int action_index, float out[4]) {
  if (w.ro.policy_kind == B2S_POLICY_AIMED) policy_aimed(w, e, w.ro.seed, action_index, w.num_episodes[e], out);
  else policy_sample(w, e, w.ro.seed, action_index, w.num_episodes[e], w.ro.max_attempts, out);
}

static void episode_start(World& w, int e, const float* first_action) {
  const int N = w.Nmax, A = w.ro.num_actions, EP = w.ro.num_episodes;
  const int nm = w.num_movables[e];
  const int ep = w.ro_state[(size_t)e * 4 + 1];
  float act[4];
  if (first_action) for (int k = 0; k < 4; ++k) act[k] = first_action[(size_t)e * 4 + k];
  else rollout_policy(w, e, 0, act);
  w.ro_state[(size_t)e * 4 + 0] = 0;
  w.episode_return[e] = 0.0f; w.reward[e] = 0.0f; w.termination[e] = 0;
  observe(w, e);
  for (int i = 0; i < N; ++i) {
    w.prev_xy[((size_t)e * N + i) * 2] = (i < nm) ? bs(w, 0, e, i) : 0.0f;
    w.prev_xy[((size_t)e * N + i) * 2 + 1] = (i < nm) ? bs(w, 1, e, i) : 0.0f;
  }
  if (ep < EP) {
    if (w.ro.lengths) w.ro.lengths[(size_t)e * EP + ep] = 0;
    if (w.ro.positions)
      for (int i = 0; i < N; ++i)
        for (int k = 0; k < 3; ++k)
          w.ro.positions[((((size_t)e * EP + ep) * (A + 1)) * N + i) * 3 + k] = (i < nm) ? bs(w, k, e, i) : 0.0f;
  }
  for (int k = 0; k < 4; ++k) w.action[(size_t)e * 4 + k] = act[k];
  set_action(w, e);
}

void rollout_begin(World& w, int e, const float* first_action) {
  w.ro_state[(size_t)e * 4 + 1] = 0; w.ro_state[(size_t)e * 4 + 2] = 0;
  episode_start(w, e, first_action);
}

static void rollout_reset(World& w, int e) {
  reset_env(w, e, w.ro.reset_seed);
  w.phase[e] = B2S_PHASE_RESET_DROP;
  w.phase_state[(size_t)e * 8 + 2] = 0; w.phase_state[(size_t)e * 8 + 3] = 0;
}

/* RobotEnv.step's bookkeeping after an action (robot_env.py:237-273) + generate_episode(s) (episode_generation.py:41-61, 88-112) */
static void rollout_advance(World& w, int e) {
  const int N = w.Nmax, nm = w.num_movables[e];
  const int A = w.ro.num_actions, EP = w.ro.num_episodes;
  const int t = w.ro_state[(size_t)e * 4 + 0], ep = w.ro_state[(size_t)e * 4 + 1];
  float* prev = &w.prev_xy[(size_t)e * N * 2];
  float s0[128] = {0}, s1[128] = {0};
  for (int i = 0; i < N; ++i) {
    s0[i * 2] = prev[i * 2]; s0[i * 2 + 1] = prev[i * 2 + 1];
    s1[i * 2] = (i < nm) ? bs(w, 0, e, i) : 0.0f; s1[i * 2 + 1] = (i < nm) ? bs(w, 1, e, i) : 0.0f;
  }
  reward_env(w, e, s0, s1);
  const float r = w.reward[e];
  const bool term = w.termination[e] != 0;
  for (int i = 0; i < N * 2; ++i) prev[i] = s1[i];
  observe(w, e);
  const bool env_done = w.phase_state[(size_t)e * 8 + 4] != 0;
  if (t < A && ep < EP) {
    const size_t rec = ((size_t)e * EP + ep) * A + t;
    if (w.ro.actions) for (int k = 0; k < 4; ++k) w.ro.actions[rec * 4 + k] = w.action[(size_t)e * 4 + k];
    if (w.ro.rewards) w.ro.rewards[rec] = r;
    if (w.ro.flags) w.ro.flags[rec] = (uint8_t)((w.is_safe[e] ? 1 : 0) | (w.is_effective[e] ? 2 : 0) | (term ? 4 : 0) | (env_done ? 8 : 0));
    if (w.ro.substeps) w.ro.substeps[rec] = w.num_steps[e];
    if (w.ro.positions)
      for (int i = 0; i < N; ++i)
        for (int k = 0; k < 3; ++k)
          w.ro.positions[((((size_t)e * EP + ep) * (A + 1) + t + 1) * N + i) * 3 + k] = (i < nm) ? bs(w, k, e, i) : 0.0f;
  }
  const bool over = term || env_done || t + 1 >= A;
  w.ro_state[(size_t)e * 4 + 0] = t + 1;
  if (over) {
    if (ep < EP) {
      if (w.ro.lengths) w.ro.lengths[(size_t)e * EP + ep] = t + 1;
      if (w.ro.returns) w.ro.returns[(size_t)e * EP + ep] = w.episode_return[e];
    }
    w.num_episodes[e] += 1;
    w.ro_state[(size_t)e * 4 + 1] = ep + 1;
    if (ep + 1 < EP) { w.ro_state[(size_t)e * 4 + 2] = 0; rollout_reset(w, e); }
    return;
  }
  float act[4];
  rollout_policy(w, e, t + 1, act);
  for (int k = 0; k < 4; ++k) w.action[(size_t)e * 4 + k] = act[k];
  set_action(w, e);
}

/* 0: scene valid (env IDLE, the caller starts the episode), 1: re-sampled, 2: gave up */
static int rollout_reset_check(World& w, int e) {
  float table_z = 0.0f;
  for (int s = 0; s < w.Ns; ++s) if (w.S.static_flags[s] & B2S_STATIC_IS_TABLE) table_z = w.S.static_pose[s * 7 + 2];
  const float zmin = table_z + w.table_dz[e];
  bool bad = false;
  for (int i = 0; i < w.num_movables[e]; ++i) if (bs(w, 2, e, i) < zmin) bad = true;
  if (w.error_flags[e] & 128) bad = true;
  if (bad) {
    if (w.ro_state[(size_t)e * 4 + 2] < w.ro.max_reset_retries) { w.ro_state[(size_t)e * 4 + 2] += 1; rollout_reset(w, e); return 1; }
    w.error_flags[e] |= 256; w.phase[e] = B2S_PHASE_IDLE;
    return 2;
  }
  w.phase[e] = B2S_PHASE_IDLE;
  return 0;
}

void rollout_substep(World& w, int e) {
  const int ph = w.phase[e];
  if (ph == B2S_PHASE_IDLE) return;
  env_substep(w, e);
  if (w.phase[e] != B2S_PHASE_IDLE) return;
  if (ph == B2S_PHASE_SETTLE) rollout_advance(w, e);
  else if (ph == B2S_PHASE_RESET_WAIT) { if (rollout_reset_check(w, e) == 0) episode_start(w, e, NULL); }
}

/* ---- asynchronous stepping with the policy on the host (b2s_env_async_step) ---- */
void async_command(World& w, int e, int cmd) {
  w.async_events[e] = 0;
  if (w.phase[e] != B2S_PHASE_IDLE) return;
  if (cmd == 1) set_action(w, e);
  else if (cmd == 2) { w.ro_state[(size_t)e * 4 + 2] = 0; rollout_reset(w, e); }
}

void async_substep(World& w, int e) {
  const int ph = w.phase[e];
  if (ph == B2S_PHASE_IDLE) return;
  env_substep(w, e);
  if (w.phase[e] != B2S_PHASE_IDLE) return;
  const int N = w.Nmax, nm = w.num_movables[e];
  if (ph == B2S_PHASE_SETTLE) {
    float* prev = &w.prev_xy[(size_t)e * N * 2];
    float s0[128] = {0}, s1[128] = {0};
    for (int i = 0; i < N; ++i) {
      s0[i * 2] = prev[i * 2]; s0[i * 2 + 1] = prev[i * 2 + 1];
      s1[i * 2] = (i < nm) ? bs(w, 0, e, i) : 0.0f; s1[i * 2 + 1] = (i < nm) ? bs(w, 1, e, i) : 0.0f;
    }
    reward_env(w, e, s0, s1);
    for (int i = 0; i < N * 2; ++i) prev[i] = s1[i];
    observe(w, e);
    w.async_events[e] |= 2;
  } else if (ph == B2S_PHASE_RESET_WAIT) {
    if (rollout_reset_check(w, e) != 0) return;
    for (int i = 0; i < N; ++i) {
      w.prev_xy[((size_t)e * N + i) * 2] = (i < nm) ? bs(w, 0, e, i) : 0.0f;
      w.prev_xy[((size_t)e * N + i) * 2 + 1] = (i < nm) ? bs(w, 1, e, i) : 0.0f;
    }
    w.episode_return[e] = 0.0f; w.reward[e] = 0.0f; w.termination[e] = 0;
    observe(w, e);
    w.async_events[e] |= 4;
  }
}

}  // namespace b2o
