// b2s_rollout.cuh -- what happens BETWEEN two actions of an environment, as device functions: the PushEnv reward, the
// start of an action (waypoints, phase machine reset) and the heuristic push policy.  They are shared by the
// one-thread-per-env kernels of b2s_aux.cu (PushEnv.step, the lock-step path) and by the substep kernel, which uses
// them to run whole episodes without returning to the host: an environment that finishes an action computes its
// reward, records the transition, draws its next action and starts it in the same launch (b2s_rollout_*), so a
// batch never waits for its slowest environment between actions.
//
//   reward_eval        push_reward.get_reward_fn(...).reward_fn, is_planning=False   robovat/reward_fns/push_reward.py:272-374
//   begin_action       PushEnv._execute_action prologue + _compute_waypoints         robovat/envs/push/push_env.py:637-651, 752-786
//   policy_sample      HeuristicPushSampler._sample                                  robovat/envs/push/heuristic_push_sampler.py:66-123
//   reset_env_dev      PushEnv._reset_scene (sampling part)                          robovat/envs/push/push_env.py:331-471
//   rollout_advance    RobotEnv.step's bookkeeping + generate_episode(s)' loops     robovat/envs/robot_env.py:237-273,
//                                                                                    robovat/io/episode_generation.py:41-61, 88-112
// Include this header BEFORE any `#define W ...` shorthand: every function takes the world description as `W`.
#pragma once

#include "b2s_dev.cuh"

#define RO_BS(W, c, e, i) (W).buf.body_state[((size_t)(c) * (W).B + (e)) * (W).Nmax + (i)]

#define RO_MP(W, c, e, i) (W).mov_params[((size_t)(c) * (W).B + (e)) * (W).Nmax + (i)]

struct XfS { V3 p; Q4 q; };
__device__ __forceinline__ XfS xfs_from(const float* a) { XfS t; t.p = v3(a[0], a[1], a[2]); t.q = q4(a[3], a[4], a[5], a[6]); return t; }
__device__ __forceinline__ XfS xfs_mul(XfS a, XfS b) { XfS t; t.p = a.p + qrot(a.q, b.p); t.q = qmul(a.q, b.q); return t; }
__device__ __forceinline__ void xfs_store(XfS t, float* o) { o[0] = t.p.x; o[1] = t.p.y; o[2] = t.p.z; o[3] = t.q.x; o[4] = t.q.y; o[5] = t.q.z; o[6] = t.q.w; }

// scalar FK of every collision link + end effector (same arithmetic as the warp version in b2s_step.cu)
__device__ inline void fk_links_scalar(const DWorld& W, const float* q, const float* qd, float* lp, float* lv) {
  const DArm* arm = W.arm;
  XfS frame[B2S_NUM_JOINTS];
  V3 ax[B2S_NUM_JOINTS], org[B2S_NUM_JOINTS];
  XfS T = xfs_from(arm->base);
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
    XfS Tj = xfs_mul(T, xfs_from(arm->joint_origin[j]));
    V3 a = v3(arm->joint_axis[j][0], arm->joint_axis[j][1], arm->joint_axis[j][2]);
    ax[j] = qrot(Tj.q, a);
    org[j] = Tj.p;
    T.p = Tj.p;
    T.q = qmul(Tj.q, q_axis_angle(a, q[j]));
    frame[j] = T;
  }
  XfS base = xfs_from(arm->base);
  for (int k = 0; k < W.L; ++k) {
    int jj = arm->link_joint[k];
    XfS Tk = xfs_mul(jj < 0 ? base : frame[jj], xfs_from(arm->link_pose[k]));
    xfs_store(Tk, lp + k * 7);
    if (lv) {
      V3 v = v3(0, 0, 0), om = v3(0, 0, 0);
      for (int i = 0; i <= jj; ++i) { v = v + cross(ax[i], Tk.p - org[i]) * qd[i]; om = om + ax[i] * qd[i]; }
      float* o = lv + k * 6;
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = om.x; o[4] = om.y; o[5] = om.z;
    }
  }
  xfs_store(xfs_mul(frame[B2S_NUM_JOINTS - 1], xfs_from(arm->ee)), lp + W.L * 7);
}

struct DevRng {
  uint32_t k0, k1, c1, c2, c3, blk;
  b2s_u4 buf; int have;
  __device__ DevRng(uint64_t seed, uint32_t stream, uint32_t env, uint32_t attempt)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c1(stream), c2(env), c3(attempt), blk(0), have(0) {}
  __device__ uint32_t next() {
    if (!have) { buf = b2s_philox(k0, k1, blk++, c1, c2, c3); have = 4; }
    uint32_t r = (have == 4) ? buf.x : (have == 3) ? buf.y : (have == 2) ? buf.z : buf.w;
    --have;
    return r;
  }
  __device__ float uni(float lo, float hi) { return lo + (hi - lo) * b2s_u01(next()); }
  __device__ int below(int n) { int k = (int)(b2s_u01(next()) * (float)n); return k < n ? k : n - 1; }
};

__device__ inline void observe_env(const DWorld& W, int e) {
  const int n = W.buf.num_movables[e];
  for (int i = 0; i < W.Nmax; ++i) {
    float* o = W.buf.obs_position + ((size_t)e * W.Nmax + i) * 3;
    if (i < n) { o[0] = RO_BS(W, 0, e, i); o[1] = RO_BS(W, 1, e, i); o[2] = RO_BS(W, 2, e, i); }
    else { o[0] = o[1] = o[2] = 0.0f; }
  }
}

// RobotEnv.reset scene part for one environment (oracle/b2o_env.cpp reset_env; push_env.py:331-471), one thread
__device__ inline void reset_env_dev(const DWorld& W, int e, uint64_t seed) {
  const DLayout& d = *W.layout;
  const B2SParams& P = W.P;
  const int Nmax = W.Nmax;
  DevRng rng(seed, 0u, (uint32_t)(P.env_id_offset + e), (uint32_t)W.reset_count[e]);
  W.reset_count[e] += 1;
  float dz = rng.uni(d.table_height_range[0], d.table_height_range[1]);
  W.table_dz[e] = dz;
  int span = Nmax - d.min_movables + 1;
  int n = d.min_movables + rng.below(span > 0 ? span : 1);
  if (n > Nmax) n = Nmax;
  W.buf.num_movables[e] = n;
  float table_z = 0.0f;
  for (int s = 0; s < W.Ns; ++s) if (W.static_flags[s] & B2S_STATIC_IS_TABLE) table_z = W.static_pose[s * 7 + 2] + dz;
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
          const float(*tiles)[2] = use_target ? d.target : d.obstacle;
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
    for (int c = 0; c < 13; ++c) RO_BS(W, c, e, i) = 0.0f;
    RO_BS(W, 6, e, i) = 1.0f;
    int32_t asset = 0; float scale = 1.0f, mass = 1.0f, fric = 0.0f;
    if (i < n) {
      if (i == 0 && d.num_target > 0 && d.num_target_assets > 0) asset = W.target_assets[rng.below(d.num_target_assets)];
      else asset = W.movable_assets[rng.below(d.num_movable_assets)];
      scale = rng.uni(d.scale_range[0], d.scale_range[1]);
      mass = rng.uni(d.mass_range[0], d.mass_range[1]);
      fric = rng.uni(d.friction_range[0], d.friction_range[1]);
      Q4 q = q_from_euler(er[i], ep[i], ey[i]);
      RO_BS(W, 0, e, i) = px[i]; RO_BS(W, 1, e, i) = py[i]; RO_BS(W, 2, e, i) = pz[i];
      RO_BS(W, 3, e, i) = q.x; RO_BS(W, 4, e, i) = q.y; RO_BS(W, 5, e, i) = q.z; RO_BS(W, 6, e, i) = q.w;
    }
    RO_MP(W, 0, e, i) = __int_as_float(asset); RO_MP(W, 1, e, i) = scale; RO_MP(W, 2, e, i) = mass; RO_MP(W, 3, e, i) = fric;
    W.buf.body_mask[(size_t)e * Nmax + i] = (i < n) ? 1 : 0;
  }
  for (int j = 0; j < 7; ++j) { W.buf.joint_state[(0 * 7 + j) * W.B + e] = P.offstage_positions[j]; W.buf.joint_state[(1 * 7 + j) * W.B + e] = 0.0f; }
  W.num_steps[e] = 0;
  W.phase[e] = B2S_PHASE_IDLE;
  W.num_manifolds[e] = 0;
  W.man_parity[e] = 0;
  const int M = P.max_manifolds;
  for (int par = 0; par < 2; ++par)
    for (int k = 0; k < M; ++k) { W.man_keys[((size_t)par * W.B + e) * M + k] = -1; W.man_npts[((size_t)par * W.B + e) * M + k] = 0; }
  W.num_pairs[e] = 0;
  W.error_flags[e] = placed ? 0 : 128;     // no arrangement with MARGIN clearance in 64 rounds: the host re-samples
  W.contact_flags[e] = 0;
  for (int k = 0; k < B2S_CTRL_FLOATS; ++k) W.ctrl[(size_t)e * B2S_CTRL_FLOATS + k] = 0.0f;
  for (int k = 0; k < 4; ++k) W.ctrl_flags[(size_t)e * 4 + k] = 0;
  for (int k = 0; k < 5; ++k) W.ctrl_time[(size_t)e * 5 + k] = 0.0;
  W.ctrl_time[(size_t)e * 5 + 4] = 0.5;
  int32_t* ps = W.phase_state + (size_t)e * 8;
  ps[1] = 0; ps[2] = 0; ps[3] = 0; ps[4] = 0; ps[5] = 0;
  W.buf.is_safe[e] = 1; W.buf.is_effective[e] = 1;
  W.buf.episode_return[e] = 0.0f; W.buf.reward[e] = 0.0f; W.buf.termination[e] = 0;
  // build_colliders
  {
    int nc = 0; bool over = false;
    int32_t* cs = W.col_slot + (size_t)e * W.Hmax;
    int32_t* ch = W.col_hull + (size_t)e * W.Hmax;
    for (int s = 0; s < W.Ns; ++s) {
      if (W.static_flags[s] & B2S_STATIC_NO_COLLIDE) continue;
      const DAsset& A = W.assets[W.static_asset[s]];
      for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) { if (nc >= W.Hmax) { over = true; break; } cs[nc] = s; ch[nc] = h; ++nc; }
    }
    for (int k = 0; k < W.L; ++k) {
      const DAsset& A = W.assets[W.arm->link_asset[k]];
      for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) { if (nc >= W.Hmax) { over = true; break; } cs[nc] = W.Ns + k; ch[nc] = h; ++nc; }
    }
    for (int i = 0; i < n; ++i) {
      const DAsset& A = W.assets[__float_as_int(RO_MP(W, 0, e, i))];
      for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) { if (nc >= W.Hmax) { over = true; break; } cs[nc] = W.Ns + W.L + i; ch[nc] = h; ++nc; }
    }
    W.ncol[e] = nc;
    if (over) W.error_flags[e] |= 32;
  }
  // move_to_joint_positions(OFFSTAGE_POSITIONS)
  {
    float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
    for (int k = 0; k < 7; ++k) c[9 + k] = P.offstage_positions[k];
    c[16] = P.joint_pos_threshold; c[17] = P.joint_vel_threshold;
    W.ctrl_time[(size_t)e * 5 + 2] = 0.0; W.ctrl_time[(size_t)e * 5 + 3] = 0.0 + (double)P.limb_timeout;
    W.ctrl_flags[(size_t)e * 4 + 0] = 0; W.ctrl_flags[(size_t)e * 4 + 1] = 1; W.ctrl_flags[(size_t)e * 4 + 2] = 0;
  }
  float q[7], qd[7];
  for (int j = 0; j < 7; ++j) { q[j] = P.offstage_positions[j]; qd[j] = 0.0f; }
  fk_links_scalar(W, q, qd, W.link_poses + (size_t)e * (W.L + 1) * 7, W.link_vel + (size_t)e * W.L * 6);
  observe_env(W, e);
  for (int i = 0; i < Nmax; ++i) {
    W.prev_xy[((size_t)e * Nmax + i) * 2] = W.buf.obs_position[((size_t)e * Nmax + i) * 3];
    W.prev_xy[((size_t)e * Nmax + i) * 2 + 1] = W.buf.obs_position[((size_t)e * Nmax + i) * 3 + 1];
  }
}

// ---- reward ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ro_on_tiles(float x, float y, const float (*tiles)[2], int nt, float size, const float* off, float max_dist) {
  bool any = false;
  for (int t = 0; t < nt; ++t) {
    float tx = off[0] + tiles[t][0] * size, ty = off[1] + tiles[t][1] * size;
    if (fabsf(x - tx) <= 0.5f * max_dist && fabsf(y - ty) <= 0.5f * max_dist) any = true;
  }
  return any;
}
__device__ __forceinline__ float ro_tile_dist(float x, float y, const float (*tiles)[2], int nt, float size, const float* off) {
  float best = 3e38f;
  for (int t = 0; t < nt; ++t) {
    float dx = x - (off[0] + tiles[t][0] * size), dy = y - (off[1] + tiles[t][1] * size);
    float dd = sqrtf(dx * dx + dy * dy);
    if (dd < best) best = dd;
  }
  return best;
}
__device__ __forceinline__ float ro_clearing_score(const float* xy, int n) {
  float d1 = 0, d3 = 0;
  for (int i = 0; i < n; ++i) { d1 = d1 + fabsf(xy[i * 2] - 0.7f); d3 = d3 + fabsf(xy[i * 2 + 1] + 0.9f); }
  d1 = d1 / (float)n; d3 = d3 / (float)n;
  return -fminf(d1, d3);
}

// s0 / s1: xy of the Nmax bodies before / after the action (zero padded).  Returns the reward; *done_out = termination.
__device__ inline float reward_eval(const DWorld& W, const float* s0, const float* s1, bool* done_out) {
  const DLayout& d = *W.layout;
  const int N = W.Nmax;
  const int task = W.P.task;
  if (task == B2S_TASK_NONE) { *done_out = false; return 1.0f; }
  bool term = false, goal = false;
  float sc0 = 0.0f, sc1 = 0.0f;
  if (task == B2S_TASK_CROSSING) {
    term = !ro_on_tiles(s1[0], s1[1], d.region, d.num_region, d.tile_size, d.tile_offset, d.tile_size * 1.5f);
    goal = ro_on_tiles(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset, d.tile_size);
    sc0 = -ro_tile_dist(s0[0], s0[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
    sc1 = -ro_tile_dist(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
  } else if (task == B2S_TASK_INSERTION) {
    goal = ro_on_tiles(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset, d.tile_size);
    sc0 = -ro_tile_dist(s0[0], s0[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
    sc1 = -ro_tile_dist(s1[0], s1[1], d.goal, d.num_goal, d.tile_size, d.tile_offset);
  } else {
    goal = true;
    for (int i = 0; i < N; ++i)
      if (ro_on_tiles(s1[i * 2], s1[i * 2 + 1], d.region, d.num_region, d.tile_size * 1.25f, d.tile_offset, d.tile_size * 1.25f)) goal = false;
    sc0 = ro_clearing_score(s0, N);
    sc1 = ro_clearing_score(s1, N);
  }
  const bool goal_reached = goal && !term;
  const bool penalty = term && !goal_reached;
  float r = 0.0f;
  r = r + 100.0f * (goal_reached ? 1.0f : 0.0f);
  r = r + (-100.0f) * (penalty ? 1.0f : 0.0f);
  r = r + fabsf(sc1 - sc0) * 1.0f;
  r = r + (-1.0f);
  *done_out = term || goal_reached;
  return r;
}

// ---- start of an action (one thread) --------------------------------------------------------------------------
__device__ inline void begin_action(const DWorld& W, int e) {
  const B2SParams& P = W.P;
  float off[3], rng[3];
  for (int k = 0; k < 3; ++k) { off[k] = 0.5f * (P.cspace_high[k] + P.cspace_low[k]); rng[k] = 0.5f * (P.cspace_high[k] - P.cspace_low[k]); }
  const Q4 down = q_from_euler(B2S_PI, 0.0f, 0.0f);
  for (int g = 0; g < W.G; ++g) {                       // _compute_all_waypoints (push_env.py:735-750)
    const float* a = W.buf.action + ((size_t)e * W.G + g) * 4;
    float x = a[0] * rng[0] + off[0], y = a[1] * rng[1] + off[1];
    float z = P.finger_tip_offset + off[2];
    float x2 = fminf(P.cspace_high[0], fmaxf(P.cspace_low[0], x + a[2] * P.translation_x));
    float y2 = fminf(P.cspace_high[1], fmaxf(P.cspace_low[1], y + a[3] * P.translation_y));
    float* wp = W.waypoints + ((size_t)e * W.G + g) * 14;
    wp[0] = x; wp[1] = y; wp[2] = z; wp[3] = down.x; wp[4] = down.y; wp[5] = down.z; wp[6] = down.w;
    wp[7] = x2; wp[8] = y2; wp[9] = z; wp[10] = down.x; wp[11] = down.y; wp[12] = down.z; wp[13] = down.w;
  }
  W.buf.is_safe[e] = 1; W.buf.is_effective[e] = 1;
  W.phase[e] = B2S_PHASE_INITIAL;
  int32_t* ps = W.phase_state + (size_t)e * 8;
  ps[1] = 0; ps[2] = 0; ps[3] = 0; ps[4] = 0; ps[5] = 0; ps[6] += 1;
  for (int i = 0; i < W.Nmax; ++i) {
    float* s = W.status + (((size_t)e * 2 + 0) * W.Nmax + i) * 4;
    if (i < W.buf.num_movables[e]) {
      s[0] = RO_BS(W, 0, e, i); s[1] = RO_BS(W, 1, e, i); s[2] = RO_BS(W, 2, e, i);
      s[3] = yaw_from_q(q4(RO_BS(W, 3, e, i), RO_BS(W, 4, e, i), RO_BS(W, 5, e, i), RO_BS(W, 6, e, i)));
    } else { s[0] = s[1] = s[2] = s[3] = 0.0f; }
  }
}

// ---- heuristic push policy (one warp: the attempts of the rejection sampler run 32 at a time) -----------------
// The reference draws start ~ U[-1,1]^2, a direction (num_episodes * 42) mod 2 pi +- pi/4 and a jitter of +-0.3 per
// axis until the start is at least 5 cm from every body and the start or the (clipped) end point lies within 1 cm of
// the target body `num_episodes mod num_bodies`; after MAX_ATTEMPS it returns the last candidate.  Attempt k draws
// from its own Philox counter (seed; stream 1, action index, global env id, episode, k), so trying 32 attempts at once
// and taking the first accepted one in index order gives exactly what the sequential loop of the oracle gives.
#define RO_START_MARGIN 0.05f      // heuristic_push_sampler.py:24
#define RO_MOTION_MARGIN 0.01f     // heuristic_push_sampler.py:25
#define RO_ANGLE_SEED 42           // heuristic_push_sampler.py:13
__device__ __forceinline__ void policy_candidate(const DWorld& W, int e, uint64_t seed, int action_index, int num_episodes,
                                                 int attempt, float base_angle, int nb, int target, float out[4], bool* accept) {
  const B2SParams& P = W.P;
  const uint32_t c1 = 1u | ((uint32_t)action_index << 8), c2 = (uint32_t)(P.env_id_offset + e);
  const uint32_t c3 = (uint32_t)attempt | ((uint32_t)num_episodes << 16);
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
  bool clear = true;
  for (int i = 0; i < nb; ++i) {
    const float dx = RO_BS(W, 0, e, i) - x0, dy = RO_BS(W, 1, e, i) - y0;
    if (!(sqrtf(dx * dx + dy * dy) > RO_START_MARGIN)) clear = false;
  }
  const float tx = RO_BS(W, 0, e, target), ty = RO_BS(W, 1, e, target);
  const float d0 = sqrtf((tx - x0) * (tx - x0) + (ty - y0) * (ty - y0));
  const float d1 = sqrtf((tx - x1) * (tx - x1) + (ty - y1) * (ty - y1));
  const bool touches = !(d0 >= RO_MOTION_MARGIN && d1 >= RO_MOTION_MARGIN);
  out[0] = sx; out[1] = sy; out[2] = mx; out[3] = my;
  *accept = clear && touches;
}

// B2S_POLICY_AIMED (synthetic workloads, bench.py): a random body, a random direction, start 8 cm behind the body, push
// through it.  One Philox block: counter (stream 2 | action index << 8, global env id, num_episodes).
__device__ inline void policy_aimed(const DWorld& W, int e, uint64_t seed, int action_index, int num_episodes, float out[4]) {
  const B2SParams& P = W.P;
  const int nb = max(1, W.buf.num_movables[e]);
  const b2s_u4 r = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), 0u, 2u | ((uint32_t)action_index << 8), (uint32_t)(P.env_id_offset + e), (uint32_t)num_episodes);
  int body = (int)(b2s_u01(r.x) * (float)nb); if (body > nb - 1) body = nb - 1;
  const float angle = -B2S_PI + (2.0f * B2S_PI) * b2s_u01(r.y);
  float sn, cs;
  b2s_sincos(angle, &sn, &cs);
  const float offx = 0.5f * (P.cspace_high[0] + P.cspace_low[0]), offy = 0.5f * (P.cspace_high[1] + P.cspace_low[1]);
  const float rngx = 0.5f * (P.cspace_high[0] - P.cspace_low[0]), rngy = 0.5f * (P.cspace_high[1] - P.cspace_low[1]);
  const float tx = RO_BS(W, 0, e, body) - 0.08f * cs, ty = RO_BS(W, 1, e, body) - 0.08f * sn;
  out[0] = fminf(1.0f, fmaxf(-1.0f, (tx - offx) / rngx));
  out[1] = fminf(1.0f, fmaxf(-1.0f, (ty - offy) / rngy));
  out[2] = cs; out[3] = sn;
}

// every lane returns the same action in out[4]
__device__ inline void policy_sample(const DWorld& W, int e, int lane, uint64_t seed, int action_index, int num_episodes, float out[4]) {
  if (W.ro.policy_kind == B2S_POLICY_AIMED) { policy_aimed(W, e, seed, action_index, num_episodes, out); return; }
  const int nb = max(1, W.buf.num_movables[e]);
  const int target = num_episodes % nb;
  const float a = (float)(num_episodes * RO_ANGLE_SEED);
  const float base_angle = a - floorf(a / (2.0f * B2S_PI)) * (2.0f * B2S_PI);
  const int max_attempts = W.ro.max_attempts;
  float cand[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int a0 = 0; a0 < max_attempts; a0 += 32) {
    const int att = a0 + lane;
    bool ok = false;
    if (att < max_attempts) policy_candidate(W, e, seed, action_index, num_episodes, att, base_angle, nb, target, cand, &ok);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int last = min(31, max_attempts - 1 - a0);                 // lane of the last attempt of this round
    const int src = m ? (__ffs(m) - 1) : last;
    if (m || a0 + 32 >= max_attempts) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = __shfl_sync(0xffffffffu, cand[k], src);
      return;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = 0.0f;                         // max_attempts <= 0
}

// ---- episodes on the device (one warp per call) -----------------------------------------------------------------
// Per-env rollout state: W.ro_state[e*4 + {0: step of the episode, 1: episode index of this rollout, 2: re-samples of
// the current reset, 3: spare}].  Records are laid out [B][EP][A]...: episode `ep` of env `e` starts at (e*EP + ep).
#define RO_STEP(W, e) (W).ro_state[(size_t)(e) * 4 + 0]
#define RO_EPISODE(W, e) (W).ro_state[(size_t)(e) * 4 + 1]
#define RO_RETRY(W, e) (W).ro_state[(size_t)(e) * 4 + 2]

// RobotEnv.reset's tail + the first policy call of an episode: the settled scene is the first observation and the
// reward's previous state (robot_env.py:224-235, push_reward.py:396-405); then action 0 starts.
__device__ inline void episode_start(const DWorld& W, int e, int lane, const float* first_action) {
  const int N = W.Nmax, A = W.ro.num_actions, EP = W.ro.num_episodes;
  const int nm = W.buf.num_movables[e];
  const int ep = RO_EPISODE(W, e);
  float act[4];
  if (first_action) { for (int k = 0; k < 4; ++k) act[k] = first_action[(size_t)e * 4 + k]; }
  else policy_sample(W, e, lane, W.ro.seed, 0, W.num_episodes[e], act);
  if (lane == 0) {
    RO_STEP(W, e) = 0;
    W.buf.episode_return[e] = 0.0f; W.buf.reward[e] = 0.0f; W.buf.termination[e] = 0;
    observe_env(W, e);
    for (int i = 0; i < N; ++i) {
      W.prev_xy[((size_t)e * N + i) * 2] = (i < nm) ? RO_BS(W, 0, e, i) : 0.0f;
      W.prev_xy[((size_t)e * N + i) * 2 + 1] = (i < nm) ? RO_BS(W, 1, e, i) : 0.0f;
    }
    if (ep < EP) {
      if (W.ro.lengths) W.ro.lengths[(size_t)e * EP + ep] = 0;
      if (W.ro.positions)
        for (int i = 0; i < N; ++i)
          for (int k = 0; k < 3; ++k)
            W.ro.positions[((((size_t)e * EP + ep) * (A + 1)) * N + i) * 3 + k] = (i < nm) ? RO_BS(W, k, e, i) : 0.0f;
    }
    for (int k = 0; k < 4; ++k) W.buf.action[(size_t)e * 4 + k] = act[k];
    begin_action(W, e);
  }
  __syncwarp();
}

// PushEnv.reset's scene part for the next episode of env e (one thread): sample, then drop (RESET_DROP: wait until
// stable with the loose thresholds, push_env.py:443-447) and the final wait (RESET_WAIT), both run by the substep kernel.
__device__ inline void rollout_reset(const DWorld& W, int e) {
  reset_env_dev(W, e, W.ro.reset_seed);
  W.phase[e] = B2S_PHASE_RESET_DROP;
  int32_t* ps = W.phase_state + (size_t)e * 8;
  ps[2] = 0; ps[3] = 0;
}

// The environment has just finished an action (finish_action left it IDLE): RobotEnv.step's bookkeeping
// (robot_env.py:237-273), the transition record, and what generate_episode / generate_episodes do next
// (episode_generation.py:41-61, 88-112): next action, or next episode (reset), or nothing (the rollout's last episode).
__device__ inline void rollout_advance(const DWorld& W, int e, int lane) {
  const int N = W.Nmax;
  const int nm = W.buf.num_movables[e];
  const int A = W.ro.num_actions, EP = W.ro.num_episodes;
  const int t = RO_STEP(W, e), ep = RO_EPISODE(W, e);
  float* prev = W.prev_xy + (size_t)e * N * 2;
  int what = 0;                                      // 0 next action, 1 next episode, 2 rollout over for this env
  __syncwarp();
  if (lane == 0) {
    float s0[64], s1[64];
    for (int i = 0; i < N; ++i) {
      s0[i * 2] = prev[i * 2]; s0[i * 2 + 1] = prev[i * 2 + 1];
      s1[i * 2] = (i < nm) ? RO_BS(W, 0, e, i) : 0.0f; s1[i * 2 + 1] = (i < nm) ? RO_BS(W, 1, e, i) : 0.0f;
    }
    bool term = false;
    const float r = reward_eval(W, s0, s1, &term);
    W.buf.reward[e] = r; W.buf.termination[e] = term ? 1 : 0; W.buf.episode_return[e] += r;
    for (int i = 0; i < N * 2; ++i) prev[i] = s1[i];
    observe_env(W, e);
    const bool env_done = W.phase_state[(size_t)e * 8 + 4] != 0;
    if (t < A && ep < EP) {
      const size_t rec = ((size_t)e * EP + ep) * A + t;
      if (W.ro.actions) for (int k = 0; k < 4; ++k) W.ro.actions[rec * 4 + k] = W.buf.action[(size_t)e * 4 + k];
      if (W.ro.rewards) W.ro.rewards[rec] = r;
      if (W.ro.flags) W.ro.flags[rec] = (uint8_t)((W.buf.is_safe[e] ? 1 : 0) | (W.buf.is_effective[e] ? 2 : 0) | (term ? 4 : 0) | (env_done ? 8 : 0));
      if (W.ro.substeps) W.ro.substeps[rec] = W.num_steps[e];
      if (W.ro.positions)
        for (int i = 0; i < N; ++i)
          for (int k = 0; k < 3; ++k)
            W.ro.positions[((((size_t)e * EP + ep) * (A + 1) + t + 1) * N + i) * 3 + k] = (i < nm) ? RO_BS(W, k, e, i) : 0.0f;
    }
    const bool over = term || env_done || t + 1 >= A;
    RO_STEP(W, e) = t + 1;
    if (over) {
      if (ep < EP) {
        if (W.ro.lengths) W.ro.lengths[(size_t)e * EP + ep] = t + 1;
        if (W.ro.returns) W.ro.returns[(size_t)e * EP + ep] = W.buf.episode_return[e];
      }
      W.num_episodes[e] += 1;
      RO_EPISODE(W, e) = ep + 1;
      if (ep + 1 < EP) { RO_RETRY(W, e) = 0; rollout_reset(W, e); what = 1; }
      else what = 2;
    }
  }
  what = __shfl_sync(0xffffffffu, what, 0);
  if (what != 0) return;
  float act[4];
  policy_sample(W, e, lane, W.ro.seed, t + 1, W.num_episodes[e], act);
  if (lane == 0) {
    for (int k = 0; k < 4; ++k) W.buf.action[(size_t)e * 4 + k] = act[k];
    begin_action(W, e);
  }
  __syncwarp();
}

// The final wait of a reset has ended: re-sample scenes whose bodies fell off the table or that found no placement
// (push_env.py:460-468; Simulator.reset_scene's retry loop).  Returns 0 when the scene is valid (the env is IDLE and
// the caller starts the episode), 1 when it was re-sampled, 2 when the env gave up (error bit8, IDLE).
__device__ inline int rollout_reset_check(const DWorld& W, int e, int lane) {
  int bad = 0;
  __syncwarp();
  if (lane == 0) {
    float table_z = 0.0f;
    for (int s = 0; s < W.Ns; ++s) if (W.static_flags[s] & B2S_STATIC_IS_TABLE) table_z = W.static_pose[s * 7 + 2];
    const float zmin = table_z + W.table_dz[e];
    for (int i = 0; i < W.buf.num_movables[e]; ++i) if (RO_BS(W, 2, e, i) < zmin) bad = 1;
    if (W.error_flags[e] & 128) bad = 1;
    if (bad) {
      if (RO_RETRY(W, e) < W.ro.max_reset_retries) { RO_RETRY(W, e) += 1; rollout_reset(W, e); }
      else { W.error_flags[e] |= 256; W.phase[e] = B2S_PHASE_IDLE; bad = 2; }   // no valid scene: the env stops
    } else {
      W.phase[e] = B2S_PHASE_IDLE;
    }
  }
  return __shfl_sync(0xffffffffu, bad, 0);
}

// ---- asynchronous stepping with the policy on the host (b2s_env_async_step) -----------------------------------
// W.async_events[e]: bit1 an action finished since the last call (reward / flags / PoseObs row are fresh), bit2 a
// reset finished (PoseObs row = first observation of the episode).  One thread.
#define B2S_ASYNC_ACTION_DONE 2
#define B2S_ASYNC_RESET_DONE 4
__device__ inline void async_action_done(const DWorld& W, int e) {
  const int N = W.Nmax, nm = W.buf.num_movables[e];
  float* prev = W.prev_xy + (size_t)e * N * 2;
  float s0[64], s1[64];
  for (int i = 0; i < N; ++i) {
    s0[i * 2] = prev[i * 2]; s0[i * 2 + 1] = prev[i * 2 + 1];
    s1[i * 2] = (i < nm) ? RO_BS(W, 0, e, i) : 0.0f; s1[i * 2 + 1] = (i < nm) ? RO_BS(W, 1, e, i) : 0.0f;
  }
  bool term = false;
  const float r = reward_eval(W, s0, s1, &term);
  W.buf.reward[e] = r; W.buf.termination[e] = term ? 1 : 0; W.buf.episode_return[e] += r;
  for (int i = 0; i < N * 2; ++i) prev[i] = s1[i];
  observe_env(W, e);
  W.async_events[e] |= B2S_ASYNC_ACTION_DONE;
}
__device__ inline void async_reset_done(const DWorld& W, int e) {
  const int N = W.Nmax, nm = W.buf.num_movables[e];
  for (int i = 0; i < N; ++i) {
    W.prev_xy[((size_t)e * N + i) * 2] = (i < nm) ? RO_BS(W, 0, e, i) : 0.0f;
    W.prev_xy[((size_t)e * N + i) * 2 + 1] = (i < nm) ? RO_BS(W, 1, e, i) : 0.0f;
  }
  W.buf.episode_return[e] = 0.0f; W.buf.reward[e] = 0.0f; W.buf.termination[e] = 0;
  observe_env(W, e);
  W.async_events[e] |= B2S_ASYNC_RESET_DONE;
}
