// b2s_aux.cu -- the kernels around the substep loop: scene reset, action -> waypoints,
// PoseObs, PushEnv reward, robot commands, FK/IK entry points, contact queries, SE(3) ops.
// One thread per environment (these run once per env.step, not per substep).
#include "b2s_dev.cuh"

#define BSX(c, e, i) W.buf.body_state[((size_t)(c) * W.B + (e)) * W.Nmax + (i)]
#define MPX(c, e, i) W.mov_params[((size_t)(c) * W.B + (e)) * W.Nmax + (i)]

struct XfS { V3 p; Q4 q; };
__device__ __forceinline__ XfS xfs_from(const float* a) { XfS t; t.p = v3(a[0], a[1], a[2]); t.q = q4(a[3], a[4], a[5], a[6]); return t; }
__device__ __forceinline__ XfS xfs_mul(XfS a, XfS b) { XfS t; t.p = a.p + qrot(a.q, b.p); t.q = qmul(a.q, b.q); return t; }
__device__ __forceinline__ void xfs_store(XfS t, float* o) { o[0] = t.p.x; o[1] = t.p.y; o[2] = t.p.z; o[3] = t.q.x; o[4] = t.q.y; o[5] = t.q.z; o[6] = t.q.w; }

// scalar FK of every collision link + end effector (same arithmetic as the warp version in b2s_step.cu)
__device__ void fk_links_scalar(const DWorld& W, const float* q, const float* qd, float* lp, float* lv) {
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

__device__ void observe_env(const DWorld& W, int e) {
  const int n = W.buf.num_movables[e];
  for (int i = 0; i < W.Nmax; ++i) {
    float* o = W.buf.obs_position + ((size_t)e * W.Nmax + i) * 3;
    if (i < n) { o[0] = BSX(0, e, i); o[1] = BSX(1, e, i); o[2] = BSX(2, e, i); }
    else { o[0] = o[1] = o[2] = 0.0f; }
  }
}

// RobotEnv.reset scene part (oracle/b2o_env.cpp reset_env; push_env.py:331-471)
__global__ void k_reset(const __grid_constant__ DWorld W, const uint8_t* mask, uint64_t seed) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  if (mask && !mask[e]) return;
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
    for (int c = 0; c < 13; ++c) BSX(c, e, i) = 0.0f;
    BSX(6, e, i) = 1.0f;
    int32_t asset = 0; float scale = 1.0f, mass = 1.0f, fric = 0.0f;
    if (i < n) {
      if (i == 0 && d.num_target > 0 && d.num_target_assets > 0) asset = W.target_assets[rng.below(d.num_target_assets)];
      else asset = W.movable_assets[rng.below(d.num_movable_assets)];
      scale = rng.uni(d.scale_range[0], d.scale_range[1]);
      mass = rng.uni(d.mass_range[0], d.mass_range[1]);
      fric = rng.uni(d.friction_range[0], d.friction_range[1]);
      Q4 q = q_from_euler(er[i], ep[i], ey[i]);
      BSX(0, e, i) = px[i]; BSX(1, e, i) = py[i]; BSX(2, e, i) = pz[i];
      BSX(3, e, i) = q.x; BSX(4, e, i) = q.y; BSX(5, e, i) = q.z; BSX(6, e, i) = q.w;
    }
    MPX(0, e, i) = __int_as_float(asset); MPX(1, e, i) = scale; MPX(2, e, i) = mass; MPX(3, e, i) = fric;
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
      const DAsset& A = W.assets[__float_as_int(MPX(0, e, i))];
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

// PushEnv._execute_action prologue: waypoints + start status (oracle set_action)
__global__ void k_set_action(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  const B2SParams& P = W.P;
  const float* a = W.buf.action + (size_t)e * 4;
  float off[3], rng[3];
  for (int k = 0; k < 3; ++k) { off[k] = 0.5f * (P.cspace_high[k] + P.cspace_low[k]); rng[k] = 0.5f * (P.cspace_high[k] - P.cspace_low[k]); }
  float x = a[0] * rng[0] + off[0], y = a[1] * rng[1] + off[1];
  float z = P.finger_tip_offset + off[2];
  float x2 = fminf(P.cspace_high[0], fmaxf(P.cspace_low[0], x + a[2] * P.translation_x));
  float y2 = fminf(P.cspace_high[1], fmaxf(P.cspace_low[1], y + a[3] * P.translation_y));
  Q4 down = q_from_euler(B2S_PI, 0.0f, 0.0f);
  float* wp = W.waypoints + (size_t)e * 14;
  wp[0] = x; wp[1] = y; wp[2] = z; wp[3] = down.x; wp[4] = down.y; wp[5] = down.z; wp[6] = down.w;
  wp[7] = x2; wp[8] = y2; wp[9] = z; wp[10] = down.x; wp[11] = down.y; wp[12] = down.z; wp[13] = down.w;
  W.buf.is_safe[e] = 1; W.buf.is_effective[e] = 1;
  W.phase[e] = B2S_PHASE_INITIAL;
  int32_t* ps = W.phase_state + (size_t)e * 8;
  ps[1] = 0; ps[2] = 0; ps[3] = 0; ps[4] = 0; ps[5] = 0; ps[6] += 1;
  for (int i = 0; i < W.Nmax; ++i) {
    float* s = W.status + (((size_t)e * 2 + 0) * W.Nmax + i) * 4;
    if (i < W.buf.num_movables[e]) {
      s[0] = BSX(0, e, i); s[1] = BSX(1, e, i); s[2] = BSX(2, e, i);
      s[3] = yaw_from_q(q4(BSX(3, e, i), BSX(4, e, i), BSX(5, e, i), BSX(6, e, i)));
    } else { s[0] = s[1] = s[2] = s[3] = 0.0f; }
  }
}

// end of RobotEnv.reset: the settled xy of the movables is the reward's "previous state" of the first step
__global__ void k_begin_episode(const __grid_constant__ DWorld W, const uint8_t* mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  if (mask && !mask[e]) return;
  const int n = W.buf.num_movables[e];
  for (int i = 0; i < W.Nmax; ++i) {
    W.prev_xy[((size_t)e * W.Nmax + i) * 2] = (i < n) ? BSX(0, e, i) : 0.0f;
    W.prev_xy[((size_t)e * W.Nmax + i) * 2 + 1] = (i < n) ? BSX(1, e, i) : 0.0f;
  }
}

__global__ void k_observe(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  observe_env(W, e);
}

// ---- PushEnv reward (oracle reward(); push_reward.py:272-374, is_planning=False) ----
__device__ bool on_tiles(float x, float y, const float (*tiles)[2], int nt, float size, const float* off, float max_dist) {
  bool any = false;
  for (int t = 0; t < nt; ++t) {
    float tx = off[0] + tiles[t][0] * size, ty = off[1] + tiles[t][1] * size;
    if (fabsf(x - tx) <= 0.5f * max_dist && fabsf(y - ty) <= 0.5f * max_dist) any = true;
  }
  return any;
}
__device__ float tile_dist(float x, float y, const float (*tiles)[2], int nt, float size, const float* off) {
  float best = 3e38f;
  for (int t = 0; t < nt; ++t) {
    float dx = x - (off[0] + tiles[t][0] * size), dy = y - (off[1] + tiles[t][1] * size);
    float dd = sqrtf(dx * dx + dy * dy);
    if (dd < best) best = dd;
  }
  return best;
}
__device__ float clearing_score(const float* xy, int n) {
  float d1 = 0, d3 = 0;
  for (int i = 0; i < n; ++i) { d1 = d1 + fabsf(xy[i * 2] - 0.7f); d3 = d3 + fabsf(xy[i * 2 + 1] + 0.9f); }
  d1 = d1 / (float)n; d3 = d3 / (float)n;
  return -fminf(d1, d3);
}

__global__ void k_reward(const __grid_constant__ DWorld W, const float* prev_xy, const float* next_xy) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  const DLayout& d = *W.layout;
  const int N = W.Nmax;
  float* cur = W.prev_xy + (size_t)e * N * 2;      // becomes "previous" for the next call
  float nx[128];
  if (!next_xy) {
    const int n = W.buf.num_movables[e];
    for (int i = 0; i < N; ++i) { nx[i * 2] = (i < n) ? BSX(0, e, i) : 0.0f; nx[i * 2 + 1] = (i < n) ? BSX(1, e, i) : 0.0f; }
  } else {
    for (int i = 0; i < N * 2; ++i) nx[i] = next_xy[(size_t)e * N * 2 + i];
  }
  float s0[128];
  for (int i = 0; i < N * 2; ++i) s0[i] = prev_xy ? prev_xy[(size_t)e * N * 2 + i] : cur[i];
  const float* s1 = nx;
  const int task = W.P.task;
  float r = 0.0f;
  bool term = false, goal = false;
  float sc0 = 0.0f, sc1 = 0.0f;
  if (task == B2S_TASK_NONE) {
    W.buf.reward[e] = 1.0f; W.buf.termination[e] = 0; W.buf.episode_return[e] += 1.0f;
  } else {
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
    } else {
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
    W.buf.reward[e] = r;
    W.buf.termination[e] = (term || goal_reached) ? 1 : 0;
    W.buf.episode_return[e] += r;
  }
  for (int i = 0; i < N * 2; ++i) cur[i] = nx[i];
}

// ---- robot commands (sawyer_sim.py:186-308) ----
// cmd 0: move_to_gripper_pose, 1: move_to_joint_positions, 2: reset_targets, 3: is_limb_ready,
// 4: latch motor targets (data = q [B][7], out reinterpreted as qd [B][7] or NULL)
__device__ bool joints_reached_scalar(const DWorld& W, int e) {
  const int32_t* f = W.ctrl_flags + (size_t)e * 4;
  if (!f[1]) return true;
  const float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  for (int j = 0; j < 7; ++j) {
    float q = W.buf.joint_state[(0 * 7 + j) * W.B + e], qd = W.buf.joint_state[(1 * 7 + j) * W.B + e];
    bool pr = fabsf(c[9 + j] - q) < c[16];
    bool vr = f[2] ? true : (fabsf(0.0f - qd) < c[17]);
    if (!(pr && vr)) return false;
  }
  return true;
}

__global__ void k_arm_cmd(const __grid_constant__ DWorld W, int cmd, const float* data, const uint8_t* mask, uint8_t* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  if (mask && !mask[e]) return;
  float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  int32_t* f = W.ctrl_flags + (size_t)e * 4;
  double* T = W.ctrl_time + (size_t)e * 5;
  const double now = W.P.time_step * (double)W.num_steps[e];
  if (cmd == 0) {
    f[0] = 0; f[1] = 0;
    for (int k = 0; k < 7; ++k) c[k] = data[(size_t)e * 7 + k];
    c[7] = W.P.joint_pos_threshold; c[8] = W.P.joint_vel_threshold;
    T[0] = now; T[1] = now + (double)W.P.limb_timeout;
    f[0] = 1;
  } else if (cmd == 1) {
    f[0] = 0; f[1] = 0;
    for (int k = 0; k < 7; ++k) c[9 + k] = data[(size_t)e * 7 + k];
    c[16] = W.P.joint_pos_threshold; c[17] = W.P.joint_vel_threshold;
    T[2] = now; T[3] = now + (double)W.P.limb_timeout;
    f[1] = 1; f[2] = 0;
  } else if (cmd == 2) {
    f[0] = 0; f[1] = 0;
  } else if (cmd == 4) {
    const float* qd = (const float*)out;
    for (int k = 0; k < 7; ++k) { c[18 + k] = data[(size_t)e * 7 + k]; c[25 + k] = qd ? qd[(size_t)e * 7 + k] : 0.0f; }
    f[3] = 1;
  } else {
    if (!f[0] || now >= T[1]) f[0] = 0;
    if (!f[1] || now >= T[3] || joints_reached_scalar(W, e)) f[1] = 0;
    out[e] = (!f[0] && !f[1]) ? 1 : 0;
  }
}

__global__ void k_rebuild_colliders(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
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
  const int n = W.buf.num_movables[e];
  for (int i = 0; i < n; ++i) {
    const DAsset& A = W.assets[__float_as_int(MPX(0, e, i))];
    for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) { if (nc >= W.Hmax) { over = true; break; } cs[nc] = W.Ns + W.L + i; ch[nc] = h; ++nc; }
  }
  W.ncol[e] = nc;
  if (over) W.error_flags[e] |= 32;
  // the collider numbering changed: cached manifolds are keyed by collider index, drop them
  W.num_manifolds[e] = 0;
}

// Deals the environments to the blocks of the substep kernel before every launch.  The cost of an environment is
// persistent over hundreds of substeps (a pushed body needs 20-50 solver iterations per substep, a resting one
// 3-6; an idle environment none), the solve of one environment is a sequential chain, and a block advances in
// lock-step rounds, so a round lasts as long as the block's slowest solve.  The environments are ranked by the
// solver work of their last substep (iterations x colours, counting sort); the expensive ones are concentrated
// in a few blocks that get at most one environment per warp (one wave per stage, and all of its solves are
// equally long, so nobody waits), the cheap ones are dealt round-robin over the remaining blocks, which take
// two waves per stage and mostly short solves.  Results do not depend on the deal (environments never interact).
// Measured on the mid-push workload (4096 envs x 100 substeps): contiguous blocks 40.5 ms with the slowest block
// at 1.41x the median; cost-ranked round-robin over all blocks the same; this two-class deal 36.7 ms, slowest
// block 1.13x the median.  What remains are sporadic 50-iteration solves (a replaced contact point loses its
// warm start), which no deal can predict.
#define HEAVY_KEY 60       // iterations x colours of the last substep from which an environment counts as expensive (a resting
                           // scene has none: then the deal is a plain round-robin, which is the best for uniform work)
__global__ void __launch_bounds__(1024) k_assign_envs(const __grid_constant__ DWorld W, int mode, int nblocks) {
  __shared__ int hist[256];
  __shared__ int base[256];
  __shared__ int s_heavy;
  const int E = W.envs_per_block, Wn = W.P.warps_per_block;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < nblocks * E; i += blockDim.x) W.env_map[i] = -1;
  __syncthreads();
  for (int e = threadIdx.x; e < W.B; e += blockDim.x) {
    const int32_t* st = W.solver_stats + (size_t)e * 4;
    int key = 1 + min(254, st[1] * st[2]);
    if (mode == MODE_ENV && W.phase[e] == B2S_PHASE_IDLE) key = 0;
    atomicAdd(&hist[255 - key], 1);                 // bin 0 = most expensive
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0, heavy = 0;
    for (int i = 0; i < 256; ++i) { base[i] = acc; acc += hist[i]; if (255 - i >= HEAVY_KEY) heavy = acc; }
    s_heavy = heavy;
  }
  __syncthreads();
  // blocks of the expensive class: one environment per warp; bounded by what the other blocks can still take
  int Hb = 0;
  // (only when the stepping environments need more than one wave per block anyway: a sparse launch -- the tail of
  // a batched PushEnv.step -- is fastest with the environments spread one per block)
  const int stepping = W.B - hist[255];
  // B2S_HW expensive environments per block of that class (default: one per warp), topped up with B2S_LF of the
  // cheapest ones (default: none)
#ifndef B2S_HW
#define B2S_HW Wn
#endif
#ifndef B2S_LF
#define B2S_LF 0
#endif
  const int HW = min(B2S_HW, Wn), LF = max(0, min(B2S_LF, E - HW));
  if (E > Wn && nblocks > 1 && stepping > nblocks * Wn) {
    const int hb_max = (nblocks * E - W.B) / (E - HW - LF);
    Hb = min(min((s_heavy + HW - 1) / HW, hb_max), nblocks - 1);
    if (Hb < 0) Hb = 0;
  }
  const int hcap = Hb * HW, Lb = nblocks - Hb;
  const int nl = W.B - min(hcap, W.B), fill = min(Hb * LF, nl);
  for (int e = threadIdx.x; e < W.B; e += blockDim.x) {
    const int32_t* st = W.solver_stats + (size_t)e * 4;
    int key = 1 + min(254, st[1] * st[2]);
    if (mode == MODE_ENV && W.phase[e] == B2S_PHASE_IDLE) key = 0;
    const int p = atomicAdd(&base[255 - key], 1);
    int block, slot;
    if (p < hcap) { block = p % Hb; slot = p / Hb; }
    else {
      const int q = p - hcap, qr = nl - 1 - q;          // qr: rank from the cheap end
      if (qr < fill) { block = qr % Hb; slot = HW + qr / Hb; }
      else { block = Hb + q % Lb; slot = q / Lb; }
    }
    W.env_map[(size_t)block * E + slot] = e;
  }
}

// scalar DLS IK (same arithmetic as arm_ik in b2s_step.cu / oracle arm_ik)
__global__ void k_ik(const __grid_constant__ DWorld W, const float* pose, const float* q_start, float* q_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  const DArm* arm = W.arm;
  const B2SParams& P = W.P;
  float q[B2S_NUM_JOINTS];
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q[j] = q_start[(size_t)j * W.B + e];
  XfS tgt = xfs_from(pose + (size_t)e * 7);
  XfS eel = xfs_from(arm->ee);
  const float res2 = P.ik_residual * P.ik_residual;
  for (int it = 0; it < P.ik_max_iters; ++it) {
    V3 ax[B2S_NUM_JOINTS], org[B2S_NUM_JOINTS];
    XfS T = xfs_from(arm->base);
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
      XfS Tj = xfs_mul(T, xfs_from(arm->joint_origin[j]));
      V3 a = v3(arm->joint_axis[j][0], arm->joint_axis[j][1], arm->joint_axis[j][2]);
      ax[j] = qrot(Tj.q, a); org[j] = Tj.p;
      T.p = Tj.p; T.q = qmul(Tj.q, q_axis_angle(a, q[j]));
    }
    XfS ee = xfs_mul(T, eel);
    V3 ep = tgt.p - ee.p;
    V3 er = q_to_rotvec(qmul(tgt.q, qconj(ee.q)));
    if (len2(ep) < res2 && len2(er) < res2) break;
    float J[6][B2S_NUM_JOINTS];
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
      V3 jv = cross(ax[j], ee.p - org[j]);
      J[0][j] = jv.x; J[1][j] = jv.y; J[2][j] = jv.z; J[3][j] = ax[j].x; J[4][j] = ax[j].y; J[5][j] = ax[j].z;
    }
    float A[36], y[6] = {ep.x, ep.y, ep.z, er.x, er.y, er.z};
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        float s = 0.0f;
        for (int j = 0; j < B2S_NUM_JOINTS; ++j) s = s + J[r][j] * J[c][j];
        if (r == c) s = s + P.ik_damping * P.ik_damping;
        A[r * 6 + c] = s;
      }
    if (!b2s_chol6_solve(A, y)) break;
    float dq[B2S_NUM_JOINTS], m = 0.0f;
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
      float s = 0.0f;
      for (int r = 0; r < 6; ++r) s = s + J[r][j] * y[r];
      dq[j] = s; m = fmaxf(m, fabsf(s));
    }
    float k = (m > P.ik_max_step) ? (P.ik_max_step / m) : 1.0f;
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) q[j] = q[j] + dq[j] * k;
  }
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q_out[(size_t)j * W.B + e] = fminf(arm->upper[j], fmaxf(arm->lower[j], q[j]));
}

__global__ void k_fk(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  float q[7], qd[7];
  for (int j = 0; j < 7; ++j) { q[j] = W.buf.joint_state[(0 * 7 + j) * W.B + e]; qd[j] = W.buf.joint_state[(1 * 7 + j) * W.B + e]; }
  fk_links_scalar(W, q, qd, W.link_poses + (size_t)e * (W.L + 1) * 7, W.link_vel + (size_t)e * W.L * 6);
}

__global__ void k_query_contacts(const __grid_constant__ DWorld W, uint8_t* arm_table, uint8_t* arm_movable) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  int cf = W.contact_flags[e];
  if (arm_table) arm_table[e] = cf & 1;
  if (arm_movable) arm_movable[e] = (cf >> 1) & 1;
}

// copy the current manifold buffer of every env into flat export arrays
__global__ void k_export_manifolds(const __grid_constant__ DWorld W, int32_t* keys, int32_t* npts, float* pts) {
  const int M = W.P.max_manifolds;
  const size_t total = (size_t)W.B * M;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int e = (int)(i / M), k = (int)(i % M);
    size_t src = ((size_t)W.man_parity[e] * W.B + e) * M + k;
    const bool live = k < W.num_manifolds[e];       // slots past the count hold leftovers of earlier substeps
    int n = live ? W.man_npts[src] : 0;
    keys[i] = live ? W.man_keys[src] : -1; npts[i] = n;
    for (int t = 0; t < 4 * B2S_CP_FLOATS; ++t) pts[i * 4 * B2S_CP_FLOATS + t] = (t < n * B2S_CP_FLOATS) ? W.man_pts[src * 4 * B2S_CP_FLOATS + t] : 0.0f;
  }
}

// ---- SE(3) element kernels (robovat.math / third_party.transformations parity) ----
__global__ void k_se3(int op, const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op == 0) { Q4 q = q_from_euler(a[i * 3], a[i * 3 + 1], a[i * 3 + 2]); out[i * 4] = q.x; out[i * 4 + 1] = q.y; out[i * 4 + 2] = q.z; out[i * 4 + 3] = q.w; }
  else if (op == 1) { V3 eu = euler_from_q(q4(a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3])); out[i * 3] = eu.x; out[i * 3 + 1] = eu.y; out[i * 3 + 2] = eu.z; }
  else if (op == 2) {
    M3 m = q_to_m3(q4(a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3]));
    float* o = out + i * 9;
    o[0] = m.r0.x; o[1] = m.r0.y; o[2] = m.r0.z; o[3] = m.r1.x; o[4] = m.r1.y; o[5] = m.r1.z; o[6] = m.r2.x; o[7] = m.r2.y; o[8] = m.r2.z;
  } else if (op == 3) {
    Q4 q = qmul(q4(a[i * 4], a[i * 4 + 1], a[i * 4 + 2], a[i * 4 + 3]), q4(b[i * 4], b[i * 4 + 1], b[i * 4 + 2], b[i * 4 + 3]));
    out[i * 4] = q.x; out[i * 4 + 1] = q.y; out[i * 4 + 2] = q.z; out[i * 4 + 3] = q.w;
  } else if (op == 4) {
    // Pose.inverse (robovat/math/pose.py:161-172): p' = -p.R, R' = R^T
    Q4 q = q4(a[i * 7 + 3], a[i * 7 + 4], a[i * 7 + 5], a[i * 7 + 6]);
    V3 p = v3(a[i * 7], a[i * 7 + 1], a[i * 7 + 2]);
    V3 pi = mtmul(q_to_m3(q), -p);
    Q4 qi = qconj(q);
    float* o = out + i * 7;
    o[0] = pi.x; o[1] = pi.y; o[2] = pi.z; o[3] = qi.x; o[4] = qi.y; o[5] = qi.z; o[6] = qi.w;
  } else {
    // Pose.transform (pose.py:174-189): position = a.p + R_a b.p ; orientation = R_a R_b
    Q4 qa = q4(a[i * 7 + 3], a[i * 7 + 4], a[i * 7 + 5], a[i * 7 + 6]);
    Q4 qb = q4(b[i * 7 + 3], b[i * 7 + 4], b[i * 7 + 5], b[i * 7 + 6]);
    V3 p = v3(a[i * 7], a[i * 7 + 1], a[i * 7 + 2]) + qrot(qa, v3(b[i * 7], b[i * 7 + 1], b[i * 7 + 2]));
    Q4 q = qmul(qa, qb);
    float* o = out + i * 7;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
  }
}

static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

void b2s_launch_assign_envs(const DWorld& W, int mode, cudaStream_t s) {
  k_assign_envs<<<1, 1024, 0, s>>>(W, mode, W.num_blocks);
}
void b2s_launch_reset(const DWorld& W, const uint8_t* mask, uint64_t seed, cudaStream_t s) { k_reset<<<blocks_for(W.B, 64), 64, 0, s>>>(W, mask, seed); }
void b2s_launch_set_action(const DWorld& W, cudaStream_t s) { k_set_action<<<blocks_for(W.B, 128), 128, 0, s>>>(W); }
void b2s_launch_begin_episode(const DWorld& W, const uint8_t* mask, cudaStream_t s) { k_begin_episode<<<blocks_for(W.B, 128), 128, 0, s>>>(W, mask); }
void b2s_launch_observe(const DWorld& W, cudaStream_t s) { k_observe<<<blocks_for(W.B, 128), 128, 0, s>>>(W); }
void b2s_launch_reward(const DWorld& W, const float* p, const float* n, cudaStream_t s) { k_reward<<<blocks_for(W.B, 128), 128, 0, s>>>(W, p, n); }
void b2s_launch_arm_cmd(const DWorld& W, int cmd, const float* data, const uint8_t* mask, uint8_t* out, cudaStream_t s) {
  k_arm_cmd<<<blocks_for(W.B, 128), 128, 0, s>>>(W, cmd, data, mask, out);
}
void b2s_launch_rebuild_colliders(const DWorld& W, cudaStream_t s) { k_rebuild_colliders<<<blocks_for(W.B, 64), 64, 0, s>>>(W); }
void b2s_launch_ik(const DWorld& W, const float* pose, const float* qs, float* qo, cudaStream_t s) { k_ik<<<blocks_for(W.B, 64), 64, 0, s>>>(W, pose, qs, qo); }
void b2s_launch_fk(const DWorld& W, cudaStream_t s) { k_fk<<<blocks_for(W.B, 64), 64, 0, s>>>(W); }
void b2s_launch_query_contacts(const DWorld& W, uint8_t* at, uint8_t* am, cudaStream_t s) { k_query_contacts<<<blocks_for(W.B, 128), 128, 0, s>>>(W, at, am); }
void b2s_launch_export_manifolds(const DWorld& W, int32_t* keys, int32_t* npts, float* pts, cudaStream_t s) {
  k_export_manifolds<<<148 * 4, 128, 0, s>>>(W, keys, npts, pts);
}
void b2s_launch_se3(int op, const float* a, const float* b, float* out, int n, cudaStream_t s) { k_se3<<<blocks_for(n, 128), 128, 0, s>>>(op, a, b, out, n); }
