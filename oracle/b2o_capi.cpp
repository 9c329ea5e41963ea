/* b2o_capi.cpp -- C entry points of the CPU oracle (libb2o.so), loaded with
 * ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h).  Same argument meaning as the
 * b2s_* functions in include/b2s.h, but every pointer is a HOST pointer and the
 * work runs on the CPU (optionally on several host threads, one environment at a time each, which are
 * independent exactly like the reference's one-pybullet-world-per-process,
 * tools/parallel_run.py:61-78).
 */
#include <string.h>

#include <new>

#include <atomic>
#include <thread>
#include <vector>

#include "b2o_world.h"

using namespace b2o;

static thread_local std::string g_err;
static int g_threads = 1;

extern "C" {

const char* b2o_last_error(void) { return g_err.c_str(); }
int b2o_set_threads(int n) { g_threads = n < 1 ? 1 : n; return 0; }
int b2o_max_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

}  // extern "C"
/* environments are independent: hand them out to g_threads host threads */
template <class F>
static void parallel_envs(int B, F f) {
  if (g_threads <= 1 || B <= 1) { for (int e = 0; e < B; ++e) f(e); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  int nt = g_threads < B ? g_threads : B;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([&]() { for (;;) { int e = next.fetch_add(1); if (e >= B) break; f(e); } });
  for (auto& th : pool) th.join();
}
extern "C" {

int b2o_create(const B2SParams* p, World** out) {
  if (!p || !out || p->num_envs <= 0 || p->max_movables <= 0 || p->max_movables > 64) { g_err = "b2o_create: bad params"; return B2S_E_INVALID; }
  World* w = new (std::nothrow) World();
  if (!w) return B2S_E_INVALID;
  w->P = *p;
  w->B = p->num_envs; w->Nmax = p->max_movables;
  w->Ns = 0; w->L = 0; w->NB = 0; w->Hmax = p->max_colliders;
  w->cam_per_env = 0; w->substeps_executed = 0;
  w->substeps_executed_env.assign(p->num_envs, 0);
  *out = w;
  return 0;
}

int b2o_destroy(World* w) { delete w; return 0; }

int b2o_load_scene(World* w, const B2SSceneDesc* d) {
  if (!w || !d) return B2S_E_INVALID;
  Scene& S = w->S;
  S.d = *d;
  S.verts.resize(d->num_verts);
  for (int i = 0; i < d->num_verts; ++i) S.verts[i] = v3(d->verts[i * 3], d->verts[i * 3 + 1], d->verts[i * 3 + 2]);
  S.hulls.resize(d->num_hulls);
  for (int h = 0; h < d->num_hulls; ++h) {
    S.hulls[h].voff = d->hull_vert_off[h]; S.hulls[h].vcnt = d->hull_vert_cnt[h]; S.hulls[h].margin = d->hull_margin[h];
    S.hulls[h].poff = d->hull_plane_off ? d->hull_plane_off[h] : 0;
    S.hulls[h].pcnt = d->hull_plane_cnt ? d->hull_plane_cnt[h] : 0;
    if (S.hulls[h].vcnt < 1 || S.hulls[h].vcnt > 64) { g_err = "hull vertex count must be 1..64"; return B2S_E_INVALID; }
  }
  S.planes.assign(d->planes, d->planes + (size_t)d->num_planes * 4);
  S.assets.resize(d->num_assets);
  for (int a = 0; a < d->num_assets; ++a) { S.assets[a].hoff = d->asset_hull_off[a]; S.assets[a].hcnt = d->asset_hull_cnt[a]; }
  S.static_asset.assign(d->static_asset, d->static_asset + d->num_statics);
  S.static_pose.assign(d->static_pose, d->static_pose + (size_t)d->num_statics * 7);
  S.static_friction.assign(d->static_friction, d->static_friction + d->num_statics);
  S.static_flags.assign(d->static_flags, d->static_flags + d->num_statics);
  S.movable_assets.assign(d->movable_assets, d->movable_assets + d->num_movable_assets);
  S.target_assets.assign(d->target_assets, d->target_assets + d->num_target_assets);
  derive_scene(S);
  const int B = w->B, N = w->Nmax;
  w->Ns = d->num_statics; w->L = d->num_links; w->NB = w->Ns + w->L + N;
  const B2SParams& P = w->P;
  w->body_state.assign((size_t)13 * B * N, 0.0f);
  w->joint_state.assign((size_t)2 * 7 * B, 0.0f);
  const int G = P.num_goal_steps > 0 ? P.num_goal_steps : 1;
  w->action.assign((size_t)B * G * 4, 0.0f);
  w->obs_position.assign((size_t)B * N * 3, 0.0f);
  w->num_movables.assign(B, 0); w->body_mask.assign((size_t)B * N, 0);
  w->depth.assign((size_t)B * P.cam_height * P.cam_width, 0.0f);
  w->segmask.assign((size_t)B * P.cam_height * P.cam_width, 255);
  w->point_cloud.assign((size_t)B * N * P.num_points * 3, 0.0f);
  w->reward.assign(B, 0.0f); w->termination.assign(B, 0); w->is_safe.assign(B, 1); w->is_effective.assign(B, 1);
  w->episode_return.assign(B, 0.0f);
  w->man_keys.assign((size_t)B * P.max_manifolds, -1); w->man_npts.assign((size_t)B * P.max_manifolds, 0);
  w->man_pts.assign((size_t)B * P.max_manifolds * 4 * B2S_CP_FLOATS, 0.0f);
  w->num_manifolds.assign(B, 0);
  w->pair_keys.assign((size_t)B * P.max_pairs, 0); w->num_pairs.assign(B, 0);
  w->phase.assign(B, B2S_PHASE_IDLE); w->num_steps.assign(B, 0);
  w->ctrl.assign((size_t)B * B2S_CTRL_FLOATS, 0.0f); w->ctrl_flags.assign((size_t)B * 4, 0); w->ctrl_time.assign((size_t)B * 5, 0.0);
  w->link_poses.assign((size_t)B * (w->L + 1) * 7, 0.0f); w->link_vel.assign((size_t)B * w->L * 6, 0.0f);
  w->mov_params.assign((size_t)4 * B * N, 0.0f);
  w->table_dz.assign(B, 0.0f); w->error_flags.assign(B, 0);
  w->waypoints.assign((size_t)B * G * 14, 0.0f); w->status.assign((size_t)B * 2 * N * 4, 0.0f);
  w->contact_flags.assign(B, 0); w->phase_state.assign((size_t)B * 8, 0); w->solver_stats.assign((size_t)B * 4, 0);
  for (int e = 0; e < B; ++e) w->phase_state[(size_t)e * 8] = -1;
  w->ncol.assign(B, 0); w->col_slot.assign((size_t)B * w->Hmax, 0); w->col_hull.assign((size_t)B * w->Hmax, 0);
  w->reset_count.assign(B, 0);
  memset(&w->ro, 0, sizeof(w->ro));
  w->ro_state.assign((size_t)B * 4, 0); w->num_episodes.assign(B, 0); w->async_events.assign(B, 0);
  w->prev_xy.assign((size_t)B * N * 2, 0.0f);
  w->cam.assign((size_t)B * 21, 0.0f);
  return 0;
}

int b2o_reset(World* w, const uint8_t* mask, uint64_t seed) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e]) reset_env(*w, e, seed);
  return 0;
}

static void run_env(World* w, int e, int n, int mode, float lin, float ang, int max_steps) {
  /* mode 0: raw substeps; 1: env substeps (phase machine); 2: settle; 3: rollout substeps; 4: async substeps */
  if (mode == 0) { for (int i = 0; i < n; ++i) substep(*w, e); return; }
  if (mode == 3) { for (int i = 0; i < n; ++i) { if (w->phase[e] == B2S_PHASE_IDLE) break; rollout_substep(*w, e); } return; }
  if (mode == 4) { for (int i = 0; i < n; ++i) { if (w->phase[e] == B2S_PHASE_IDLE) break; async_substep(*w, e); } return; }
  if (mode == 1) { for (int i = 0; i < n; ++i) { if (w->phase[e] == B2S_PHASE_IDLE) break; env_substep(*w, e); } return; }
  int steps = 0, stable = 0;
  while (1) {
    substep(*w, e);
    ++steps;
    if (steps < w->P.stable_check_after) continue;
    bool ok = true;
    for (int i = 0; i < w->num_movables[e]; ++i) {
      size_t B = w->B, N = w->Nmax;
      V3 v = v3(w->body_state[(7 * B + e) * N + i], w->body_state[(8 * B + e) * N + i], w->body_state[(9 * B + e) * N + i]);
      V3 o = v3(w->body_state[(10 * B + e) * N + i], w->body_state[(11 * B + e) * N + i], w->body_state[(12 * B + e) * N + i]);
      if (len(v) >= lin || len(o) >= ang) { ok = false; break; }
    }
    if (ok) ++stable;
    if (stable >= w->P.stable_min_steps || steps >= max_steps) break;
  }
}

static void run_all(World* w, int n, int mode, float lin, float ang, int max_steps, const uint8_t* mask = nullptr) {
  const int B = w->B;
  /* substep() counts per env (no race between the host threads); the total is summed here */
  int64_t base = w->substeps_executed;
  std::vector<int64_t> added(B, 0);
  parallel_envs(B, [&](int e) {
    if (mask && !mask[e]) return;
    const int64_t before = w->substeps_executed_env[e];
    run_env(w, e, n, mode, lin, ang, max_steps);
    added[e] = w->substeps_executed_env[e] - before;
  });
  int64_t add = 0;
  for (int e = 0; e < B; ++e) add += added[e];
  w->substeps_executed = base + add;
}

int b2o_step(World* w, int n) { run_all(w, n, 0, 0, 0, 0); return 0; }
int b2o_settle(World* w, float lin, float ang, int max_steps) { run_all(w, 0, 2, lin, ang, max_steps); return 0; }
int b2o_settle_masked(World* w, const uint8_t* mask, float lin, float ang, int max_steps) { run_all(w, 0, 2, lin, ang, max_steps, mask); return 0; }
/* end of RobotEnv.reset (robot_env.py:224-235): the settled scene is the first step's prev_obs_data */
int b2o_begin_episode(World* w, const uint8_t* mask) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e])
    for (int i = 0; i < w->Nmax; ++i) {
      const bool live = i < w->num_movables[e];
      w->prev_xy[((size_t)e * w->Nmax + i) * 2] = live ? w->body_state[((size_t)0 * w->B + e) * w->Nmax + i] : 0.0f;
      w->prev_xy[((size_t)e * w->Nmax + i) * 2 + 1] = live ? w->body_state[((size_t)1 * w->B + e) * w->Nmax + i] : 0.0f;
    }
  return 0;
}
int b2o_set_action(World* w) { w->ro.enabled = 0; for (int e = 0; e < w->B; ++e) set_action(*w, e); return 0; }
int b2o_env_substeps(World* w, int n, int* unfinished) {
  run_all(w, n, 1, 0, 0, 0);
  if (unfinished) { int u = 0; for (int e = 0; e < w->B; ++e) u += (w->phase[e] != B2S_PHASE_IDLE); *unfinished = u; }
  return 0;
}
/* b2s_rollout_begin / b2s_rollout_run: scalars, flags / substeps / lengths from *r; the real-typed record arrays are
 * passed separately because they are double in the double-precision build */
int b2o_rollout_begin(World* w, const B2SRollout* r, const float* first_action, float* actions, float* rewards, float* positions,
                      float* returns) {
  if (!r || r->num_actions < 1 || r->num_episodes < 1 || r->max_attempts < 1 || r->max_attempts > 65535) { g_err = "b2o_rollout_begin: bad rollout"; return B2S_E_INVALID; }
  World::Rollout& ro = w->ro;
  ro.enabled = 1; ro.num_actions = r->num_actions; ro.max_attempts = r->max_attempts; ro.num_episodes = r->num_episodes;
  ro.max_reset_retries = r->max_reset_retries; ro.drop_max_steps = r->drop_max_steps; ro.policy_kind = r->policy_kind;
  ro.drop_lin = r->drop_lin_threshold; ro.drop_ang = r->drop_ang_threshold;
  ro.seed = r->seed; ro.reset_seed = r->reset_seed;
  ro.actions = actions; ro.rewards = rewards; ro.positions = positions; ro.returns = returns;
  ro.flags = r->flags; ro.substeps = r->substeps; ro.lengths = r->lengths;
  for (int e = 0; e < w->B; ++e) rollout_begin(*w, e, first_action);
  return 0;
}
int b2o_rollout_run(World* w, int n, int* unfinished) {
  run_all(w, n, 3, 0, 0, 0);
  if (unfinished) { int u = 0; for (int e = 0; e < w->B; ++e) u += (w->phase[e] != B2S_PHASE_IDLE); *unfinished = u; }
  return 0;
}
int b2o_env_async_step(World* w, const uint8_t* command, int n, uint64_t reset_seed, uint8_t* status) {
  World::Rollout& ro = w->ro;
  if (ro.enabled != 2) { memset(&ro, 0, sizeof(ro)); ro.enabled = 2; ro.max_reset_retries = 8; ro.drop_lin = 0.1f; ro.drop_ang = 0.1f; ro.drop_max_steps = 500; }
  ro.reset_seed = reset_seed;
  for (int e = 0; e < w->B; ++e) async_command(*w, e, command ? command[e] : 0);
  if (n > 0) run_all(w, n, 4, 0, 0, 0);
  if (status)
    for (int e = 0; e < w->B; ++e)
      status[e] = (uint8_t)((w->phase[e] == B2S_PHASE_IDLE ? 1 : 0) | (w->async_events[e] & 6) | (w->phase_state[(size_t)e * 8 + 4] ? 8 : 0));
  return 0;
}
/* the device policy for one observation (testing): positions are read from the world's body state */
int b2o_policy_sample(World* w, uint64_t seed, int action_index, const int32_t* num_episodes, int max_attempts, float* out /*[B][4]*/) {
  for (int e = 0; e < w->B; ++e) policy_sample(*w, e, seed, action_index, num_episodes[e], max_attempts, out + (size_t)e * 4);
  return 0;
}
int b2o_arm_move_to_gripper_pose(World* w, const float* pose, const uint8_t* mask) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e]) { arm_reset_targets(*w, e); arm_set_link_target(*w, e, pose + (size_t)e * 7); }
  return 0;
}
int b2o_arm_move_to_joint_positions(World* w, const float* q, const uint8_t* mask) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e]) { arm_reset_targets(*w, e); arm_set_joint_target(*w, e, q + (size_t)e * 7); }
  return 0;
}
int b2o_arm_reset_targets(World* w, const uint8_t* mask) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e]) arm_reset_targets(*w, e);
  return 0;
}
int b2o_set_motor_targets(World* w, const float* q, const float* qd, const uint8_t* mask) {
  for (int e = 0; e < w->B; ++e) if (!mask || mask[e]) {
    float* c = &w->ctrl[(size_t)e * B2S_CTRL_FLOATS];
    for (int k = 0; k < 7; ++k) { c[18 + k] = q[(size_t)e * 7 + k]; c[25 + k] = qd ? qd[(size_t)e * 7 + k] : 0.0f; }
    w->ctrl_flags[(size_t)e * 4 + 3] = 1;
  }
  return 0;
}
int b2o_rebuild_colliders(World* w) {
  for (int e = 0; e < w->B; ++e) { build_colliders(*w, e); w->num_manifolds[e] = 0; }
  return 0;
}
int b2o_arm_is_ready(World* w, uint8_t* out) { for (int e = 0; e < w->B; ++e) out[e] = (uint8_t)arm_is_ready(*w, e); return 0; }
int b2o_inverse_kinematics(World* w, const float* pose, const float* q_start, float* q_out) {
  for (int e = 0; e < w->B; ++e) {
    float qs[7], qo[7];
    for (int j = 0; j < 7; ++j) qs[j] = q_start[(size_t)j * w->B + e];
    arm_ik(*w, pose + (size_t)e * 7, qs, qo);
    for (int j = 0; j < 7; ++j) q_out[(size_t)j * w->B + e] = qo[j];
  }
  return 0;
}
int b2o_forward_kinematics(World* w) {
  for (int e = 0; e < w->B; ++e) {
    float q[7], qd[7];
    for (int j = 0; j < 7; ++j) { q[j] = w->joint_state[(0 * 7 + j) * w->B + e]; qd[j] = w->joint_state[(1 * 7 + j) * w->B + e]; }
    arm_fk(*w, q, qd, &w->link_poses[(size_t)e * (w->L + 1) * 7], &w->link_vel[(size_t)e * w->L * 6]);
  }
  return 0;
}
int b2o_query_contacts(World* w, uint8_t* arm_table, uint8_t* arm_movable) {
  for (int e = 0; e < w->B; ++e) { if (arm_table) arm_table[e] = w->contact_flags[e] & 1; if (arm_movable) arm_movable[e] = (w->contact_flags[e] >> 1) & 1; }
  return 0;
}
int b2o_observe(World* w) { for (int e = 0; e < w->B; ++e) observe(*w, e); return 0; }
int b2o_reward(World* w, const float* prev_xy, const float* next_xy) {
  std::vector<float> cur;
  if (!next_xy) {
    cur.resize((size_t)w->B * w->Nmax * 2);
    for (int e = 0; e < w->B; ++e) {
      observe(*w, e);
      for (int i = 0; i < w->Nmax; ++i) { cur[((size_t)e * w->Nmax + i) * 2] = w->obs_position[((size_t)e * w->Nmax + i) * 3]; cur[((size_t)e * w->Nmax + i) * 2 + 1] = w->obs_position[((size_t)e * w->Nmax + i) * 3 + 1]; }
    }
    next_xy = cur.data();
  }
  const float* prev = prev_xy ? prev_xy : w->prev_xy.data();
  for (int e = 0; e < w->B; ++e) reward(*w, e, prev, next_xy);
  if (!prev_xy || true) memcpy(w->prev_xy.data(), next_xy, sizeof(float) * (size_t)w->B * w->Nmax * 2);
  return 0;
}
int b2o_set_camera(World* w, const float* K, const float* R, const float* t, int per_env) {
  w->cam_per_env = per_env;
  for (int e = 0; e < w->B; ++e) {
    size_t o = per_env ? (size_t)e : 0;
    memcpy(&w->cam[(size_t)e * 21], K + o * 9, sizeof(float) * 9); memcpy(&w->cam[(size_t)e * 21 + 9], R + o * 9, sizeof(float) * 9);
    memcpy(&w->cam[(size_t)e * 21 + 18], t + o * 3, sizeof(float) * 3);
  }
  return 0;
}
int b2o_render(World* w) {
  parallel_envs(w->B, [&](int e) { render(*w, e); });
  return 0;
}
int b2o_point_cloud(World* w, uint64_t seed) { for (int e = 0; e < w->B; ++e) point_cloud(*w, e, seed); return 0; }
int64_t b2o_substeps_executed(World* w) { return w->substeps_executed; }

/* the shared SE(3) leaf math on the CPU (ops as b2s_se3_*: 0 quat_from_euler, 1 euler_from_quat,
 * 2 matrix_from_quat, 3 quat_multiply, 4 pose_inverse, 5 pose_transform) */
int b2o_se3(int op, const float* a, const float* b, float* out, int n) {
  for (int i = 0; i < n; ++i) {
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
      Q4 q = q4(a[i * 7 + 3], a[i * 7 + 4], a[i * 7 + 5], a[i * 7 + 6]);
      V3 pi = mtmul(q_to_m3(q), -v3(a[i * 7], a[i * 7 + 1], a[i * 7 + 2]));
      Q4 qi = qconj(q);
      float* o = out + i * 7;
      o[0] = pi.x; o[1] = pi.y; o[2] = pi.z; o[3] = qi.x; o[4] = qi.y; o[5] = qi.z; o[6] = qi.w;
    } else {
      Q4 qa = q4(a[i * 7 + 3], a[i * 7 + 4], a[i * 7 + 5], a[i * 7 + 6]);
      Q4 qb = q4(b[i * 7 + 3], b[i * 7 + 4], b[i * 7 + 5], b[i * 7 + 6]);
      V3 p = v3(a[i * 7], a[i * 7 + 1], a[i * 7 + 2]) + qrot(qa, v3(b[i * 7], b[i * 7 + 1], b[i * 7 + 2]));
      Q4 q = qmul(qa, qb);
      float* o = out + i * 7;
      o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
    }
  }
  return 0;
}

/* ids < 100: the B2S_ARR_* arrays; >= 100: the arrays that are caller-owned B2SBuffers on the CUDA side */
int b2o_array(World* w, int which, void** ptr, int64_t* bytes) {
#define RET(v) { *ptr = (void*)(v).data(); *bytes = (int64_t)((v).size() * sizeof((v)[0])); return 0; }
  switch (which) {
    case B2S_ARR_MANIFOLD_KEYS: RET(w->man_keys) case B2S_ARR_MANIFOLD_NPTS: RET(w->man_npts)
    case B2S_ARR_MANIFOLD_PTS: RET(w->man_pts) case B2S_ARR_NUM_MANIFOLDS: RET(w->num_manifolds)
    case B2S_ARR_PAIR_KEYS: RET(w->pair_keys) case B2S_ARR_NUM_PAIRS: RET(w->num_pairs)
    case B2S_ARR_PHASE: RET(w->phase) case B2S_ARR_NUM_STEPS: RET(w->num_steps)
    case B2S_ARR_CTRL: RET(w->ctrl) case B2S_ARR_CTRL_FLAGS: RET(w->ctrl_flags)
    case B2S_ARR_LINK_POSES: RET(w->link_poses) case B2S_ARR_MOV_PARAMS: RET(w->mov_params)
    case B2S_ARR_TABLE_DZ: RET(w->table_dz) case B2S_ARR_ERROR_FLAGS: RET(w->error_flags)
    case B2S_ARR_WAYPOINTS: RET(w->waypoints) case B2S_ARR_STATUS: RET(w->status)
    case B2S_ARR_CONTACT_FLAGS: RET(w->contact_flags) case B2S_ARR_PHASE_STATE: RET(w->phase_state)
    case B2S_ARR_SOLVER_STATS: RET(w->solver_stats)
    case B2S_ARR_CTRL_TIME: RET(w->ctrl_time) case B2S_ARR_LINK_VEL: RET(w->link_vel)
    case B2S_ARR_NUM_EPISODES: RET(w->num_episodes) case B2S_ARR_ROLLOUT_STATE: RET(w->ro_state)
    case B2S_ARR_NUM_COLLIDERS: RET(w->ncol) case B2S_ARR_COL_SLOT: RET(w->col_slot) case B2S_ARR_COL_HULL: RET(w->col_hull)
    case 100: RET(w->body_state) case 101: RET(w->joint_state) case 102: RET(w->action)
    case 103: RET(w->obs_position) case 104: RET(w->num_movables) case 105: RET(w->body_mask)
    case 106: RET(w->depth) case 107: RET(w->segmask) case 108: RET(w->point_cloud)
    case 109: RET(w->reward) case 110: RET(w->termination) case 111: RET(w->is_safe)
    case 112: RET(w->is_effective) case 113: RET(w->episode_return)
  }
#undef RET
  g_err = "b2o_array: unknown id";
  return B2S_E_INVALID;
}

}  // extern "C"
