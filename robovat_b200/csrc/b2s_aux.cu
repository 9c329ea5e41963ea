// b2s_aux.cu -- the kernels around the substep loop: scene reset, action -> waypoints,
// PoseObs, PushEnv reward, robot commands, FK/IK entry points, contact queries, SE(3) ops.
// One thread per environment (these run once per env.step, not per substep).
#include "b2s_dev.cuh"
#include "b2s_rollout.cuh"

#define BSX(c, e, i) W.buf.body_state[((size_t)(c) * W.B + (e)) * W.Nmax + (i)]
#define MPX(c, e, i) W.mov_params[((size_t)(c) * W.B + (e)) * W.Nmax + (i)]

// RobotEnv.reset scene part (oracle/b2o_env.cpp reset_env; push_env.py:331-471)
__global__ void k_reset(const __grid_constant__ DWorld W, const uint8_t* mask, uint64_t seed) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  if (mask && !mask[e]) return;
  reset_env_dev(W, e, seed);
}

// PushEnv._execute_action prologue: waypoints + start status (oracle set_action)
__global__ void k_set_action(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  begin_action(W, e);
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

// ---- PushEnv reward (reward_eval in b2s_rollout.cuh; push_reward.py:272-374, is_planning=False) ----
__global__ void k_reward(const __grid_constant__ DWorld W, const float* prev_xy, const float* next_xy) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  const int N = W.Nmax;
  float* cur = W.prev_xy + (size_t)e * N * 2;      // becomes "previous" for the next call
  float nx[128], s0[128];
  if (!next_xy) {
    const int n = W.buf.num_movables[e];
    for (int i = 0; i < N; ++i) { nx[i * 2] = (i < n) ? BSX(0, e, i) : 0.0f; nx[i * 2 + 1] = (i < n) ? BSX(1, e, i) : 0.0f; }
  } else {
    for (int i = 0; i < N * 2; ++i) nx[i] = next_xy[(size_t)e * N * 2 + i];
  }
  for (int i = 0; i < N * 2; ++i) s0[i] = prev_xy ? prev_xy[(size_t)e * N * 2 + i] : cur[i];
  bool done = false;
  const float r = reward_eval(W, s0, nx, &done);
  W.buf.reward[e] = r;
  W.buf.termination[e] = done ? 1 : 0;
  W.buf.episode_return[e] += r;
  for (int i = 0; i < N * 2; ++i) cur[i] = nx[i];
}

// start of a device-side episode in every env (one warp per env): step counter, first observation, first action
__global__ void k_rollout_begin(const __grid_constant__ DWorld W, const float* first_action) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= W.B) return;
  if (lane == 0) { RO_EPISODE(W, e) = 0; RO_RETRY(W, e) = 0; }
  __syncwarp();
  episode_start(W, e, lane, first_action);
}

// b2s_env_async_step, before the substeps: start the host's action in the envs it addressed (1) or re-sample their scene (2)
__global__ void k_async_commands(const __grid_constant__ DWorld W, const uint8_t* command) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  W.async_events[e] = 0;
  const int cmd = command ? command[e] : 0;
  if (W.phase[e] != B2S_PHASE_IDLE) return;            // a busy env ignores commands
  if (cmd == 1) begin_action(W, e);
  else if (cmd == 2) { RO_RETRY(W, e) = 0; rollout_reset(W, e); }
}
// ... and after them: bit0 ready for a command, bit1/bit2 the events, bit3 unsafe at 'done' (RobotEnv._done, push_env.py:719-721)
__global__ void k_async_status(const __grid_constant__ DWorld W, uint8_t* status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= W.B) return;
  status[e] = (uint8_t)((W.phase[e] == B2S_PHASE_IDLE ? 1 : 0) | (W.async_events[e] & 6) | (W.phase_state[(size_t)e * 8 + 4] ? 8 : 0));
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

// The deal of a FREE-RUNNING launch (MODE_ENV; b2s_rollout_run / b2s_env_async_step_free).  Nobody waits for the slowest
// block there, so a block's cost need not match the others' -- what counts is that the warps of ONE block finish their
// stages together: environments are ranked by a running mean of their solver work and dealt in consecutive runs, the
// expensive blocks make fewer rounds with every warp on a long solve, the cheap ones many short rounds.
// Only blocks that are resident together take part (run_blocks <= SM count): a block that started after the launch's
// total was reached would stop after one round and its environments would starve.  When the world needs more blocks
// than that (large scenes: fewer environments fit a block), every launch steps a window of the environments that
// advances round-robin over the env index, so all of them get the same share of launches.
__global__ void __launch_bounds__(1024) k_assign_envs_free(const __grid_constant__ DWorld W, int total_blocks, int run_blocks, int per,
                                                           int free_chunk) {
  __shared__ int hist[256];
  __shared__ int base[256];
  __shared__ int s_stepping;
  const int E = W.envs_per_block;
  const int capacity = run_blocks * per;
  const int offset = (int)(W.free_target[1] % (unsigned long long)W.B);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < total_blocks * E; i += blockDim.x) W.env_map[i] = -1;
  __syncthreads();
  for (int e = threadIdx.x; e < W.B; e += blockDim.x) {
    const int rot = (e - offset + W.B) % W.B;
    if (rot < capacity && W.phase[e] != B2S_PHASE_IDLE) atomicAdd(&hist[254 - min(254, (int)W.work_ema[e])], 1);   // bin 0 = most expensive
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < 256; ++i) { base[i] = acc; acc += hist[i]; }
    s_stepping = acc;
    W.free_target[0] = *W.substeps + (unsigned long long)acc * (unsigned long long)free_chunk;
    W.free_target[1] = (unsigned long long)((offset + (capacity < W.B ? capacity : 0)) % W.B);
  }
  __syncthreads();
  // consecutive runs of the ranking, of equal length (+-1) over all blocks that run: block b holds the ranks
  // [ceil(b S / nb), ceil((b + 1) S / nb)) -- 4096 environments on 148 SMs are 27 or 28 per block, and few running
  // environments (the end of a rollout) end up one per block
  const int S = s_stepping, nb = max(1, min(run_blocks, S));
  for (int e = threadIdx.x; e < W.B; e += blockDim.x) {
    const int rot = (e - offset + W.B) % W.B;
    if (!(rot < capacity && W.phase[e] != B2S_PHASE_IDLE)) continue;
    const int p = atomicAdd(&base[254 - min(254, (int)W.work_ema[e])], 1);
#ifdef B2S_DEAL_STRIPED
    W.env_map[(size_t)(p % nb) * E + (p / nb)] = e;         // ranks dealt round-robin: every block gets its share of the expensive ones
#else
    const int b = (int)(((long long)p * nb) / S);
    const int first = (int)(((long long)b * S + nb - 1) / nb);
    W.env_map[(size_t)b * E + (p - first)] = e;
#endif
  }
}

// number of environments whose action / episode is still in flight (what b2s_env_substeps & co. report to the host)
__global__ void k_count_running(const __grid_constant__ DWorld W) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = e < W.B && W.phase[e] != B2S_PHASE_IDLE;
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(W.unfinished, __popc(m));
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

int b2s_launch_assign_envs(const DWorld& W, int mode, cudaStream_t s, int free_chunk, int resident_blocks) {
  if (free_chunk > 0) {
    const int run_blocks = W.num_blocks < resident_blocks ? W.num_blocks : resident_blocks;
    const int per = (W.B + W.num_blocks - 1) / W.num_blocks;
    k_assign_envs_free<<<1, 1024, 0, s>>>(W, W.num_blocks, run_blocks, per, free_chunk);
    return run_blocks;
  }
  k_assign_envs<<<1, 1024, 0, s>>>(W, mode, W.num_blocks);
  return W.num_blocks;
}
void b2s_launch_count_running(const DWorld& W, cudaStream_t s) {
  cudaMemsetAsync(W.unfinished, 0, sizeof(int), s);
  k_count_running<<<blocks_for(W.B, 256), 256, 0, s>>>(W);
}
void b2s_launch_reset(const DWorld& W, const uint8_t* mask, uint64_t seed, cudaStream_t s) { k_reset<<<blocks_for(W.B, 64), 64, 0, s>>>(W, mask, seed); }
void b2s_launch_set_action(const DWorld& W, cudaStream_t s) { k_set_action<<<blocks_for(W.B, 128), 128, 0, s>>>(W); }
void b2s_launch_begin_episode(const DWorld& W, const uint8_t* mask, cudaStream_t s) { k_begin_episode<<<blocks_for(W.B, 128), 128, 0, s>>>(W, mask); }
void b2s_launch_rollout_begin(const DWorld& W, const float* first_action, cudaStream_t s) {
  k_rollout_begin<<<blocks_for(W.B * 32, 128), 128, 0, s>>>(W, first_action);
}
void b2s_launch_async_commands(const DWorld& W, const uint8_t* c, cudaStream_t s) { k_async_commands<<<blocks_for(W.B, 128), 128, 0, s>>>(W, c); }
void b2s_launch_async_status(const DWorld& W, uint8_t* st, cudaStream_t s) { k_async_status<<<blocks_for(W.B, 128), 128, 0, s>>>(W, st); }
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
