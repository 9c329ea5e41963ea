// b2s_step.cu -- the per-substep hot path, one environment per warp.
//
// Replaces, for B environments at once, what the reference does per substep in
// Python + pybullet: Simulator.step (simulator.py:94-103),
// ControllableBody.update (controllable_body.py:387-413), pybullet.stepSimulation
// (bullet_physics.py:106-109: gravity, AABB broad phase, GJK/EPA narrow phase,
// persistent manifolds, PGS contact/friction solve, integration) and the PushEnv
// phase machine (push_env.py:631-937).
//
// Design: one environment is stepped by one warp at a time.  A block of 16 warps owns up to E environments (dealt
// to the blocks before every launch, k_assign_envs) and advances them in lock-step rounds of three stages
// separated by block barriers -- scene (controller, FK, body table, AABBs, broad phase), narrow phase, solve +
// integration + phase machine -- with dynamic hand-out of the work inside a stage: environments in the first and
// the last stage, single candidate PAIRS in the narrow phase (a warp takes UNITS_PER_WARP pairs per grab and works
// them side by side in units of UW lanes, b2s_dev.cuh).  All per-substep intermediates (body table,
// collider AABBs, pair list, contact list) live in the block's shared memory; only the persistent state (13
// floats per movable, 14 per arm, the <=4-point manifolds) goes back to HBM/L2.  Lanes split the work inside a
// unit: one hull vertex per lane in the GJK/EPA support function (exact max via redux on order-preserving
// keys), one collider per lane for AABBs, ballot-compacted pair lists, one contact per lane for Jacobian rows,
// colouring and the Gauss-Seidel sweeps (rows in registers, the body's velocity handed from lane to lane with
// shuffles; one BODY per lane when movables touch each other), which run the oracle's colour order exactly.  The
// residual test is a warp reduce.  All fp32 arithmetic is ordered exactly as in oracle/ (no FMA contraction), so
// every output matches bit for bit.  Why stages and barriers at all when the environments never interact: the
// kernel is ~10x the instruction cache, see the note above k_substeps.
// Episodes: in MODE_ENV an environment that finishes an action can go straight on (b2s_rollout.cuh: reward, record,
// next action from the device policy, or scene reset + drop + settle and the next episode), or wait for the host
// (asynchronous stepping).  A launch either gives every environment exactly n substeps, or is free-running: it ends
// when it has executed a total of substeps, all blocks stopping within one round of each other.
#include <mutex>

#include "b2s_dev.cuh"
#include "b2s_rollout.cuh"

#define LD3(p) v3((p)[0], (p)[1], (p)[2])
#define ST3(p, v) { (p)[0] = (v).x; (p)[1] = (v).y; (p)[2] = (v).z; }

__device__ __forceinline__ M3 ldm3(const float* p) { M3 m; m.r0 = LD3(p); m.r1 = LD3(p + 3); m.r2 = LD3(p + 6); return m; }
__device__ __forceinline__ void stm3(float* p, M3 m) { ST3(p, m.r0); ST3(p + 3, m.r1); ST3(p + 6, m.r2); }

// body record offsets
#define BO_POS 0
#define BO_R 3
#define BO_VEL 12
#define BO_ANG 15
#define BO_INVM 18
#define BO_INVI 19
#define BO_FRIC 28
#define BO_TYPE 29
#define BO_QUAT 30
// collider record offsets
#define CO_HULL 0
#define CO_SLOT 1
#define CO_TYPE 2
#define CO_SCALE 3
#define CO_MARGIN 4
#define CO_RAD 5
#define CO_AMIN 6
#define CO_AMAX 9

extern __shared__ float b2s_smem[];

// The world description lives in constant memory (uploaded before every launch): device functions read
// it through the constant bank instead of chasing a generic pointer to the kernel parameter block.
__constant__ DWorld g_W;
#define W g_W

// Stage timing for tuning (compiled only with -DB2S_PROF; see tools/profile_step.py): W.prof[0..2] = duration of
// stage A/B/C summed over blocks and rounds (thread 0, barrier to barrier), [3..5] = busy time of the warps inside
// the stage summed over warps, [6] = rounds, [7] = longest single environment of stage C summed over rounds.
#ifdef B2S_PROF
__device__ __forceinline__ long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; }
#define PROF_STAGE(i) { const long long now_ = prof_now(); if (lane == 0) atomicAdd(W.prof + 3 + (i), (unsigned long long)(now_ - pstart_)); }
#define PROF_MARK(i) { const long long now_ = prof_now(); if (threadIdx.x == 0) { atomicAdd(W.prof + (i), (unsigned long long)(now_ - pstart_)); W.prof[8 + blockIdx.x * 4 + (i)] += (unsigned long long)(now_ - pstart_); } pstart_ = now_; }
#define PROF_SEC(i) { const long long now_ = prof_now(); if (lane == 0) atomicAdd(W.prof + 8 + 4096 + (i), (unsigned long long)(now_ - psec_)); psec_ = now_; }
#define PROF_SEC0() long long psec_ = prof_now();
#else
#define PROF_STAGE(i)
#define PROF_MARK(i)
#define PROF_SEC(i)
#define PROF_SEC0()
#endif

struct Xf { V3 p; Q4 q; };
__device__ __forceinline__ Xf xf_from(const float* a) { Xf t; t.p = v3(a[0], a[1], a[2]); t.q = q4(a[3], a[4], a[5], a[6]); return t; }
__device__ __forceinline__ Xf xf_mul(Xf a, Xf b) { Xf t; t.p = a.p + qrot(a.q, b.p); t.q = qmul(a.q, b.q); return t; }
__device__ __forceinline__ void xf_store(Xf t, float* o) { o[0] = t.p.x; o[1] = t.p.y; o[2] = t.p.z; o[3] = t.q.x; o[4] = t.q.y; o[5] = t.q.z; o[6] = t.q.w; }

struct WarpSmem {
  float* body; float* col; int* pairs; int* cmk; float* con; float* stage; float* fk; float* sx;
};

// `sw` packs (environment slot of the block) | (warp in block) << 16
__device__ __forceinline__ int* env_meta(int slot) {
  return (int*)(b2s_smem + (size_t)W.envs_per_block * W.sm.words_env + (size_t)W.P.warps_per_block * W.sm.words_warp +
                (size_t)slot * META_WORDS);
}
__device__ __forceinline__ WarpSmem carve(int sw) {
  // both bases derive from the __shared__ array, so accesses compile to LDS/STS, not generic LD/ST
  float* eb = b2s_smem + (size_t)(sw & 0xffff) * W.sm.words_env;
  float* wb = b2s_smem + (size_t)W.envs_per_block * W.sm.words_env + (size_t)(sw >> 16) * W.sm.words_warp;
  WarpSmem s;
  s.body = eb + W.sm.body; s.col = eb + W.sm.col; s.pairs = (int*)(eb + W.sm.pairs); s.cmk = (int*)(eb + W.sm.cmk);
  s.con = wb + W.sm.con;
  s.stage = wb + W.sm.stage; s.fk = wb + W.sm.fk; s.sx = wb + W.sm.simplex;
  return s;
}

__device__ __forceinline__ M3 inv_inertia_world(M3 R, V3 d) {
  V3 a0 = vmul(R.r0, d), a1 = vmul(R.r1, d), a2 = vmul(R.r2, d);
  M3 m;
  m.r0 = v3(dot(a0, R.r0), dot(a0, R.r1), dot(a0, R.r2));
  m.r1 = v3(dot(a1, R.r0), dot(a1, R.r1), dot(a1, R.r2));
  m.r2 = v3(dot(a2, R.r0), dot(a2, R.r1), dot(a2, R.r2));
  return m;
}

// ----------------------------------------------------------------- arm ------

// serial chain (uniform over the warp): frames after each joint, world joint axes and origins, written
// to the warp's shared FK area  fk[0..49) frames 7x7, fk[49..70) axes, fk[70..91) origins
__device__ __noinline__ void fk_chain(const DArm* __restrict__ arm, const float* q, int lane, int wib) {
  float* fk = carve(wib).fk;
  Xf T = xf_from(arm->base);
  __syncwarp();
#pragma unroll 1
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
    Xf Tj = xf_mul(T, xf_from(arm->joint_origin[j]));
    V3 a = v3(arm->joint_axis[j][0], arm->joint_axis[j][1], arm->joint_axis[j][2]);
    V3 axw = qrot(Tj.q, a);
    T.p = Tj.p;
    T.q = qmul(Tj.q, q_axis_angle(a, q[j]));
    if (lane == 0) { xf_store(T, fk + j * 7); ST3(fk + 49 + j * 3, axw); ST3(fk + 70 + j * 3, Tj.p); }
  }
  __syncwarp();
}

// damped-least-squares IK, uniform over the warp (every lane computes the same values)
__device__ __noinline__ void arm_ik(const float* target, const float* q_start, float* q_out, int lane, int wib) {
  const float* fk = carve(wib).fk;
  const DArm* arm = W.arm;
  const B2SParams& P = W.P;
  float q[B2S_NUM_JOINTS];
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q[j] = q_start[j];
  Xf tgt = xf_from(target);
  Xf eel = xf_from(arm->ee);
  const float res2 = P.ik_residual * P.ik_residual;
  for (int it = 0; it < P.ik_max_iters; ++it) {
    fk_chain(arm, q, lane, wib);
    Xf ee = xf_mul(xf_from(fk + 6 * 7), eel);
    V3 ep = tgt.p - ee.p;
    V3 er = q_to_rotvec(qmul(tgt.q, qconj(ee.q)));
    if (len2(ep) < res2 && len2(er) < res2) break;
    float J[6][B2S_NUM_JOINTS];
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
      V3 axj = LD3(fk + 49 + j * 3);
      V3 jv = cross(axj, ee.p - LD3(fk + 70 + j * 3));
      J[0][j] = jv.x; J[1][j] = jv.y; J[2][j] = jv.z;
      J[3][j] = axj.x; J[4][j] = axj.y; J[5][j] = axj.z;
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
      dq[j] = s;
      m = fmaxf(m, fabsf(s));
    }
    float k = (m > P.ik_max_step) ? (P.ik_max_step / m) : 1.0f;
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) q[j] = q[j] + dq[j] * k;
  }
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q_out[j] = fminf(arm->upper[j], fmaxf(arm->lower[j], q[j]));
}

// FK of all collision links + end effector into global link_poses/link_vel and (optionally) the
// shared body table.  Chain is uniform; links are spread over lanes.
__device__ __noinline__ void arm_fk_links(int e, int lane, const float* q, const float* qd, int wib) {
  const WarpSmem S_ = carve(wib);
  float* fk = S_.fk;
  float* body = S_.body;
  const DArm* arm = W.arm;
  fk_chain(arm, q, lane, wib);
  const int L = W.L;
  if (lane < L) {
    int k = lane;
    int jj = arm->link_joint[k];
    Xf base = (jj < 0) ? xf_from(arm->base) : xf_from(fk + jj * 7);
    Xf T = xf_mul(base, xf_from(arm->link_pose[k]));
    V3 v = v3(0, 0, 0), om = v3(0, 0, 0);
    for (int i = 0; i <= jj; ++i) {
      V3 a = LD3(fk + 49 + i * 3), o = LD3(fk + 70 + i * 3);
      v = v + cross(a, T.p - o) * qd[i];
      om = om + a * qd[i];
    }
    if (W.P.export_debug) {
      float* lp = W.link_poses + ((size_t)e * (L + 1) + k) * 7;
      xf_store(T, lp);
      float* lv = W.link_vel + ((size_t)e * L + k) * 6;
      lv[0] = v.x; lv[1] = v.y; lv[2] = v.z; lv[3] = om.x; lv[4] = om.y; lv[5] = om.z;
    }
    if (body) {
      float* b = body + (W.Ns + k) * BODY_STRIDE;
      ST3(b + BO_POS, T.p);
      stm3(b + BO_R, q_to_m3(T.q));
      ST3(b + BO_VEL, v); ST3(b + BO_ANG, om);
      b[BO_INVM] = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) b[BO_INVI + i] = 0.0f;
      b[BO_FRIC] = arm->friction;
      b[BO_TYPE] = __int_as_float(B2S_TYPE_KINEMATIC);
      b[BO_QUAT] = T.q.x; b[BO_QUAT + 1] = T.q.y; b[BO_QUAT + 2] = T.q.z; b[BO_QUAT + 3] = T.q.w;
    }
  } else if (lane == L && W.P.export_debug) {
    Xf T = xf_mul(xf_from(fk + 6 * 7), xf_from(arm->ee));
    xf_store(T, W.link_poses + ((size_t)e * (L + 1) + L) * 7);
  }
  __syncwarp();
}

__device__ __forceinline__ bool joints_reached(int e, int lane, const float* c, int f1, int f2) {
  if (!f1) return true;
  bool ok = true;
  if (lane < 7) {
    float q = W.buf.joint_state[(0 * 7 + lane) * W.B + e], qd = W.buf.joint_state[(1 * 7 + lane) * W.B + e];
    bool pr = fabsf(c[9 + lane] - q) < c[16];
    bool vr = f2 ? true : (fabsf(0.0f - qd) < c[17]);
    ok = pr && vr;
  }
  return __all_sync(FULL, ok);
}

// ControllableBody.update + POSITION_CONTROL motor (oracle/b2o_arm.cpp arm_update)
__device__ __noinline__ void stage_arm(int e, int lane, int wib) {
  const B2SParams& P = W.P;
  float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  int32_t* f = W.ctrl_flags + (size_t)e * 4;
  double* T = W.ctrl_time + (size_t)e * 5;
  int f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3];
  const int n = W.num_steps[e];
  const double now = P.time_step * (double)n;
  bool ik_updated = false;
  __syncwarp();
  if (f0 && n % P.check_done_interval == 0) { if (now >= T[1]) f0 = 0; }
  if (f0 && (n % P.ik_interval == 0 || !f1)) {
    float q[7], qo[7], tgt[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) { q[j] = W.buf.joint_state[(0 * 7 + j) * W.B + e]; tgt[j] = c[j]; }
    arm_ik(tgt, q, qo, lane, wib);
    double t0 = T[0], t1 = T[1];
    float th0 = c[7], th1 = c[8];
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 7; ++j) c[9 + j] = qo[j];
      c[16] = th0; c[17] = th1;
      T[2] = t0; T[3] = t1;
    }
    __syncwarp();
    f1 = 1; f2 = 0;
    ik_updated = true;
    if (joints_reached(e, lane, c, f1, f2)) f0 = 0;
  }
  if (f1 && (n % P.check_done_interval == 0 || ik_updated)) {
    bool done = (now >= T[3]);
    bool reached = joints_reached(e, lane, c, f1, f2);
    if (done || reached) f1 = 0;
  }
  if (f1) {
    if (lane < 7) { c[18 + lane] = c[9 + lane]; c[25 + lane] = 0.0f; }
    f3 = 1;
  }
  __syncwarp();
  const float dt = (float)P.time_step;
  {
    const int j = lane < 7 ? lane : 0;
    float q = W.buf.joint_state[(0 * 7 + j) * W.B + e], qd = W.buf.joint_state[(1 * 7 + j) * W.B + e];
    float v = 0.0f, ratio = 1.0f;
    if (f3 && lane < 7) {
      v = (P.position_gain * (c[18 + j] - q) / dt + qd) + P.velocity_gain * (c[25 + j] - qd);
      if (P.clamp_joint_velocity) {
        float vm = P.limb_velocity_ratio * W.arm->max_vel[j];
        float a = fabsf(v);
        if (a > vm) ratio = vm / a;
      }
    }
    // common scale = min over joints (ratios are positive: bit order == value order)
    float scale = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(ratio)));
    v = v * scale;
    if (f3 && lane < 7) {
      float qn = q + v * dt;
      if (qn > W.arm->upper[j]) v = (W.arm->upper[j] - q) / dt;
      if (qn < W.arm->lower[j]) v = (W.arm->lower[j] - q) / dt;
    }
    if (lane < 7) W.buf.joint_state[(1 * 7 + j) * W.B + e] = v;
  }
  if (lane == 0) { f[0] = f0; f[1] = f1; f[2] = f2; f[3] = f3; }
  __syncwarp();
}

// SawyerSim.is_limb_ready with its side effects (oracle arm_is_ready)
__device__ __noinline__ int arm_is_ready(int e, int lane) {
  float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  int32_t* f = W.ctrl_flags + (size_t)e * 4;
  double* T = W.ctrl_time + (size_t)e * 5;
  __syncwarp();
  int f0 = f[0], f1 = f[1], f2 = f[2];
  const double now = W.P.time_step * (double)W.num_steps[e];
  if (!f0 || now >= T[1]) f0 = 0;
  bool jd = (!f1) || (now >= T[3]) || joints_reached(e, lane, c, f1, f2);
  if (jd) f1 = 0;
  __syncwarp();
  if (lane == 0) { f[0] = f0; f[1] = f1; }
  __syncwarp();
  return (!f0 && !f1) ? 1 : 0;
}

__device__ void arm_set_link_target(int e, int lane, const float* pose) {
  float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  double* T = W.ctrl_time + (size_t)e * 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) c[k] = pose[k];
    c[7] = W.P.joint_pos_threshold; c[8] = W.P.joint_vel_threshold;
    double now = W.P.time_step * (double)W.num_steps[e];
    T[0] = now; T[1] = now + (double)W.P.limb_timeout;
    W.ctrl_flags[(size_t)e * 4 + 0] = 1;
    W.ctrl_flags[(size_t)e * 4 + 1] = 0;          // reset_targets() precedes every set
  }
  __syncwarp();
}
__device__ void arm_set_joint_target(int e, int lane, const float* q) {
  float* c = W.ctrl + (size_t)e * B2S_CTRL_FLOATS;
  double* T = W.ctrl_time + (size_t)e * 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) c[9 + k] = q[k];
    c[16] = W.P.joint_pos_threshold; c[17] = W.P.joint_vel_threshold;
    double now = W.P.time_step * (double)W.num_steps[e];
    T[2] = now; T[3] = now + (double)W.P.limb_timeout;
    W.ctrl_flags[(size_t)e * 4 + 0] = 0;
    W.ctrl_flags[(size_t)e * 4 + 1] = 1;
    W.ctrl_flags[(size_t)e * 4 + 2] = 0;
  }
  __syncwarp();
}

// -------------------------------------------------------------- support ----
// Narrow-phase units.  A unit (one candidate pair) is processed by UW consecutive lanes: 4 by default (B2S_HALF = 3,
// eight pairs per warp at once; -DB2S_HALF=0..2 for 32 / 16 / 8 lanes).  The support function spreads the hull's
// vertices over the unit's lanes; everything else in GJK/EPA is scalar work replicated on them.
// UL = lane within the unit, UB = first lane of the unit, UM = member mask of the unit, UH = unit index in the warp.
// unit macros (UW, UL, UB, UM, UH): b2s_dev.cuh


// pose of the collider's body is read from the shared body table through `b` (broadcast LDS): keeping pos/R
// in the struct made it 20 words, which the compiler parked in local memory around the noinline EPA calls
struct ColRef { const float* b; float scale, margin; int voff, vcnt; V3 cen; };
__device__ __forceinline__ V3 cr_pos(const ColRef& c) { return LD3(c.b + BO_POS); }
__device__ __forceinline__ M3 cr_R(const ColRef& c) { return ldm3(c.b + BO_R); }

__device__ __forceinline__ ColRef col_ref(const float* col, const float* body, int c) {
  const float* cr = col + c * COL_STRIDE;
  const float* b = body + __float_as_int(cr[CO_SLOT]) * BODY_STRIDE;
  ColRef r;
  r.b = b;
  r.scale = cr[CO_SCALE]; r.margin = cr[CO_MARGIN];
  const DHull* H = W.hulls + __float_as_int(cr[CO_HULL]);
  r.voff = H->voff; r.vcnt = H->vcnt;
  r.cen = (LD3(cr + CO_AMIN) + LD3(cr + CO_AMAX)) * 0.5f;
  return r;
}

// argmax_i v_i . d over the hull's vertices (first maximum), vertices dealt round-robin to the unit's lanes
__device__ __forceinline__ int support(const ColRef& c, V3 d, V3* p, int lane) {
  const M3 R = cr_R(c);
  V3 dl = mtmul(R, d);
  float best = 0.0f;
  int bi = 0x7fffffff;
  bool has = false;
  for (int i = UL; i < c.vcnt; i += UW) {
    float4 v = __ldg(W.verts + c.voff + i);
    float t = dot(v3(v.x, v.y, v.z), dl);
    if (!has || t > best) { best = t; bi = i; has = true; }
  }
  unsigned key = has ? f2ord(best + 0.0f) : 0u;
  unsigned m = __reduce_max_sync(UM, key);
  unsigned cand = (has && key == m) ? (unsigned)bi : 0x7fffffffu;
  int idx = (int)__reduce_min_sync(UM, cand);
  float4 v = __ldg(W.verts + c.voff + idx);
  *p = cr_pos(c) + mmul(R, v3(v.x, v.y, v.z) * c.scale);
  return idx;
}

// ------------------------------------------------------------------ GJK ----
// simplex in shared memory: w[4][3] a[4][3] b[4][3] ia[4] ib[4] n  (45 words)
#define SX_W 0
#define SX_A 12
#define SX_B 24
#define SX_IA 36
#define SX_IB 40
#define SX_N 44

// Device form of b2s_closest_simplex (include/b2s_geom.h): the same arithmetic and the same decisions, but the
// triangle routine exists ONCE (noinline, canonical vertex roles 0,1,2) and the callers remap its weights.  The
// header version inlines it five times with different index triples: 1570 SASS instructions (25 KB, most of the
// 32 KB instruction cache) against ~600 here, and the GJK loop is the hottest loop of the narrow phase.
struct TriRes { V3 v; float t0, t1, t2; int used; };
__device__ __noinline__ void closest_tri_dev(V3 a, V3 b, V3 c, TriRes* o) {
  b2s_simplex_result t;
  b2s_closest_triangle(a, b, c, 0, 1, 2, &t);
  o->v = t.v; o->t0 = t.bary[0]; o->t1 = t.bary[1]; o->t2 = t.bary[2]; o->used = t.used;
}
// writes the weights of the triangle (i0, i1, i2 are compile-time constants at every call site)
#define TRI_TO_RESULT(T, i0, i1, i2, R)                                                        \
  {                                                                                            \
    (R)->v = (T).v;                                                                            \
    (R)->bary[0] = (R)->bary[1] = (R)->bary[2] = (R)->bary[3] = 0.0f;                          \
    if ((T).used & 1) (R)->bary[i0] = (T).t0;                                                  \
    if ((T).used & 2) (R)->bary[i1] = (T).t1;                                                  \
    if ((T).used & 4) (R)->bary[i2] = (T).t2;                                                  \
    (R)->used = (((T).used & 1) << (i0)) | ((((T).used >> 1) & 1) << (i1)) | ((((T).used >> 2) & 1) << (i2)); \
  }
#ifdef B2S_SIMPLEX_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
void closest_simplex_dev(const V3* w, int n, b2s_simplex_result* r) {
  r->inside = 0;
  r->degenerate = 0;
  if (n == 1) {
    r->v = w[0];
    r->bary[0] = 1.0f; r->bary[1] = r->bary[2] = r->bary[3] = 0.0f;
    r->used = 1;
    return;
  }
  if (n == 2) { b2s_closest_segment(w[0], w[1], 0, 1, r); return; }
  TriRes t;
  if (n == 3) { closest_tri_dev(w[0], w[1], w[2], &t); TRI_TO_RESULT(t, 0, 1, 2, r); return; }
  const int o0 = b2s_origin_outside_plane(w[0], w[1], w[2], w[3]);
  const int o1 = b2s_origin_outside_plane(w[0], w[2], w[3], w[1]);
  const int o2 = b2s_origin_outside_plane(w[0], w[3], w[1], w[2]);
  const int o3 = b2s_origin_outside_plane(w[1], w[3], w[2], w[0]);
  if (o0 < 0 || o1 < 0 || o2 < 0 || o3 < 0) {
    r->degenerate = 1;
    closest_tri_dev(w[1], w[2], w[3], &t); TRI_TO_RESULT(t, 1, 2, 3, r);
    return;
  }
  if (!o0 && !o1 && !o2 && !o3) {
    r->inside = 1;
    r->v = v3(0.0f, 0.0f, 0.0f);
    r->bary[0] = r->bary[1] = r->bary[2] = r->bary[3] = 0.25f;
    r->used = 15;
    return;
  }
  float best = 3.0e38f;
  if (o0) { closest_tri_dev(w[0], w[1], w[2], &t); float d = len2(t.v); if (d < best) { best = d; TRI_TO_RESULT(t, 0, 1, 2, r); } }
  if (o1) { closest_tri_dev(w[0], w[2], w[3], &t); float d = len2(t.v); if (d < best) { best = d; TRI_TO_RESULT(t, 0, 2, 3, r); } }
  if (o2) { closest_tri_dev(w[0], w[3], w[1], &t); float d = len2(t.v); if (d < best) { best = d; TRI_TO_RESULT(t, 0, 3, 1, r); } }
  if (o3) { closest_tri_dev(w[1], w[3], w[2], &t); float d = len2(t.v); if (d < best) { best = d; TRI_TO_RESULT(t, 1, 3, 2, r); } }
}

// The simplex lives in shared memory in PHYSICAL slots that never move; `perm` (2 bits per entry) maps
// the logical order (= the oracle's compacted order) to physical slots and `ids` packs (ia | ib << 8)
// of the logical entries, so the duplicate test and the compaction are register-only integer work.
// world position of hull vertex i (the point `support` returns for that index)
__device__ __forceinline__ V3 vertex_world(const ColRef& c, int i) {
  const float4 v = __ldg(W.verts + c.voff + i);
  return cr_pos(c) + mmul(cr_R(c), v3(v.x, v.y, v.z) * c.scale);
}

// sx[SX_CACHE..+2] (in/out): the simplex of the pair's previous call as vertex indices -- n, then (ia | ib << 8) of
// the entries 0,1 and 2,3 packed 16 bits each.  Like Bullet's cached separating axis it warm-starts the descent: a
// resting pair confirms its closest features with one support query instead of rebuilding them (4.9 -> ~1.5
// iterations per pair on the bench scene).  n = 0: cold start.  (oracle/b2o_physics.cpp gjk)
#define SX_CACHE 45
__device__ __noinline__ int gjk(const ColRef& A_in, const ColRef& B_in, float limit, float* sx, V3* v_out, V3* pa,
                   V3* pb, int lane) {
  const ColRef A = A_in, B = B_in;      // register copies: the references point into the caller's local memory
  V3 v = A.cen - B.cen;
  if (len2(v) < 1e-12f) v = v3(1.0f, 0.0f, 0.0f);
  int n = 0;
  unsigned perm = 0;
  unsigned long long ids = 0ull;
  bool have = false;
  float bary0 = 0, bary1 = 0, bary2 = 0, bary3 = 0;
  int status = 1;
  __syncwarp(UM);
  const int cn = __float_as_int(sx[SX_CACHE]);
  bool warm = cn > 0;
  if (warm) {
    ids = (unsigned long long)(unsigned)__float_as_int(sx[SX_CACHE + 1]) | ((unsigned long long)(unsigned)__float_as_int(sx[SX_CACHE + 2]) << 32);
    __syncwarp(UM);
    if (UL < cn) {
      const unsigned id = (unsigned)((ids >> (16 * UL)) & 0xffffull);
      const int ia = (int)(id & 255u), ib = (int)(id >> 8);
      const V3 a = vertex_world(A, ia), b = vertex_world(B, ib);
      ST3(sx + SX_W + UL * 3, a - b); ST3(sx + SX_A + UL * 3, a); ST3(sx + SX_B + UL * 3, b);
      sx[SX_IA + UL] = __int_as_float(ia); sx[SX_IB + UL] = __int_as_float(ib);
    }
    n = cn;
    perm = 0xE4u;                        // logical entry k in physical slot k
  }
  __syncwarp(UM);
  if (UL == 0) { sx[SX_CACHE] = __int_as_float(0); sx[SX_CACHE + 1] = __int_as_float(0); sx[SX_CACHE + 2] = __int_as_float(0); }
  __syncwarp(UM);
  for (int it = 0; it < W.P.gjk_max_iters; ++it) {
    float vv = 0.0f;
    if (!warm) {
      V3 a, b;
      int ia = support(A, -v, &a, lane);
      int ib = support(B, v, &b, lane);
      V3 ww = a - b;
      vv = dot(v, v);
      const float vw = dot(v, ww);
      if (vw > 0.0f && vw * vw > (limit * limit) * vv) return 0;
      const unsigned id = (unsigned)ia | ((unsigned)ib << 8);
      bool dup = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) if (k < n && (unsigned)((ids >> (16 * k)) & 0xffffu) == id) dup = true;
      if (dup) break;
      if (have && (vv - vw) <= vv * 1e-6f) break;
      // lowest free physical slot
      unsigned usedp = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) if (k < n) usedp |= 1u << ((perm >> (2 * k)) & 3u);
      const int ps = __ffs(~usedp) - 1;
      __syncwarp(UM);
      if (UL == 0) {
        ST3(sx + SX_W + ps * 3, ww); ST3(sx + SX_A + ps * 3, a); ST3(sx + SX_B + ps * 3, b);
        sx[SX_IA + ps] = __int_as_float(ia); sx[SX_IB + ps] = __int_as_float(ib);
      }
      __syncwarp(UM);
      perm = (perm & ~(3u << (2 * n))) | ((unsigned)ps << (2 * n));
      ids = (ids & ~(0xffffull << (16 * n))) | ((unsigned long long)id << (16 * n));
      n = n + 1;
    }
    V3 wl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wl[k] = LD3(sx + SX_W + ((perm >> (2 * k)) & 3u) * 3);
    b2s_simplex_result r;
    closest_simplex_dev(wl, n, &r);
    if (r.inside) { status = 2; break; }
    // logical compaction (order preserving): registers only
    unsigned nperm = 0; unsigned long long nids = 0ull;
    float nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0;
    int m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < n && (r.used & (1 << k))) {
        nperm |= ((perm >> (2 * k)) & 3u) << (2 * m);
        nids |= ((ids >> (16 * k)) & 0xffffull) << (16 * m);
        float bk = r.bary[k];
        if (m == 0) nb0 = bk; else if (m == 1) nb1 = bk; else if (m == 2) nb2 = bk; else nb3 = bk;
        ++m;
      }
    }
    float nv = len2(r.v);
    // no progress (converged to rounding, or the new vertex made a flat simplex whose sub-simplex search lost ground):
    // the previous simplex stays the answer -- perm / ids / n / weights still describe it, its slots were not touched
    if (!warm && have && nv >= vv && nv >= 1e-14f) {
      n = n - 1;                                   // drop the vertex appended above (logical entry n)
      ids &= ~(0xffffull << (16 * n));
      perm &= ~(3u << (2 * n));
      break;
    }
    perm = nperm; ids = nids; n = m;
    bary0 = nb0; bary1 = nb1; bary2 = nb2; bary3 = nb3;
    if (nv < 1e-14f) { status = 2; break; }
    v = r.v;
    have = true;
    if (warm) { warm = false; --it; }    // the warm start is not one of the gjk_max_iters iterations
  }
  if (status == 2) {
    // EPA wants the simplex in logical order at slots 0..n-1
    float rec[11];
    const int src = (UL < 4) ? (int)((perm >> (2 * UL)) & 3u) : 0;
    __syncwarp(UM);
    if (UL < 4) {
#pragma unroll
      for (int t = 0; t < 3; ++t) { rec[t] = sx[SX_W + src * 3 + t]; rec[3 + t] = sx[SX_A + src * 3 + t]; rec[6 + t] = sx[SX_B + src * 3 + t]; }
      rec[9] = sx[SX_IA + src]; rec[10] = sx[SX_IB + src];
    }
    __syncwarp(UM);
    if (UL < 4) {
#pragma unroll
      for (int t = 0; t < 3; ++t) { sx[SX_W + UL * 3 + t] = rec[t]; sx[SX_A + UL * 3 + t] = rec[3 + t]; sx[SX_B + UL * 3 + t] = rec[6 + t]; }
      sx[SX_IA + UL] = rec[9]; sx[SX_IB + UL] = rec[10];
    }
    if (UL == 0) sx[SX_N] = __int_as_float(n);
    __syncwarp(UM);
    return 2;
  }
  if (!have) return 0;
  V3 xa = v3(0, 0, 0), xb = v3(0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < n) {
      const int ph = (perm >> (2 * k)) & 3u;
      float bk = (k == 0) ? bary0 : (k == 1) ? bary1 : (k == 2) ? bary2 : bary3;
      xa = xa + LD3(sx + SX_A + ph * 3) * bk;
      xb = xb + LD3(sx + SX_B + ph * 3) * bk;
    }
  }
  *pa = xa; *pb = xb; *v_out = v;
  __syncwarp(UM);
  if (UL == 0) {
    sx[SX_CACHE] = __int_as_float(n);
    sx[SX_CACHE + 1] = __int_as_float((int)(unsigned)(ids & 0xffffffffull));
    sx[SX_CACHE + 2] = __int_as_float((int)(unsigned)(ids >> 32));
  }
  __syncwarp(UM);
  return 1;
}

// ------------------------------------------------------------------ EPA ----
// scratch (aliases the contact-row area): W[48][3] PA[48][3] PB[48][3] IA[48] IB[48]
//   faces: idx[192] (i0 | i1<<8 | i2<<16 | alive<<24), n[192][3], d[192]; edges[192]; newfaces n/d staged in place
#define EP_W 0
#define EP_PA (EPA_MAXV * 3)
#define EP_PB (EPA_MAXV * 6)
#define EP_IA (EPA_MAXV * 9)
#define EP_IB (EPA_MAXV * 10)
#define EP_FI (EPA_MAXV * 11)
#define EP_FN (EP_FI + EPA_MAXF)
#define EP_FD (EP_FN + EPA_MAXF * 3)
#define EP_ED (EP_FD + EPA_MAXF)
#define EP_VIS (EP_ED + EPA_MAXF)
#define EP_WORDS (EP_VIS + EPA_MAXF)

__device__ __forceinline__ bool epa_face_plane(const float* ep, int i0, int i1, int i2, V3* n, float* d) {
  V3 p0 = LD3(ep + EP_W + i0 * 3);
  V3 nn = cross(LD3(ep + EP_W + i1 * 3) - p0, LD3(ep + EP_W + i2 * 3) - p0);
  float l2 = len2(nn);
  if (l2 < 1e-20f) return false;
  nn = nn * (1.0f / sqrtf(l2));
  *n = nn; *d = dot(nn, p0);
  return true;
}

// add a support point to the GJK simplex (uniform); returns false when it is already there
__device__ __noinline__ bool sx_add(const ColRef& A, const ColRef& B, float* sx, int* n, V3 d, int lane) {
  V3 a, b;
  int ia = support(A, d, &a, lane);
  int ib = support(B, -d, &b, lane);
  for (int k = 0; k < *n; ++k) if (__float_as_int(sx[SX_IA + k]) == ia && __float_as_int(sx[SX_IB + k]) == ib) return false;
  __syncwarp(UM);
  if (UL == 0) {
    int k = *n;
    ST3(sx + SX_W + k * 3, a - b); ST3(sx + SX_A + k * 3, a); ST3(sx + SX_B + k * 3, b);
    sx[SX_IA + k] = __int_as_float(ia); sx[SX_IB + k] = __int_as_float(ib);
  }
  __syncwarp(UM);
  *n = *n + 1;
  return true;
}

__device__ __noinline__ bool epa_complete(const ColRef& A, const ColRef& B, float* sx, int* np, int lane) {
  int n = *np;
  if (n == 1) {
    for (int k = 0; k < 6 && n == 1; ++k) {
      V3 d = v3(k == 0 ? 1.f : k == 1 ? -1.f : 0.f, k == 2 ? 1.f : k == 3 ? -1.f : 0.f, k == 4 ? 1.f : k == 5 ? -1.f : 0.f);
      sx_add(A, B, sx, &n, d, lane);
    }
    if (n == 1) { *np = n; return false; }
  }
  if (n == 2) {
    V3 d = LD3(sx + SX_W + 3) - LD3(sx + SX_W);
    for (int k = 0; k < 6 && n == 2; k += 2) {
      V3 axis = v3(k == 0 ? 1.f : 0.f, k == 2 ? 1.f : 0.f, k == 4 ? 1.f : 0.f);
      V3 dir = cross(d, axis);
      if (len2(dir) < 1e-12f * len2(d)) continue;
      if (!sx_add(A, B, sx, &n, dir, lane)) sx_add(A, B, sx, &n, -dir, lane);
      if (n == 3) {
        V3 w0 = LD3(sx + SX_W);
        V3 nn = cross(LD3(sx + SX_W + 3) - w0, LD3(sx + SX_W + 6) - w0);
        if (len2(nn) < 1e-20f) n = 2;
      }
    }
    if (n == 2) { *np = n; return false; }
  }
  if (n == 3) {
    V3 w0 = LD3(sx + SX_W);
    V3 nn = cross(LD3(sx + SX_W + 3) - w0, LD3(sx + SX_W + 6) - w0);
    if (len2(nn) < 1e-20f) { *np = n; return false; }
    if (!sx_add(A, B, sx, &n, nn, lane)) { if (!sx_add(A, B, sx, &n, -nn, lane)) { *np = n; return false; } }
    V3 e3 = LD3(sx + SX_W + 9) - w0;
    float vol = dot(e3, nn);
    if (vol * vol < 1e-12f * len2(nn) * len2(e3)) {
      n = 3;
      if (!sx_add(A, B, sx, &n, -nn, lane)) { *np = n; return false; }
      e3 = LD3(sx + SX_W + 9) - w0;
      vol = dot(e3, nn);
      if (vol * vol < 1e-12f * len2(nn) * len2(e3)) { *np = n; return false; }
    }
  }
  *np = n;
  return n == 4;
}

__device__ __noinline__ int epa(const ColRef& A, const ColRef& B, float* sx, float* ep, V3* n_out, float* depth,
                   V3* pa, V3* pb, int lane) {
  int sn = __float_as_int(sx[SX_N]);
  if (sn < 4 && !epa_complete(A, B, sx, &sn, lane)) return 0;
  __syncwarp(UM);
  if (UL < 4) {
    ST3(ep + EP_W + UL * 3, LD3(sx + SX_W + UL * 3));
    ST3(ep + EP_PA + UL * 3, LD3(sx + SX_A + UL * 3));
    ST3(ep + EP_PB + UL * 3, LD3(sx + SX_B + UL * 3));
    ep[EP_IA + UL] = sx[SX_IA + UL]; ep[EP_IB + UL] = sx[SX_IB + UL];
  }
  __syncwarp(UM);
  int nv = 4, nf = 4;
  bool bad = false;
  if (UL < 4) {
    const int t0 = (UL == 3) ? 1 : 0, t1 = (UL == 0) ? 1 : (UL == 1) ? 3 : (UL == 2) ? 2 : 3;
    const int t2 = (UL == 0) ? 2 : (UL == 1) ? 1 : (UL == 2) ? 3 : 2, t3 = (UL == 0) ? 3 : (UL == 1) ? 2 : (UL == 2) ? 1 : 0;
    V3 n; float d;
    int i0 = t0, i1 = t1, i2 = t2;
    if (!epa_face_plane(ep, i0, i1, i2, &n, &d)) bad = true;
    else {
      if (dot(n, LD3(ep + EP_W + t3 * 3)) - d > 0.0f) { int t = i1; i1 = i2; i2 = t; n = -n; d = -d; }
      ep[EP_FI + UL] = __int_as_float(i0 | (i1 << 8) | (i2 << 16) | (1 << 24));
      ST3(ep + EP_FN + UL * 3, n); ep[EP_FD + UL] = d;
    }
  }
  if (__any_sync(UM, bad)) return 0;
  __syncwarp(UM);
  int best = 0;
  for (int it = 0; it < W.P.epa_max_iters; ++it) {
    // closest alive face (first minimum)
    float bd = 3e38f; int bf = 0x7fffffff;
    for (int f = UL; f < nf; f += UW) {
      int fi = __float_as_int(ep[EP_FI + f]);
      float d = ep[EP_FD + f];
      if ((fi >> 24) && d < bd) { bd = d; bf = f; }
    }
    unsigned key = (bf != 0x7fffffff) ? f2ord(bd + 0.0f) : 0xffffffffu;
    unsigned mk = __reduce_min_sync(UM, key);
    if (mk == 0xffffffffu) return 0;
    unsigned cand = (key == mk && bf != 0x7fffffff) ? (unsigned)bf : 0x7fffffffu;
    best = (int)__reduce_min_sync(UM, cand);
    bd = ep[EP_FD + best];
    V3 n = LD3(ep + EP_FN + best * 3);
    V3 a, b;
    int ia = support(A, n, &a, lane);
    int ib = support(B, -n, &b, lane);
    V3 ww = a - b;
    float s = dot(ww, n);
    if (s - bd < 1e-6f) break;
    bool dup = false;
    for (int k = UL; k < nv; k += UW) if (__float_as_int(ep[EP_IA + k]) == ia && __float_as_int(ep[EP_IB + k]) == ib) dup = true;
    if (__any_sync(UM, dup) || nv >= EPA_MAXV) break;
    // visibility
    for (int f = UL; f < nf; f += UW) {
      int fi = __float_as_int(ep[EP_FI + f]);
      int vis = (fi >> 24) && (dot(LD3(ep + EP_FN + f * 3), ww) - ep[EP_FD + f] > 0.0f);
      ep[EP_VIS + f] = __int_as_float(vis);
    }
    __syncwarp(UM);
    // horizon (sequential, lane 0)
    int ne = 0;
    if (UL == 0) {
      for (int f = 0; f < nf; ++f) {
        if (!__float_as_int(ep[EP_VIS + f])) continue;
        int fi = __float_as_int(ep[EP_FI + f]);
        int i0 = fi & 255, i1 = (fi >> 8) & 255, i2 = (fi >> 16) & 255;
        int e0[3] = {i0, i1, i2}, e1[3] = {i1, i2, i0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          int rev = e1[k] | (e0[k] << 8);
          int found = -1;
          for (int j = 0; j < ne; ++j) if (__float_as_int(ep[EP_ED + j]) == rev) { found = j; break; }
          if (found >= 0) { for (int j = found; j + 1 < ne; ++j) ep[EP_ED + j] = ep[EP_ED + j + 1]; --ne; }
          else if (ne < EPA_MAXF) { ep[EP_ED + ne] = __int_as_float(e0[k] | (e1[k] << 8)); ++ne; }
        }
      }
    }
    ne = __shfl_sync(UM, ne, 0, UW);
    __syncwarp(UM);
    if (ne < 3 || nf + ne > EPA_MAXF) break;
    if (UL == 0) {
      ST3(ep + EP_W + nv * 3, ww); ST3(ep + EP_PA + nv * 3, a); ST3(ep + EP_PB + nv * 3, b);
      ep[EP_IA + nv] = __int_as_float(ia); ep[EP_IB + nv] = __int_as_float(ib);
    }
    __syncwarp(UM);
    // new faces: one horizon edge per lane; staged after the current faces, committed only if all are sound
    bool ok = true;
    for (int j = UL; j < ne; j += UW) {
      int ed = __float_as_int(ep[EP_ED + j]);
      int u = ed & 255, vtx = (ed >> 8) & 255;
      V3 fn; float fd;
      if (!epa_face_plane(ep, u, vtx, nv, &fn, &fd)) ok = false;
      else {
        ep[EP_FI + nf + j] = __int_as_float(u | (vtx << 8) | (nv << 16) | (1 << 24));
        ST3(ep + EP_FN + (nf + j) * 3, fn); ep[EP_FD + nf + j] = fd;
      }
    }
    if (!__all_sync(UM, ok)) break;
    for (int f = UL; f < nf; f += UW)
      if (__float_as_int(ep[EP_VIS + f])) ep[EP_FI + f] = __int_as_float(__float_as_int(ep[EP_FI + f]) & 0x00ffffff);
    __syncwarp(UM);
    nf += ne;
    ++nv;
  }
  __syncwarp(UM);
  int fi = __float_as_int(ep[EP_FI + best]);
  int i0 = fi & 255, i1 = (fi >> 8) & 255, i2 = (fi >> 16) & 255;
  V3 fn = LD3(ep + EP_FN + best * 3);
  float fd = ep[EP_FD + best];
  V3 p = fn * fd;
  b2s_simplex_result r;
  b2s_closest_triangle(LD3(ep + EP_W + i0 * 3) - p, LD3(ep + EP_W + i1 * 3) - p, LD3(ep + EP_W + i2 * 3) - p, 0, 1, 2, &r);
  *pa = (LD3(ep + EP_PA + i0 * 3) * r.bary[0] + LD3(ep + EP_PA + i1 * 3) * r.bary[1]) + LD3(ep + EP_PA + i2 * 3) * r.bary[2];
  *pb = (LD3(ep + EP_PB + i0 * 3) * r.bary[0] + LD3(ep + EP_PB + i1 * 3) * r.bary[1]) + LD3(ep + EP_PB + i2 * 3) * r.bary[2];
  *n_out = fn;
  *depth = fd;
  __syncwarp(UM);
  return 1;
}

__device__ int collide_pair(const ColRef& A, const ColRef& B, float threshold, float* sx, float* ep,
                            V3* pA, V3* pB, V3* normal, float* distance, int lane) {
  float msum = A.margin + B.margin;
  V3 v, pa, pb;
  int st = gjk(A, B, msum + threshold, sx, &v, &pa, &pb, lane);
  if (st == 0) return 0;
  V3 n;
  float dist;
  if (st == 1) {
    float l = len(v);
    n = v * (1.0f / l);
    dist = l - msum;
  } else {
    V3 no;
    float depth;
    if (!epa(A, B, sx, ep, &no, &depth, &pa, &pb, lane)) {
      V3 c = A.cen - B.cen;
      float l2 = len2(c);
      n = (l2 < 1e-12f) ? v3(0.0f, 0.0f, 1.0f) : c * (1.0f / sqrtf(l2));
      pa = A.cen; pb = pa;
      dist = -msum;
    } else {
      n = -no;
      dist = -depth - msum;
    }
  }
  if (!(dist < threshold)) return 0;
  *pA = pa - n * A.margin;
  *pB = pb + n * B.margin;
  *normal = n;
  *distance = dist;
  return 1;
}

// ------------------------------------------------------------ substep -------

#define BS(c, i) W.buf.body_state[((size_t)(c) * W.B + e) * W.Nmax + (i)]
#define MPAR(c, i) W.mov_params[((size_t)(c) * W.B + e) * W.Nmax + (i)]

// stages 1-6 of oracle substep(): controller, body table, colliders, broad phase, narrow phase, rows
// stage A: controller + FK, body table, collider AABBs, broad phase -> number of candidate pairs
__device__ __noinline__ int stage_scene(int e, int lane, int wib) {
  const WarpSmem S = carve(wib);
  const B2SParams& P = W.P;
  const float dt = (float)P.time_step;
  const int Ns = W.Ns, L = W.L, NB = W.NB, Nmax = W.Nmax;
  const int nm = W.buf.num_movables[e];
  PROF_SEC0()

  stage_arm(e, lane, wib);
  {
    float q[7], qd[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) { q[j] = W.buf.joint_state[(0 * 7 + j) * W.B + e]; qd[j] = W.buf.joint_state[(1 * 7 + j) * W.B + e]; }
    arm_fk_links(e, lane, q, qd, wib);
  }
  PROF_SEC(8)
  // body table: statics + movables (links were written by arm_fk_links)
  const V3 g = v3(P.gravity[0], P.gravity[1], P.gravity[2]);
  const float ld = fmaxf(0.0f, 1.0f - P.linear_damping * dt), ad = fmaxf(0.0f, 1.0f - P.angular_damping * dt);
  for (int s = lane; s < NB; s += 32) {
    float* b = S.body + s * BODY_STRIDE;
    if (s < Ns) {
      const float* sp = W.static_pose + s * 7;
      float dz = (W.static_flags[s] & B2S_STATIC_ON_TABLE) ? W.table_dz[e] : 0.0f;
      Q4 qq = q4(sp[3], sp[4], sp[5], sp[6]);
      ST3(b + BO_POS, v3(sp[0], sp[1], sp[2] + dz));
      stm3(b + BO_R, q_to_m3(qq));
      ST3(b + BO_VEL, v3(0, 0, 0)); ST3(b + BO_ANG, v3(0, 0, 0));
      b[BO_INVM] = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) b[BO_INVI + i] = 0.0f;
      b[BO_FRIC] = W.static_friction[s];
      b[BO_TYPE] = __int_as_float(B2S_TYPE_STATIC);
      b[BO_QUAT] = qq.x; b[BO_QUAT + 1] = qq.y; b[BO_QUAT + 2] = qq.z; b[BO_QUAT + 3] = qq.w;
    } else if (s >= Ns + L) {
      const int i = s - Ns - L;
      if (i >= nm) {
        ST3(b + BO_POS, v3(0, 0, 0));
        M3 I3; I3.r0 = v3(1, 0, 0); I3.r1 = v3(0, 1, 0); I3.r2 = v3(0, 0, 1);
        stm3(b + BO_R, I3);
        ST3(b + BO_VEL, v3(0, 0, 0)); ST3(b + BO_ANG, v3(0, 0, 0));
        b[BO_INVM] = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) b[BO_INVI + k] = 0.0f;
        b[BO_FRIC] = 0.0f; b[BO_TYPE] = __int_as_float(B2S_TYPE_STATIC);
        b[BO_QUAT] = 0; b[BO_QUAT + 1] = 0; b[BO_QUAT + 2] = 0; b[BO_QUAT + 3] = 1;
      } else {
        V3 pos = v3(BS(0, i), BS(1, i), BS(2, i));
        Q4 qq = q4(BS(3, i), BS(4, i), BS(5, i), BS(6, i));
        V3 v = v3(BS(7, i), BS(8, i), BS(9, i));
        V3 om = v3(BS(10, i), BS(11, i), BS(12, i));
        V3 vel = (v + g * dt) * ld;
        V3 ang = om * ad;
        int asset = __float_as_int(MPAR(0, i));
        float sc = MPAR(1, i), mass = MPAR(2, i);
        const DAsset* As = W.assets + asset;
        V3 h = v3(As->half[0], As->half[1], As->half[2]) * sc;
        float k3 = mass * (1.0f / 3.0f);
        V3 I = v3(k3 * (h.y * h.y + h.z * h.z), k3 * (h.x * h.x + h.z * h.z), k3 * (h.x * h.x + h.y * h.y));
        M3 R = q_to_m3(qq);
        ST3(b + BO_POS, pos); stm3(b + BO_R, R);
        ST3(b + BO_VEL, vel); ST3(b + BO_ANG, ang);
        b[BO_INVM] = 1.0f / mass;
        stm3(b + BO_INVI, inv_inertia_world(R, v3(1.0f / I.x, 1.0f / I.y, 1.0f / I.z)));
        b[BO_FRIC] = MPAR(3, i);
        b[BO_TYPE] = __int_as_float(B2S_TYPE_DYNAMIC);
        b[BO_QUAT] = qq.x; b[BO_QUAT + 1] = qq.y; b[BO_QUAT + 2] = qq.z; b[BO_QUAT + 3] = qq.w;
      }
    }
  }
  __syncwarp();

  PROF_SEC(9)
  // colliders + AABBs (one collider per lane)
  const int nc = W.ncol[e];
  int first_dyn = nc, arm0 = nc, arm1 = 0;
  for (int c0 = 0; c0 < nc; c0 += 32) {
    int c = c0 + lane;
    int lfd = nc, la0 = nc, la1 = 0;
    if (c < nc) {
      float* cr = S.col + c * COL_STRIDE;
      int slot = W.col_slot[(size_t)e * W.Hmax + c], hull = W.col_hull[(size_t)e * W.Hmax + c];
      const DHull* H = W.hulls + hull;
      const float* b = S.body + slot * BODY_STRIDE;
      int type = __float_as_int(b[BO_TYPE]);
      unsigned flags = (slot < Ns) ? W.static_flags[slot] : 0u;
      V3 pos = LD3(b + BO_POS);
      M3 R = ldm3(b + BO_R);
      float scale = (slot >= Ns + L) ? MPAR(1, slot - Ns - L) : 1.0f;
      float rad = H->rad * scale + H->margin;
      V3 cen = pos + mmul(R, v3(H->lc[0], H->lc[1], H->lc[2]) * scale);
      V3 hs = v3(H->lh[0], H->lh[1], H->lh[2]) * scale;
      float pad = H->margin + P.breaking_factor * rad;
      V3 ext = v3((fabsf(R.r0.x) * hs.x + fabsf(R.r0.y) * hs.y) + fabsf(R.r0.z) * hs.z + pad,
                  (fabsf(R.r1.x) * hs.x + fabsf(R.r1.y) * hs.y) + fabsf(R.r1.z) * hs.z + pad,
                  (fabsf(R.r2.x) * hs.x + fabsf(R.r2.y) * hs.y) + fabsf(R.r2.z) * hs.z + pad);
      cr[CO_HULL] = __int_as_float(hull); cr[CO_SLOT] = __int_as_float(slot);
      cr[CO_TYPE] = __int_as_float(type | ((int)flags << 8));
      cr[CO_SCALE] = scale; cr[CO_MARGIN] = H->margin; cr[CO_RAD] = rad;
      ST3(cr + CO_AMIN, cen - ext); ST3(cr + CO_AMAX, cen + ext);
      if (type == B2S_TYPE_DYNAMIC) lfd = c;
      if (type == B2S_TYPE_KINEMATIC) { la0 = c; la1 = c + 1; }
    }
    first_dyn = min(first_dyn, (int)__reduce_min_sync(FULL, (unsigned)lfd));
    arm0 = min(arm0, (int)__reduce_min_sync(FULL, (unsigned)la0));
    arm1 = max(arm1, (int)__reduce_max_sync(FULL, (unsigned)la1));
  }
  __syncwarp();

  PROF_SEC(10)
  // broad phase: sorted pair keys, ballot compaction keeps the sequential order
  int np = 0;
  bool pair_over = false;
  const unsigned lt = (1u << lane) - 1u;
  for (int a = arm0; a < arm1; ++a) {
    const float* ca = S.col + a * COL_STRIDE;
    V3 amin = LD3(ca + CO_AMIN), amax = LD3(ca + CO_AMAX);
    for (int b0 = 0; b0 < a; b0 += 32) {
      int b = b0 + lane;
      bool pred = false;
      if (b < a) {
        const float* cb = S.col + b * COL_STRIDE;
        int tf = __float_as_int(cb[CO_TYPE]);
        if ((tf & 255) == B2S_TYPE_STATIC && ((tf >> 8) & B2S_STATIC_IS_TABLE)) {
          V3 bmin = LD3(cb + CO_AMIN), bmax = LD3(cb + CO_AMAX);
          pred = amin.x <= bmax.x && bmin.x <= amax.x && amin.y <= bmax.y && bmin.y <= amax.y && amin.z <= bmax.z && bmin.z <= amax.z;
        }
      }
      unsigned m = __ballot_sync(FULL, pred);
      if (pred) { int pos = np + __popc(m & lt); if (pos < P.max_pairs) S.pairs[pos] = (a << 16) | b; }
      np += __popc(m);
    }
  }
  for (int a = first_dyn; a < nc; ++a) {
    const float* ca = S.col + a * COL_STRIDE;
    V3 amin = LD3(ca + CO_AMIN), amax = LD3(ca + CO_AMAX);
    int sa = __float_as_int(ca[CO_SLOT]);
    for (int b0 = 0; b0 < a; b0 += 32) {
      int b = b0 + lane;
      bool pred = false;
      if (b < a) {
        const float* cb = S.col + b * COL_STRIDE;
        if (__float_as_int(cb[CO_SLOT]) != sa) {
          V3 bmin = LD3(cb + CO_AMIN), bmax = LD3(cb + CO_AMAX);
          pred = amin.x <= bmax.x && bmin.x <= amax.x && amin.y <= bmax.y && bmin.y <= amax.y && amin.z <= bmax.z && bmin.z <= amax.z;
        }
      }
      unsigned m = __ballot_sync(FULL, pred);
      if (pred) { int pos = np + __popc(m & lt); if (pos < P.max_pairs) S.pairs[pos] = (a << 16) | b; }
      np += __popc(m);
    }
  }
  if (np > P.max_pairs) { pair_over = true; np = P.max_pairs; }
  __syncwarp();
  if (P.export_debug) for (int p = lane; p < np; p += 32) W.pair_keys[(size_t)e * P.max_pairs + p] = S.pairs[p];
  if (lane == 0) { W.num_pairs[e] = np; if (pair_over) W.error_flags[e] |= 1; }
  __syncwarp();
  PROF_SEC(11)
  (void)Nmax;
  return np;
}

// stage B: narrow phase + persistent manifolds (ping-pong buffers in HBM/L2) -> contact list.
// The unit of work is one candidate PAIR of one environment (any warp of the block takes any pair): refresh of the
// pair's cached manifold, GJK/EPA, manifold update.  The result goes to a staging record of the pair; the warp
// that later solves the environment (stage C) merges the records in pair order into the new side of the
// ping-pong buffers, which reproduces the oracle's sequential compaction exactly.  With environments as the
// unit, a block with few active environments -- the last 11 of the 24 launches of a bench step carry < 6 % of
// the environments -- spent 3-4 pairs' worth of latency in this stage; now it is one pair's worth.
#define PS_WORDS 68                      // staging record: the manifold (64 words, GJK cache in 13..15), point count at 64
// The records of the first sm.ps_cap pairs of an environment live in its shared-memory region (up to 4: the PushEnv
// scenes have 3.3 candidate pairs on average; more shared memory for them costs more in L1 than it saves), the rest in
// an L2-resident global array.  The record is written in stage B by whichever warp took the pair and read in stage C
// by the warp that solves the environment: through global memory every substep paid for that hand-over with ~850 B
// of stores per environment, most of this kernel's DRAM write traffic (the speed is the same: measured 21.5 M
// substeps/s either way).
__device__ __forceinline__ float* pair_stage(int sw, int p) {
  if (p < W.sm.ps_cap) return b2s_smem + (size_t)(sw & 0xffff) * W.sm.words_env + W.sm.pstage + p * PS_WORDS;
  return W.pair_stage + (((size_t)blockIdx.x * W.envs_per_block + (sw & 0xffff)) * W.P.max_pairs + p) * PS_WORDS;
}

__device__ __noinline__ void stage_narrow_pair(int e, int lane, int wib, int p) {
  const WarpSmem S = carve(wib);
  const B2SParams& P = W.P;
  const unsigned lt = (1u << UL) - 1u;
  const int M = P.max_manifolds;
  const int par = W.man_parity[e];
  const size_t obase = ((size_t)par * W.B + e) * M;
  const int old_n = W.num_manifolds[e];
  float* stg = S.stage + UH * 64;   // [4][16] of this unit
  float* sxu = S.sx + UH * 48;
  PROF_SEC0()
  float* epa_scr = W.epa_scratch + (((size_t)blockIdx.x * W.P.warps_per_block + (wib >> 16)) * UNITS_PER_WARP + UH) * EP_WORDS;
  const int key = S.pairs[p];
  const int a = key >> 16, b = key & 0xffff;
  ColRef A = col_ref(S.col, S.body, a), Bc = col_ref(S.col, S.body, b);
  const float* ca = S.col + a * COL_STRIDE;
  const float* cb = S.col + b * COL_STRIDE;
  const float threshold = P.breaking_factor * fminf(ca[CO_RAD], cb[CO_RAD]);
  // old manifold lookup
  int found = 0x7fffffff;
  for (int k = UL; k < old_n; k += UW) if (W.man_keys[obase + k] == key) found = min(found, k);
  found = (int)__reduce_min_sync(UM, (unsigned)found);
  int n = 0;
  __syncwarp(UM);
  if (found != 0x7fffffff) {
    n = W.man_npts[obase + found];
    const float* src = W.man_pts + (obase + found) * 4 * B2S_CP_FLOATS;
    for (int i = UL; i < n * B2S_CP_FLOATS; i += UW) {
      const float s0 = src[i];
      stg[i] = s0;
      if (i >= 13 && i < 16) sxu[SX_CACHE + i - 13] = s0;             // GJK simplex of the previous substep
    }
  } else if (UL < 3) sxu[SX_CACHE + UL] = __int_as_float(0);
  __syncwarp(UM);
  // refresh (one point per lane), compaction keeps the order
  {
    float pt[B2S_CP_FLOATS];
    bool keep = false;
    if (UL < n) {
#pragma unroll
      for (int t = 0; t < B2S_CP_FLOATS; ++t) pt[t] = stg[UL * B2S_CP_FLOATS + t];
      V3 wA = cr_pos(A) + mmul(cr_R(A), v3(pt[0], pt[1], pt[2]));
      V3 wB = cr_pos(Bc) + mmul(cr_R(Bc), v3(pt[3], pt[4], pt[5]));
      V3 nn = v3(pt[6], pt[7], pt[8]);
      float dist = dot(wA - wB, nn);
      if (!(dist > threshold)) {
        V3 proj = wA - nn * dist;
        V3 dd = wB - proj;
        if (!(len2(dd) > threshold * threshold)) { keep = true; pt[9] = dist; }
      }
    }
    const unsigned km = (__ballot_sync(UM, keep) >> UB) & (UM >> UB);
    __syncwarp(UM);
    if (keep) {
      int dst = __popc(km & lt);
#pragma unroll
      for (int t = 0; t < B2S_CP_FLOATS; ++t) stg[dst * B2S_CP_FLOATS + t] = pt[t];
    }
    n = __popc(km);
    __syncwarp(UM);
  }
  PROF_SEC(4)
  V3 pA, pB, nrm;
  float dist;
  const int hit_ = collide_pair(A, Bc, threshold, sxu, epa_scr, &pA, &pB, &nrm, &dist, lane);
  PROF_SEC(5)
  if (hit_) {
    V3 lA = mtmul(cr_R(A), pA - cr_pos(A));
    V3 lB = mtmul(cr_R(Bc), pB - cr_pos(Bc));
    // manifold_add (uniform decisions, lane 0 writes)
    int nearest = -1;
    float shortest = threshold * threshold;
    for (int k = 0; k < n; ++k) {
      V3 d = LD3(stg + k * B2S_CP_FLOATS) - lA;
      float dd = len2(d);
      if (dd < shortest) { shortest = dd; nearest = k; }
    }
    int idx; bool keepl = false;
    if (nearest >= 0) { idx = nearest; keepl = true; }
    else if (n < 4) { idx = n; n = n + 1; }
    else {
      V3 PP[4]; float DD[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { PP[k] = LD3(stg + k * B2S_CP_FLOATS); DD[k] = stg[k * B2S_CP_FLOATS + 9]; }
      idx = b2s_manifold_replace_index(PP, DD, lA, dist);
    }
    __syncwarp(UM);
    if (UL == 0) {
      float* pp = stg + idx * B2S_CP_FLOATS;
      pp[0] = lA.x; pp[1] = lA.y; pp[2] = lA.z; pp[3] = lB.x; pp[4] = lB.y; pp[5] = lB.z;
      pp[6] = nrm.x; pp[7] = nrm.y; pp[8] = nrm.z; pp[9] = dist;
      if (!keepl) { pp[10] = 0.0f; pp[11] = 0.0f; pp[12] = 0.0f; }
      pp[13] = 0.0f; pp[14] = 0.0f; pp[15] = 0.0f;
    }
    __syncwarp(UM);
  }
  float* dst = pair_stage(wib, p);
  if (n > 0) {
    for (int i = UL; i < n * B2S_CP_FLOATS; i += UW)
      dst[i] = (i >= 13 && i < 16) ? sxu[SX_CACHE + i - 13] : stg[i];
  }
  if (UL == 0) dst[64] = __int_as_float(n);
  __syncwarp(UM);
  PROF_SEC(6)
}

// merge of the staged pairs of one environment, in pair order (start of stage C, by the warp that solves it)
__device__ __noinline__ void stage_narrow_merge(int e, int lane, int wib, int np, int* nc_out, int* newn_out) {
  const WarpSmem S = carve(wib);
  const B2SParams& P = W.P;
  const int M = P.max_manifolds;
  const int par = W.man_parity[e];
  const size_t nbase = ((size_t)(par ^ 1) * W.B + e) * M;
  int newn = 0, ncon = 0, cflags = 0;
  bool man_over = false, con_over = false;
  PROF_SEC0()
  for (int p = 0; p < np; ++p) {
    const float* src = pair_stage(wib, p);
    const int n = __float_as_int(src[64]);
    if (n <= 0) continue;
    if (newn < M) {
      const int key = S.pairs[p];
      const float* ca = S.col + (key >> 16) * COL_STRIDE;
      const float* cb = S.col + (key & 0xffff) * COL_STRIDE;
      float* dst = W.man_pts + (nbase + newn) * 4 * B2S_CP_FLOATS;
      if (lane < n * B2S_CP_FLOATS) dst[lane] = src[lane];
      if (lane + 32 < n * B2S_CP_FLOATS) dst[lane + 32] = src[lane + 32];
      if (lane == 0) { W.man_keys[nbase + newn] = key; W.man_npts[nbase + newn] = n; }
      const int tA = __float_as_int(ca[CO_TYPE]) & 255, tfB = __float_as_int(cb[CO_TYPE]);
      const int tB = tfB & 255;
      if (tA == B2S_TYPE_KINEMATIC && ((tfB >> 8) & B2S_STATIC_IS_TABLE)) cflags |= 1;
      if (tA == B2S_TYPE_DYNAMIC && tB == B2S_TYPE_KINEMATIC) cflags |= 2;
      if (tA == B2S_TYPE_DYNAMIC || tB == B2S_TYPE_DYNAMIC) {
        if (lane < n) { if (ncon + lane < P.max_contacts) S.cmk[ncon + lane] = (newn << 2) | lane; }
        if (ncon + n > P.max_contacts) { con_over = true; ncon = P.max_contacts; } else ncon += n;
      }
      ++newn;
    } else man_over = true;
  }
  __syncwarp();
  if (lane == 0) {
    W.num_manifolds[e] = newn;
    W.man_parity[e] = par ^ 1;
    W.contact_flags[e] = cflags;
    int ef = (man_over ? 2 : 0) | (con_over ? 8 : 0);
    if (ef) W.error_flags[e] |= ef;
  }
  __syncwarp();
  PROF_SEC(7)
  *nc_out = ncon;
  *newn_out = newn;
}

// ---- a long solve does not hold its block (free-running launches) ------------------------------------------------
// The rounds of a block are separated by three block barriers (top of the round, end of stage A, end of stage B), and a
// Gauss-Seidel solve is one warp's sequential chain: a 50-iteration solve (5 % of the env-substeps, 80-150 us against
// a median of 18) used to be the end of its round, with the other 15 warps waiting at the barrier.  In a free-running
// launch nobody needs that environment's result: which environment gets how many substeps depends on the schedule
// anyway.  So a warp whose solve has run B2S_LONG_ITERS (20) iterations looks, after every further iteration, whether all
// other warps are waiting at the next barrier (arrival counter in shared memory); if so it takes its environment out of
// the hand-out (META_ACTIVE = 0), arrives at the barrier from where it stands -- a barrier counts arrivals, not program
// counters -- and goes on solving while the block starts its next round without that environment.  When the solve is
// done the warp finishes the environment's substep, passes the barriers that are left until the block is in a stage C
// again, takes up stage-C work there, and the environment rejoins at the top of the following round.
// Warp 0 never does this (thread 0 resets the hand-out counters after the barriers).  What an environment computes is
// unchanged; only the schedule is.  Used for large scenes (substep_post_big, k_substeps<true>: 1.12 -> 1.62 M substeps/s
// on the crossing config); with register-resident rows it measured neutral (the deal already groups the expensive
// environments) and its plumbing cost 1.2 %, so k_substeps<false> is the plain kernel.
#ifndef B2S_LONG_ITERS
#define B2S_LONG_ITERS 20
#endif
__shared__ int sh_arrived;   // barrier arrivals of this block since the start of the launch (free-running launches)
__shared__ int sh_long;      // warps that are in a solve past B2S_LONG_ITERS iterations
__shared__ int sh_stop;      // free-running launch: the launch has done its total of substeps, every block stops after its round

struct BarState {
  int g;          // block barriers this warp has passed in the round loop: g % 3 == 0 <=> the next one is a top-of-round barrier
  int enabled;    // 1: free-running launch with run-ahead, not warp 0, not a B2S_PROF build; 0: free-running (arrivals are counted); -1: exact launch
  int long_on;    // registered in sh_long
  int stopped;    // passed a top-of-round barrier at which the launch ended: finish the environment, then leave
  int slot;       // environment slot being solved (its META_ACTIVE is cleared when the warp first passes a barrier)
  int passed;     // barriers passed inside the current solve
};

// one barrier of the round loop from wherever the warp stands; `any` is what a top-of-round barrier reduces
__device__ __forceinline__ int bar_pass(BarState& bs, int lane, int any) {
  if (bs.enabled >= 0 && lane == 0) atomicAdd(&sh_arrived, 1);
  int r = 1;
  if (bs.g % 3 == 0) r = __syncthreads_or(any); else __syncthreads();
  bs.g += 1;
  return r;
}
// called after every iteration of a solve: lets the block go on if everybody else is waiting for this warp
__device__ __forceinline__ void long_solve_poll(BarState& bs, int lane, int it) {
  if (it + 1 < B2S_LONG_ITERS) return;                 // (first: `bs` lives in local memory, and short solves never read it)
  if (bs.enabled <= 0 || bs.stopped) return;
  const int Wn = blockDim.x >> 5;
  if (!bs.long_on) { if (lane == 0) atomicAdd(&sh_long, 1); bs.long_on = 1; __syncwarp(); }
  const int arrived = *(volatile int*)&sh_arrived - bs.g * Wn;
  const int lg = *(volatile int*)&sh_long;
  if (arrived < Wn - lg) return;
  if (!bs.passed) { if (lane == 0) env_meta(bs.slot)[META_ACTIVE] = 0; __syncwarp(); }
  const bool top = (bs.g % 3 == 0);
  bar_pass(bs, lane, 1);
  bs.passed += 1;
  if (lane == 0) atomicAdd(W.prof + 8 + 4096 + 15, 1ull);     // B2S_ARR_PROF[8 + 4096 + 15]: barriers passed from inside a solve
  if (top && *(volatile int*)&sh_stop) bs.stopped = 1;
}
__device__ __forceinline__ void long_solve_end(BarState& bs, int lane) {
  if (bs.long_on) { if (lane == 0) atomicSub(&sh_long, 1); bs.long_on = 0; __syncwarp(); }
}

// ---- body-centric solve, any number of contacts (max_movables <= 32) ---------------------------------------
// Same scheme as substep_post_reg below (read its header first) for scenes that do not fit one contact per lane
// and one body slot per lane: config #3 has 8 multi-hull movables on 18 tile bodies, ~120 contact points and ~30
// colours per environment.  Differences: rows are built 32 contacts at a time; lane i of the sweeps is MOVABLE i
// (the only dynamic bodies), not body slot i; lambda and the body indices of a contact live in its record
// (76 words in the warp's global scratch) instead of shared memory; the sequential greedy colouring reads the
// contact's bodies from a small shared array instead of shuffling them out of the contact lanes.  The row
// arithmetic, the colouring rule and the sweep order are the oracle's, so results are bit-identical.
// (This path replaced a contact-per-lane solver with all rows in shared memory: 50 KB per warp, i.e. 4 warps
// per SM on config #3, whose solve stage took 76 % of the substep.  Measured on config #3, 4096 envs x 50 substeps:
// 470 ms -> 305 ms; the sweeps are still 60 % of the warp time there (~1 500 cycles per colour step: the records of
// 16 warps, 580 KB, do not fit the L1 the block leaves; prefetching them three colours ahead changed nothing).)
#define RR_B 48                          // word offset of the B side of the rows in a contact record: per row angB, iangB,
                                         //   dir / mass of B (9 words)
#define RB_WORDS 124                     // record, in words: rows [0,48): row r = float4 (dir.xyz, angA.x) (angA.yz, iangA.xy)
#define RB_LAM 76                        //   (iangA.z, 1/d, d, c = bias + dir . vB + angB . wB of a body B that is not dynamic)
#define RB_LAMB 80                       //   (dir.xyz / mass of A, mu); B side [48,75); lambda [76,79), the B lane's lambda
#define RB_INFO 84                       //   [80,83); movable index of A | B << 8 at 84, bias of the normal row at 85;
#define RB_BIAS 85                       //   torsional rows from 88: float4 (1/d of the three rows, combined spinning coeff.),
#define RB_TQ 88                         //   float4 (axis . wB of the three rows, combined rolling coeff.; 0 = rows off),
#define RB_TL 96                         //   float4 impulses of the A lane, float4 impulses of the B lane, then -- only read
#define RB_TLB 100                       //   in a colour that joins two movables -- I_A^-1 axis [104,113) and I_B^-1 axis
#define RB_TIA 104                       //   [113,122): what the partner lane applies
#define RB_TIB 113

template <bool RA>
__device__ __noinline__ void substep_post_big(int e, int lane, int wib, int C, int newn, BarState* bsp) {
  const WarpSmem S = carve(wib);
  const B2SParams& P = W.P;
  const float dt = (float)P.time_step;
  const int Ns = W.Ns, L = W.L;
  const int nm = W.buf.num_movables[e];
  const int par = W.man_parity[e];          // already flipped: current buffer
  const size_t nbase = ((size_t)par * W.B + e) * P.max_manifolds;
  const int nrows = 1 + P.friction_dirs;
  const bool tors_world = P.rolling_friction > 0.0f;     // warp uniform; b2s_create checked friction_dirs == 2
  float* rr = W.row_scratch + ((size_t)blockIdx.x * W.P.warps_per_block + (wib >> 16)) * ((size_t)P.max_contacts * RB_WORDS);
  unsigned short* T = (unsigned short*)S.con;                 // [64 colours][32 movables]: contact | side << 14 | coupled << 15
  unsigned short* cinfo = (unsigned short*)(S.con + 1024);    // [C] movable of A (31 = none) | movable of B << 5 | dA << 10 | dB << 11
  unsigned char* ccol = (unsigned char*)(S.con + 1024 + (P.max_contacts + 1) / 2);   // [C] colour, 0xff = none
  const int mov0 = Ns + L;
  PROF_SEC0()
  __syncwarp();
  // rows, 32 contacts at a time
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    if (c < C) {
      const int mk = S.cmk[c];
      const int m = mk >> 2, k = mk & 3;
      const int key = W.man_keys[nbase + m];
      const int a = key >> 16, b = key & 0xffff;
      const int sA = __float_as_int(S.col[a * COL_STRIDE + CO_SLOT]), sB = __float_as_int(S.col[b * COL_STRIDE + CO_SLOT]);
      const float* bA = S.body + sA * BODY_STRIDE;
      const float* bB = S.body + sB * BODY_STRIDE;
      const float* p = W.man_pts + ((nbase + m) * 4 + k) * B2S_CP_FLOATS;
      V3 posA = LD3(bA + BO_POS), posB = LD3(bB + BO_POS);
      V3 wA = posA + mmul(ldm3(bA + BO_R), v3(p[0], p[1], p[2]));
      V3 wB = posB + mmul(ldm3(bB + BO_R), v3(p[3], p[4], p[5]));
      V3 n = v3(p[6], p[7], p[8]);
      V3 rA = wA - posA, rB = wB - posB;
      V3 t1, t2;
      plane_space(n, &t1, &t2);
      const V3 velB = LD3(bB + BO_VEL), angvB = LD3(bB + BO_ANG);
      if (P.friction_dirs == 1) {
        V3 rel = (LD3(bA + BO_VEL) + cross(LD3(bA + BO_ANG), rA)) - (velB + cross(angvB, rB));
        V3 lat = rel - n * dot(rel, n);
        float l2 = len2(lat);
        if (l2 > 1e-12f) t1 = lat * (1.0f / sqrtf(l2));
      }
      const float imA = bA[BO_INVM], imB = bB[BO_INVM];
      const bool dA = __float_as_int(bA[BO_TYPE]) == B2S_TYPE_DYNAMIC, dB = __float_as_int(bB[BO_TYPE]) == B2S_TYPE_DYNAMIC;
      const M3 iA = ldm3(bA + BO_INVI), iB = ldm3(bB + BO_INVI);
      float* rec1 = rr + (size_t)c * RB_WORDS;
      float4* rec = (float4*)rec1;
      const float pen = p[9] + P.linear_slop;
      const float bias0 = (pen > 0.0f) ? -(pen / dt) : -(pen * P.erp2 / dt);
      const float mu = bA[BO_FRIC] * bB[BO_FRIC];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const V3 dir = (r == 0) ? n : (r == 1 ? t1 : t2);
        const V3 angA = cross(rA, dir), angB = cross(rB, dir);
        const V3 iangA = mmul(iA, angA), iangB = mmul(iB, angB);
        const float d = ((imA + imB) + dot(iangA, angA)) + dot(iangB, angB);
        const float inv_d = (d > 0.0f && r < nrows) ? 1.0f / d : 0.0f;
        float l0 = (r == 0 || P.friction_dirs == 2) ? p[10 + r] * P.warmstart : 0.0f;
        if (r >= nrows) l0 = 0.0f;
        rec[r * 4 + 0] = make_float4(dir.x, dir.y, dir.z, angA.x);
        rec[r * 4 + 1] = make_float4(angA.y, angA.z, iangA.x, iangA.y);
        const V3 dirMA = dir * imA, dirMB = dir * imB;
        rec[r * 4 + 2] = make_float4(iangA.z, inv_d, d, ((r == 0) ? bias0 : 0.0f) + (dot(dir, velB) + dot(angB, angvB)));
        rec[r * 4 + 3] = make_float4(dirMA.x, dirMA.y, dirMA.z, mu);
        ST3(rec1 + RR_B + r * 9, angB); ST3(rec1 + RR_B + r * 9 + 3, iangB); ST3(rec1 + RR_B + r * 9 + 6, dirMB);
        rec1[RB_LAM + r] = l0; rec1[RB_LAMB + r] = l0;
      }
      rec1[RB_BIAS] = bias0;
      {
        float mu_s = 0.0f, mu_r = 0.0f;
        if (tors_world) {
          const float rollB = dB ? P.rolling_friction : 0.0f, spinB = dB ? P.spinning_friction : 0.0f;
          mu_s = fminf(10.0f, P.spinning_friction * bB[BO_FRIC] + spinB * bA[BO_FRIC]);
          mu_r = fminf(10.0f, P.rolling_friction * bB[BO_FRIC] + rollB * bA[BO_FRIC]);
          const bool tors = mu_r > 0.0f;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const V3 ax = (r == 0) ? n : (r == 1 ? t1 : t2);
            const V3 tiA = mmul(iA, ax), tiB = mmul(iB, ax);
            const float d = dot(tiA, ax) + dot(tiB, ax);
            ST3(rec1 + RB_TIA + r * 3, tiA); ST3(rec1 + RB_TIB + r * 3, tiB);
            rec1[RB_TQ + r] = (d > 0.0f && tors) ? 1.0f / d : 0.0f;
            rec1[RB_TQ + 4 + r] = dot(ax, angvB);
          }
          if (!tors) mu_r = 0.0f;
        }
        rec1[RB_TQ + 3] = mu_s; rec1[RB_TQ + 7] = mu_r;
        rec[RB_TL / 4] = make_float4(0, 0, 0, 0); rec[RB_TLB / 4] = make_float4(0, 0, 0, 0);
      }
      const int iA_ = dA ? sA - mov0 : 31, iB_ = dB ? sB - mov0 : 31;
      rec1[RB_INFO] = __int_as_float(iA_ | (iB_ << 8));
      cinfo[c] = (unsigned short)(iA_ | (iB_ << 5) | ((int)dA << 10) | ((int)dB << 11));
      if (dB && !dA) W.error_flags[e] |= 64;     // cannot happen: movable colliders are numbered last
    }
  }
  __syncwarp();
  PROF_SEC(0)
  // greedy colouring in contact order; lane i keeps the colour mask of movable i
  unsigned long long used = 0ull;
  int ncolours = 0;
  bool col_over = false;
  for (int i = 0; i < C; ++i) {
    const unsigned info = cinfo[i];
    const int iA_ = info & 31, iB_ = (info >> 5) & 31;
    const bool idA = (info >> 10) & 1, idB = (info >> 11) & 1;
    const unsigned long long mA = __shfl_sync(FULL, used, iA_), mB = __shfl_sync(FULL, used, iB_);
    const unsigned long long mask = (idA ? mA : 0ull) | (idB ? mB : 0ull);
    if (mask == ~0ull) { col_over = true; if (lane == 0) ccol[i] = 0xff; continue; }
    const int k = __ffsll((long long)~mask) - 1;
    if (lane == 0) ccol[i] = (unsigned char)k;
    if (k + 1 > ncolours) ncolours = k + 1;
    if ((idA && lane == iA_) || (idB && lane == iB_)) used |= 1ull << k;
  }
  if (col_over && lane == 0) W.error_flags[e] |= 16;
  for (int i = lane; i < ncolours * 16; i += 32) ((unsigned*)T)[i] = 0xffffffffu;
  __syncwarp();
  unsigned long long coupled = 0ull;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    unsigned cplk = 0xffu;
    if (c < C && ccol[c] != 0xff) {
      const unsigned info = cinfo[c];
      const int k = ccol[c];
      const bool dA = (info >> 10) & 1, dB = (info >> 11) & 1;
      const unsigned cpl = (dA && dB) ? 0x8000u : 0u;
      if (dA) T[k * 32 + (info & 31)] = (unsigned short)(c | cpl);
      if (dB) T[k * 32 + ((info >> 5) & 31)] = (unsigned short)(c | 0x4000u | cpl);
      if (cpl) cplk = k;
    }
    // colours that hold a contact between two dynamic bodies (warp-uniform mask)
    for (unsigned mm = __ballot_sync(FULL, cplk != 0xffu); mm; mm &= mm - 1) coupled |= 1ull << __shfl_sync(FULL, cplk, __ffs(mm) - 1);
  }
  __syncwarp();
  PROF_SEC(1)
  // ---- sweeps: lane = movable index
  const bool dyn = lane < nm;
  float* myb = S.body + (mov0 + (dyn ? lane : 0)) * BODY_STRIDE;
  V3 vel = LD3(myb + BO_VEL), ang = LD3(myb + BO_ANG);
  float maxres = 0.0f;
  int iters = 0;
  // torsional rows of a contact, right after its friction rows (PASS 2): angular velocities only.  The lane keeps the
  // inverse inertia of its movable in registers and forms I^-1 axis itself (the oracle's mmul on the same operands); the
  // record supplies 1/d, the constant of a kinematic body B and the impulses as three 128-bit words.  In a colour with a
  // contact between two movables both lanes of that contact compute the same impulses and apply their own side (and the
  // partner's side from the record, to carry its velocity through the rows).
  const M3 myI = ldm3(myb + BO_INVI);
#define BIG_TORS_ROWS(SIDEB, CPL)                                                                                 \
  {                                                                                                               \
    float4* rect = (float4*)rec1;                                                                                 \
    const float tot = ml[0];                                                                                      \
    const float4 qd = rect[RB_TQ / 4], qk = rect[RB_TQ / 4 + 1];                                                  \
    if (qk.w > 0.0f && tot > 0.0f) {                                                                              \
      float4 tl4 = rect[((SIDEB) ? RB_TLB : RB_TL) / 4];                                                          \
      float tl[3] = {tl4.x, tl4.y, tl4.z};                                                                        \
      const float tinvd[3] = {qd.x, qd.y, qd.z}, tk[3] = {qk.x, qk.y, qk.z};                                      \
      V3 wA = (SIDEB) ? ow : ang, wB = (SIDEB) ? ang : ow;                                                        \
      _Pragma("unroll")                                                                                           \
      for (int r = 0; r < 3; ++r) {                                                                               \
        const float4 q0 = rect[r * 4];                                                                            \
        const V3 ax = v3(q0.x, q0.y, q0.z);                                                                       \
        const float mu_c = (r == 0) ? qd.w : qk.w;                                                                \
        const float tlim = fminf(mu_c * tot, mu_c);                                                               \
        float dl = (((CPL) ? dot(ax, wB) : tk[r]) - dot(ax, wA)) * tinvd[r];                                      \
        dl = fminf(tlim - tl[r], fmaxf((0.0f - tlim) - tl[r], dl));                                               \
        tl[r] = tl[r] + dl;                                                                                       \
        const V3 mine_i = mmul(myI, ax);                                                                          \
        if (!(SIDEB)) {                                                                                           \
          wA = vmad(wA, mine_i, dl);                                                                              \
          if (CPL) wB = vmad(wB, LD3(rec1 + RB_TIB + r * 3), -dl);                                                \
        } else {                                                                                                  \
          wB = vmad(wB, mine_i, -dl);                                                                             \
          wA = vmad(wA, LD3(rec1 + RB_TIA + r * 3), dl);                                                          \
        }                                                                                                         \
      }                                                                                                           \
      rect[((SIDEB) ? RB_TLB : RB_TL) / 4] = make_float4(tl[0], tl[1], tl[2], 0.0f);                              \
      if (SIDEB) { ang = wB; ow = wA; } else { ang = wA; ow = wB; }                                               \
    }                                                                                                             \
  }
#define BIG_STEP(PASS)                                                                                            \
  {                                                                                                               \
    const unsigned t = dyn ? (unsigned)T[k * 32 + lane] : 0xffffu;                                                \
    const bool has = t != 0xffffu;                                                                                \
    const int c = t & 0x3fff;                                                                                     \
    const bool sideB = (t & 0x4000u) != 0u, cpl = (t & 0x8000u) != 0u;                                            \
    float* rec1 = rr + (size_t)(has ? c : 0) * RB_WORDS;                                                          \
    V3 ov = v3(0, 0, 0), ow = v3(0, 0, 0);                                                                        \
    if ((coupled >> k) & 1ull) {                                                                                  \
      int src = lane;                                                                                             \
      if (has && cpl) { const int info = __float_as_int(rec1[RB_INFO]); src = sideB ? (info & 255) : (info >> 8); } \
      ov = v3(__shfl_sync(FULL, vel.x, src), __shfl_sync(FULL, vel.y, src), __shfl_sync(FULL, vel.z, src));       \
      ow = v3(__shfl_sync(FULL, ang.x, src), __shfl_sync(FULL, ang.y, src), __shfl_sync(FULL, ang.z, src));       \
    }                                                                                                             \
    if (has) {                                                                                                    \
      const float4* rec = (const float4*)rec1;                                                                    \
      float* ml = rec1 + (sideB ? RB_LAMB : RB_LAM);                                                              \
      float lim = 0.0f;                                                                                           \
      if (PASS == 2) lim = rec[7].w * ml[0];                                                                      \
      const float bias = (PASS == 1) ? rec1[RB_BIAS] : 0.0f;                                                      \
      _Pragma("unroll")                                                                                           \
      for (int r = (PASS == 2) ? 1 : 0; r < ((PASS == 1) ? 1 : 3); ++r) {                                         \
        if (r >= nrows) break;                                                                                    \
        const float4 q0 = rec[r * 4], q1 = rec[r * 4 + 1], q2 = rec[r * 4 + 2], q3 = rec[r * 4 + 3];              \
        const V3 dir = v3(q0.x, q0.y, q0.z), angA = v3(q0.w, q1.x, q1.y), iangA = v3(q1.z, q1.w, q2.x);           \
        const V3 dirMA = v3(q3.x, q3.y, q3.z);                                                                    \
        V3 angB = v3(0, 0, 0), iangB = v3(0, 0, 0), dirMB = v3(0, 0, 0);                                          \
        if (cpl || sideB) { angB = LD3(rec1 + RR_B + r * 9); iangB = LD3(rec1 + RR_B + r * 9 + 3); dirMB = LD3(rec1 + RR_B + r * 9 + 6); } \
        const float l = ml[r];                                                                                    \
        float dl = l;                                                                                             \
        if (PASS != 0) {                                                                                          \
          const V3 vA = sideB ? ov : vel, wA = sideB ? ow : ang;                                                  \
          const float a = dot(dir, vA) + dot(angA, wA);                                                           \
          float cc = q2.w;                                                                                        \
          if (cpl) cc = bias + (dot(dir, sideB ? vel : ov) + dot(angB, sideB ? ang : ow));                        \
          dl = (cc - a) * q2.y;                                                                                   \
          dl = (PASS == 1) ? fmaxf(0.0f - l, dl) : fminf(lim - l, fmaxf((0.0f - lim) - l, dl));                   \
          ml[r] = l + dl;                                                                                         \
          const float res = dl * q2.z;                                                                            \
          maxres = fmaxf(maxres, res * res);                                                                      \
        }                                                                                                         \
        if (!sideB) {                                                                                             \
          vel = vmad(vel, dirMA, dl); ang = vmad(ang, iangA, dl);                                                 \
          if (cpl) { ov = vmad(ov, dirMB, -dl); ow = vmad(ow, iangB, -dl); }          /* what the partner does */ \
        } else {                                                                                                  \
          vel = vmad(vel, dirMB, -dl); ang = vmad(ang, iangB, -dl);                                               \
          ov = vmad(ov, dirMA, dl); ow = vmad(ow, iangA, dl);                                                     \
        }                                                                                                         \
      }                                                                                                           \
      if (PASS == 2 && tors_world) BIG_TORS_ROWS(sideB, cpl)                                                      \
    }                                                                                                             \
  }
  // a colour without a contact between two dynamic bodies (the rule) takes the A-side-only step
#define BIG_FAST(PASS)                                                                                            \
  {                                                                                                               \
    const unsigned t = dyn ? (unsigned)T[k * 32 + lane] : 0xffffu;                                                \
    if (t != 0xffffu) {                                                                                           \
      float* rec1 = rr + (size_t)(t & 0x3fff) * RB_WORDS;                                                         \
      const float4* rec = (const float4*)rec1;                                                                    \
      float* ml = rec1 + RB_LAM;                                                                                  \
      float lim = 0.0f;                                                                                           \
      if (PASS == 2) lim = rec[7].w * ml[0];                                                                      \
      _Pragma("unroll")                                                                                           \
      for (int r = (PASS == 2) ? 1 : 0; r < ((PASS == 1) ? 1 : 3); ++r) {                                         \
        if (r >= nrows) break;                                                                                    \
        const float4 q0 = rec[r * 4], q1 = rec[r * 4 + 1], q2 = rec[r * 4 + 2], q3 = rec[r * 4 + 3];              \
        const V3 dir = v3(q0.x, q0.y, q0.z), angA = v3(q0.w, q1.x, q1.y), iangA = v3(q1.z, q1.w, q2.x);           \
        const V3 dirMA = v3(q3.x, q3.y, q3.z);                                                                    \
        const float l = ml[r];                                                                                    \
        float dl = l;                                                                                             \
        if (PASS != 0) {                                                                                          \
          const float a = dot(dir, vel) + dot(angA, ang);                                                         \
          dl = (q2.w - a) * q2.y;                                                                                 \
          dl = (PASS == 1) ? fmaxf(0.0f - l, dl) : fminf(lim - l, fmaxf((0.0f - lim) - l, dl));                   \
          ml[r] = l + dl;                                                                                         \
          const float res = dl * q2.z;                                                                            \
          maxres = fmaxf(maxres, res * res);                                                                      \
        }                                                                                                         \
        vel = vmad(vel, dirMA, dl); ang = vmad(ang, iangA, dl);                                                   \
      }                                                                                                           \
      if (PASS == 2 && tors_world) { V3 ow = v3(0, 0, 0); BIG_TORS_ROWS(false, false) (void)ow; }                 \
    }                                                                                                             \
  }
#define BIG_PASS(PASS)                                                                                            \
  {                                                                                                               \
    _Pragma("unroll 1")                                                                                           \
    for (int k = 0; k < ncolours; ++k) {                                                                          \
      if ((coupled >> k) & 1ull) BIG_STEP(PASS) else BIG_FAST(PASS)                                               \
    }                                                                                                             \
    __syncwarp();                                                                                                 \
  }
  BIG_PASS(0)
  for (int it = 0; it < P.solver_iterations && C > 0; ++it) {
    maxres = 0.0f;
    BIG_PASS(1)
    BIG_PASS(2)
    iters = it + 1;
    unsigned mx = __reduce_max_sync(FULL, __float_as_uint(maxres));
    if (__uint_as_float(mx) <= P.residual_threshold) break;
    if (RA) long_solve_poll(*bsp, lane, it);
  }
  if (RA && iters >= B2S_LONG_ITERS) long_solve_end(*bsp, lane);
#undef BIG_PASS
#undef BIG_TORS_ROWS
#undef BIG_FAST
#undef BIG_STEP
  PROF_SEC(2)
  if (dyn) { ST3(myb + BO_VEL, vel); ST3(myb + BO_ANG, ang); }
  __syncwarp();
  for (int c = lane; c < C; c += 32) {
    const int mk = S.cmk[c];
    const float* rec1 = rr + (size_t)c * RB_WORDS;
    float* p = W.man_pts + ((nbase + (mk >> 2)) * 4 + (mk & 3)) * B2S_CP_FLOATS;
    p[10] = rec1[RB_LAM]; p[11] = rec1[RB_LAM + 1]; p[12] = rec1[RB_LAM + 2];
  }
  if (lane == 0) {
    int32_t* st = W.solver_stats + (size_t)e * 4;
    st[0] = C * nrows; st[1] = ncolours; st[2] = iters; st[3] = C;
  }
  __syncwarp();
  bool bad = false;
  for (int i = lane; i < nm; i += 32) {
    const float* b = S.body + (Ns + L + i) * BODY_STRIDE;
    V3 ang_ = LD3(b + BO_ANG), vel_ = LD3(b + BO_VEL);
    float wl = len(ang_);
    if (wl * dt > B2S_HALF_PI) ang_ = ang_ * (B2S_HALF_PI / (wl * dt));
    V3 pos = LD3(b + BO_POS) + vel_ * dt;
    Q4 qq = q_integrate(q4(b[BO_QUAT], b[BO_QUAT + 1], b[BO_QUAT + 2], b[BO_QUAT + 3]), ang_, dt);
    BS(0, i) = pos.x; BS(1, i) = pos.y; BS(2, i) = pos.z;
    BS(3, i) = qq.x; BS(4, i) = qq.y; BS(5, i) = qq.z; BS(6, i) = qq.w;
    BS(7, i) = vel_.x; BS(8, i) = vel_.y; BS(9, i) = vel_.z;
    BS(10, i) = ang_.x; BS(11, i) = ang_.y; BS(12, i) = ang_.z;
    float chk = (pos.x + pos.y) + pos.z;
    if (!(fabsf(chk) < 1e6f)) bad = true;
  }
  if (__any_sync(FULL, bad) && lane == 0) W.error_flags[e] |= 4;
  if (lane < 7) {
    float q = W.buf.joint_state[(0 * 7 + lane) * W.B + e], qd = W.buf.joint_state[(1 * 7 + lane) * W.B + e];
    W.buf.joint_state[(0 * 7 + lane) * W.B + e] = q + qd * dt;
  }
  __syncwarp();
  if (lane == 0) W.num_steps[e] += 1;
  __syncwarp();
  PROF_SEC(3)
  (void)newn;
}

// ---- register-resident solve (max_contacts <= 32 and NB <= 32) -------------------------------------------
// Row build, colouring AND the Gauss-Seidel sweeps run one CONTACT per lane: the rows of a contact never leave the
// registers they were built in.  The sweep of one environment is a chain of dependent row updates (12 per body and
// iteration for a body resting on four points, 24 with the torsional rows) and 5% of the env-substeps run all 50
// iterations, so its latency -- not its throughput -- sets the length of the block's solve stage: a row update is two
// dot products, a subtraction, a multiplication, the clamp of the increment and one fused multiply-add on registers
// (the constant terms, the mass scaling and the bounds are folded into per-row values off the chain).  Order of
// operations per body = colour order = the oracle's order, and every row update performs the oracle's IEEE operations
// on the same operands, so results stay bit-identical.
#define SOLVE_T_WORDS 512               // byte table [64 colours][32 slots] in the warp's `con` scratch

__device__ __noinline__ void substep_post_reg(int e, int lane, int wib, int C, int newn) {
  const WarpSmem S = carve(wib);
  const B2SParams& P = W.P;
  const float dt = (float)P.time_step;
  const int Ns = W.Ns, L = W.L;
  const int nm = W.buf.num_movables[e];
  const int par = W.man_parity[e];          // already flipped: current buffer
  const size_t nbase = ((size_t)par * W.B + e) * P.max_manifolds;
  const int nrows = 1 + P.friction_dirs;
  const bool act = lane < C;
  PROF_SEC0()
  unsigned char* T = (unsigned char*)S.con;
  int sA = 0, sB = 0, mk = 0;
  bool dA = false, dB = false;
  // the three rows of this lane's contact (row r: dir, dirM = dir / mass of A, angA = rA x dir, iangA = I_A^-1 angA, 1/d,
  // d, c = bias + (dir . vB + angB . wB) for a body B that is not dynamic) and its scalars
  V3 rdir[3], rdirM[3], rangA[3], riangA[3];
  float rinvd[3], rd[3], rc[3], rl[3];
  float imB = 0.0f, bias0 = 0.0f, mu = 0.0f;
  V3 rangB[3], riangB[3];          // B side of the rows: only used by environments in which movables touch each other
  // torsional rows of this lane's contact (spinning about rdir[0], rolling about rdir[1] and rdir[2]; angular only, no
  // warm start, not part of the residual test: see the oracle): I_A^-1 axis, 1/d, axis . wB of a kinematic body B,
  // impulse; combined spinning / rolling coefficient
  V3 tiA[3];
  float tinvd[3], tk[3], tl[3], mu_s = 0.0f, mu_r = 0.0f;
  const bool tors_world = P.rolling_friction > 0.0f;     // warp uniform; b2s_create checked friction_dirs == 2
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    rdir[r] = rdirM[r] = rangA[r] = riangA[r] = rangB[r] = riangB[r] = tiA[r] = v3(0, 0, 0);
    rinvd[r] = rd[r] = rc[r] = rl[r] = tinvd[r] = tk[r] = tl[r] = 0.0f;
  }
  __syncwarp();
  if (act) {
    mk = S.cmk[lane];
    const int key = W.man_keys[nbase + (mk >> 2)];
    const int a = key >> 16, b = key & 0xffff;
    sA = __float_as_int(S.col[a * COL_STRIDE + CO_SLOT]); sB = __float_as_int(S.col[b * COL_STRIDE + CO_SLOT]);
    dA = __float_as_int(S.body[sA * BODY_STRIDE + BO_TYPE]) == B2S_TYPE_DYNAMIC;
    dB = __float_as_int(S.body[sB * BODY_STRIDE + BO_TYPE]) == B2S_TYPE_DYNAMIC;
  }
  PROF_SEC(0)
  // greedy colouring in contact order; lane s keeps the colour mask of body slot s
  unsigned long long used = 0ull;
  int mycol = -1, ncolours = 0;
  bool col_over = false;
  for (int i = 0; i < C; ++i) {
    const int isA = __shfl_sync(FULL, sA, i), isB = __shfl_sync(FULL, sB, i);
    const bool idA = __shfl_sync(FULL, (int)dA, i) != 0, idB = __shfl_sync(FULL, (int)dB, i) != 0;
    unsigned long long mA = __shfl_sync(FULL, used, isA), mB = __shfl_sync(FULL, used, isB);
    unsigned long long mask = (idA ? mA : 0ull) | (idB ? mB : 0ull);
    if (mask == ~0ull) { col_over = true; continue; }
    const int k = __ffsll((long long)~mask) - 1;
    if (lane == i) mycol = k;
    if (k + 1 > ncolours) ncolours = k + 1;
    if ((idA && lane == isA) || (idB && lane == isB)) used |= 1ull << k;
  }
  if (col_over && lane == 0) W.error_flags[e] |= 16;
  // T[k][slot] = contact of colour k touching dynamic body `slot`: index | side << 5 | coupled << 6, 0xff = none
  for (int i = lane; i < ncolours * 8; i += 32) ((unsigned*)T)[i] = 0xffffffffu;
  __syncwarp();
  if (act && mycol >= 0) {
    const unsigned cpl = (dA && dB) ? 64u : 0u;
    if (dA) T[mycol * 32 + sA] = (unsigned char)(lane | cpl);
    if (dB) T[mycol * 32 + sB] = (unsigned char)(lane | 32u | cpl);
  }
  // colours that hold a contact between two dynamic bodies need the velocity exchange (warp uniform)
  unsigned long long coupled = 0ull;
  for (int k = 0; k < ncolours; ++k)
    if (__any_sync(FULL, act && mycol == k && dA && dB)) coupled |= 1ull << k;
  if (act && dB && !dA) W.error_flags[e] |= 64;   // cannot happen: movable colliders are numbered last
  __syncwarp();
  // rows of this lane's contact (after the colouring, which only needs the body slots: fewer live registers there)
  if (act) {
    const int m = mk >> 2, k = mk & 3;
    const float* bA = S.body + sA * BODY_STRIDE;
    const float* bB = S.body + sB * BODY_STRIDE;
    const float* p = W.man_pts + ((nbase + m) * 4 + k) * B2S_CP_FLOATS;
    V3 posA = LD3(bA + BO_POS), posB = LD3(bB + BO_POS);
    V3 wA = posA + mmul(ldm3(bA + BO_R), v3(p[0], p[1], p[2]));
    V3 wB = posB + mmul(ldm3(bB + BO_R), v3(p[3], p[4], p[5]));
    V3 n = v3(p[6], p[7], p[8]);
    V3 rA = wA - posA, rB = wB - posB;
    V3 t1, t2;
    plane_space(n, &t1, &t2);
    const V3 velB = LD3(bB + BO_VEL), angvB = LD3(bB + BO_ANG);
    if (P.friction_dirs == 1) {
      V3 rel = (LD3(bA + BO_VEL) + cross(LD3(bA + BO_ANG), rA)) - (velB + cross(angvB, rB));
      V3 lat = rel - n * dot(rel, n);
      float l2 = len2(lat);
      if (l2 > 1e-12f) t1 = lat * (1.0f / sqrtf(l2));
    }
    const float imA = bA[BO_INVM];
    imB = bB[BO_INVM];
    const M3 iA = ldm3(bA + BO_INVI), iB = ldm3(bB + BO_INVI);
    const float pen = p[9] + P.linear_slop;
    bias0 = (pen > 0.0f) ? -(pen / dt) : -(pen * P.erp2 / dt);
    mu = bA[BO_FRIC] * bB[BO_FRIC];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const V3 dir = (r == 0) ? n : (r == 1 ? t1 : t2);
      const V3 angA = cross(rA, dir), angB = cross(rB, dir);
      const V3 iangA = mmul(iA, angA), iangB = mmul(iB, angB);
      const float d = ((imA + imB) + dot(iangA, angA)) + dot(iangB, angB);
      const float inv_d = (d > 0.0f && r < nrows) ? 1.0f / d : 0.0f;
      float l0 = (r == 0 || P.friction_dirs == 2) ? p[10 + r] * P.warmstart : 0.0f;
      if (r >= nrows) l0 = 0.0f;
      rdir[r] = dir; rdirM[r] = dir * imA; rangA[r] = angA; riangA[r] = iangA; rinvd[r] = inv_d; rd[r] = d;
      rc[r] = ((r == 0) ? bias0 : 0.0f) + (dot(dir, velB) + dot(angB, angvB)); rl[r] = l0;
      rangB[r] = angB; riangB[r] = iangB;
    }
    if (tors_world) {
      const float rollB = dB ? P.rolling_friction : 0.0f, spinB = dB ? P.spinning_friction : 0.0f;
      mu_s = fminf(10.0f, P.spinning_friction * bB[BO_FRIC] + spinB * bA[BO_FRIC]);
      mu_r = fminf(10.0f, P.rolling_friction * bB[BO_FRIC] + rollB * bA[BO_FRIC]);
      const bool tors = mu_r > 0.0f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const V3 ax = rdir[r];
        tiA[r] = mmul(iA, ax);
        const V3 tiB = mmul(iB, ax);
        const float d = dot(tiA[r], ax) + dot(tiB, ax);
        tinvd[r] = (d > 0.0f && tors) ? 1.0f / d : 0.0f;
        tk[r] = dot(ax, angvB);
      }
      if (!tors) mu_r = 0.0f;
    }
  }
  __syncwarp();     // the contact lanes rejoin the others here: the sweeps below are warp-synchronous (shuffles in every
                    // colour step), and a warp that enters them split pays the divergent-shuffle path on every one
  PROF_SEC(1)
  float maxres = 0.0f;
  int iters = 0;
  if (coupled == 0ull) {
    // ---- the rule: no contact joins two dynamic bodies.  The sweeps stay on the CONTACT lanes: the rows never leave
    // the registers they were built in, and every lane carries a copy of the velocity of its (dynamic) body A.  In
    // colour step k the lanes whose contact has colour k update their rows -- at most one contact per body, that is
    // what a colour is -- and then every lane takes the velocity of its body from the lane that just changed it
    // (T[k][body]: one byte load and six shuffles).  Per body the row updates happen in colour order with the
    // oracle's IEEE operations on the same operands as before, so the results are bit-identical; per row there is no
    // load, no address arithmetic and no table walk left, which is what the latency of a 50-iteration solve -- the
    // thing a block's solve stage waits for -- was made of.
    const bool mine = act && dA && mycol >= 0;
    const float* myb = S.body + sA * BODY_STRIDE;
    V3 vel = LD3(myb + BO_VEL), ang = LD3(myb + BO_ANG);
    const unsigned char* Tm = T + sA;
    // Every colour step is branch-free: all lanes compute the update of "their" row, the lanes whose contact is not of
    // colour k discard it with selects (a divergent branch per step cost more than the arithmetic it skipped), and the
    // byte of T that names the exchange partner is loaded before the arithmetic, so its latency hides behind it.
#define XCHG_SRC(k) const unsigned t_ = (act && dA) ? (unsigned)Tm[(k) * 32] : 0xffu; const int src_ = (t_ != 0xffu) ? (int)(t_ & 31u) : lane;
#define XCHG()                                                                                                    \
    {                                                                                                             \
      vel.x = __shfl_sync(FULL, vel.x, src_); vel.y = __shfl_sync(FULL, vel.y, src_); vel.z = __shfl_sync(FULL, vel.z, src_); \
      ang.x = __shfl_sync(FULL, ang.x, src_); ang.y = __shfl_sync(FULL, ang.y, src_); ang.z = __shfl_sync(FULL, ang.z, src_); \
    }
#define KEEP(on, nv, na) { vel.x = (on) ? (nv).x : vel.x; vel.y = (on) ? (nv).y : vel.y; vel.z = (on) ? (nv).z : vel.z; \
                           ang.x = (on) ? (na).x : ang.x; ang.y = (on) ? (na).y : ang.y; ang.z = (on) ? (na).z : ang.z; }
    // one row update on (v, w) -> (v, w), lambda and the residual in temporaries.  The chain from (v, w) to the new
    // (v, w) is two dot products, a subtraction, a multiplication, the clamp of the increment and one fused
    // multiply-add; everything else (the bounds of the increment, the new lambda, the residual) hangs off it
#define ROW(r, LO, HI, v, w, lnew, res2)                                                                          \
      {                                                                                                           \
        const float lo_ = (LO) - rl[r], hi_ = (HI) - rl[r];                                                       \
        const float a_ = dot(rdir[r], v) + dot(rangA[r], w);                                                      \
        float dl = (rc[r] - a_) * rinvd[r];                                                                       \
        dl = fminf(hi_, fmaxf(lo_, dl));                                                                          \
        lnew = rl[r] + dl;                                                                                        \
        const float res = dl * rd[r];                                                                             \
        res2 = res * res;                                                                                         \
        v = vmad(v, rdirM[r], dl); w = vmad(w, riangA[r], dl);                                                    \
      }
    // warm start
#pragma unroll 1
    for (int k = 0; k < ncolours; ++k) {
      XCHG_SRC(k)
      const bool on = mine && mycol == k;
      V3 v = vel, w = ang;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (r < nrows) { v = vmad(v, rdirM[r], rl[r]); w = vmad(w, riangA[r], rl[r]); }
      KEEP(on, v, w)
      XCHG()
    }
    for (int it = 0; it < P.solver_iterations && C > 0; ++it) {
      maxres = 0.0f;
#pragma unroll 1
      for (int k = 0; k < ncolours; ++k) {                 // normal rows
        XCHG_SRC(k)
        const bool on = mine && mycol == k;
        V3 v = vel, w = ang;
        float l0n, r0;
        {
          const float lo_ = 0.0f - rl[0];
          const float a_ = dot(rdir[0], v) + dot(rangA[0], w);
          float dl = (rc[0] - a_) * rinvd[0];
          dl = fmaxf(lo_, dl);
          l0n = rl[0] + dl;
          const float res = dl * rd[0];
          r0 = res * res;
          v = vmad(v, rdirM[0], dl); w = vmad(w, riangA[0], dl);
        }
        rl[0] = on ? l0n : rl[0];
        maxres = on ? fmaxf(maxres, r0) : maxres;
        KEEP(on, v, w)
        XCHG()
      }
#pragma unroll 1
      for (int k = 0; k < ncolours; ++k) {                 // friction rows
        XCHG_SRC(k)
        const bool on = mine && mycol == k;
        const float lim = mu * rl[0];
        V3 v = vel, w = ang;
        float l1n, r1, l2n = rl[2], r2 = 0.0f;
        ROW(1, 0.0f - lim, lim, v, w, l1n, r1)
        if (nrows > 2) ROW(2, 0.0f - lim, lim, v, w, l2n, r2)
        rl[1] = on ? l1n : rl[1]; rl[2] = on ? l2n : rl[2];
        maxres = on ? fmaxf(fmaxf(maxres, r1), r2) : maxres;
        if (tors_world) {                                    // torsional rows of the same contact: angular velocity only
          const bool ont = on && mu_r > 0.0f && rl[0] > 0.0f;
          V3 wt = w;
          float tn[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float mu_c = (r == 0) ? mu_s : mu_r;
            const float tlim = fminf(mu_c * rl[0], mu_c);
            const float lo_ = (0.0f - tlim) - tl[r], hi_ = tlim - tl[r];
            float dl = (tk[r] - dot(rdir[r], wt)) * tinvd[r];
            dl = fminf(hi_, fmaxf(lo_, dl));
            tn[r] = tl[r] + dl;
            wt = vmad(wt, tiA[r], dl);
          }
#pragma unroll
          for (int r = 0; r < 3; ++r) tl[r] = ont ? tn[r] : tl[r];
          w.x = ont ? wt.x : w.x; w.y = ont ? wt.y : w.y; w.z = ont ? wt.z : w.z;
        }
        KEEP(on, v, w)
        XCHG()
      }
      iters = it + 1;
      unsigned mx = __reduce_max_sync(FULL, __float_as_uint(maxres));
      if (__uint_as_float(mx) <= P.residual_threshold) break;
    }
#undef KEEP
#undef XCHG_SRC
#undef ROW
#undef XCHG
    PROF_SEC(2)
    if (act && dA) {
      float* wb = S.body + sA * BODY_STRIDE;           // every lane of a body holds the same final velocity
      ST3(wb + BO_VEL, vel); ST3(wb + BO_ANG, ang);
    }
    __syncwarp();
    if (act) {
      float* p = W.man_pts + ((nbase + (mk >> 2)) * 4 + (mk & 3)) * B2S_CP_FLOATS;
      p[10] = rl[0]; p[11] = rl[1]; p[12] = rl[2];
    }
  } else {
    // ---- movables touch each other (5 % of the env-substeps, but the long solves: 57 us on average against 10).
    // Still one CONTACT per lane with its rows in registers; the velocities of the bodies stay in the shared body
    // table: the lanes whose contact has colour k load the velocities of their one or two dynamic bodies, update their
    // row(s) -- body B takes the opposite impulse -- and store them back; a warp barrier separates the colours.  No
    // two contacts of a colour share a dynamic body, so the order inside a colour is immaterial, and every row update
    // performs the oracle's IEEE operations on the same operands: bit-identical to the sequential sweep.
    const bool mine = act && dA && mycol >= 0;
    float* vA = S.body + sA * BODY_STRIDE + BO_VEL;        // vel3 then ang3 (BO_ANG = BO_VEL + 3)
    float* vB = S.body + sB * BODY_STRIDE + BO_VEL;
    // one row update: (va, wa) of body A, (vb, wb) of body B when it is dynamic (else its constant terms k1, k2)
#define CROW(r, BIAS, LO, HI)                                                                                     \
      {                                                                                                           \
        const float lo_ = (LO) - rl[r], hi_ = (HI) - rl[r];                                                       \
        const float a_ = dot(rdir[r], va) + dot(rangA[r], wa);                                                    \
        const float c_ = dB ? (BIAS) + (dot(rdir[r], vb) + dot(rangB[r], wb)) : rc[r];                            \
        float dl = (c_ - a_) * rinvd[r];                                                                          \
        dl = fminf(hi_, fmaxf(lo_, dl));                                                                          \
        rl[r] = rl[r] + dl;                                                                                       \
        const float res = dl * rd[r];                                                                             \
        maxres = fmaxf(maxres, res * res);                                                                        \
        va = vmad(va, rdirM[r], dl); wa = vmad(wa, riangA[r], dl);                                                \
        if (dB) { vb = vmad(vb, rdir[r] * imB, -dl); wb = vmad(wb, riangB[r], -dl); }                             \
      }
#define CLOAD V3 va = LD3(vA), wa = LD3(vA + 3), vb = v3(0, 0, 0), wb = v3(0, 0, 0); if (dB) { vb = LD3(vB); wb = LD3(vB + 3); }
#define CSTORE { ST3(vA, va); ST3(vA + 3, wa); if (dB) { ST3(vB, vb); ST3(vB + 3, wb); } }
    // warm start
#pragma unroll 1
    for (int k = 0; k < ncolours; ++k) {
      if (mine && mycol == k) {
        CLOAD
#pragma unroll
        for (int r = 0; r < 3; ++r)
          if (r < nrows) {
            va = vmad(va, rdirM[r], rl[r]); wa = vmad(wa, riangA[r], rl[r]);
            if (dB) { vb = vmad(vb, rdir[r] * imB, -rl[r]); wb = vmad(wb, riangB[r], -rl[r]); }
          }
        CSTORE
      }
      __syncwarp();
    }
    for (int it = 0; it < P.solver_iterations && C > 0; ++it) {
      maxres = 0.0f;
#pragma unroll 1
      for (int k = 0; k < ncolours; ++k) {                 // normal rows
        if (mine && mycol == k) {
          CLOAD
          CROW(0, bias0, 0.0f, __int_as_float(0x7f800000))
          CSTORE
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int k = 0; k < ncolours; ++k) {                 // friction rows
        if (mine && mycol == k) {
          CLOAD
          const float lim = mu * rl[0];
          CROW(1, 0.0f, 0.0f - lim, lim)
          if (nrows > 2) CROW(2, 0.0f, 0.0f - lim, lim)
          if (tors_world && mu_r > 0.0f && rl[0] > 0.0f) {   // torsional rows of the same contact
            M3 iB = {v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0)};
            if (dB) iB = ldm3(S.body + sB * BODY_STRIDE + BO_INVI);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const float mu_c = (r == 0) ? mu_s : mu_r;
              const float tlim = fminf(mu_c * rl[0], mu_c);
              const float lo_ = (0.0f - tlim) - tl[r], hi_ = tlim - tl[r];
              float dl = ((dB ? dot(rdir[r], wb) : tk[r]) - dot(rdir[r], wa)) * tinvd[r];
              dl = fminf(hi_, fmaxf(lo_, dl));
              tl[r] = tl[r] + dl;
              wa = vmad(wa, tiA[r], dl);
              if (dB) wb = vmad(wb, mmul(iB, rdir[r]), -dl);
            }
          }
          CSTORE
        }
        __syncwarp();
      }
      iters = it + 1;
      unsigned mx = __reduce_max_sync(FULL, __float_as_uint(maxres));
      if (__uint_as_float(mx) <= P.residual_threshold) break;
    }
#undef CROW
#undef CLOAD
#undef CSTORE
    PROF_SEC(2)
    if (act) {
      float* p = W.man_pts + ((nbase + (mk >> 2)) * 4 + (mk & 3)) * B2S_CP_FLOATS;
      p[10] = rl[0]; p[11] = rl[1]; p[12] = rl[2];
    }
  }
  if (lane == 0) {
    int32_t* st = W.solver_stats + (size_t)e * 4;
    st[0] = C * nrows; st[1] = ncolours; st[2] = iters; st[3] = C;
  }
  __syncwarp();
  bool bad = false;
  for (int i = lane; i < nm; i += 32) {
    const float* b = S.body + (Ns + L + i) * BODY_STRIDE;
    V3 ang = LD3(b + BO_ANG), vel = LD3(b + BO_VEL);
    float wl = len(ang);
    if (wl * dt > B2S_HALF_PI) ang = ang * (B2S_HALF_PI / (wl * dt));
    V3 pos = LD3(b + BO_POS) + vel * dt;
    Q4 qq = q_integrate(q4(b[BO_QUAT], b[BO_QUAT + 1], b[BO_QUAT + 2], b[BO_QUAT + 3]), ang, dt);
    BS(0, i) = pos.x; BS(1, i) = pos.y; BS(2, i) = pos.z;
    BS(3, i) = qq.x; BS(4, i) = qq.y; BS(5, i) = qq.z; BS(6, i) = qq.w;
    BS(7, i) = vel.x; BS(8, i) = vel.y; BS(9, i) = vel.z;
    BS(10, i) = ang.x; BS(11, i) = ang.y; BS(12, i) = ang.z;
    float chk = (pos.x + pos.y) + pos.z;
    if (!(fabsf(chk) < 1e6f)) bad = true;
  }
  if (__any_sync(FULL, bad) && lane == 0) W.error_flags[e] |= 4;
  if (lane < 7) {
    float q = W.buf.joint_state[(0 * 7 + lane) * W.B + e], qd = W.buf.joint_state[(1 * 7 + lane) * W.B + e];
    W.buf.joint_state[(0 * 7 + lane) * W.B + e] = q + qd * dt;
  }
  __syncwarp();
  if (lane == 0) W.num_steps[e] += 1;
  __syncwarp();
  PROF_SEC(3)
  (void)newn;
}

// -------------------------------------------------------- phase machine -----

__device__ void movable_status(int e, int lane, int which) {
  for (int i = lane; i < W.Nmax; i += 32) {
    float* s = W.status + (((size_t)e * 2 + which) * W.Nmax + i) * 4;
    if (i < W.buf.num_movables[e]) {
      s[0] = BS(0, i); s[1] = BS(1, i); s[2] = BS(2, i);
      s[3] = yaw_from_q(q4(BS(3, i), BS(4, i), BS(5, i), BS(6, i)));
    } else { s[0] = s[1] = s[2] = s[3] = 0.0f; }
  }
  __syncwarp();
}

__device__ __noinline__ void phase_logic(int e, int lane, int wib) {
  float* fk = carve(wib).fk;
  const B2SParams& P = W.P;
  int32_t* ps = W.phase_state + (size_t)e * 8;
  int ph = W.phase[e];
  const int nsteps = W.num_steps[e];
  bool interrupt = ps[5] != 0;
  const int ps0 = ps[0];
  bool ready, reset_t = false;
  if (interrupt) ready = true;
  else {
    int lr = arm_is_ready(e, lane);
    if (lr && (P.time_step * (double)nsteps >= W.ctrl_time[(size_t)e * 5 + 4])) { reset_t = true; ready = true; }
    else if (ps0 < 0) ready = true;
    else if (nsteps >= ps0) { reset_t = true; ready = true; }
    else ready = false;
  }
  if (reset_t && lane == 0) { W.ctrl_flags[(size_t)e * 4] = 0; W.ctrl_flags[(size_t)e * 4 + 1] = 0; }
  __syncwarp();
  float ee[7];
  {
    float q[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) q[j] = W.buf.joint_state[(0 * 7 + j) * W.B + e];
    fk_chain(W.arm, q, lane, wib);
    xf_store(xf_mul(xf_from(fk + 6 * 7), xf_from(W.arm->ee)), ee);
  }
  int ps1 = ps[1];
  int new_ps0 = ps0;
  if (ready) {
    if (interrupt && ph != B2S_PHASE_POST && ph != B2S_PHASE_OFFSTAGE && ph != B2S_PHASE_DONE) ph = B2S_PHASE_POST;
    else if (ph == B2S_PHASE_POST && P.num_goal_steps > 0 && ps1 < P.num_goal_steps) ph = B2S_PHASE_PRE;   // next goal step (push_env.py:803-806)
    else ph = ph + 1;
    new_ps0 = nsteps + (ph == B2S_PHASE_MOTION ? P.max_motion_steps : ph == B2S_PHASE_OFFSTAGE ? P.max_offstage_steps : P.max_phase_steps);
    // waypoints[num_waypoints]; an interrupt can take the arm to 'post' once more after the last goal step, and the
    // reference would then index past its list in 'pre' -- which it never reaches (post -> offstage), as here
    const float* wp = W.waypoints + ((size_t)e * W.G + min(ps1, W.G - 1)) * 14;
    float pose[7];
    if (ph == B2S_PHASE_PRE) {
#pragma unroll
      for (int k = 0; k < 7; ++k) pose[k] = wp[k];
      pose[2] = P.gripper_safe_height;
      arm_set_link_target(e, lane, pose);
    } else if (ph == B2S_PHASE_START) {
#pragma unroll
      for (int k = 0; k < 7; ++k) pose[k] = wp[k];
      arm_set_link_target(e, lane, pose);
    } else if (ph == B2S_PHASE_MOTION) {
#pragma unroll
      for (int k = 0; k < 7; ++k) pose[k] = wp[7 + k];
      arm_set_link_target(e, lane, pose);
    } else if (ph == B2S_PHASE_POST) {
      ps1 += 1;
#pragma unroll
      for (int k = 0; k < 7; ++k) pose[k] = ee[k];
      pose[2] = P.gripper_safe_height;
      arm_set_link_target(e, lane, pose);
    } else if (ph == B2S_PHASE_OFFSTAGE) {
      float q[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) q[k] = P.offstage_positions[k];
      arm_set_joint_target(e, lane, q);
    }
  }
  interrupt = false;
  const int cf = W.contact_flags[e];
  if (ph == B2S_PHASE_MOTION && (cf & 1)) interrupt = true;
  bool safe = true;
  if (ph == B2S_PHASE_PRE) { if (cf & 2) safe = false; }
  else if (ph == B2S_PHASE_START) {
    if (cf & 2) {
      float start_z = P.finger_tip_offset + 0.5f * (P.cspace_high[2] + P.cspace_low[2]);
      float dist = ee[2] - start_z;
      if (!(fabsf(dist) <= 0.01f)) safe = false;
    }
  } else if (ph == B2S_PHASE_DONE) {
    if (cf & 2) safe = false;
    else {
      bool out = false;
      for (int i = lane; i < W.buf.num_movables[e]; i += 32) {
        float x = BS(0, i), y = BS(1, i);
        if (x < P.table_workspace_low[0] || x > P.table_workspace_high[0] || y < P.table_workspace_low[1] || y > P.table_workspace_high[1]) out = true;
      }
      if (__any_sync(FULL, out)) safe = false;
    }
  }
  if (!safe) interrupt = true;
  __syncwarp();
  if (lane == 0) {
    if (!safe) W.buf.is_safe[e] = 0;
    if (interrupt && ph == B2S_PHASE_DONE) ps[4] = 1;
    ps[0] = new_ps0; ps[1] = ps1; ps[5] = interrupt ? 1 : 0;
    W.phase[e] = ph;
  }
  __syncwarp();
  (void)fk;
}

__device__ bool all_stable(int e, int lane, float lin, float ang) {
  bool moving = false;
  for (int i = lane; i < W.buf.num_movables[e]; i += 32) {
    float lv = len(v3(BS(7, i), BS(8, i), BS(9, i)));
    float av = len(v3(BS(10, i), BS(11, i), BS(12, i)));
    if (lv >= lin || av >= ang) moving = true;
  }
  return !__any_sync(FULL, moving);
}

__device__ __noinline__ void finish_action(int e, int lane) {
  const B2SParams& P = W.P;
  movable_status(e, lane, 1);
  // sums are sequential over bodies in the oracle: lane 0 does them (Nmax is small)
  if (lane == 0) {
    float dp = 0.0f, da = 0.0f;
    for (int i = 0; i < W.buf.num_movables[e]; ++i) {
      const float* s0 = W.status + (((size_t)e * 2 + 0) * W.Nmax + i) * 4;
      const float* s1 = W.status + (((size_t)e * 2 + 1) * W.Nmax + i) * 4;
      dp = dp + len(v3(s1[0] - s0[0], s1[1] - s0[1], s1[2] - s0[2]));
      da = da + fabsf(b2s_wrap_pi(s1[3] - s0[3]));
    }
    W.buf.is_effective[e] = (dp <= P.min_delta_position && da <= P.min_delta_angle) ? 0 : 1;
    W.phase[e] = B2S_PHASE_IDLE;
  }
  __syncwarp();
}

// between two actions of a device-side episode (b2s_rollout.cuh); out of line: it runs once per ~2000 substeps
__device__ __noinline__ void rollout_next(int e, int lane) { rollout_advance(g_W, e, lane); }
__device__ __noinline__ void rollout_after_reset(int e, int lane) {
  if (rollout_reset_check(g_W, e, lane) != 0) return;
  if (g_W.ro.enabled == RO_ASYNC) { if (lane == 0) async_reset_done(g_W, e); __syncwarp(); }
  else episode_start(g_W, e, lane, nullptr);
}
__device__ __noinline__ void async_after_action(int e, int lane) { if (lane == 0) async_action_done(g_W, e); __syncwarp(); }

// ----------------------------------------------------------- the kernel -----

// Hand-out of environments inside a stage: dynamic (shared counter) when the solver rows live in registers;
// with rows in per-warp shared memory an environment must stay with one warp for the whole substep.
__device__ __forceinline__ int grab_slot(int* counter, int lane, int wib, int E, bool first) {
  (void)wib; (void)first;
  int slot = 0;
  if (lane == 0) slot = atomicAdd(counter, 1);
  return __shfl_sync(FULL, slot, 0);
}

// Next candidate pair of the block: unit u of the concatenated pair lists of the active environments.
// Returns the environment slot (-1: none left) and the pair index in *p_out.
#if B2S_HALF
// nu <= UNITS_PER_WARP units per grab: unit k of the warp (lanes k*UW..) takes pair u0 + k; a unit may come back with
// -1.  nu follows the block's pair count of this round (block_pairs, summed in stage A): a block with few pairs -- a
// sparse launch, a single environment -- spreads them over its warps instead of serialising four on one.
__device__ __forceinline__ int grab_pair(int* counter, int lane, int E, int* p_out, int block_pairs, int Wn) {
#ifdef B2S_FIXED_NU
  const int nu = UNITS_PER_WARP;
#else
  const int nu = max(1, min(UNITS_PER_WARP, (block_pairs + Wn - 1) / Wn));
#endif
  int u0 = 0;
  if (lane == 0) u0 = atomicAdd(counter, nu);
  u0 = __shfl_sync(FULL, u0, 0);
  int base = 0;
  int res = -1, pidx = 0;
  for (int s0 = 0; s0 < E; s0 += 32) {
    const int slot = s0 + lane;
    int cnt = 0;
    if (slot < E) { const int* m = env_meta(slot); if (m[META_ACTIVE]) cnt = m[META_NP]; }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    const int total = __shfl_sync(FULL, incl, 31);
#pragma unroll
    for (int k = 0; k < UNITS_PER_WARP; ++k) {
      const int u = u0 + k;
      if (k < nu && u >= base && u < base + total) {  // uniform over the warp
        const int l = __ffs(__ballot_sync(FULL, u < base + incl)) - 1;
        const int pk = u - base - __shfl_sync(FULL, incl - cnt, l);
        if (UH == k) { res = s0 + l; pidx = pk; }
      }
    }
    base += total;
    if (u0 + nu - 1 < base) break;
  }
  *p_out = pidx;
  return res;
}
#else
__device__ __forceinline__ int grab_pair(int* counter, int lane, int E, int* p_out) {
  int u = 0;
  if (lane == 0) u = atomicAdd(counter, 1);
  u = __shfl_sync(FULL, u, 0);
  int base = 0;
  for (int s0 = 0; s0 < E; s0 += 32) {
    const int slot = s0 + lane;
    int cnt = 0;
    if (slot < E) { const int* m = env_meta(slot); if (m[META_ACTIVE]) cnt = m[META_NP]; }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    const int total = __shfl_sync(FULL, incl, 31);
    if (u < base + total) {
      const int l = __ffs(__ballot_sync(FULL, u < base + incl)) - 1;
      *p_out = u - base - __shfl_sync(FULL, incl - cnt, l);
      return s0 + l;
    }
    base += total;
  }
  return -1;
}

#endif

// Block = Wn warps stepping E environments.  Every substep runs three stages separated by block barriers
// (scene -> narrow phase -> solve/integrate/phase logic); inside a stage the warps take environments from
// a shared counter, so a warp stuck on a long solve does not hold the others back, and all warps of the
// block execute the same code region at the same time (the kernel is ~10x the 32 KB instruction cache;
// walking it together is what keeps instruction fetch from dominating).  Measured alternatives, all slower
// on 4096 envs x 100 substeps mid-push: no barriers + a ready ring of environments
// (2.0x slower: 16 warps in 16 code regions), warps leaving the barrier protocol during long solves
// (1.15-1.25x slower: persistent stragglers become the tail of the launch), two 8-warp blocks per SM (1.2x).
// RA: the instantiation whose long solves run ahead (large scenes, free-running launches of rollouts / asynchronous
// stepping).  The other one keeps the barrier state out of its way: plain barriers, nothing address-taken.
template <bool RA>
__global__ void __launch_bounds__(B2S_BLOCK_THREADS, B2S_MIN_BLOCKS) k_substeps(int n, int mode, float lin, float ang, int max_steps,
                                                                                 const uint8_t* __restrict__ env_mask, int free_run) {
  __shared__ int s_cnt[5];    // hand-out counters of the three stages, candidate pairs of the block in this round, active slots
  __shared__ unsigned char s_order[256];   // stage C hand-out order of the slots: most solver work in the previous substep first
#ifdef B2S_PROF
  __shared__ int s_maxc;
  if (threadIdx.x == 0) s_maxc = 0;
#endif
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int Wn = blockDim.x >> 5;
  const int E = W.envs_per_block;
  const int e0 = blockIdx.x * E;
  const B2SParams& P = W.P;
  for (int slot = wib; slot < E; slot += Wn)
    if (lane < META_WORDS) env_meta(slot)[lane] = 0;
  if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0; sh_stop = 0; sh_arrived = 0; sh_long = 0; }
  __syncthreads();
  int done_steps = 0;
  int any_next = 0;
  BarState bs;
  bs.g = 0; bs.long_on = 0; bs.stopped = 0; bs.slot = 0; bs.passed = 0;
#ifdef B2S_PROF
  bs.enabled = free_run ? 0 : -1;
#else
  bs.enabled = free_run ? ((RA && wib != 0) ? 1 : 0) : -1;
#endif
  if (!RA) bs.enabled = -1;                            // nobody polls: no arrival counting either
  int pending_slot = -1, pending_active = 0;   // environment of a long solve: it rejoins the hand-out at the next top of a round
  for (int s = 0;; ++s) {
    int any = any_next;
    if (pending_slot >= 0) {
      if (lane == 0) env_meta(pending_slot)[META_ACTIVE] = pending_active;
      pending_slot = -1;
    }
    if (s == 0) {
      for (int slot = wib; slot < E; slot += Wn) {
        const int e = W.env_map[e0 + slot];
        const bool valid = e >= 0;
        if (lane == 0) env_meta(slot)[META_ENV] = e;
        bool active;
        if (mode == MODE_RAW) active = valid && n > 0;
        else if (mode == MODE_ENV) active = valid && n > 0 && W.phase[valid ? e : 0] != B2S_PHASE_IDLE;
        else active = valid && (!env_mask || env_mask[valid ? e : 0]);
        if (lane == 0) {
          env_meta(slot)[META_ACTIVE] = active ? 1 : 0;
          const int32_t* st = W.solver_stats + (size_t)(valid ? e : 0) * 4;
          env_meta(slot)[META_COST] = valid ? st[1] * st[2] * 4 + st[3] : 0;
        }
        any |= active ? 1 : 0;
      }
    }
    any_next = 0;
    // Free-running launch (environments of a rollout are independent: nobody has to wait for anybody).  Every warp
    // publishes the substeps it has done to the global counter once per round; when the launch has executed its total,
    // all blocks stop within one round of each other, whatever their environments cost -- no SM idles at the end of a
    // launch behind the slowest block.  Which environment got how many substeps depends on the schedule; what an
    // environment computes does not.
    if (free_run && s > 0) {
      if (lane == 0 && done_steps) { atomicAdd(W.substeps, (unsigned long long)done_steps); }
      done_steps = 0;
      if (threadIdx.x == 0) sh_stop = (*(volatile unsigned long long*)W.substeps >= *(volatile unsigned long long*)W.free_target) ? 1 : 0;
    }
    if (!(RA ? bar_pass(bs, lane, any) : __syncthreads_or(any))) break;
    if (free_run && sh_stop) break;
    if (threadIdx.x == 0) s_cnt[2] = 0;
    // Stage C ends with the warp that finishes last, and an environment's solve is one warp's sequential chain: a long
    // solve that is picked up late IS the end of the stage.  So the slots are handed out longest first, ranked by the
    // solver work (colours x iterations) of their previous substep, which the next one resembles.
    // (warp 0 ranks: the one warp that is never inside a long solve.  s_cnt[4] = active slots of this round: stage C hands
    // out exactly those, it must not look at META_ACTIVE again -- the environment of a long solve is switched back on by
    // its warp while others may still be in stage C)
    if (wib == 0) {
      int nact = 0;
      for (int t0 = 0; t0 < E; t0 += 32) {
        const int t = t0 + lane;
        const bool valid = t < E;
        const int act = valid ? env_meta(t)[META_ACTIVE] : 0;
        const int my = act ? env_meta(t)[META_COST] : -1;
        int rank = 0;
        for (int j = 0; j < E; ++j) {
          const int cj = env_meta(j)[META_ACTIVE] ? env_meta(j)[META_COST] : -1;
          rank += (cj > my || (cj == my && j < t)) ? 1 : 0;
        }
        if (valid) s_order[rank] = (unsigned char)t;
        nact += __popc(__ballot_sync(FULL, act != 0));
      }
      if (lane == 0) s_cnt[4] = nact;
    }
#ifdef B2S_PROF
    long long pstart_ = prof_now();
    if (threadIdx.x == 0) { atomicAdd(W.prof + 6, 1ull); W.prof[8 + blockIdx.x * 4 + 3] += 1ull; }
#endif
    // ---- stage A: controller + FK, body table, colliders, broad phase
    bool first0 = true, first1 = true, first2 = true;
    for (;;) {
      const int slot = grab_slot(&s_cnt[0], lane, wib, E, first0);
      first0 = false;
      if (slot >= E) break;
      int* meta = env_meta(slot);
      if (!meta[META_ACTIVE]) continue;
      const int np = stage_scene(meta[META_ENV], lane, slot | (wib << 16));
      if (lane == 0) { meta[META_NP] = np; atomicAdd(&s_cnt[3], np); }
    }
    PROF_STAGE(0)
    if (RA) bar_pass(bs, lane, 0); else __syncthreads();
    PROF_MARK(0)
    if (threadIdx.x == 0) s_cnt[0] = 0;
    // ---- stage B: narrow phase, one candidate pair per grab
    for (;;) {
      int pp = 0;
#if B2S_HALF
      const int slot = grab_pair(&s_cnt[1], lane, E, &pp, s_cnt[3], Wn);
      if (__all_sync(FULL, slot < 0)) break;
      if (slot >= 0) stage_narrow_pair(env_meta(slot)[META_ENV], lane, slot | (wib << 16), pp);
      __syncwarp();
#else
      const int slot = grab_pair(&s_cnt[1], lane, E, &pp);
      if (slot < 0) break;
      stage_narrow_pair(env_meta(slot)[META_ENV], lane, slot | (wib << 16), pp);
#endif
    }
    PROF_STAGE(1)
    if (RA) bar_pass(bs, lane, 0); else __syncthreads();
    PROF_MARK(1)
    if (threadIdx.x == 0) { s_cnt[1] = 0; s_cnt[3] = 0; }
    // ---- stage C: solve, integrate, phase machine / settle bookkeeping
    for (;;) {
      const int turn = grab_slot(&s_cnt[2], lane, wib, E, first2);
      first2 = false;
      if (turn >= s_cnt[4]) break;            // inactive slots rank last
      const int slot = s_order[turn];
      int* meta = env_meta(slot);
#ifdef B2S_PROF
      const long long penv_ = prof_now();
#endif
      const int e = meta[META_ENV];
      const int sw = slot | (wib << 16);
      const int ph = (mode == MODE_ENV) ? W.phase[e] : B2S_PHASE_IDLE;
      int C = 0, newn = 0;
      stage_narrow_merge(e, lane, sw, meta[META_NP], &C, &newn);
      bool ran_ahead = false;                      // the block went on without this environment: bs.g says where it is now
      if (RA) {
        bs.slot = slot; bs.passed = 0;
        substep_post_big<true>(e, lane, sw, C, newn, &bs);
        ran_ahead = bs.passed > 0;
      } else if (W.reg_rows) substep_post_reg(e, lane, sw, C, newn);
      else substep_post_big<false>(e, lane, sw, C, newn, nullptr);
      if (ran_ahead) s = (bs.g - 1) / 3;
      PROF_SEC0()
      ++done_steps;
      if (lane == 0) {
        const int32_t* st = W.solver_stats + (size_t)e * 4;
        meta[META_COST] = st[1] * st[2] * 4 + st[3];
        // what the next launch's deal ranks the environments by (k_assign_envs_free)
        if (free_run) W.work_ema[e] = 0.75f * W.work_ema[e] + 0.25f * (float)(st[1] * st[2]);
      }
      bool nxt = (s + 1 < n);
      if (mode == MODE_ENV) {
        int32_t* ps = W.phase_state + (size_t)e * 8;
        if (ph < B2S_PHASE_DONE) {
          if (W.num_steps[e] % P.steps_check == 0) {
          phase_logic(e, lane, sw);
          if (W.phase[e] == B2S_PHASE_DONE) {
            if (lane == 0) { W.phase[e] = B2S_PHASE_SETTLE; ps[2] = 0; ps[3] = 0; }
            __syncwarp();
          }
          }
        } else {
          // wait_until_stable: after 'done' (SETTLE), or of a rollout's reset (RESET_DROP with the loose thresholds, then RESET_WAIT)
          const bool drop = (ph == B2S_PHASE_RESET_DROP);
          const float slin = drop ? W.ro.drop_lin : P.stable_lin_threshold, sang = drop ? W.ro.drop_ang : P.stable_ang_threshold;
          const int smax = drop ? W.ro.drop_max_steps : P.stable_max_steps;
          int s2 = ps[2] + 1, s3 = ps[3];
          __syncwarp();
          bool fin = false;
          if (s2 >= P.stable_check_after) {
            if (all_stable(e, lane, slin, sang)) s3 += 1;
            if (s3 >= P.stable_min_steps || s2 >= smax) fin = true;
          }
          if (lane == 0) { ps[2] = s2; ps[3] = s3; }
          __syncwarp();
          if (fin) {
            if (ph == B2S_PHASE_SETTLE) {
              finish_action(e, lane);
              if (W.ro.enabled == RO_EPISODES) rollout_next(e, lane);   // reward, record, next action or next episode: the env goes on in this launch
              else if (W.ro.enabled == RO_ASYNC) async_after_action(e, lane);   // reward + observation, then the env waits for the host
            } else if (drop) {
              if (lane == 0) { W.phase[e] = B2S_PHASE_RESET_WAIT; ps[2] = 0; ps[3] = 0; }
              __syncwarp();
            } else {
              rollout_after_reset(e, lane);
            }
          }
        }
        nxt = nxt && (W.phase[e] != B2S_PHASE_IDLE);
      } else if (mode == MODE_SETTLE) {   // Simulator.wait_until_stable for this env
        const int steps = meta[META_SETTLE_STEPS] + 1;
        int stable = meta[META_SETTLE_STABLE];
        __syncwarp();
        bool fin = false;
        if (steps >= P.stable_check_after) {
          if (all_stable(e, lane, lin, ang)) ++stable;
          if (stable >= P.stable_min_steps || steps >= max_steps) fin = true;
        }
        if (lane == 0) { meta[META_SETTLE_STEPS] = steps; meta[META_SETTLE_STABLE] = stable; meta[META_FINISHED] = fin ? 1 : 0; }
        __syncwarp();
        nxt = !fin;
      }
      // activity of this environment in the next substep, published by the warp that just stepped it (after a long
      // solve: at the next top of a round, the environment must not appear in the middle of one)
      if (ran_ahead) { pending_slot = slot; pending_active = nxt ? 1 : 0; }
      else if (lane == 0) meta[META_ACTIVE] = nxt ? 1 : 0;
      any_next |= nxt ? 1 : 0;
      PROF_SEC(12)
      if (RA && bs.stopped) break;
      if (RA && ran_ahead) { while (bs.g % 3 != 0) bar_pass(bs, lane, 0); }     // sit out what is left of the block's stages A and B
#ifdef B2S_PROF
      {
        // histogram of the stage-C time of an environment (2 us bins) and of its start time within the stage (4 us bins)
        const long long now2_ = prof_now();
        if (lane == 0) {
          atomicMax(&s_maxc, (int)(now2_ - penv_));
          atomicAdd(W.prof + 8 + 4096 + 16 + min(63, (int)((now2_ - penv_) / 2000)), 1ull);
          atomicAdd(W.prof + 8 + 4096 + 16 + 64 + min(31, (int)((penv_ - pstart_) / 4000)), 1ull);
          if (turn == 0) atomicAdd(W.prof + 8 + 4096 + 16 + 96 + min(31, (int)((now2_ - penv_) / 4000)), 1ull);
        }
      }
#endif
    }
    if (RA && bs.stopped) break;
#ifdef B2S_PROF
    PROF_STAGE(2)
    __syncthreads();
    PROF_MARK(2)
    if (threadIdx.x == 0) { atomicAdd(W.prof + 7, (unsigned long long)s_maxc); s_maxc = 0; }
#endif
  }
  if (lane == 0 && done_steps) atomicAdd(W.substeps, (unsigned long long)done_steps);
}

#undef W
// The world description of a launch sits in the device's constant bank (g_W).  One copy per device serves every
// world and stream of the process, so launches that do not follow each other on one stream are ordered here:
// before g_W is overwritten for a launch on another stream (or for another world), that stream waits for the last
// kernel that read it.  Worlds on different devices never meet (a __constant__ symbol exists once per device).
struct DevLaunch { cudaEvent_t done; cudaStream_t stream; bool have; };
static DevLaunch g_launch[B2S_MAX_DEVICES];
static size_t g_smem_configured[B2S_MAX_DEVICES], g_smem_configured_ra[B2S_MAX_DEVICES];
static std::mutex g_launch_mutex;

void b2s_launch_substeps(const DWorld& W, int n, int mode, float lin, float ang, int max_steps, const uint8_t* env_mask, cudaStream_t s,
                         int free_chunk, bool run_ahead) {
  std::lock_guard<std::mutex> lock(g_launch_mutex);
  const int wpb = W.P.warps_per_block;
  const int blocks = W.num_blocks;
  size_t smem = b2s_smem_bytes(W);
  // long solves run ahead in the free-running launches of large scenes (rows in records: the solves are long and uneven);
  // with register-resident rows it measured neutral and the plumbing cost 1.2 %
  const bool ra = free_chunk > 0 && mode == MODE_ENV && run_ahead && !W.reg_rows;
  if (ra) b2s_opt_in_smem(k_substeps<true>, smem, g_smem_configured_ra); else b2s_opt_in_smem(k_substeps<false>, smem, g_smem_configured);
  int dev = 0;
  cudaGetDevice(&dev);
  DevLaunch* L = (dev >= 0 && dev < B2S_MAX_DEVICES) ? &g_launch[dev] : nullptr;
  if (L && L->have && L->stream != s) cudaStreamWaitEvent(s, L->done, 0);
  if (mode != MODE_ENV) free_chunk = 0;
  int resident = 1 << 30;
  if (free_chunk > 0 && dev >= 0 && dev < B2S_MAX_DEVICES) {
    // a free-running launch only uses blocks that are resident together (k_assign_envs_free)
    static int g_resident[B2S_MAX_DEVICES];
    static size_t g_resident_key[B2S_MAX_DEVICES];
    const size_t key = smem * 64 + (size_t)wpb;
    if (g_resident_key[dev] != key) {
      int occ = 1, sms = 148;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_substeps<false>, wpb * 32, smem);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      g_resident[dev] = (occ > 0 ? occ : 1) * (sms > 0 ? sms : 148);
      g_resident_key[dev] = key;
    }
    resident = g_resident[dev];
  }
  const int run_blocks = b2s_launch_assign_envs(W, mode, s, free_chunk, resident);
  cudaMemcpyToSymbolAsync(g_W, &W, sizeof(DWorld), 0, cudaMemcpyHostToDevice, s);
  if (ra) k_substeps<true><<<run_blocks, wpb * 32, smem, s>>>(n, mode, lin, ang, max_steps, env_mask, 2);
  else k_substeps<false><<<run_blocks, wpb * 32, smem, s>>>(n, mode, lin, ang, max_steps, env_mask, free_chunk > 0 ? 1 : 0);
  if (L) {
    if (!L->have) { if (cudaEventCreateWithFlags(&L->done, cudaEventDisableTiming) == cudaSuccess) L->have = true; }
    if (L->have) { cudaEventRecord(L->done, s); L->stream = s; }
  }
}

size_t b2s_smem_bytes(const DWorld& W) {
  return ((size_t)W.envs_per_block * (W.sm.words_env + META_WORDS) + (size_t)W.P.warps_per_block * W.sm.words_warp) * 4;
}
