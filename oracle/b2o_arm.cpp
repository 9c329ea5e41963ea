/* b2o_arm.cpp -- CPU oracle: Sawyer limb kinematics, DLS inverse kinematics and
 * the per-substep controller state machine.
 *
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h).
 *
 * Control flow restated from robovat/simulation/controllable_body.py:
 *   ControllableBody.update            :387-413
 *   _check_link_target_done            :415-432
 *   _check_joint_target_done           :434-456
 *   _update_position_control           :458-466  (-> setJointMotorControlArray, POSITION_CONTROL)
 *   _update_ik                         :468-499  (-> calculateInverseKinematics)
 *   check_joints_reached               :501-537
 *   is_ready                           :565-595
 *   LinkTarget.set / JointTarget.set   :97-128 / :186-217
 * The arm is kinematic: the POSITION_CONTROL motor's velocity
 * target  v* = kp (q* - q)/dt + qd + kd (qd* - qd)  [upstream-recall of
 * btMultiBodyJointMotor] is taken as the joint velocity of the substep.
 */
#include <math.h>
#include <string.h>

#include "b2o_world.h"

namespace b2o {

struct Xf { V3 p; Q4 q; };
static Xf xf_from(const float* a) { Xf t; t.p = v3(a[0], a[1], a[2]); t.q = q4(a[3], a[4], a[5], a[6]); return t; }
#ifdef B2S_F64
static Xf xf_from(const abi_float* a) { Xf t; t.p = v3(a[0], a[1], a[2]); t.q = q4(a[3], a[4], a[5], a[6]); return t; }
#endif
static Xf xf_mul(Xf a, Xf b) { Xf t; t.p = a.p + qrot(a.q, b.p); t.q = qmul(a.q, b.q); return t; }
static void xf_store(Xf t, float* o) { o[0] = t.p.x; o[1] = t.p.y; o[2] = t.p.z; o[3] = t.q.x; o[4] = t.q.y; o[5] = t.q.z; o[6] = t.q.w; }

/* frames after each joint, world joint axes and origins */
static void fk_chain(const World& w, const float* q, Xf* frame, V3* axis_w, V3* origin_w) {
  const B2SSceneDesc& d = w.S.d;
  Xf T = xf_from(d.arm_base_pose);
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
    Xf Tj = xf_mul(T, xf_from(d.joint_origin[j]));
    V3 ax = v3(d.joint_axis[j][0], d.joint_axis[j][1], d.joint_axis[j][2]);
    axis_w[j] = qrot(Tj.q, ax);
    origin_w[j] = Tj.p;
    T.p = Tj.p;
    T.q = qmul(Tj.q, q_axis_angle(ax, q[j]));
    frame[j] = T;
  }
}

void arm_fk(const World& w, const float* q, const float* qd, float* link_poses, float* link_vel) {
  const B2SSceneDesc& d = w.S.d;
  Xf frame[B2S_NUM_JOINTS];
  V3 ax[B2S_NUM_JOINTS], org[B2S_NUM_JOINTS];
  fk_chain(w, q, frame, ax, org);
  Xf base = xf_from(d.arm_base_pose);
  for (int k = 0; k < w.L; ++k) {
    int jj = d.link_joint[k];
    Xf T = xf_mul(jj < 0 ? base : frame[jj], xf_from(d.link_pose[k]));
    xf_store(T, link_poses + k * 7);
    if (link_vel) {
      V3 v = v3(0, 0, 0), om = v3(0, 0, 0);
      for (int i = 0; i <= jj; ++i) {
        v = v + cross(ax[i], T.p - org[i]) * qd[i];
        om = om + ax[i] * qd[i];
      }
      float* o = link_vel + k * 6;
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = om.x; o[4] = om.y; o[5] = om.z;
    }
  }
  xf_store(xf_mul(frame[B2S_NUM_JOINTS - 1], xf_from(d.ee_pose)), link_poses + w.L * 7);
}

/* damped least squares from q_start (BulletPhysics.compute_inverse_kinematics,
 * bullet_physics.py:1203-1262 -> pybullet.calculateInverseKinematics, no null space) */
void arm_ik(const World& w, const float* target_pose, const float* q_start, float* q_out) {
  const B2SSceneDesc& d = w.S.d;
  const B2SParams& P = w.P;
  float q[B2S_NUM_JOINTS];
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q[j] = q_start[j];
  Xf tgt = xf_from(target_pose);
  Xf eel = xf_from(d.ee_pose);
  const float res2 = P.ik_residual * P.ik_residual;
  for (int it = 0; it < P.ik_max_iters; ++it) {
    Xf frame[B2S_NUM_JOINTS];
    V3 ax[B2S_NUM_JOINTS], org[B2S_NUM_JOINTS];
    fk_chain(w, q, frame, ax, org);
    Xf ee = xf_mul(frame[B2S_NUM_JOINTS - 1], eel);
    V3 ep = tgt.p - ee.p;
    V3 er = q_to_rotvec(qmul(tgt.q, qconj(ee.q)));
    if (len2(ep) < res2 && len2(er) < res2) break;
    float J[6][B2S_NUM_JOINTS];
    for (int j = 0; j < B2S_NUM_JOINTS; ++j) {
      V3 jv = cross(ax[j], ee.p - org[j]);
      J[0][j] = jv.x; J[1][j] = jv.y; J[2][j] = jv.z;
      J[3][j] = ax[j].x; J[4][j] = ax[j].y; J[5][j] = ax[j].z;
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
  for (int j = 0; j < B2S_NUM_JOINTS; ++j) q_out[j] = fminf(d.joint_upper[j], fmaxf(d.joint_lower[j], q[j]));
}

/* ---- controller state ------------------------------------------------------
 * ctrl floats: [0..7) link target pose, [7] link pos thr, [8] link vel thr,
 *              [9..16) joint target q*, [16] joint pos thr, [17] joint vel thr,
 *              [18..25) motor q*, [25..32) motor qd*
 * ctrl_flags : link_active, joint_active, joint_qd_none, motor_active
 * ctrl_time  : link start, link stop, joint start, joint stop, gripper ready */
static float* C_(World& w, int e) { return &w.ctrl[(size_t)e * B2S_CTRL_FLOATS]; }
static int32_t* F_(World& w, int e) { return &w.ctrl_flags[(size_t)e * 4]; }
static double* T_(World& w, int e) { return &w.ctrl_time[(size_t)e * 5]; }
static double now(const World& w, int e) { return w.P.time_step * (double)w.num_steps[e]; }

void arm_reset_targets(World& w, int e) { F_(w, e)[0] = 0; F_(w, e)[1] = 0; }

void arm_set_link_target(World& w, int e, const float* pose) {
  float* c = C_(w, e);
  for (int k = 0; k < 7; ++k) c[k] = pose[k];
  c[7] = w.P.joint_pos_threshold; c[8] = w.P.joint_vel_threshold;
  T_(w, e)[0] = now(w, e); T_(w, e)[1] = T_(w, e)[0] + (double)w.P.limb_timeout;
  F_(w, e)[0] = 1;
}

void arm_set_joint_target(World& w, int e, const float* q) {
  float* c = C_(w, e);
  for (int k = 0; k < 7; ++k) c[9 + k] = q[k];
  c[16] = w.P.joint_pos_threshold; c[17] = w.P.joint_vel_threshold;
  T_(w, e)[2] = now(w, e); T_(w, e)[3] = T_(w, e)[2] + (double)w.P.limb_timeout;
  F_(w, e)[1] = 1; F_(w, e)[2] = 0;
}

static bool joints_reached(World& w, int e) {
  if (!F_(w, e)[1]) return true;
  const float* c = C_(w, e);
  for (int j = 0; j < 7; ++j) {
    float q = w.joint_state[(0 * 7 + j) * w.B + e], qd = w.joint_state[(1 * 7 + j) * w.B + e];
    bool pr = fabsf(c[9 + j] - q) < c[16];
    bool vr = F_(w, e)[2] ? true : (fabsf(0.0f - qd) < c[17]);
    if (!(pr && vr)) return false;
  }
  return true;
}
static bool link_done(World& w, int e) {
  if (!F_(w, e)[0]) return true;               /* stop_time is None */
  return now(w, e) >= T_(w, e)[1];
}
static bool joint_done(World& w, int e) {
  if (!F_(w, e)[1]) return true;
  if (now(w, e) >= T_(w, e)[3]) return true;
  return joints_reached(w, e);
}

void arm_update(World& w, int e) {
  const B2SParams& P = w.P;
  const B2SSceneDesc& d = w.S.d;
  float* c = C_(w, e);
  int32_t* f = F_(w, e);
  const int n = w.num_steps[e];
  bool ik_updated = false;
  if (f[0] && n % P.check_done_interval == 0) { if (link_done(w, e)) f[0] = 0; }
  if (f[0] && (n % P.ik_interval == 0 || !f[1])) {
    float q[7], qo[7];
    for (int j = 0; j < 7; ++j) q[j] = w.joint_state[(0 * 7 + j) * w.B + e];
    arm_ik(w, c, q, qo);
    for (int j = 0; j < 7; ++j) c[9 + j] = qo[j];
    c[16] = c[7]; c[17] = c[8];
    T_(w, e)[2] = T_(w, e)[0]; T_(w, e)[3] = T_(w, e)[1];
    f[1] = 1; f[2] = 0;
    ik_updated = true;
    if (joints_reached(w, e)) f[0] = 0;          /* LinkTarget.pop() with an empty queue */
  }
  if (f[1] && (n % P.check_done_interval == 0 || ik_updated)) { if (joint_done(w, e)) f[1] = 0; }
  if (f[1]) {                                   /* _update_position_control */
    for (int j = 0; j < 7; ++j) { c[18 + j] = c[9 + j]; c[25 + j] = 0.0f; }
    f[3] = 1;
  }
  /* motor: persists with its last command, like a pybullet joint motor */
  const float dt = (float)P.time_step;
  float vj[7], scale = 1.0f;
  for (int j = 0; j < 7; ++j) {
    float q = w.joint_state[(0 * 7 + j) * w.B + e], qd = w.joint_state[(1 * 7 + j) * w.B + e];
    float v = 0.0f;
    if (f[3]) {
      v = (P.position_gain * (c[18 + j] - q) / dt + qd) + P.velocity_gain * (c[25 + j] - qd);
      if (P.clamp_joint_velocity) {
        /* one common scale for all joints keeps the joint-space direction of the move */
        float vm = P.limb_velocity_ratio * d.joint_max_velocity[j];
        float a = fabsf(v);
        if (a > vm) scale = fminf(scale, vm / a);
      }
    }
    vj[j] = v;
  }
  for (int j = 0; j < 7; ++j) {
    float q = w.joint_state[(0 * 7 + j) * w.B + e];
    float v = vj[j] * scale;
    if (f[3]) {
      float qn = q + v * dt;
      if (qn > d.joint_upper[j]) v = (d.joint_upper[j] - q) / dt;
      if (qn < d.joint_lower[j]) v = (d.joint_lower[j] - q) / dt;
    }
    w.joint_state[(1 * 7 + j) * w.B + e] = v;
  }
}

/* SawyerSim.is_limb_ready -> ControllableBody.is_ready(limb joints), side effects included */
int arm_is_ready(World& w, int e) {
  int32_t* f = F_(w, e);
  if (link_done(w, e)) f[0] = 0;
  if (joint_done(w, e)) f[1] = 0;
  return (!f[0] && !f[1]) ? 1 : 0;
}

}  // namespace b2o
