/* b2o_obs.cpp -- CPU oracle: depth / segmentation raster and segmented point cloud.
 *
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h).  PARITY UNPINNED vs pybullet's renderer.
 *
 * Restates BulletCamera._frames (robovat/simulation/camera/bullet_camera.py:188-235):
 *   - calibration x_cam = R x_world + t, intrinsics K (set_calibration :237-258);
 *   - the OpenGL round trip (projection matrix with the principal point mirrored + 180 degree image flip,
 *     :28-83, :217-220) is a plain pinhole image again, so pixel (row v, col u) looks along
 *     K^-1 [u, v, 1] -- the same convention Camera.deproject_depth_image / project_point use
 *     (robovat/perception/camera/camera.py:170-244: integer pixel coordinates are pixel centres);
 *   - depth is the eye-space z of the hit (the z_e linearisation :225-229), clipped to [near, far];
 *     pixels that hit nothing read the far plane (z_b = 1 -> z_e = far) and segmentation 255
 *     (-1 cast to uint8, :212);
 *   - the segmentation value is the body unique id in the reference's loading order
 *     (ground, table, movables, tiles, arm).
 * Geometry = the convex hulls' face planes (no collision margin).  One ray per pixel, every collider.
 *
 * Point cloud: SegmentedPointCloudObs.get_observation (robovat/observations/camera_obs.py:182-212):
 * deproject, group pixels by movable, sample P points per body (zeros when the body is invisible).
 */
#include <math.h>
#include <string.h>

#include <vector>

#include "b2o_world.h"

namespace b2o {

struct RayCollider { int pbeg, pend, uid; };

static int body_uid(const World& w, int e, int slot) {
  const int n = w.num_movables[e];
  int first_tile = w.Ns;
  for (int s = 0; s < w.Ns; ++s) if (w.S.static_flags[s] & B2S_STATIC_IS_TILE) { first_tile = s; break; }
  if (slot < w.Ns) return (slot < first_tile) ? slot : slot + n;
  if (slot < w.Ns + w.L) return w.Ns + n;
  return first_tile + (slot - w.Ns - w.L);
}

/* camera-space planes of every hull of every body of env e */
static void build_planes(World& w, int e, std::vector<float>& planes, std::vector<RayCollider>& cols) {
  const Scene& S = w.S;
  const float* cam = &w.cam[(size_t)e * 21];
  M3 R; R.r0 = v3(cam[9], cam[10], cam[11]); R.r1 = v3(cam[12], cam[13], cam[14]); R.r2 = v3(cam[15], cam[16], cam[17]);
  const V3 t = v3(cam[18], cam[19], cam[20]);
  float q[7], qd[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 7; ++j) q[j] = w.joint_state[(0 * 7 + j) * w.B + e];
  std::vector<float> lp((w.L + 1) * 7);
  arm_fk(w, q, qd, lp.data(), NULL);
  auto add_body = [&](int slot, int asset, V3 pos, Q4 quat, float scale) {
    const Asset& A = S.assets[asset];
    const M3 Rb = q_to_m3(quat);
    for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) {
      const Hull& H = S.hulls[h];
      RayCollider rc;
      rc.pbeg = (int)planes.size() / 4; rc.uid = body_uid(w, e, slot);
      for (int p = H.poff; p < H.poff + H.pcnt; ++p) {
        V3 nl = v3(S.planes[p * 4], S.planes[p * 4 + 1], S.planes[p * 4 + 2]);
        V3 nw = mmul(Rb, nl);
        float dw = S.planes[p * 4 + 3] * scale + dot(nw, pos);
        V3 nc = mmul(R, nw);
        float dc = dw + dot(nc, t);
        planes.push_back(nc.x); planes.push_back(nc.y); planes.push_back(nc.z); planes.push_back(dc);
      }
      rc.pend = (int)planes.size() / 4;
      cols.push_back(rc);
    }
  };
  for (int s = 0; s < w.Ns; ++s) {
    const float* sp = &S.static_pose[s * 7];
    float dz = (S.static_flags[s] & B2S_STATIC_ON_TABLE) ? w.table_dz[e] : 0.0f;
    add_body(s, S.static_asset[s], v3(sp[0], sp[1], sp[2] + dz), q4(sp[3], sp[4], sp[5], sp[6]), 1.0f);
  }
  for (int k = 0; k < w.L; ++k)
    add_body(w.Ns + k, S.d.link_asset[k], v3(lp[k * 7], lp[k * 7 + 1], lp[k * 7 + 2]), q4(lp[k * 7 + 3], lp[k * 7 + 4], lp[k * 7 + 5], lp[k * 7 + 6]), 1.0f);
  const size_t B = w.B, N = w.Nmax;
  for (int i = 0; i < w.num_movables[e]; ++i) {
    float af = w.mov_params[(0 * B + e) * N + i];
    int32_t asset; memcpy(&asset, &af, 4);
    add_body(w.Ns + w.L + i, asset,
             v3(w.body_state[(0 * B + e) * N + i], w.body_state[(1 * B + e) * N + i], w.body_state[(2 * B + e) * N + i]),
             q4(w.body_state[(3 * B + e) * N + i], w.body_state[(4 * B + e) * N + i], w.body_state[(5 * B + e) * N + i], w.body_state[(6 * B + e) * N + i]),
             w.mov_params[(1 * B + e) * N + i]);
  }
}

void render(World& w, int e) {
  const B2SParams& P = w.P;
  const int H = P.cam_height, Wd = P.cam_width;
  const float* cam = &w.cam[(size_t)e * 21];
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  std::vector<float> planes;
  std::vector<RayCollider> cols;
  build_planes(w, e, planes, cols);
  float* depth = &w.depth[(size_t)e * H * Wd];
  uint8_t* seg = &w.segmask[(size_t)e * H * Wd];
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < Wd; ++u) {
      const float dy = ((float)v - cy) / fy;
      const float dx = (((float)u - cx) - sk * dy) / fx;
      const V3 dir = v3(dx, dy, 1.0f);
      float best = P.cam_far;
      int uid = 255;
      for (size_t c = 0; c < cols.size(); ++c) {
        float t0 = P.cam_near, t1 = best;
        bool miss = false;
        for (int p = cols[c].pbeg; p < cols[c].pend && !miss; ++p) {
          const float* pl = &planes[(size_t)p * 4];
          float den = dot(v3(pl[0], pl[1], pl[2]), dir);
          float dc = pl[3];
          if (den < 0.0f) { float tt = dc / den; if (tt > t0) t0 = tt; }
          else if (den > 0.0f) { float tt = dc / den; if (tt < t1) t1 = tt; }
          else if (dc < 0.0f) miss = true;
          if (t0 > t1) miss = true;
        }
        if (!miss && t0 < best) { best = t0; uid = cols[c].uid & 255; }
      }
      depth[(size_t)v * Wd + u] = best;
      seg[(size_t)v * Wd + u] = (uint8_t)uid;
    }
}

void point_cloud(World& w, int e, uint64_t seed) {
  const B2SParams& P = w.P;
  const int H = P.cam_height, Wd = P.cam_width, NP = P.num_points, N = w.Nmax;
  const float* cam = &w.cam[(size_t)e * 21];
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  M3 R; R.r0 = v3(cam[9], cam[10], cam[11]); R.r1 = v3(cam[12], cam[13], cam[14]); R.r2 = v3(cam[15], cam[16], cam[17]);
  const V3 t = v3(cam[18], cam[19], cam[20]);
  const float* depth = &w.depth[(size_t)e * H * Wd];
  const uint8_t* seg = &w.segmask[(size_t)e * H * Wd];
  float* out = &w.point_cloud[(size_t)e * N * NP * 3];
  for (int i = 0; i < N; ++i) {
    float* o = out + (size_t)i * NP * 3;
    memset(o, 0, sizeof(float) * NP * 3);
    if (i >= w.num_movables[e]) continue;
    const int uid = body_uid(w, e, w.Ns + w.L + i) & 255;
    std::vector<int> idx;
    for (int k = 0; k < H * Wd; ++k) {
      if (seg[k] != uid) continue;
      if (P.use_crop) {                                /* OBS.CROP_MIN / CROP_MAX on the world-frame cloud (camera_obs.py:187-193) */
        const int v = k / Wd, u = k % Wd;
        const float z = depth[k];
        const float dy = ((float)v - cy) / fy;
        const float dx = (((float)u - cx) - sk * dy) / fx;
        const V3 x = mtmul(R, v3(dx * z, dy * z, z) - t);
        if (!(x.x >= P.crop_min[0] && x.y >= P.crop_min[1] && x.z >= P.crop_min[2] &&
              x.x <= P.crop_max[0] && x.y <= P.crop_max[1] && x.z <= P.crop_max[2])) continue;
      }
      idx.push_back(k);
    }
    const int n = (int)idx.size();
    if (n == 0) continue;
    /* np.random.choice(n, P, replace = n < P) (perception/point_cloud_utils.py:23-39): Philox stream 3 */
    uint32_t ctr = 0;
    b2s_u4 r = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), 0u, 3u + 16u * (uint32_t)i, (uint32_t)(P.env_id_offset + e), 0u);
    const float u0 = b2s_u01(r.x);
    for (int j = 0; j < NP; ++j) {
      int k;
      if (n >= NP) k = (int)(((float)j + u0) * (float)n / (float)NP);       /* distinct pixels, random phase */
      else {
        if ((j & 3) == 0) r = b2s_philox((uint32_t)seed, (uint32_t)(seed >> 32), ++ctr, 3u + 16u * (uint32_t)i, (uint32_t)(P.env_id_offset + e), 0u);
        uint32_t bits = ((j & 3) == 0) ? r.x : ((j & 3) == 1) ? r.y : ((j & 3) == 2) ? r.z : r.w;
        k = (int)(b2s_u01(bits) * (float)n);
      }
      if (k >= n) k = n - 1;
      const int pix = idx[k];
      const int v = pix / Wd, u = pix % Wd;
      const float z = depth[pix];
      const float dy = ((float)v - cy) / fy;
      const float dx = (((float)u - cx) - sk * dy) / fx;
      V3 xc = v3(dx * z, dy * z, z);
      V3 xw = mtmul(R, xc - t);                       /* Camera.deproject_depth_image, camera.py:213-244 */
      o[j * 3] = xw.x; o[j * 3 + 1] = xw.y; o[j * 3 + 2] = xw.z;
    }
  }
}

}  // namespace b2o
