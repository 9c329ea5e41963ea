// b2s_obs.cu -- depth / segmentation raster (K8) and segmented point cloud (K9).
//
// Replaces BulletCamera._frames (robovat/simulation/camera/bullet_camera.py:188-235: pybullet.getCameraImage
// = Bullet's CPU TinyRenderer, 180 degree flip, depth linearisation) and SegmentedPointCloudObs.get_observation
// (robovat/observations/camera_obs.py:182-212: numpy deprojection + per-body grouping + random down-sampling).
//
// Raster: direct pinhole ray casting, no OpenGL round trip.  A block renders 32x8 pixel tiles of one
// environment (its share of the image): the block first transforms every hull face plane of the scene into the CAMERA frame in shared
// memory (so a pixel's ray/plane test is one 3-term dot product and one divide), then every thread clips its
// ray against the convex hulls, with a conservative bounding-sphere reject per hull.  Outputs are written
// coalesced: 4 B depth + 1 B segmentation per pixel is the only HBM traffic that scales with the image
// (81 920 B per 128x128 frame), which is what bounds this kernel once culling is good enough.
// Arithmetic per plane and per pixel is ordered exactly as in oracle/b2o_obs.cpp: depth and masks are bit-exact.
#include "b2s_dev.cuh"

#define TILE_W 32
#define TILE_H 8

struct RayCol { int pbeg, pend, uid; float cx, cy, cz, r2; };

__device__ __forceinline__ int body_uid(const DWorld& W, int n, int first_tile, int slot) {
  if (slot < W.Ns) return (slot < first_tile) ? slot : slot + n;
  if (slot < W.Ns + W.L) return W.Ns + n;
  return first_tile + (slot - W.Ns - W.L);
}

extern __shared__ float4 ray_smem[];

__global__ void __launch_bounds__(TILE_W* TILE_H) k_render(const __grid_constant__ DWorld W, int tiles_x, int num_tiles, int max_cols) {
  const int e = blockIdx.y;
  const int tid = threadIdx.x;
  const B2SParams& P = W.P;
  const int H = P.cam_height, Wd = P.cam_width;
  float4* planes = ray_smem;                                  // [max_ray_planes]
  RayCol* cols = (RayCol*)(ray_smem + W.max_ray_planes);      // [max_cols]
  __shared__ int s_ncol;
  __shared__ int s_slot[256], s_hull[256];
  const float* cam = W.cam + (size_t)e * 21;
  M3 R; R.r0 = v3(cam[9], cam[10], cam[11]); R.r1 = v3(cam[12], cam[13], cam[14]); R.r2 = v3(cam[15], cam[16], cam[17]);
  const V3 t = v3(cam[18], cam[19], cam[20]);
  const int n = W.buf.num_movables[e];
  int first_tile = W.Ns;
  for (int s = 0; s < W.Ns; ++s) if (W.static_flags[s] & B2S_STATIC_IS_TILE) { first_tile = s; break; }
  // 1. collider list in the oracle's order: statics (incl. visual-only), arm links, movables
  if (tid == 0) {
    int nc = 0, np = 0;
    auto push = [&](int slot, int asset) {
      const DAsset& A = W.assets[asset];
      for (int h = A.hoff; h < A.hoff + A.hcnt && nc < max_cols; ++h) {
        s_slot[nc] = slot; s_hull[nc] = h;
        cols[nc].pbeg = np; np += W.hulls[h].pcnt; cols[nc].pend = np;
        ++nc;
      }
    };
    for (int s = 0; s < W.Ns; ++s) push(s, W.static_asset[s]);
    for (int k = 0; k < W.L; ++k) push(W.Ns + k, W.arm->link_asset[k]);
    for (int i = 0; i < n; ++i) push(W.Ns + W.L + i, __float_as_int(W.mov_params[((size_t)0 * W.B + e) * W.Nmax + i]));
    s_ncol = nc;
  }
  __syncthreads();
  const int ncol = s_ncol;
  // 2. camera-space planes + bounding spheres, one collider per thread
  for (int c = tid; c < ncol; c += blockDim.x) {
    const int slot = s_slot[c];
    const DHull& Hh = W.hulls[s_hull[c]];
    V3 pos; Q4 quat; float scale = 1.0f;
    if (slot < W.Ns) {
      const float* sp = W.static_pose + slot * 7;
      float dz = (W.static_flags[slot] & B2S_STATIC_ON_TABLE) ? W.table_dz[e] : 0.0f;
      pos = v3(sp[0], sp[1], sp[2] + dz); quat = q4(sp[3], sp[4], sp[5], sp[6]);
    } else if (slot < W.Ns + W.L) {
      const float* lp = W.link_poses + ((size_t)e * (W.L + 1) + (slot - W.Ns)) * 7;
      pos = v3(lp[0], lp[1], lp[2]); quat = q4(lp[3], lp[4], lp[5], lp[6]);
    } else {
      const int i = slot - W.Ns - W.L;
      const size_t B = W.B, N = W.Nmax;
      const float* bs = W.buf.body_state;
      pos = v3(bs[(0 * B + e) * N + i], bs[(1 * B + e) * N + i], bs[(2 * B + e) * N + i]);
      quat = q4(bs[(3 * B + e) * N + i], bs[(4 * B + e) * N + i], bs[(5 * B + e) * N + i], bs[(6 * B + e) * N + i]);
      scale = W.mov_params[((size_t)1 * B + e) * N + i];
    }
    const M3 Rb = q_to_m3(quat);
    for (int p = 0; p < Hh.pcnt; ++p) {
      float4 pl = W.planes[Hh.poff + p];
      V3 nw = mmul(Rb, v3(pl.x, pl.y, pl.z));
      float dw = pl.w * scale + dot(nw, pos);
      V3 nc = mmul(R, nw);
      float dc = dw + dot(nc, t);
      planes[cols[c].pbeg + p] = make_float4(nc.x, nc.y, nc.z, dc);
    }
    V3 cc = mmul(R, pos) + t;
    float rad = Hh.rad * scale * 1.001f + 1e-4f;               // conservative: the hull lies inside this sphere
    cols[c].cx = cc.x; cols[c].cy = cc.y; cols[c].cz = cc.z; cols[c].r2 = rad * rad;
    cols[c].uid = body_uid(W, n, first_tile, slot) & 255;
  }
  __syncthreads();
  // 3. one ray per thread and tile; the block walks its share of the image's tiles, so the scene set-up above
  //    (one block per 256 pixels before: 64 x per 128x128 frame) is paid once per block: 1.85 -> 1.03 ms for 2048
  //    envs at 128x128.  What is left is ALU work -- every ray is clipped against the hulls whose bounding sphere it
  //    crosses (always ground and table, plus the arm links in view), six planes and up to six IEEE divisions
  //    each; a per-tile cull by the spheres' pixel rectangles changed nothing (1.06 ms) and was dropped again.
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int u = tx * TILE_W + (tid % TILE_W), v = ty * TILE_H + (tid / TILE_W);
  if (u >= Wd || v >= H) continue;
  const float dy = ((float)v - cy) / fy;
  const float dx = (((float)u - cx) - sk * dy) / fx;
  const V3 dir = v3(dx, dy, 1.0f);
  const float inv_d2 = 1.0f / dot(dir, dir);
  float best = P.cam_far;
  int uid = 255;
  for (int c = 0; c < ncol; ++c) {
    const RayCol rc = cols[c];
    // conservative sphere reject (never decides a hit; slack absorbs rounding)
    const V3 cc = v3(rc.cx, rc.cy, rc.cz);
    const float along = dot(cc, dir);
    const float perp2 = dot(cc, cc) - along * along * inv_d2;
    if (perp2 > rc.r2 * 1.01f + 1e-6f) continue;
    float t0 = P.cam_near, t1 = best;
    bool miss = false;
    for (int p = rc.pbeg; p < rc.pend && !miss; ++p) {
      const float4 pl = planes[p];
      float den = dot(v3(pl.x, pl.y, pl.z), dir);
      float dc = pl.w;
      if (den < 0.0f) { float tt = dc / den; if (tt > t0) t0 = tt; }
      else if (den > 0.0f) { float tt = dc / den; if (tt < t1) t1 = tt; }
      else if (dc < 0.0f) miss = true;
      if (t0 > t1) miss = true;
    }
    if (!miss && t0 < best) { best = t0; uid = rc.uid; }
  }
  W.buf.depth[((size_t)e * H + v) * Wd + u] = best;
  W.buf.segmask[((size_t)e * H + v) * Wd + u] = (uint8_t)uid;
  }
}

// ---- segmented point cloud: one warp per (env, movable) ------------------------------------------------
extern __shared__ int pc_smem[];

__global__ void k_point_cloud(const __grid_constant__ DWorld W, uint64_t seed) {
  const int e = blockIdx.x, i = blockIdx.y;
  const int lane = threadIdx.x;
  const B2SParams& P = W.P;
  const int H = P.cam_height, Wd = P.cam_width, NP = P.num_points, N = W.Nmax;
  const int npix = H * Wd, nchunk = (npix + 31) / 32;
  float* o = W.buf.point_cloud + ((size_t)e * N + i) * NP * 3;
  for (int k = lane; k < NP * 3; k += 32) o[k] = 0.0f;
  const int nm = W.buf.num_movables[e];
  if (i >= nm) return;
  int first_tile = W.Ns;
  for (int s = 0; s < W.Ns; ++s) if (W.static_flags[s] & B2S_STATIC_IS_TILE) { first_tile = s; break; }
  const int uid = body_uid(W, nm, first_tile, W.Ns + W.L + i) & 255;
  const uint8_t* seg = W.buf.segmask + (size_t)e * npix;
  const float* depth = W.buf.depth + (size_t)e * npix;
  int* prefix = pc_smem;                      // [nchunk + 1] exclusive prefix of matches per 32-pixel chunk
  const float* cam = W.cam + (size_t)e * 21;
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  M3 R; R.r0 = v3(cam[9], cam[10], cam[11]); R.r1 = v3(cam[12], cam[13], cam[14]); R.r2 = v3(cam[15], cam[16], cam[17]);
  const V3 t = v3(cam[18], cam[19], cam[20]);
  // Camera.deproject_depth_image for one pixel (camera.py:213-244)
  auto deproject = [&](int pix) {
    const int v = pix / Wd, u = pix % Wd;
    const float z = depth[pix];
    const float dy = ((float)v - cy) / fy;
    const float dx = (((float)u - cx) - sk * dy) / fx;
    return mtmul(R, v3(dx * z, dy * z, z) - t);
  };
  // a pixel belongs to body i if the segmentation says so and, with OBS.CROP_MIN/MAX, its point lies inside the crop box
  // (the reference crops the whole cloud before it groups by label, camera_obs.py:187-201)
  const bool crop = P.use_crop != 0;
  auto match = [&](int pix) {
    if (pix >= npix || seg[pix] != uid) return false;
    if (!crop) return true;
    const V3 x = deproject(pix);
    return x.x >= P.crop_min[0] && x.y >= P.crop_min[1] && x.z >= P.crop_min[2] &&
           x.x <= P.crop_max[0] && x.y <= P.crop_max[1] && x.z <= P.crop_max[2];
  };
  // pass 1: chunk counts via ballot, running total kept by all lanes
  int total = 0;
  for (int c = 0; c < nchunk; ++c) {
    const bool m = match(c * 32 + lane);
    unsigned b = __ballot_sync(FULL, m);
    if (lane == 0) prefix[c] = total;
    total += __popc(b);
  }
  if (lane == 0) prefix[nchunk] = total;
  __syncwarp();
  const int n = total;
  if (n == 0) return;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), env = (uint32_t)(P.env_id_offset + e);
  const float u0 = b2s_u01(b2s_philox(k0, k1, 0u, 3u + 16u * (uint32_t)i, env, 0u).x);
  for (int j = lane; j < NP; j += 32) {
    int k;
    if (n >= NP) k = (int)(((float)j + u0) * (float)n / (float)NP);
    else {
      b2s_u4 r = b2s_philox(k0, k1, (uint32_t)(j / 4 + 1), 3u + 16u * (uint32_t)i, env, 0u);
      uint32_t bits = ((j & 3) == 0) ? r.x : ((j & 3) == 1) ? r.y : ((j & 3) == 2) ? r.z : r.w;
      k = (int)(b2s_u01(bits) * (float)n);
    }
    if (k >= n) k = n - 1;
    // chunk holding the k-th match: last c with prefix[c] <= k
    int lo = 0, hi = nchunk - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (prefix[mid] <= k) lo = mid; else hi = mid - 1; }
    int rem = k - prefix[lo], pix = lo * 32;
    for (int q = 0; q < 32; ++q) {
      int pp = lo * 32 + q;
      if (match(pp)) { if (rem == 0) { pix = pp; break; } --rem; }
    }
    const V3 xw = deproject(pix);
    o[j * 3] = xw.x; o[j * 3 + 1] = xw.y; o[j * 3 + 2] = xw.z;
  }
}

void b2s_launch_render(const DWorld& W, cudaStream_t s) {
  const int tiles_x = (W.P.cam_width + TILE_W - 1) / TILE_W, tiles_y = (W.P.cam_height + TILE_H - 1) / TILE_H;
  const int max_cols = W.max_ray_cols;
  size_t smem = (size_t)W.max_ray_planes * sizeof(float4) + (size_t)max_cols * sizeof(RayCol);
  static size_t configured[B2S_MAX_DEVICES];
  b2s_opt_in_smem(k_render, smem, configured);
  b2s_launch_fk(W, s);                      // link poses of the current joint state
  // blocks per environment: enough blocks to fill the GPU a few times over, at most one per tile
  const int num_tiles = tiles_x * tiles_y;
  int per_env = (8 * 148 + W.B - 1) / W.B;
  per_env = per_env < 1 ? 1 : (per_env > num_tiles ? num_tiles : per_env);
#ifdef B2S_RENDER_PER_TILE
  per_env = num_tiles;                      // tuning: the previous mapping, one block per tile
#endif
  k_render<<<dim3(per_env, W.B), TILE_W * TILE_H, smem, s>>>(W, tiles_x, num_tiles, max_cols);
}

void b2s_launch_point_cloud(const DWorld& W, uint64_t seed, cudaStream_t s) {
  const int nchunk = (W.P.cam_height * W.P.cam_width + 31) / 32;
  size_t smem = (size_t)(nchunk + 1) * sizeof(int);
  static size_t configured[B2S_MAX_DEVICES];
  b2s_opt_in_smem(k_point_cloud, smem, configured);
  k_point_cloud<<<dim3(W.B, W.Nmax), 32, smem, s>>>(W, seed);
}

// staged stepping: one launch per substep (per-substep launch/timing granularity for profiling)
void b2s_launch_staged(const DWorld& W, int n, cudaStream_t s, int64_t* launches) {
  for (int i = 0; i < n; ++i) b2s_launch_substeps(W, 1, MODE_RAW, 0, 0, 0, nullptr, s);
  *launches = n;
}
