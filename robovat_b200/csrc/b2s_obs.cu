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

#define TILE_W 16            // a warp shades 16 x 2 pixels, a block of 8 warps one 16 x 16 tile
#define TILE_H 16

struct RayCol { int pbeg, pend, uid; float umin, umax, vmin, vmax; };   // planes, body uid, pixel rectangle of the hull

__device__ __forceinline__ int body_uid(const DWorld& W, int n, int first_tile, int slot) {
  if (slot < W.Ns) return (slot < first_tile) ? slot : slot + n;
  if (slot < W.Ns + W.L) return W.Ns + n;
  return first_tile + (slot - W.Ns - W.L);
}

extern __shared__ float4 ray_smem[];

// full clip of the ray `dir` against hull rc; returns true and the entry depth when it hits in front of `best`
__device__ __forceinline__ bool ray_clip(const float4* planes, const RayCol& rc, V3 dir, float near, float best, float* t_hit) {
  float t0 = near, t1 = best;
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
  *t_hit = t0;
  return !miss && t0 < best;
}

// Classification of hull rc against one tile (pixel rectangle [u0, u1] x [v0, v1]):
//   CLS_SKIP  no ray of the tile can hit it: the pixel rectangle of its projected vertices does not meet the tile
//   face f    every ray of the tile enters it through face f and leaves it later; the pixel then needs ONE plane,
//             t = d_f / (n_f . dir), instead of the clip against all of them
//   CLS_FULL  anything else: the pixel clips its ray against all planes of the hull
// "face f" is decided at the four outer corners of the tile (a quadrilateral that contains every pixel centre), with
// margins that survive the rounding of the per-pixel evaluation.  For two planes i, f of one hull the sign of
// t_f - t_i along a ray is the sign of d_i (n_f . dir) - d_f (n_i . dir), which is affine in the pixel coordinates as
// long as n_i . dir and n_f . dir keep their signs, and an affine function takes its extrema over the quadrilateral at
// its corners: a gap g at the four corners is a gap of at least g * (min/max of |n_f . dir|) * (min/max of |n_i . dir|)
// inside.  A plane whose half-space contains the camera (d_i > 0) can only cut rays short (it never raises the entry
// depth): for it the exit depth d_i / max(n_i . dir) is compared, whatever its sign pattern over the tile.  CLS_MARGIN
// is orders of magnitude above the rounding of a three-term dot product and a division, so "face f" pixels get bit
// for bit the depth and the hit decision of the full clip (the GPU tests compare with the oracle, which clips every
// ray against every hull).
#define CLS_SKIP 0xfe
#define CLS_FULL 0xff
#define CLS_MAXP 8            // hulls with more face planes are never classified "face f" (boxes fill the screen, not they)
#define CLS_MARGIN 2e-3f
__device__ int classify_tile(const float4* planes, const RayCol& rc, float fx, float sk, float cx, float fy, float cy, float near,
                             int u0, int u1, int v0, int v1, float* tin, float* rf_out) {
  // pixel rectangle of the hull's projected vertices (+-1.5 pixels; everything when the hull reaches the camera plane)
  if (rc.umax < (float)u0 || rc.umin > (float)u1 || rc.vmax < (float)v0 || rc.vmin > (float)v1) return CLS_SKIP;
  const int np = rc.pend - rc.pbeg;
  if (np > CLS_MAXP) return CLS_FULL;
  float den[4][CLS_MAXP];
  int face = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float uc = (k & 1) ? (float)u1 + 0.5f : (float)u0 - 0.5f, vc = (k & 2) ? (float)v1 + 0.5f : (float)v0 - 0.5f;
    const float dy = (vc - cy) / fy, dx = ((uc - cx) - sk * dy) / fx;
    const V3 dir = v3(dx, dy, 1.0f);
    float best = -3e38f;
    int f = -1;
#pragma unroll
    for (int p = 0; p < CLS_MAXP; ++p) {
      if (p >= np) break;
      const float4 pl = planes[rc.pbeg + p];
      const float d = dot(v3(pl.x, pl.y, pl.z), dir);
      den[k][p] = d;
      if (d < -1e-4f) { const float tt = pl.w / d; if (tt > best) { best = tt; f = p; } }
    }
    if (f < 0) return CLS_FULL;
    if (k == 0) face = f; else if (f != face) return CLS_FULL;
    tin[k] = best;
  }
  const float tin_max = fmaxf(fmaxf(tin[0], tin[1]), fmaxf(tin[2], tin[3])), tin_min = fminf(fminf(tin[0], tin[1]), fminf(tin[2], tin[3]));
  const float m = CLS_MARGIN * (1.0f + fabsf(tin_max));
  if (!(tin_min > near + m)) return CLS_FULL;
  float rf = 1.0f;
#pragma unroll
  for (int p = 0; p < CLS_MAXP; ++p) {
    if (p >= np) break;
    const float d0 = den[0][p], d1 = den[1][p], d2 = den[2][p], d3 = den[3][p];
    const float lo = fminf(fminf(d0, d1), fminf(d2, d3)), hi = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
    if (p == face) { rf = hi / lo; continue; }             // both negative: min |den| / max |den|
  }
#pragma unroll
  for (int p = 0; p < CLS_MAXP; ++p) {
    if (p >= np) break;
    if (p == face) continue;
    const float4 pl = planes[rc.pbeg + p];
    const float d0 = den[0][p], d1 = den[1][p], d2 = den[2][p], d3 = den[3][p];
    const float lo = fminf(fminf(d0, d1), fminf(d2, d3)), hi = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
    if (pl.w > 0.0f) {
      // the camera is inside this half-space: where the plane faces the ray it is passed at a negative depth, elsewhere
      // it ends the ray at d / den >= d / hi
      if (hi > 0.0f && !(pl.w / hi - tin_max > m)) return CLS_FULL;
      continue;
    }
    float gap, ratio;
    if (hi < -1e-4f) {            // entering everywhere: must stay below face f
      gap = fminf(fminf(tin[0] - pl.w / d0, tin[1] - pl.w / d1), fminf(tin[2] - pl.w / d2, tin[3] - pl.w / d3));
      ratio = hi / lo;
    } else if (lo > 1e-4f) {      // leaving everywhere: must stay above face f
      gap = fminf(fminf(pl.w / d0 - tin[0], pl.w / d1 - tin[1]), fminf(pl.w / d2 - tin[2], pl.w / d3 - tin[3]));
      ratio = lo / hi;
    } else return CLS_FULL;
    if (!(gap * ratio * rf > m)) return CLS_FULL;
  }
  *rf_out = rf;
  return face;
}

// Per-environment raster scene in global memory (L2 resident, written by k_render_setup, read by k_render_shade):
//   float4  planes[max_ray_planes]        camera-space face planes of every hull
//   RayCol  cols[max_cols]
//   int     cnt[num_tiles]                hulls the pixels of a tile look at
//   ushort  list[num_tiles][max_cols]     hull | class << 8, hull order
__host__ __device__ __forceinline__ size_t ray_env_bytes(int max_planes, int max_cols, int num_tiles) {
  size_t b = (size_t)max_planes * 16 + (size_t)max_cols * sizeof(RayCol) + (size_t)num_tiles * 4 + (size_t)num_tiles * max_cols * 2;
  return (b + 15) & ~(size_t)15;
}

// Scene set-up, one small block per environment: collider list, camera-space planes + bounding spheres, classification of
// every (tile, hull), per-tile hull lists.  All of it is a few thousand operations per environment; as the prologue of
// the shading blocks it kept 256 threads waiting behind one (list) or fourteen (planes) of them.
#define SETUP_T 64
__global__ void __launch_bounds__(SETUP_T) k_render_setup(const __grid_constant__ DWorld W, unsigned char* scratch, int tiles_x, int num_tiles,
                                                          int max_cols) {
  const int e = blockIdx.x;
  const int lane = threadIdx.x;
  __shared__ int s_ncol;
  const B2SParams& P = W.P;
  const int H = P.cam_height, Wd = P.cam_width;
  unsigned char* base = scratch + (size_t)e * ray_env_bytes(W.max_ray_planes, max_cols, num_tiles);
  float4* planes = (float4*)base;
  RayCol* cols = (RayCol*)(planes + W.max_ray_planes);
  int* cnt = (int*)(cols + max_cols);
  unsigned short* list = (unsigned short*)(cnt + num_tiles);
  const float* cam = W.cam + (size_t)e * 21;
  M3 R; R.r0 = v3(cam[9], cam[10], cam[11]); R.r1 = v3(cam[12], cam[13], cam[14]); R.r2 = v3(cam[15], cam[16], cam[17]);
  const V3 t = v3(cam[18], cam[19], cam[20]);
  const int n = W.buf.num_movables[e];
  int first_tile = W.Ns;
  for (int s = 0; s < W.Ns; ++s) if (W.static_flags[s] & B2S_STATIC_IS_TILE) { first_tile = s; break; }
  // 1. collider list in the oracle's order: statics (incl. visual-only), arm links, movables.  Until step 2 replaces them
  //    uid holds the body slot and the upper half of pend the hull index (< 65536 hulls and planes per environment).
  int ncol = 0;
  if (lane == 0) {
    int nc = 0, np = 0;
    auto push = [&](int slot, int asset) {
      const DAsset& A = W.assets[asset];
      for (int h = A.hoff; h < A.hoff + A.hcnt && nc < max_cols; ++h) {
        cols[nc].uid = slot;
        cols[nc].pbeg = np; np += W.hulls[h].pcnt; cols[nc].pend = np | (h << 16);
        ++nc;
      }
    };
    for (int s = 0; s < W.Ns; ++s) push(s, W.static_asset[s]);
    for (int k = 0; k < W.L; ++k) push(W.Ns + k, W.arm->link_asset[k]);
    for (int i = 0; i < n; ++i) push(W.Ns + W.L + i, __float_as_int(W.mov_params[((size_t)0 * W.B + e) * W.Nmax + i]));
    ncol = nc;
  }
  if (lane == 0) s_ncol = ncol;
  __syncthreads();
  ncol = s_ncol;
  // 2. camera-space planes + pixel rectangles, one collider per thread
  const float* camk = cam;
  for (int c = lane; c < ncol; c += SETUP_T) {
    const int slot = cols[c].uid;
    const DHull& Hh = W.hulls[cols[c].pend >> 16];
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
    const int pbeg = cols[c].pbeg;
    cols[c].pend = cols[c].pend & 0xffff;
    for (int p = 0; p < Hh.pcnt; ++p) {
      float4 pl = W.planes[Hh.poff + p];
      V3 nw = mmul(Rb, v3(pl.x, pl.y, pl.z));
      float dw = pl.w * scale + dot(nw, pos);
      V3 nc = mmul(R, nw);
      float dc = dw + dot(nc, t);
      planes[pbeg + p] = make_float4(nc.x, nc.y, nc.z, dc);
    }
    // pixel rectangle of the hull: its vertices projected through K (a convex hull projects inside the rectangle of its
    // projected vertices as long as all of them are in front of the camera)
    float ulo = 3e38f, uhi = -3e38f, vlo = 3e38f, vhi = -3e38f;
    bool behind = false;
    for (int i = 0; i < Hh.vcnt; ++i) {
      const float4 vv = W.verts[Hh.voff + i];
      const V3 xc = mmul(R, pos + mmul(Rb, v3(vv.x, vv.y, vv.z) * scale)) + t;
      if (!(xc.z > P.cam_near)) { behind = true; break; }
      const float Y = xc.y / xc.z, X = xc.x / xc.z;
      const float pu = camk[0] * X + camk[1] * Y + camk[2], pv = camk[4] * Y + camk[5];
      ulo = fminf(ulo, pu); uhi = fmaxf(uhi, pu); vlo = fminf(vlo, pv); vhi = fmaxf(vhi, pv);
    }
    if (behind || Hh.vcnt < 3) { ulo = vlo = -3e38f; uhi = vhi = 3e38f; }
    cols[c].umin = ulo - 1.5f; cols[c].umax = uhi + 1.5f; cols[c].vmin = vlo - 1.5f; cols[c].vmax = vhi + 1.5f;
    cols[c].uid = body_uid(W, n, first_tile, slot) & 255;
  }
  __syncthreads();
  // 3. one thread per tile: classify every hull against the tile and list the ones its pixels have to look at, in hull
  //    order (the nearest-hit rule breaks ties by that order)
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  for (int tile = lane; tile < num_tiles; tile += SETUP_T) {
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int u0 = tx * TILE_W, v0 = ty * TILE_H;
    // Of the hulls that cover the whole tile through one face, one that lies behind another at all four corners (same
    // affine argument, same margin, scaled by both faces' min/max ratios of |n . dir|) is hidden in every pixel of the
    // tile and leaves the list: the ground under the table, mostly.
    unsigned short* lst = list + (size_t)tile * max_cols;
    int k = 0, bestk = -1;
    float bt[4] = {0, 0, 0, 0}, brf = 1.0f;
    for (int c = 0; c < ncol; ++c) {
      float tin[4], rf = 1.0f;
      const int cl = classify_tile(planes, cols[c], fx, sk, cx, fy, cy, P.cam_near, u0, min(u0 + TILE_W, Wd) - 1, v0, min(v0 + TILE_H, H) - 1, tin, &rf);
      if (cl == CLS_SKIP) continue;
      if (cl != CLS_FULL) {
        bool take = true;
        if (bestk >= 0) {
          const float behind = fminf(fminf(tin[0] - bt[0], tin[1] - bt[1]), fminf(tin[2] - bt[2], tin[3] - bt[3]));
          const float front = fminf(fminf(bt[0] - tin[0], bt[1] - tin[1]), fminf(bt[2] - tin[2], bt[3] - tin[3]));
          const float tmax = fmaxf(fmaxf(fmaxf(tin[0], tin[1]), fmaxf(tin[2], tin[3])), fmaxf(fmaxf(bt[0], bt[1]), fmaxf(bt[2], bt[3])));
          const float m = CLS_MARGIN * (1.0f + fabsf(tmax));
          if (behind * rf * brf > m) continue;                       // hidden by the nearest single-face hull so far
          if (front * rf * brf > m) lst[bestk] = 0xffffu;            // hides it
          else take = false;                                         // they cross inside the tile: both stay
        }
        if (take) { bestk = k; bt[0] = tin[0]; bt[1] = tin[1]; bt[2] = tin[2]; bt[3] = tin[3]; brf = rf; }
      }
      lst[k++] = (unsigned short)(c | (cl << 8));
    }
    int kk = 0;
    for (int i = 0; i < k; ++i) { const unsigned short en = lst[i]; if (en != 0xffffu) lst[kk++] = en; }
    cnt[tile] = kk;
  }
}

// Shading: block = 256 threads = one 32x8 tile at a time; it renders the tiles [band * tiles_per_band, ...) of
// environment e from the scene k_render_setup left in global memory (planes and colliders are staged in shared memory).
__global__ void __launch_bounds__(TILE_W* TILE_H) k_render_shade(const __grid_constant__ DWorld W, const unsigned char* __restrict__ scratch,
                                                               int tiles_x, int num_tiles, int max_cols, int tiles_per_band) {
  const int e = blockIdx.y;
  const int tid = threadIdx.x;
  const B2SParams& P = W.P;
  const int H = P.cam_height, Wd = P.cam_width;
  const unsigned char* base = scratch + (size_t)e * ray_env_bytes(W.max_ray_planes, max_cols, num_tiles);
  const int* gcnt = (const int*)(base + (size_t)W.max_ray_planes * 16 + (size_t)max_cols * sizeof(RayCol));
  const unsigned short* glist = (const unsigned short*)(gcnt + num_tiles);
  float4* planes = ray_smem;                                  // [max_ray_planes] + cols [max_cols], one contiguous copy
  const RayCol* cols = (const RayCol*)(ray_smem + W.max_ray_planes);
  {
    const int words = (W.max_ray_planes * 16 + max_cols * (int)sizeof(RayCol)) / 4;
    const float* src = (const float*)base;
    float* dst = (float*)ray_smem;
    for (int i = tid; i < words; i += blockDim.x) dst[i] = src[i];
  }
  const float* cam = W.cam + (size_t)e * 21;
  const float fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
  // ray directions: dy of every image row and, when the intrinsics have no skew (dx then depends on the column only),
  // dx of every column, computed once per block with the per-pixel formula (two IEEE divisions per pixel otherwise)
  float* dyrow = (float*)ray_smem + (W.max_ray_planes * 16 + max_cols * (int)sizeof(RayCol) + 15) / 16 * 4;
  float* dxcol = dyrow + H;
  for (int i = tid; i < H; i += blockDim.x) dyrow[i] = ((float)i - cy) / fy;
  const bool no_skew = (sk == 0.0f);
  if (no_skew) for (int i = tid; i < Wd; i += blockDim.x) dxcol[i] = (((float)i - cx) - sk * 0.0f) / fx;
  __syncthreads();
  const int tile0 = blockIdx.x * tiles_per_band;
  const int ntile = min(tiles_per_band, num_tiles - tile0);
  // one ray per thread and tile.  A pixel looks at the hulls of its tile's list only -- two or three instead of all
  // (ground, table, a dozen arm links, the movables) -- and needs one plane and one division for the hulls that
  // cover the tile through a single face (table top, ground), the full clip for the others.
  int tx = tile0 % tiles_x, ty = tile0 / tiles_x;
  for (int tl = 0; tl < ntile; ++tl, ++tx) {
    if (tx == tiles_x) { tx = 0; ++ty; }
    const int tile = tile0 + tl;
    const int u = tx * TILE_W + (tid % TILE_W), v = ty * TILE_H + (tid / TILE_W);
    if (u >= Wd || v >= H) continue;
    const float dy = dyrow[v];
    const float dx = no_skew ? dxcol[u] : (((float)u - cx) - sk * dy) / fx;
    const V3 dir = v3(dx, dy, 1.0f);
    float best = P.cam_far;
    int uid = 255;
    const int k1 = gcnt[tile];
    const unsigned short* lst = glist + (size_t)tile * max_cols;
    for (int k = 0; k < k1; ++k) {
      const int ent = lst[k];
      const RayCol rc = cols[ent & 255];
      const int cl = ent >> 8;
      if (cl != CLS_FULL) {
        const float4 pl = planes[rc.pbeg + cl];
        const float tt = pl.w / dot(v3(pl.x, pl.y, pl.z), dir);
        const float t0 = fmaxf(P.cam_near, tt);
        if (t0 < best) { best = t0; uid = rc.uid; }
      } else {
        float t0;
        if (ray_clip(planes, rc, dir, P.cam_near, best, &t0)) { best = t0; uid = rc.uid; }
      }
    }
    W.buf.depth[((size_t)e * H + v) * Wd + u] = best;
    W.buf.segmask[((size_t)e * H + v) * Wd + u] = (uint8_t)uid;
  }
}

// ---- segmented point cloud: one warp per (env, movable) ------------------------------------------------
extern __shared__ int pc_smem[];

template <bool CROP>
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
  auto match = [&](int pix) {
    if (pix >= npix || seg[pix] != uid) return false;
    if (!CROP) return true;
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

size_t b2s_render_scratch_bytes(const DWorld& W) {
  const int tiles_x = (W.P.cam_width + TILE_W - 1) / TILE_W, tiles_y = (W.P.cam_height + TILE_H - 1) / TILE_H;
  return (size_t)W.B * ray_env_bytes(W.max_ray_planes, W.max_ray_cols, tiles_x * tiles_y);
}

void b2s_launch_render(const DWorld& W, cudaStream_t s) {
  const int tiles_x = (W.P.cam_width + TILE_W - 1) / TILE_W, tiles_y = (W.P.cam_height + TILE_H - 1) / TILE_H;
  const int max_cols = W.max_ray_cols;
  const int num_tiles = tiles_x * tiles_y;
  b2s_launch_fk(W, s);                      // link poses of the current joint state
  k_render_setup<<<W.B, SETUP_T, 0, s>>>(W, W.ray_scratch, tiles_x, num_tiles, max_cols);
  // bands of tiles per block: enough blocks to fill the GPU a few times over, at most one per tile
  int per_env = (8 * 148 + W.B - 1) / W.B;
  per_env = per_env < 1 ? 1 : (per_env > num_tiles ? num_tiles : per_env);
  const int tiles_per_band = (num_tiles + per_env - 1) / per_env;
  const int bands = (num_tiles + tiles_per_band - 1) / tiles_per_band;
  size_t smem = (size_t)W.max_ray_planes * sizeof(float4) + (size_t)max_cols * sizeof(RayCol) + 32 + (size_t)(W.P.cam_height + W.P.cam_width) * 4;
  static size_t configured[B2S_MAX_DEVICES];
  b2s_opt_in_smem(k_render_shade, smem, configured);
  k_render_shade<<<dim3(bands, W.B), TILE_W * TILE_H, smem, s>>>(W, W.ray_scratch, tiles_x, num_tiles, max_cols, tiles_per_band);
}

void b2s_launch_point_cloud(const DWorld& W, uint64_t seed, cudaStream_t s) {
  const int nchunk = (W.P.cam_height * W.P.cam_width + 31) / 32;
  size_t smem = (size_t)(nchunk + 1) * sizeof(int);
  static size_t configured[B2S_MAX_DEVICES];
  if (W.P.use_crop) {
    static size_t configured_crop[B2S_MAX_DEVICES];
    b2s_opt_in_smem(k_point_cloud<true>, smem, configured_crop);
    k_point_cloud<true><<<dim3(W.B, W.Nmax), 32, smem, s>>>(W, seed);
  } else {
    b2s_opt_in_smem(k_point_cloud<false>, smem, configured);
    k_point_cloud<false><<<dim3(W.B, W.Nmax), 32, smem, s>>>(W, seed);
  }
}

// staged stepping: one launch per substep (per-substep launch/timing granularity for profiling)
void b2s_launch_staged(const DWorld& W, int n, cudaStream_t s, int64_t* launches) {
  for (int i = 0; i < n; ++i) b2s_launch_substeps(W, 1, MODE_RAW, 0, 0, 0, nullptr, s);
  *launches = n;
}
