// b2s_api.cu -- host side of the C-ABI declared in include/b2s.h.
//
// Error convention: 0 / negative code + thread-local message; nothing throws
// across the boundary; CUDA errors are mapped to B2S_E_CUDA.  There is no CPU
// path behind these entry points: without a CUDA device b2s_create fails.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "b2s_dev.cuh"

struct B2SWorld {
  DWorld d;
  int device;
  bool scene_loaded, buffers_bound;
  std::vector<void*> allocs;
  int64_t launches;
  // export staging for b2s_array(MANIFOLD_*)
  int32_t* exp_keys; int32_t* exp_npts; float* exp_pts;
  size_t arr_bytes[B2S_ARR_COUNT];
  void* arr_ptr[B2S_ARR_COUNT];
  int* unfinished_pinned;          // [4] ring of read-backs
  cudaEvent_t ring_event[4];
};

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(expr)                                                                                 \
  do {                                                                                           \
    cudaError_t err_ = (expr);                                                                   \
    if (err_ != cudaSuccess) return fail(B2S_E_CUDA, "%s: %s", #expr, cudaGetErrorString(err_)); \
  } while (0)
#define NEED(w) \
  if (!(w)) return fail(B2S_E_INVALID, "%s: world is NULL", __func__)
// every entry point works on the world's device whatever the caller's current device is, and leaves the
// caller's current device as it found it (torch keeps its own notion of it)
struct DeviceGuard {
  int prev; bool changed;
  explicit DeviceGuard(int dev) : prev(dev), changed(false) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
#define NEED_READY(w)                                                                            \
  NEED(w);                                                                                       \
  if (!(w)->scene_loaded) return fail(B2S_E_STATE, "%s: call b2s_load_scene first", __func__);   \
  if (!(w)->buffers_bound) return fail(B2S_E_STATE, "%s: call b2s_bind_buffers first", __func__); \
  DeviceGuard device_guard_((w)->device)

template <class T>
static int dalloc(B2SWorld* w, T** p, size_t n, int fill = 0) {
  void* q = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  CU(cudaMalloc(&q, bytes));
  CU(cudaMemset(q, fill, bytes));
  w->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}
template <class T>
static int upload(B2SWorld* w, const T** p, const T* host, size_t n) {
  T* q = nullptr;
  int rc = dalloc(w, &q, n);
  if (rc) return rc;
  if (n) CU(cudaMemcpy(q, host, n * sizeof(T), cudaMemcpyHostToDevice));
  *p = q;
  return 0;
}

static int check_launch(B2SWorld* w, const char* what, int n = 1) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(B2S_E_CUDA, "%s launch: %s", what, cudaGetErrorString(err));
  w->launches += n;
  return 0;
}

extern "C" {

int b2s_version(void) { return B2S_VERSION; }
const char* b2s_last_error(void) { return g_err.c_str(); }
int b2s_sizeof(int which) {
  return which == 0 ? (int)sizeof(B2SParams) : which == 1 ? (int)sizeof(B2SSceneDesc) : which == 2 ? (int)sizeof(B2SBuffers) :
         which == 3 ? (int)sizeof(B2SRollout) : -1;
}

int b2s_default_params(B2SParams* p) {
  if (!p) return fail(B2S_E_INVALID, "b2s_default_params: NULL");
  memset(p, 0, sizeof(*p));
  p->solver_iterations = 50; p->friction_dirs = 2; p->gjk_max_iters = 32; p->epa_max_iters = 32;
  p->ik_max_iters = 20; p->ik_interval = 10; p->check_done_interval = 100;
  p->steps_check = 20; p->max_phase_steps = 3000; p->max_motion_steps = 4000; p->max_offstage_steps = 4000;
  p->stable_check_after = 100; p->stable_min_steps = 100; p->stable_max_steps = 2000;
  p->clamp_joint_velocity = 1; p->warps_per_block = B2S_BLOCK_THREADS / 32;
  p->time_step = 1e-3;
  p->gravity[0] = 0; p->gravity[1] = 0; p->gravity[2] = -9.8f;
  p->erp2 = 0.08f; p->linear_slop = 1e-5f; p->warmstart = 0.85f; p->residual_threshold = 1e-7f;
  p->linear_damping = 0.04f; p->angular_damping = 0.04f; p->breaking_factor = 0.02f;
  p->ik_damping = 0.1f; p->ik_residual = 1e-4f; p->ik_max_step = 0.78539816f;
  p->position_gain = 0.05f; p->velocity_gain = 1.0f;
  p->joint_pos_threshold = 0.008726640f; p->joint_vel_threshold = 0.05f; p->limb_timeout = 15.0f;
  p->limb_velocity_ratio = 0.5f; p->stable_lin_threshold = 0.005f; p->stable_ang_threshold = 0.005f;
  p->cam_near = 0.02f; p->cam_far = 100.0f;
  p->rolling_friction = 0.001f; p->spinning_friction = 0.001f;
  return 0;
}

int b2s_create(const B2SParams* p, int device, B2SWorld** out) {
  if (!p || !out) return fail(B2S_E_INVALID, "b2s_create: NULL argument");
  if (p->num_envs <= 0 || p->max_movables <= 0 || p->max_movables > 32) return fail(B2S_E_INVALID, "b2s_create: num_envs must be > 0 and max_movables 1..32 (one movable per lane in the solve)");
  if (p->max_pairs <= 0 || p->max_manifolds <= 0 || p->max_contacts <= 0 || p->max_colliders <= 0)
    return fail(B2S_E_INVALID, "b2s_create: capacities must be positive");
  if (p->max_colliders > 65535) return fail(B2S_E_INVALID, "b2s_create: max_colliders > 65535");
  if (p->warps_per_block < 0 || p->warps_per_block * 32 > B2S_BLOCK_THREADS) return fail(B2S_E_INVALID, "b2s_create: warps_per_block must be 0 (build default) or 1..%d for this build", B2S_BLOCK_THREADS / 32);
  if (p->friction_dirs != 1 && p->friction_dirs != 2) return fail(B2S_E_INVALID, "b2s_create: friction_dirs must be 1 or 2");
  if (!(p->time_step > 0)) return fail(B2S_E_INVALID, "b2s_create: time_step must be > 0");
  if (p->rolling_friction < 0 || p->spinning_friction < 0) return fail(B2S_E_INVALID, "b2s_create: rolling_friction / spinning_friction must be >= 0");
  if (p->rolling_friction > 0 && p->friction_dirs != 2) return fail(B2S_E_INVALID, "b2s_create: torsional friction rows (rolling_friction > 0) need friction_dirs == 2");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(B2S_E_CUDA, "b2s_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(B2S_E_INVALID, "b2s_create: device %d of %d", device, ndev);
  B2SWorld* w = new B2SWorld();
  memset(&w->d, 0, sizeof(w->d));
  w->d.P = *p;
  if (w->d.P.warps_per_block == 0) w->d.P.warps_per_block = B2S_BLOCK_THREADS / 32;
  w->d.B = p->num_envs; w->d.Nmax = p->max_movables; w->d.Hmax = p->max_colliders;
  w->d.G = p->num_goal_steps > 0 ? p->num_goal_steps : 1;
  w->device = device; w->scene_loaded = false; w->buffers_bound = false; w->launches = 0;
  w->exp_keys = nullptr; w->exp_npts = nullptr; w->exp_pts = nullptr; w->unfinished_pinned = nullptr;
  memset(w->arr_bytes, 0, sizeof(w->arr_bytes)); memset(w->arr_ptr, 0, sizeof(w->arr_ptr));
  *out = w;
  return 0;
}

int b2s_destroy(B2SWorld* w) {
  if (!w) return 0;
  DeviceGuard device_guard_(w->device);
  for (void* p : w->allocs) cudaFree(p);
  if (w->unfinished_pinned) { cudaFreeHost(w->unfinished_pinned); for (int i = 0; i < 4; ++i) cudaEventDestroy(w->ring_event[i]); }
  delete w;
  return 0;
}

int b2s_get_params(const B2SWorld* w, B2SParams* out) {
  if (!w || !out) return fail(B2S_E_INVALID, "b2s_get_params: NULL");
  *out = w->d.P;
  return 0;
}

int b2s_load_scene(B2SWorld* w, const B2SSceneDesc* s) {
  NEED(w);
  if (!s) return fail(B2S_E_INVALID, "b2s_load_scene: scene is NULL");
  if (w->scene_loaded) return fail(B2S_E_STATE, "b2s_load_scene: scene already loaded");
  if (s->num_links < 1 || s->num_links > B2S_MAX_LINKS) return fail(B2S_E_INVALID, "b2s_load_scene: num_links out of range");
  if (s->num_statics < 0 || s->num_hulls <= 0 || s->num_assets <= 0 || s->num_verts <= 0) return fail(B2S_E_INVALID, "b2s_load_scene: empty hull library");
  if (s->num_movable_assets <= 0) return fail(B2S_E_INVALID, "b2s_load_scene: no movable assets (reference asserts len(movable_paths) > 0, push_env.py:107)");
  DeviceGuard device_guard_(w->device);
  DWorld& d = w->d;
  const B2SParams& P = d.P;
  // hull library: derived quantities (local AABB, bounding radius, inertia box)
  std::vector<float4> verts(s->num_verts);
  for (int i = 0; i < s->num_verts; ++i) verts[i] = make_float4(s->verts[i * 3], s->verts[i * 3 + 1], s->verts[i * 3 + 2], 0.0f);
  std::vector<DHull> hulls(s->num_hulls);
  for (int h = 0; h < s->num_hulls; ++h) {
    DHull& H = hulls[h];
    H.voff = s->hull_vert_off[h]; H.vcnt = s->hull_vert_cnt[h]; H.margin = s->hull_margin[h];
    H.poff = s->hull_plane_off ? s->hull_plane_off[h] : 0; H.pcnt = s->hull_plane_cnt ? s->hull_plane_cnt[h] : 0;
    if (H.vcnt < 1 || H.vcnt > 64 || H.voff < 0 || H.voff + H.vcnt > s->num_verts) return fail(B2S_E_INVALID, "b2s_load_scene: hull %d has %d vertices (1..64 allowed)", h, H.vcnt);
    V3 mn = v3(verts[H.voff].x, verts[H.voff].y, verts[H.voff].z), mx = mn;
    float r2 = 0.0f;
    for (int i = 0; i < H.vcnt; ++i) {
      V3 v = v3(verts[H.voff + i].x, verts[H.voff + i].y, verts[H.voff + i].z);
      mn = v3(fminf(mn.x, v.x), fminf(mn.y, v.y), fminf(mn.z, v.z));
      mx = v3(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y), fmaxf(mx.z, v.z));
      r2 = fmaxf(r2, len2(v));
    }
    V3 lc = (mn + mx) * 0.5f, lh = (mx - mn) * 0.5f;
    H.lc[0] = lc.x; H.lc[1] = lc.y; H.lc[2] = lc.z; H.lh[0] = lh.x; H.lh[1] = lh.y; H.lh[2] = lh.z;
    H.rad = sqrtf(r2);
  }
  std::vector<DAsset> assets(s->num_assets);
  for (int a = 0; a < s->num_assets; ++a) {
    DAsset& A = assets[a];
    A.hoff = s->asset_hull_off[a]; A.hcnt = s->asset_hull_cnt[a]; A.pad = 0;
    if (A.hcnt < 1 || A.hoff < 0 || A.hoff + A.hcnt > s->num_hulls) return fail(B2S_E_INVALID, "b2s_load_scene: asset %d hull range", a);
    V3 mn = v3(3e38f, 3e38f, 3e38f), mx = v3(-3e38f, -3e38f, -3e38f);
    for (int h = A.hoff; h < A.hoff + A.hcnt; ++h) {
      const DHull& H = hulls[h];
      V3 lc = v3(H.lc[0], H.lc[1], H.lc[2]), lh = v3(H.lh[0], H.lh[1], H.lh[2]);
      V3 lo = (lc - lh) - v3(H.margin, H.margin, H.margin), hi = (lc + lh) + v3(H.margin, H.margin, H.margin);
      mn = v3(fminf(mn.x, lo.x), fminf(mn.y, lo.y), fminf(mn.z, lo.z));
      mx = v3(fmaxf(mx.x, hi.x), fmaxf(mx.y, hi.y), fmaxf(mx.z, hi.z));
    }
    V3 half = (mx - mn) * 0.5f;
    A.half[0] = half.x; A.half[1] = half.y; A.half[2] = half.z;
  }
  std::vector<float4> planes(std::max(s->num_planes, 1));
  for (int i = 0; i < s->num_planes; ++i) planes[i] = make_float4(s->planes[i * 4], s->planes[i * 4 + 1], s->planes[i * 4 + 2], s->planes[i * 4 + 3]);
  int rc;
  if ((rc = upload(w, &d.verts, verts.data(), verts.size()))) return rc;
  if ((rc = upload(w, &d.hulls, hulls.data(), hulls.size()))) return rc;
  if ((rc = upload(w, &d.assets, assets.data(), assets.size()))) return rc;
  if ((rc = upload(w, &d.planes, planes.data(), planes.size()))) return rc;
  if ((rc = upload(w, &d.static_asset, s->static_asset, s->num_statics))) return rc;
  if ((rc = upload(w, &d.static_pose, s->static_pose, (size_t)s->num_statics * 7))) return rc;
  if ((rc = upload(w, &d.static_friction, s->static_friction, s->num_statics))) return rc;
  if ((rc = upload(w, &d.static_flags, s->static_flags, s->num_statics))) return rc;
  if ((rc = upload(w, &d.movable_assets, s->movable_assets, s->num_movable_assets))) return rc;
  if ((rc = upload(w, &d.target_assets, s->target_assets, std::max(s->num_target_assets, 0)))) return rc;
  DArm arm;
  memset(&arm, 0, sizeof(arm));
  memcpy(arm.base, s->arm_base_pose, sizeof(arm.base));
  memcpy(arm.joint_origin, s->joint_origin, sizeof(arm.joint_origin));
  memcpy(arm.joint_axis, s->joint_axis, sizeof(arm.joint_axis));
  memcpy(arm.lower, s->joint_lower, sizeof(arm.lower)); memcpy(arm.upper, s->joint_upper, sizeof(arm.upper));
  memcpy(arm.max_vel, s->joint_max_velocity, sizeof(arm.max_vel));
  memcpy(arm.ee, s->ee_pose, sizeof(arm.ee));
  arm.num_links = s->num_links;
  memcpy(arm.link_joint, s->link_joint, sizeof(arm.link_joint)); memcpy(arm.link_asset, s->link_asset, sizeof(arm.link_asset));
  memcpy(arm.link_pose, s->link_pose, sizeof(arm.link_pose));
  arm.friction = s->arm_friction;
  for (int k = 0; k < s->num_links; ++k)
    if (s->link_joint[k] < -1 || s->link_joint[k] >= B2S_NUM_JOINTS || s->link_asset[k] < 0 || s->link_asset[k] >= s->num_assets)
      return fail(B2S_E_INVALID, "b2s_load_scene: link %d joint/asset out of range", k);
  if ((rc = upload(w, &d.arm, &arm, 1))) return rc;
  DLayout lay;
  memset(&lay, 0, sizeof(lay));
  lay.tile_size = s->tile_size; memcpy(lay.tile_offset, s->tile_offset, 8);
  lay.num_region = s->num_region; lay.num_goal = s->num_goal; lay.num_target = s->num_target; lay.num_obstacle = s->num_obstacle;
  memcpy(lay.region, s->region, sizeof(lay.region)); memcpy(lay.goal, s->goal, sizeof(lay.goal));
  memcpy(lay.target, s->target, sizeof(lay.target)); memcpy(lay.obstacle, s->obstacle, sizeof(lay.obstacle));
  memcpy(lay.scale_range, s->scale_range, 8); memcpy(lay.mass_range, s->mass_range, 8); memcpy(lay.friction_range, s->friction_range, 8);
  memcpy(lay.pose_x, s->pose_x, 8); memcpy(lay.pose_y, s->pose_y, 8); memcpy(lay.pose_z, s->pose_z, 8);
  memcpy(lay.pose_roll, s->pose_roll, 8); memcpy(lay.pose_pitch, s->pose_pitch, 8); memcpy(lay.pose_yaw, s->pose_yaw, 8);
  lay.placement_margin = s->placement_margin; lay.min_movables = s->min_movables;
  memcpy(lay.table_height_range, s->table_height_range, 8); lay.safe_drop_height = s->safe_drop_height;
  lay.num_movable_assets = s->num_movable_assets; lay.num_target_assets = s->num_target_assets;
  if ((rc = upload(w, &d.layout, &lay, 1))) return rc;

  d.Ns = s->num_statics; d.L = s->num_links; d.NB = d.Ns + d.L + d.Nmax;
  {
    // raster capacities: every static (visual-only ones too), every arm link, Nmax x the largest movable asset
    auto asset_planes = [&](int a) { int n = 0; for (int h = assets[a].hoff; h < assets[a].hoff + assets[a].hcnt; ++h) n += hulls[h].pcnt; return n; };
    int cols = 0, pls = 0, mc = 0, mp = 0;
    for (int i = 0; i < s->num_statics; ++i) { cols += assets[s->static_asset[i]].hcnt; pls += asset_planes(s->static_asset[i]); }
    for (int k = 0; k < s->num_links; ++k) { cols += assets[s->link_asset[k]].hcnt; pls += asset_planes(s->link_asset[k]); }
    for (int i = 0; i < s->num_movable_assets; ++i) { mc = std::max(mc, assets[s->movable_assets[i]].hcnt); mp = std::max(mp, asset_planes(s->movable_assets[i])); }
    for (int i = 0; i < s->num_target_assets; ++i) { mc = std::max(mc, assets[s->target_assets[i]].hcnt); mp = std::max(mp, asset_planes(s->target_assets[i])); }
    d.max_ray_cols = cols + d.Nmax * mc;
    d.max_ray_planes = std::max(1, pls + d.Nmax * mp);
    if (d.max_ray_cols > 256) return fail(B2S_E_CAPACITY, "b2s_load_scene: %d hulls per environment exceed the raster's 256", d.max_ray_cols);
    if (d.max_ray_planes > 65535 || s->num_hulls > 32767) return fail(B2S_E_CAPACITY, "b2s_load_scene: the raster indexes at most 65535 face planes per environment and 32767 hulls");
  }
  const size_t B = d.B, N = d.Nmax, M = P.max_manifolds;
#define ALLOC(field, count, fill) if ((rc = dalloc(w, &d.field, (count), (fill)))) return rc
  ALLOC(man_keys, 2 * B * M, 0xff); ALLOC(man_npts, 2 * B * M, 0); ALLOC(man_pts, 2 * B * M * 4 * B2S_CP_FLOATS, 0);
  ALLOC(man_parity, B, 0); ALLOC(num_manifolds, B, 0);
  ALLOC(pair_keys, B * P.max_pairs, 0); ALLOC(num_pairs, B, 0);
  ALLOC(phase, B, 0); ALLOC(num_steps, B, 0);
  ALLOC(ctrl, B * B2S_CTRL_FLOATS, 0); ALLOC(ctrl_flags, B * 4, 0); ALLOC(ctrl_time, B * 5, 0);
  ALLOC(link_poses, B * (d.L + 1) * 7, 0); ALLOC(link_vel, B * d.L * 6, 0);
  ALLOC(mov_params, 4 * B * N, 0); ALLOC(table_dz, B, 0); ALLOC(error_flags, B, 0);
  ALLOC(waypoints, B * d.G * 14, 0); ALLOC(status, B * 2 * N * 4, 0);
  ALLOC(contact_flags, B, 0); ALLOC(phase_state, B * 8, 0); ALLOC(solver_stats, B * 4, 0);
  ALLOC(ncol, B, 0); ALLOC(col_slot, B * d.Hmax, 0); ALLOC(col_hull, B * d.Hmax, 0);
  ALLOC(reset_count, B, 0); ALLOC(prev_xy, B * N * 2, 0); ALLOC(cam, B * 21, 0);
  ALLOC(ro_state, B * 4, 0); ALLOC(num_episodes, B, 0); ALLOC(async_events, B, 0); ALLOC(work_ema, B, 0);
  ALLOC(substeps, 1, 0); ALLOC(free_target, 2, 0); ALLOC(unfinished, 1, 0); ALLOC(prof, 8 + 4 * 1024 + 16 + 128, 0);
  if ((rc = dalloc(w, &w->exp_keys, B * M, 0))) return rc;
  if ((rc = dalloc(w, &w->exp_npts, B * M, 0))) return rc;
  if ((rc = dalloc(w, &w->exp_pts, B * M * 4 * B2S_CP_FLOATS, 0))) return rc;
#undef ALLOC
  {
    std::vector<int32_t> ph(B, B2S_PHASE_IDLE), ps(B * 8, 0);
    for (size_t e = 0; e < B; ++e) ps[e * 8] = -1;
    CU(cudaMemcpy(d.phase, ph.data(), B * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d.phase_state, ps.data(), B * 8 * 4, cudaMemcpyHostToDevice));
  }
  CU(cudaMallocHost((void**)&w->unfinished_pinned, 4 * sizeof(int)));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreateWithFlags(&w->ring_event[i], cudaEventDisableTiming));
  // shared-memory carve-up (see SmemLayout)
  SmemLayout& sm = d.sm;
  int o = 0;
  auto take = [&](int words) { int at = o; o += (words + 1) & ~1; return at; };   // keep 8-byte alignment
  d.reg_rows = (P.max_contacts <= 32 && d.NB <= 32) ? 1 : 0;
  sm.body = take(d.NB * BODY_STRIDE);
  sm.col = take(d.Hmax * COL_STRIDE);
  sm.pairs = take(P.max_pairs);
  sm.cmk = take(P.max_contacts);
  sm.pstage = o; sm.ps_cap = 0;               // sized below, once the rest of the block's shared memory is known
  sm.words_env = o;
  o = 0;
  const int x0 = o;                 // region shared by the narrow-phase scratch and the solver rows
  sm.stage = take(4 * B2S_CP_FLOATS * UNITS_PER_WARP);
  sm.fk = take(FK_WORDS);
  sm.simplex = take(48 * UNITS_PER_WARP);
  sm.con = x0;
  // solve-stage scratch: small path = colour table (bytes) + lambda x2 + slots; big path = colour table (16 bit) +
  // bodies and colour of every contact
  o = std::max(o, x0 + (d.reg_rows ? 512 + 8 * 32 : 1024 + (P.max_contacts + 1) / 2 + (P.max_contacts + 3) / 4 + 2));
  sm.words_warp = o;
  {
    // large scenes (rows in shared memory): fewer warps per block so the block still fits 220 KB
    while (d.P.warps_per_block > 1 &&
           ((size_t)d.P.warps_per_block * (sm.words_env + META_WORDS + sm.words_warp)) * 4 > 220 * 1024) d.P.warps_per_block -= 1;
    const int wpb = d.P.warps_per_block;
    // environments per block: with register-resident rows any warp can run any stage of any environment of
    // its block (dynamic hand-out), so a block may own more environments than warps; otherwise one per warp
    int maxE = 4 * wpb;
    while (maxE > 1 && ((size_t)maxE * (sm.words_env + META_WORDS) + (size_t)wpb * sm.words_warp) * 4 > 220 * 1024) --maxE;
    int E, whole_waves = 0;
    if (P.envs_per_block > 0) E = P.envs_per_block;
    else {
      // fill whole waves of SMs (one block per SM): B = 4096 -> 27 or 28 per block -> 148 blocks
      int dev_sms = 148;
      cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, w->device);
      const int waves = (d.B + dev_sms * maxE - 1) / (dev_sms * maxE);
      E = (d.B + dev_sms * waves - 1) / (dev_sms * waves);
      whole_waves = std::min(dev_sms * waves, d.B);
    }
    if (E < 1) E = 1;
    if (E > maxE) E = maxE;
    d.num_blocks = std::max((d.B + E - 1) / E, whole_waves);     // both deals (b2s_aux.cu) spread the envs evenly over them
    // spare slots: the deal (k_assign_envs) gives the blocks that hold the expensive environments at most one
    // environment per warp and lets the cheap blocks take the rest
    // (B2S_EXTRA_SLOTS: tuning knob for variant builds -- more spare slots let the deal form more, smaller blocks of
    // expensive environments, see B2S_HW in b2s_aux.cu)
#ifndef B2S_EXTRA_SLOTS
#define B2S_EXTRA_SLOTS 0
#endif
    if (P.envs_per_block <= 0 && E > wpb) E = std::min(maxE, E + std::max(4, E / 7) + B2S_EXTRA_SLOTS);
    d.envs_per_block = E;
    {
      // staging records of the narrow phase in shared memory: as many pairs per environment as fit in what is left of
      // B2S_SMEM_BUDGET (the rest of the SM's 256 KB stays L1: hull vertices, manifolds and scene tables live there)
#ifndef B2S_SMEM_BUDGET
#define B2S_SMEM_BUDGET (200 * 1024)
#endif
#ifndef B2S_PS_CAP_MAX
#define B2S_PS_CAP_MAX 4
#endif
      const size_t used = ((size_t)E * (sm.words_env + META_WORDS) + (size_t)wpb * sm.words_warp) * 4;
      int cap = used < (size_t)B2S_SMEM_BUDGET ? (int)(((size_t)B2S_SMEM_BUDGET - used) / ((size_t)E * 68 * 4)) : 0;
      cap = std::min(std::min(cap, (int)P.max_pairs), B2S_PS_CAP_MAX);
      if (cap < 2) cap = 0;
      sm.ps_cap = cap;
      sm.words_env += cap * 68;
    }
    const size_t blocks = d.num_blocks;
    if ((rc = dalloc(w, &d.epa_scratch, blocks * wpb * UNITS_PER_WARP * (size_t)(EPA_MAXV * 11 + EPA_MAXF * 7), 0))) return rc;
    if ((rc = dalloc(w, &d.pair_stage, blocks * (size_t)E * P.max_pairs * 68, 0))) return rc;
    if ((rc = dalloc(w, &d.env_map, blocks * (size_t)E, 0xff))) return rc;
    if ((rc = dalloc(w, &d.row_scratch, d.reg_rows ? 4 : blocks * wpb * (size_t)P.max_contacts * 124, 0))) return rc;
  }
  size_t smem = b2s_smem_bytes(d);
  if (smem > 227 * 1024) return fail(B2S_E_CAPACITY, "b2s_load_scene: %zu bytes of shared memory per block exceed 227 KB; lower warps_per_block or the capacities", smem);
  // arrays exposed through b2s_array
  auto reg = [&](int id, void* p, size_t bytes) { w->arr_ptr[id] = p; w->arr_bytes[id] = bytes; };
  reg(B2S_ARR_MANIFOLD_KEYS, w->exp_keys, B * M * 4); reg(B2S_ARR_MANIFOLD_NPTS, w->exp_npts, B * M * 4);
  reg(B2S_ARR_MANIFOLD_PTS, w->exp_pts, B * M * 4 * B2S_CP_FLOATS * 4); reg(B2S_ARR_NUM_MANIFOLDS, d.num_manifolds, B * 4);
  reg(B2S_ARR_PAIR_KEYS, d.pair_keys, B * P.max_pairs * 4); reg(B2S_ARR_NUM_PAIRS, d.num_pairs, B * 4);
  reg(B2S_ARR_PHASE, d.phase, B * 4); reg(B2S_ARR_NUM_STEPS, d.num_steps, B * 4);
  reg(B2S_ARR_CTRL, d.ctrl, B * B2S_CTRL_FLOATS * 4); reg(B2S_ARR_CTRL_FLAGS, d.ctrl_flags, B * 16);
  reg(B2S_ARR_LINK_POSES, d.link_poses, B * (d.L + 1) * 28); reg(B2S_ARR_MOV_PARAMS, d.mov_params, 4 * B * N * 4);
  reg(B2S_ARR_TABLE_DZ, d.table_dz, B * 4); reg(B2S_ARR_ERROR_FLAGS, d.error_flags, B * 4);
  reg(B2S_ARR_WAYPOINTS, d.waypoints, B * d.G * 56); reg(B2S_ARR_STATUS, d.status, B * 2 * N * 16);
  reg(B2S_ARR_CONTACT_FLAGS, d.contact_flags, B * 4); reg(B2S_ARR_PHASE_STATE, d.phase_state, B * 32);
  reg(B2S_ARR_SOLVER_STATS, d.solver_stats, B * 16); reg(B2S_ARR_CTRL_TIME, d.ctrl_time, B * 40);
  reg(B2S_ARR_LINK_VEL, d.link_vel, B * d.L * 24); reg(B2S_ARR_NUM_COLLIDERS, d.ncol, B * 4);
  reg(B2S_ARR_COL_SLOT, d.col_slot, B * d.Hmax * 4); reg(B2S_ARR_COL_HULL, d.col_hull, B * d.Hmax * 4);
  reg(B2S_ARR_PROF, d.prof, (8 + 4 * 1024 + 16 + 128) * 8);
  reg(B2S_ARR_NUM_EPISODES, d.num_episodes, B * 4); reg(B2S_ARR_ROLLOUT_STATE, d.ro_state, B * 4 * 4);
  w->scene_loaded = true;
  return 0;
}

int b2s_bind_buffers(B2SWorld* w, const B2SBuffers* b) {
  NEED(w);
  if (!b) return fail(B2S_E_INVALID, "b2s_bind_buffers: NULL");
  if (!b->body_state || !b->joint_state || !b->action || !b->obs_position || !b->num_movables || !b->body_mask || !b->reward ||
      !b->termination || !b->is_safe || !b->is_effective || !b->episode_return)
    return fail(B2S_E_INVALID, "b2s_bind_buffers: body_state, joint_state, action, obs_position, num_movables, body_mask, reward, termination, is_safe, is_effective and episode_return are required");
  w->d.buf = *b;
  w->buffers_bound = true;
  return 0;
}

int b2s_reset(B2SWorld* w, const uint8_t* mask, uint64_t seed, void* stream) {
  NEED_READY(w);
  b2s_launch_reset(w->d, mask, seed, (cudaStream_t)stream);
  return check_launch(w, "reset");
}

int b2s_settle_masked(B2SWorld* w, const uint8_t* mask, float lin, float ang, int max_steps, void* stream) {
  NEED_READY(w);
  if (max_steps < 1) return fail(B2S_E_INVALID, "b2s_settle: max_steps < 1");
  b2s_launch_substeps(w->d, 0, MODE_SETTLE, lin, ang, max_steps, mask, (cudaStream_t)stream);
  return check_launch(w, "settle", 2);      // deal of the environments + substep kernel
}
int b2s_settle(B2SWorld* w, float lin, float ang, int max_steps, void* stream) {
  return b2s_settle_masked(w, nullptr, lin, ang, max_steps, stream);
}

int b2s_begin_episode(B2SWorld* w, const uint8_t* mask, void* stream) {
  NEED_READY(w);
  b2s_launch_begin_episode(w->d, mask, (cudaStream_t)stream);
  return check_launch(w, "begin_episode");
}

int b2s_step(B2SWorld* w, int n, void* stream) {
  NEED_READY(w);
  if (n < 0) return fail(B2S_E_INVALID, "b2s_step: n < 0");
  if (n == 0) return 0;
  b2s_launch_substeps(w->d, n, MODE_RAW, 0, 0, 0, nullptr, (cudaStream_t)stream);
  return check_launch(w, "step", 2);
}

int b2s_step_staged(B2SWorld* w, int n, void* stream) {
  NEED_READY(w);
  if (n < 0) return fail(B2S_E_INVALID, "b2s_step_staged: n < 0");
  int64_t launches = 0;
  b2s_launch_staged(w->d, n, (cudaStream_t)stream, &launches);
  return check_launch(w, "step_staged", 2 * (int)launches);
}

int b2s_set_action(B2SWorld* w, void* stream) {
  NEED_READY(w);
  w->d.ro.enabled = RO_OFF;                  // lock-step path: the host decides what follows an action
  b2s_launch_set_action(w->d, (cudaStream_t)stream);
  return check_launch(w, "set_action");
}

static int env_substeps(B2SWorld* w, int n, int* unfinished_host, void* stream, bool free_running) {
  NEED_READY(w);
  if (n < 0) return fail(B2S_E_INVALID, "b2s_env_substeps: n < 0");
  cudaStream_t s = (cudaStream_t)stream;
  if (free_running && n > 0) b2s_launch_substeps(w->d, 4 * n, MODE_ENV, 0, 0, 0, nullptr, s, n, false);    // lock-step: the batch waits for its slowest environments
  else b2s_launch_substeps(w->d, n, MODE_ENV, 0, 0, 0, nullptr, s);
  int rc = check_launch(w, "env_substeps", 2);
  if (rc) return rc;
  if (unfinished_host) {
    b2s_launch_count_running(w->d, s);
    if ((rc = check_launch(w, "count_running"))) return rc;
    CU(cudaMemcpyAsync(w->unfinished_pinned, w->d.unfinished, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *unfinished_host = *w->unfinished_pinned;
  }
  return 0;
}

int b2s_env_substeps(B2SWorld* w, int n, int* unfinished_host, void* stream) { return env_substeps(w, n, unfinished_host, stream, false); }
int b2s_env_substeps_free(B2SWorld* w, int n, int* unfinished_host, void* stream) { return env_substeps(w, n, unfinished_host, stream, true); }

int b2s_env_step(B2SWorld* w, int chunk, int max_substeps, void* stream) {
  NEED_READY(w);
  if (chunk < 1) return fail(B2S_E_INVALID, "b2s_env_step: chunk < 1");
  int rc = b2s_set_action(w, stream);
  if (rc) return rc;
  int done = 0, unfinished = 1;
  while (unfinished > 0 && done < max_substeps) {
    rc = b2s_env_substeps(w, chunk, &unfinished, stream);
    if (rc) return rc;
    done += chunk;
  }
  return 0;
}

int b2s_rollout_begin(B2SWorld* w, const B2SRollout* r, void* stream) {
  NEED_READY(w);
  if (!r) return fail(B2S_E_INVALID, "b2s_rollout_begin: rollout is NULL");
  if (r->num_actions < 1) return fail(B2S_E_INVALID, "b2s_rollout_begin: num_actions < 1");
  if (r->max_attempts < 1 || r->max_attempts > 65535) return fail(B2S_E_INVALID, "b2s_rollout_begin: max_attempts must be 1..65535");
  if (w->d.Nmax > 32) return fail(B2S_E_CAPACITY, "b2s_rollout_begin: max_movables > 32");
  DRollout& ro = w->d.ro;
  if (r->policy_kind != B2S_POLICY_HEURISTIC && r->policy_kind != B2S_POLICY_AIMED) return fail(B2S_E_INVALID, "b2s_rollout_begin: unknown policy_kind");
  if (w->d.G > 1) return fail(B2S_E_UNSUPPORTED, "b2s_rollout_begin: the device policies draw one (start, motion) pair per action; NUM_GOAL_STEPS > 1 needs the host policy (b2s_env_async_step)");
  if (r->num_episodes < 1) return fail(B2S_E_INVALID, "b2s_rollout_begin: num_episodes < 1");
  if (r->max_reset_retries < 0) return fail(B2S_E_INVALID, "b2s_rollout_begin: max_reset_retries < 0");
  ro.enabled = RO_EPISODES; ro.num_actions = r->num_actions; ro.max_attempts = r->max_attempts; ro.num_episodes = r->num_episodes;
  ro.max_reset_retries = r->max_reset_retries; ro.drop_max_steps = r->drop_max_steps; ro.policy_kind = r->policy_kind; ro.free_running = r->free_running ? 1 : 0;
  ro.drop_lin = r->drop_lin_threshold; ro.drop_ang = r->drop_ang_threshold;
  ro.seed = r->seed; ro.reset_seed = r->reset_seed;
  ro.actions = r->actions; ro.rewards = r->rewards; ro.positions = r->positions; ro.flags = r->flags;
  ro.substeps = r->substeps; ro.lengths = r->lengths; ro.returns = r->returns;
  b2s_launch_rollout_begin(w->d, r->first_action, (cudaStream_t)stream);
  return check_launch(w, "rollout_begin");
}

int b2s_rollout_run(B2SWorld* w, int chunk, int max_substeps, int* unfinished_host, void* stream) {
  NEED_READY(w);
  if (chunk < 1) return fail(B2S_E_INVALID, "b2s_rollout_run: chunk < 1");
  if (w->d.ro.enabled != RO_EPISODES) return fail(B2S_E_STATE, "b2s_rollout_run: call b2s_rollout_begin first");
  cudaStream_t s = (cudaStream_t)stream;
  int launched = 0, i = 0, last = -1;
  while (launched < max_substeps) {
    const int n = (max_substeps - launched < chunk) ? (max_substeps - launched) : chunk;
    if (w->d.ro.free_running) b2s_launch_substeps(w->d, 4 * n, MODE_ENV, 0, 0, 0, nullptr, s, n);
    else b2s_launch_substeps(w->d, n, MODE_ENV, 0, 0, 0, nullptr, s);
    int rc = check_launch(w, "rollout_run", 2);
    if (rc) return rc;
    b2s_launch_count_running(w->d, s);
    if ((rc = check_launch(w, "count_running"))) return rc;
    CU(cudaMemcpyAsync(w->unfinished_pinned + (i & 3), w->d.unfinished, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(w->ring_event[i & 3], s));
    launched += n;
    ++i;
    if (i >= 3) {                            // look at the launch before the previous one: two launches stay queued
      CU(cudaEventSynchronize(w->ring_event[(i - 3) & 3]));
      if (w->unfinished_pinned[(i - 3) & 3] == 0) break;
    }
  }
  CU(cudaStreamSynchronize(s));
  last = w->unfinished_pinned[(i - 1) & 3];
  if (unfinished_host) *unfinished_host = last;
  return 0;
}

static int async_step(B2SWorld* w, const uint8_t* command, int n, uint64_t reset_seed, uint8_t* status, void* stream, bool free_running);
int b2s_env_async_step(B2SWorld* w, const uint8_t* command, int n, uint64_t reset_seed, uint8_t* status, void* stream) {
  return async_step(w, command, n, reset_seed, status, stream, false);
}
int b2s_env_async_step_free(B2SWorld* w, const uint8_t* command, int n, uint64_t reset_seed, uint8_t* status, void* stream) {
  return async_step(w, command, n, reset_seed, status, stream, true);
}
static int async_step(B2SWorld* w, const uint8_t* command, int n, uint64_t reset_seed, uint8_t* status, void* stream, bool free_running) {
  NEED_READY(w);
  if (n < 0) return fail(B2S_E_INVALID, "b2s_env_async_step: n_substeps < 0");
  if (w->d.Nmax > 32) return fail(B2S_E_CAPACITY, "b2s_env_async_step: max_movables > 32");
  cudaStream_t s = (cudaStream_t)stream;
  DRollout& ro = w->d.ro;
  if (ro.enabled != RO_ASYNC) {
    memset(&ro, 0, sizeof(ro));
    ro.enabled = RO_ASYNC; ro.max_reset_retries = 8; ro.drop_lin = 0.1f; ro.drop_ang = 0.1f; ro.drop_max_steps = 500;
  }
  ro.reset_seed = reset_seed;
  b2s_launch_async_commands(w->d, command, s);
  int rc = check_launch(w, "async_commands");
  if (rc) return rc;
  if (n > 0) {
    if (free_running) b2s_launch_substeps(w->d, 4 * n, MODE_ENV, 0, 0, 0, nullptr, s, n);
    else b2s_launch_substeps(w->d, n, MODE_ENV, 0, 0, 0, nullptr, s);
    if ((rc = check_launch(w, "env_async_step", 2))) return rc;
  }
  if (status) {
    b2s_launch_async_status(w->d, status, s);
    if ((rc = check_launch(w, "async_status"))) return rc;
  }
  return 0;
}

int b2s_arm_move_to_gripper_pose(B2SWorld* w, const float* pose, const uint8_t* mask, void* stream) {
  NEED_READY(w);
  if (!pose) return fail(B2S_E_INVALID, "b2s_arm_move_to_gripper_pose: pose is NULL");
  b2s_launch_arm_cmd(w->d, 0, pose, mask, nullptr, (cudaStream_t)stream);
  return check_launch(w, "arm_cmd");
}
int b2s_arm_move_to_joint_positions(B2SWorld* w, const float* q, const uint8_t* mask, void* stream) {
  NEED_READY(w);
  if (!q) return fail(B2S_E_INVALID, "b2s_arm_move_to_joint_positions: q is NULL");
  b2s_launch_arm_cmd(w->d, 1, q, mask, nullptr, (cudaStream_t)stream);
  return check_launch(w, "arm_cmd");
}
int b2s_arm_reset_targets(B2SWorld* w, const uint8_t* mask, void* stream) {
  NEED_READY(w);
  b2s_launch_arm_cmd(w->d, 2, nullptr, mask, nullptr, (cudaStream_t)stream);
  return check_launch(w, "arm_cmd");
}
int b2s_set_motor_targets(B2SWorld* w, const float* q, const float* qd, const uint8_t* mask, void* stream) {
  NEED_READY(w);
  if (!q) return fail(B2S_E_INVALID, "b2s_set_motor_targets: q is NULL");
  b2s_launch_arm_cmd(w->d, 4, q, mask, (uint8_t*)qd, (cudaStream_t)stream);
  return check_launch(w, "arm_cmd");
}
int b2s_rebuild_colliders(B2SWorld* w, void* stream) {
  NEED_READY(w);
  b2s_launch_rebuild_colliders(w->d, (cudaStream_t)stream);
  return check_launch(w, "rebuild_colliders");
}
int b2s_arm_is_ready(B2SWorld* w, uint8_t* out, void* stream) {
  NEED_READY(w);
  if (!out) return fail(B2S_E_INVALID, "b2s_arm_is_ready: out is NULL");
  b2s_launch_arm_cmd(w->d, 3, nullptr, nullptr, out, (cudaStream_t)stream);
  return check_launch(w, "arm_cmd");
}
int b2s_inverse_kinematics(B2SWorld* w, const float* pose, const float* q_start, float* q_out, void* stream) {
  NEED_READY(w);
  if (!pose || !q_start || !q_out) return fail(B2S_E_INVALID, "b2s_inverse_kinematics: NULL argument");
  b2s_launch_ik(w->d, pose, q_start, q_out, (cudaStream_t)stream);
  return check_launch(w, "ik");
}
int b2s_forward_kinematics(B2SWorld* w, void* stream) {
  NEED_READY(w);
  b2s_launch_fk(w->d, (cudaStream_t)stream);
  return check_launch(w, "fk");
}
int b2s_query_contacts(B2SWorld* w, uint8_t* arm_table, uint8_t* arm_movable, void* stream) {
  NEED_READY(w);
  b2s_launch_query_contacts(w->d, arm_table, arm_movable, (cudaStream_t)stream);
  return check_launch(w, "query_contacts");
}
int b2s_observe(B2SWorld* w, void* stream) {
  NEED_READY(w);
  b2s_launch_observe(w->d, (cudaStream_t)stream);
  return check_launch(w, "observe");
}

int b2s_set_camera(B2SWorld* w, const float* K, const float* R, const float* t, int per_env) {
  NEED(w);
  if (!w->scene_loaded) return fail(B2S_E_STATE, "b2s_set_camera: call b2s_load_scene first");
  if (!K || !R || !t) return fail(B2S_E_INVALID, "b2s_set_camera: NULL argument");
  const int B = w->d.B;
  std::vector<float> cam((size_t)B * 21);
  for (int e = 0; e < B; ++e) {
    size_t o = per_env ? (size_t)e : 0;
    memcpy(&cam[(size_t)e * 21], K + o * 9, 36); memcpy(&cam[(size_t)e * 21 + 9], R + o * 9, 36); memcpy(&cam[(size_t)e * 21 + 18], t + o * 3, 12);
  }
  DeviceGuard device_guard_(w->device);
  CU(cudaMemcpy(w->d.cam, cam.data(), cam.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}
int b2s_render(B2SWorld* w, void* stream) {
  NEED_READY(w);
  if (!w->d.buf.depth || !w->d.buf.segmask) return fail(B2S_E_STATE, "b2s_render: depth/segmask buffers are not bound");
  if (!w->d.ray_scratch) {                   // the raster's scene scratch exists only in worlds that render
    DeviceGuard device_guard_(w->device);
    int rc = dalloc(w, &w->d.ray_scratch, b2s_render_scratch_bytes(w->d), 0);
    if (rc) return rc;
    w->arr_ptr[B2S_ARR_RAY_SCENE] = w->d.ray_scratch; w->arr_bytes[B2S_ARR_RAY_SCENE] = b2s_render_scratch_bytes(w->d);
  }
  b2s_launch_render(w->d, (cudaStream_t)stream);
  return check_launch(w, "render", 3);
}
int b2s_point_cloud(B2SWorld* w, uint64_t seed, void* stream) {
  NEED_READY(w);
  if (!w->d.buf.depth || !w->d.buf.segmask || !w->d.buf.point_cloud) return fail(B2S_E_STATE, "b2s_point_cloud: depth/segmask/point_cloud buffers are not bound");
  b2s_launch_point_cloud(w->d, seed, (cudaStream_t)stream);
  return check_launch(w, "point_cloud");
}
int b2s_reward(B2SWorld* w, const float* prev_xy, const float* next_xy, void* stream) {
  NEED_READY(w);
  if (w->d.Nmax > 64) return fail(B2S_E_CAPACITY, "b2s_reward: max_movables > 64");
  b2s_launch_reward(w->d, prev_xy, next_xy, (cudaStream_t)stream);
  return check_launch(w, "reward");
}

// one ncclAllGather of the episode returns (replaces tools/parallel_run.py).  NCCL is resolved at run
// time from the process (torch loads libnccl.so.2), so the library itself does not link against it.
int b2s_allgather_returns(B2SWorld* w, void* nccl_comm, float* out_dev, void* stream) {
  NEED_READY(w);
  if (!nccl_comm || !out_dev) return fail(B2S_E_INVALID, "b2s_allgather_returns: NULL argument");
  typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
  static allgather_fn fn = nullptr;
  if (!fn) {
    fn = (allgather_fn)dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (allgather_fn)dlsym(h, "ncclAllGather");
    }
    if (!fn) return fail(B2S_E_UNSUPPORTED, "b2s_allgather_returns: ncclAllGather not found in this process");
  }
  const int ncclFloat32 = 7;
  int rc = fn(w->d.buf.episode_return, out_dev, (size_t)w->d.B, ncclFloat32, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return fail(B2S_E_CUDA, "ncclAllGather failed with %d", rc);
  return 0;
}

int b2s_array(B2SWorld* w, int which, void** dev_ptr, int64_t* bytes) {
  NEED(w);
  if (!w->scene_loaded) return fail(B2S_E_STATE, "b2s_array: call b2s_load_scene first");
  if (which < 0 || which >= B2S_ARR_COUNT || !dev_ptr || !bytes) return fail(B2S_E_INVALID, "b2s_array: bad id %d", which);
  if (which == B2S_ARR_MANIFOLD_KEYS || which == B2S_ARR_MANIFOLD_NPTS || which == B2S_ARR_MANIFOLD_PTS) {
    // gather the current ping-pong side into the export arrays (default stream, synchronous)
    DeviceGuard device_guard_(w->device);
    CU(cudaDeviceSynchronize());
    b2s_launch_export_manifolds(w->d, w->exp_keys, w->exp_npts, w->exp_pts, 0);
    int rc = check_launch(w, "export_manifolds");
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
  }
  *dev_ptr = w->arr_ptr[which];
  *bytes = (int64_t)w->arr_bytes[which];
  return 0;
}

int64_t b2s_launch_count(const B2SWorld* w) { return w ? w->launches : 0; }

int64_t b2s_substeps_executed(B2SWorld* w, void* stream) {
  if (!w || !w->scene_loaded) return -1;
  DeviceGuard device_guard_(w->device);
  unsigned long long v = 0;
  cudaStreamSynchronize((cudaStream_t)stream);
  if (cudaMemcpy(&v, w->d.substeps, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

static int se3(int op, const float* a, const float* b, float* out, int n, void* stream) {
  if (!a || !out || n < 0) return fail(B2S_E_INVALID, "b2s_se3: bad argument");
  if (n == 0) return 0;
  b2s_launch_se3(op, a, b, out, n, (cudaStream_t)stream);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(B2S_E_CUDA, "se3 launch: %s", cudaGetErrorString(err));
  return 0;
}
int b2s_se3_quat_from_euler(const float* e, float* q, int n, void* s) { return se3(0, e, nullptr, q, n, s); }
int b2s_se3_euler_from_quat(const float* q, float* e, int n, void* s) { return se3(1, q, nullptr, e, n, s); }
int b2s_se3_matrix_from_quat(const float* q, float* m, int n, void* s) { return se3(2, q, nullptr, m, n, s); }
int b2s_se3_quat_multiply(const float* a, const float* b, float* o, int n, void* s) { if (!b) return fail(B2S_E_INVALID, "b2s_se3: NULL"); return se3(3, a, b, o, n, s); }
int b2s_se3_pose_inverse(const float* p, float* o, int n, void* s) { return se3(4, p, nullptr, o, n, s); }
int b2s_se3_pose_transform(const float* a, const float* b, float* o, int n, void* s) { if (!b) return fail(B2S_E_INVALID, "b2s_se3: NULL"); return se3(5, a, b, o, n, s); }

}  // extern "C"
