"""Host-side asset pipeline: convex-hull library, Sawyer stand-in and scene
flattening into the SoA arrays `b2s_load_scene` takes.

Stands for what `BulletPhysics.add_body` -> `pybullet.loadURDF` does with the
URDFs produced by tools/convert_obj_to_urdf.py (reference
robovat/simulation/physics/bullet_physics.py:143-186, tools/convert_obj_to_urdf.py:211-344,
tools/templates/urdf_template.xml:9-26): every body is one link whose collision
geometry is a compound of convex hulls (<= 64 vertices each, the V-HACD cap) and
whose frame is centred on the centre of mass.

The reference ships no assets, no configs and no Sawyer URDF (README.md:48-55),
so everything here is authored and labelled as an assumption:
  * movables: simple convex prisms plus concave L/T/U/C prisms given as exact
    unions of convex pieces (and, when present, V-HACD decompositions stored
    under robovat_b200/data/);
  * Sawyer: a 7-revolute-joint stand-in with the public Sawyer joint origins,
    limits and speed limits; collision links are boxes along each link.
"""
import ctypes as C
import json
import os

import numpy as np

from robovat_b200 import _capi

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
HULL_MARGIN = 0.001      # pybullet's margin for URDF convex meshes [upstream-recall]
MAX_HULL_VERTS = 64      # bin/vhacd --maxNumVerticesPerCH default


def quat_from_euler(roll, pitch, yaw):
    """'sxyz' Euler -> [x,y,z,w] (third_party/transformations.py:1194-1248)."""
    ci, si = np.cos(roll / 2.), np.sin(roll / 2.)
    cj, sj = np.cos(pitch / 2.), np.sin(pitch / 2.)
    ck, sk = np.cos(yaw / 2.), np.sin(yaw / 2.)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc, cj * cc + sj * ss])


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def box_vertices(hx, hy, hz, center=(0., 0., 0.)):
    v = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)],
                 dtype=np.float64)
    return v + np.asarray(center, dtype=np.float64)


def prism_vertices(n, radius, half_height, phase=0.0):
    ang = phase + 2 * np.pi * np.arange(n) / n
    ring = np.stack([radius * np.cos(ang), radius * np.sin(ang)], axis=1)
    return np.concatenate([np.c_[ring, np.full(n, -half_height)], np.c_[ring, np.full(n, half_height)]])


def polygon_prism(xy, half_height):
    xy = np.asarray(xy, dtype=np.float64)
    n = len(xy)
    return np.concatenate([np.c_[xy, np.full(n, -half_height)], np.c_[xy, np.full(n, half_height)]])


def segment_box(p0, p1, half_width):
    """Box of square cross-section whose axis is the segment p0 -> p1."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    d = p1 - p0
    ln = np.linalg.norm(d)
    if ln < 1e-9:
        return box_vertices(half_width, half_width, half_width, p0)
    z = d / ln
    a = np.array([1., 0., 0.]) if abs(z[0]) < 0.9 else np.array([0., 1., 0.])
    x = np.cross(a, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    out = []
    for sx in (-1, 1):
        for sy in (-1, 1):
            for t in (p0 - z * half_width, p1 + z * half_width):
                out.append(t + sx * half_width * x + sy * half_width * y)
    return np.array(out)


def _hull_volume_centroid(points):
    from scipy.spatial import ConvexHull
    hull = ConvexHull(points)
    c0 = points[hull.vertices].mean(axis=0)
    vol, cen = 0.0, np.zeros(3)
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = points[simplex]
        v = abs(np.dot(a - c0, np.cross(b - c0, c - c0))) / 6.0
        vol += v
        cen += v * (a + b + c + c0) / 4.0
    return vol, cen / vol


def _hull_planes(points):
    """Unique outward face planes (n, d) with n.x <= d inside."""
    from scipy.spatial import ConvexHull
    hull = ConvexHull(points)
    planes = []
    for eq in hull.equations:
        n, d = eq[:3], -eq[3]
        if not any(np.allclose(n, p[:3], atol=1e-6) and abs(d - p[3]) < 1e-7 for p in planes):
            planes.append(np.r_[n, d])
    return np.array(planes)


class AssetLibrary(object):
    """Vertex/plane pools of every convex hull plus the compounds ("assets") built from them."""

    def __init__(self):
        self.verts, self.planes = [], []
        self.hull_vert_off, self.hull_vert_cnt, self.hull_margin = [], [], []
        self.hull_plane_off, self.hull_plane_cnt = [], []
        self.asset_hull_off, self.asset_hull_cnt = [], []
        self.asset_names = {}
        self.asset_com = []

    def add_asset(self, name, hulls, margin=HULL_MARGIN, center_on_com=True):
        """hulls: list of [n,3] arrays in the asset frame.  Returns the asset id."""
        hulls = [np.asarray(h, dtype=np.float64) for h in hulls]
        if center_on_com:
            vols, cens = zip(*[_hull_volume_centroid(h) for h in hulls])
            com = np.average(np.array(cens), axis=0, weights=np.array(vols))
        else:
            com = np.zeros(3)
        aid = len(self.asset_hull_off)
        self.asset_hull_off.append(len(self.hull_vert_off))
        self.asset_hull_cnt.append(len(hulls))
        self.asset_com.append(com)
        for h in hulls:
            from scipy.spatial import ConvexHull
            pts = h - com
            pts = pts[ConvexHull(pts).vertices]          # keep extreme points only
            if len(pts) > MAX_HULL_VERTS:
                raise ValueError('hull of %s has %d > %d vertices' % (name, len(pts), MAX_HULL_VERTS))
            planes = _hull_planes(pts)
            self.hull_vert_off.append(len(self.verts))
            self.hull_vert_cnt.append(len(pts))
            self.hull_margin.append(margin)
            self.hull_plane_off.append(len(self.planes))
            self.hull_plane_cnt.append(len(planes))
            self.verts.extend(pts.tolist())
            self.planes.extend(planes.tolist())
        self.asset_names[name] = aid
        return aid

    def max_hulls(self, asset_ids):
        return max(self.asset_hull_cnt[a] for a in asset_ids)


# ---- movable shapes (assumed; the reference's meshes are an external download) -------------

def convex_movables():
    """Config #2: three convex hulls (box, hexagonal prism, wedge)."""
    return {
        'box': [box_vertices(0.04, 0.04, 0.025)],
        'hex': [prism_vertices(6, 0.04, 0.025)],
        'wedge': [polygon_prism([[-0.045, -0.03], [0.045, -0.03], [0.0, 0.045]], 0.025)],
    }


def concave_movables():
    """Config #3: eight concave prisms as unions of convex pieces (exact decompositions)."""
    h = 0.025
    t = 0.02            # arm thickness (half)
    shapes = {}
    shapes['L'] = [box_vertices(0.06, t, h, (0.0, -0.04, 0)), box_vertices(t, 0.06, h, (-0.04, 0.02, 0))]
    shapes['T'] = [box_vertices(0.06, t, h, (0.0, 0.04, 0)), box_vertices(t, 0.05, h, (0.0, -0.03, 0))]
    shapes['U'] = [box_vertices(0.06, t, h, (0.0, -0.04, 0)), box_vertices(t, 0.04, h, (-0.04, 0.02, 0)),
                   box_vertices(t, 0.04, h, (0.04, 0.02, 0))]
    shapes['C'] = [box_vertices(t, 0.06, h, (-0.04, 0.0, 0)), box_vertices(0.04, t, h, (0.02, 0.04, 0)),
                   box_vertices(0.04, t, h, (0.02, -0.04, 0))]
    shapes['plus'] = [box_vertices(0.06, t, h), box_vertices(t, 0.06, h)]
    shapes['Z'] = [box_vertices(0.04, t, h, (-0.02, 0.04, 0)), box_vertices(t, 0.06, h),
                   box_vertices(0.04, t, h, (0.02, -0.04, 0))]
    shapes['V'] = [segment_box((-0.05, 0.04, 0), (0.0, -0.04, 0), t)[:, :],
                   segment_box((0.05, 0.04, 0), (0.0, -0.04, 0), t)[:, :]]
    shapes['H'] = [box_vertices(t, 0.06, h, (-0.04, 0, 0)), box_vertices(t, 0.06, h, (0.04, 0, 0)),
                   box_vertices(0.02, t, h)]
    # flatten V pieces to the common height
    for k in ('V',):
        shapes[k] = [np.c_[p[:, :2], np.clip(p[:, 2], -h, h)] for p in shapes[k]]
    return shapes


def load_vhacd_movables():
    """V-HACD decompositions generated offline by tools/make_vhacd_assets.py (bin/vhacd defaults)."""
    path = os.path.join(DATA_DIR, 'vhacd_movables.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        data = json.load(f)
    return {k: [np.array(h) for h in v] for k, v in data.items()}


# ---- Sawyer stand-in -------------------------------------------------------------------------

SAWYER = {
    # joint frame in the parent link frame: xyz, rpy  (public sawyer.urdf.xacro values, from memory)
    'joint_names': ['right_j0', 'right_j1', 'right_j2', 'right_j3', 'right_j4', 'right_j5', 'right_j6'],
    'origin_xyz': [[0, 0, 0.08], [0.081, 0.05, 0.237], [0, -0.14, 0.1425], [0, -0.042, 0.26],
                   [0, -0.125, -0.1265], [0, 0.031, 0.275], [0, -0.11, 0.1053]],
    'origin_rpy': [[0, 0, 0], [-1.57079632679, 1.57079632679, 0], [1.57079632679, 0, 0],
                   [-1.57079632679, 0, 0], [1.57079632679, 0, 0], [-1.57079632679, 0, 0],
                   [-1.57079632679, -0.17453, 3.14159265359]],
    'axis': [[0, 0, 1]] * 7,
    'lower': [-3.0503, -3.8095, -3.0426, -3.0439, -2.9761, -2.9761, -4.7124],
    'upper': [3.0503, 2.2736, 3.0426, 3.0439, 2.9761, 2.9761, 4.7124],
    'max_velocity': [1.74, 1.328, 1.957, 1.957, 3.485, 3.485, 4.545],
    'neutral': [0.0, -1.18, 0.0, 2.18, 0.0, 0.57, 3.3161],
    # right_hand (END_EFFCTOR_NAME) in the right_l6 frame
    'ee_xyz': [0, 0, 0.0245], 'ee_rpy': [0, 0, 1.57079632679],
}


def add_sawyer(lib, finger_length, link_half_width=0.04, finger_half_width=0.012):
    """Adds the arm's collision links to the library; returns the arm part of the scene."""
    s = SAWYER
    links = []      # (link_joint, asset id, pose[7])
    ident = [0, 0, 0, 0, 0, 0, 1.0]
    # pedestal/base: box under joint 0
    links.append((-1, lib.add_asset('sawyer_base', [box_vertices(0.09, 0.09, 0.04, (0, 0, 0.04))],
                                    center_on_com=False), ident))
    for j in range(6):
        nxt = np.array(s['origin_xyz'][j + 1], dtype=float)
        links.append((j, lib.add_asset('sawyer_l%d' % j, [segment_box((0, 0, 0), nxt, link_half_width)],
                                       center_on_com=False), ident))
    ee = np.array(s['ee_xyz'], dtype=float)
    links.append((6, lib.add_asset('sawyer_l6', [segment_box((0, 0, 0), ee, link_half_width)],
                                   center_on_com=False), ident))
    # pusher finger along +z of the end-effector frame (which points down when euler = (pi,0,0))
    ee_q = quat_from_euler(*s['ee_rpy'])
    R = quat_to_matrix(ee_q)
    tip0 = ee + R.dot([0, 0, 0.0])
    tip1 = ee + R.dot([0, 0, finger_length - finger_half_width])
    links.append((6, lib.add_asset('sawyer_finger', [segment_box(tip0, tip1, finger_half_width)],
                                   center_on_com=False), ident))
    return {
        'joint_origin': [list(xyz) + list(quat_from_euler(*rpy)) for xyz, rpy in
                         zip(s['origin_xyz'], s['origin_rpy'])],
        'joint_axis': s['axis'], 'lower': s['lower'], 'upper': s['upper'],
        'max_velocity': s['max_velocity'], 'ee_pose': list(ee) + list(ee_q), 'links': links,
    }


def _arr(values, ctype):
    a = np.ascontiguousarray(values, dtype={C.c_float: np.float32, C.c_int32: np.int32,
                                            C.c_uint32: np.uint32}[ctype]).ravel()
    return a, a.ctypes.data_as(C.POINTER(ctype))


class Scene(object):
    """A flattened scene: owns the numpy arrays a B2SSceneDesc points into."""

    def __init__(self, lib, statics, arm, movable_assets, target_assets, layout, sampling,
                 arm_base_pose, arm_friction):
        self.lib = lib
        self.statics = statics
        d = _capi.B2SSceneDesc()
        self._keep = []

        def bind(field, values, ctype, count_field=None, count=None):
            a, p = _arr(values, ctype)
            self._keep.append(a)
            setattr(d, field, p)
            if count_field:
                setattr(d, count_field, count if count is not None else len(values))

        bind('verts', lib.verts, C.c_float, 'num_verts', len(lib.verts))
        bind('hull_vert_off', lib.hull_vert_off, C.c_int32, 'num_hulls')
        bind('hull_vert_cnt', lib.hull_vert_cnt, C.c_int32)
        bind('hull_margin', lib.hull_margin, C.c_float)
        bind('planes', lib.planes, C.c_float, 'num_planes', len(lib.planes))
        bind('hull_plane_off', lib.hull_plane_off, C.c_int32)
        bind('hull_plane_cnt', lib.hull_plane_cnt, C.c_int32)
        bind('asset_hull_off', lib.asset_hull_off, C.c_int32, 'num_assets')
        bind('asset_hull_cnt', lib.asset_hull_cnt, C.c_int32)
        bind('static_asset', [s['asset'] for s in statics], C.c_int32, 'num_statics')
        bind('static_pose', [s['pose'] for s in statics], C.c_float)
        bind('static_friction', [s['friction'] for s in statics], C.c_float)
        bind('static_flags', [s['flags'] for s in statics], C.c_uint32)
        bind('movable_assets', movable_assets, C.c_int32, 'num_movable_assets')
        bind('target_assets', target_assets if len(target_assets) else [0], C.c_int32)
        d.num_target_assets = len(target_assets)
        d.arm_base_pose[:] = arm_base_pose
        for j in range(_capi.NUM_JOINTS):
            d.joint_origin[j][:] = arm['joint_origin'][j]
            d.joint_axis[j][:] = arm['joint_axis'][j]
        d.joint_lower[:] = arm['lower']
        d.joint_upper[:] = arm['upper']
        d.joint_max_velocity[:] = arm['max_velocity']
        d.ee_pose[:] = arm['ee_pose']
        d.num_links = len(arm['links'])
        for k, (jj, asset, pose) in enumerate(arm['links']):
            d.link_joint[k] = jj
            d.link_asset[k] = asset
            d.link_pose[k][:] = pose
        d.arm_friction = arm_friction
        d.tile_size = layout.get('size', 0.15)
        d.tile_offset[:] = layout.get('offset', [0.295, -0.485])
        for key in ('region', 'goal', 'target', 'obstacle'):
            tiles = layout.get(key) or []
            if len(tiles) > _capi.MAX_TILES:
                raise ValueError('layout.%s has more than %d tiles' % (key, _capi.MAX_TILES))
            setattr(d, 'num_' + key, len(tiles))
            arr = getattr(d, key)
            for i, t in enumerate(tiles):
                arr[i][:] = [float(t[0]), float(t[1])]
        for key in ('scale_range', 'mass_range', 'friction_range', 'pose_x', 'pose_y', 'pose_z',
                    'pose_roll', 'pose_pitch', 'pose_yaw', 'table_height_range'):
            getattr(d, key)[:] = sampling[key]
        d.placement_margin = sampling['placement_margin']
        d.min_movables = sampling['min_movables']
        d.safe_drop_height = sampling['safe_drop_height']
        self.desc = d
        self.num_links = d.num_links
        self.num_statics = d.num_statics
        static_hulls = sum(lib.asset_hull_cnt[s['asset']] for s in statics
                           if not (s['flags'] & _capi.STATIC_NO_COLLIDE))
        arm_hulls = sum(lib.asset_hull_cnt[a] for _, a, _ in arm['links'])
        self.fixed_colliders = static_hulls + arm_hulls
        # tiles that collide are separate static bodies: a movable lying across several of them has one manifold (up to
        # four points) with each, which is what the contact capacity has to hold (config.build_params)
        self.colliding_tiles = sum(1 for s in statics if (s['flags'] & _capi.STATIC_IS_TILE) and not (s['flags'] & _capi.STATIC_NO_COLLIDE))
        self.max_movable_hulls = lib.max_hulls(list(movable_assets) + list(target_assets))
