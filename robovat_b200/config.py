"""PushEnv configuration (schema reconstructed from the reference code; values authored).

The reference keeps its YAML configs outside the repository (README.md:48-55), so
only the *keys* are known (push_env.py, arm_env.py,
robot_env.py, sawyer_sim.py read them).  Every value below is an assumption and is
commented with the hint it was derived from.  `AttrDict` stands for easydict
(not installed here); `load_yaml` keeps `YamlConfig`'s plain-YAML behaviour
(robovat/utils/yaml_config.py:21-153) for user overrides.
"""
import copy
import os
import math

import numpy as np

from robovat_b200 import _capi
from robovat_b200 import assets as assets_lib
from robovat_b200 import layouts as push_layouts
from robovat_b200 import mesh_io


class AttrDict(dict):
    """Attribute access over a dict (what the reference gets from easydict)."""

    def __init__(self, d=None):
        super(AttrDict, self).__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super(AttrDict, self).__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


DEFAULT_PUSH_ENV = {
    'DEBUG': False,
    'MAX_STEPS': 8,
    'SUCCESS_THRESH': 50.0,
    'TASK_NAME': None,                 # None | data_collection | clearing | insertion | crossing
    'LAYOUT_ID': 0,
    'NUM_GOAL_STEPS': None,
    'MIN_MOVABLE_BODIES': 3,
    'MAX_MOVABLE_BODIES': 3,
    'MOVABLE_NAME': 'convex',
    'MOVABLE': {
        'CONVEX': {
            'PATHS': ['box', 'hex', 'wedge'], 'TARGET_PATHS': ['box'],
            'SCALE': [0.8, 1.2], 'MASS': [0.05, 0.3], 'FRICTION': [0.4, 1.0], 'MARGIN': 0.12,
            # uniform over the obstacle grid of crossing/0 (layouts.py:182-186)
            'POSE': {'X': [0.37, 0.82], 'Y': [-0.41, 0.49], 'Z': 0.2,
                     'ROLL': [-math.pi, math.pi], 'PITCH': [-math.pi / 2, math.pi / 2],
                     'YAW': [-math.pi, math.pi]},
        },
        'VHACD': {                     # V-HACD decompositions written by tools/make_vhacd_assets.py (reference bin/vhacd)
            'PATHS': ['urdf/L/L.urdf', 'urdf/T/T.urdf', 'urdf/U/U.urdf', 'urdf/plus/plus.urdf'],
            'TARGET_PATHS': ['urdf/L/L.urdf'],
            'SCALE': [0.8, 1.0], 'MASS': [0.05, 0.3], 'FRICTION': [0.4, 1.0], 'MARGIN': 0.12,
            'POSE': {'X': [0.37, 0.82], 'Y': [-0.41, 0.49], 'Z': 0.2,
                     'ROLL': [-math.pi, math.pi], 'PITCH': [-math.pi / 2, math.pi / 2],
                     'YAW': [-math.pi, math.pi]},
        },
        'CONCAVE': {
            'PATHS': ['L', 'T', 'U', 'C', 'plus', 'Z', 'V', 'H'], 'TARGET_PATHS': ['L'],
            'SCALE': [0.8, 1.0], 'MASS': [0.05, 0.3], 'FRICTION': [0.4, 1.0], 'MARGIN': 0.12,
            'POSE': {'X': [0.37, 0.82], 'Y': [-0.41, 0.49], 'Z': 0.2,
                     'ROLL': [-math.pi, math.pi], 'PITCH': [-math.pi / 2, math.pi / 2],
                     'YAW': [-math.pi, math.pi]},
        },
    },
    'USE_RANDOM_RGBA': False,
    'USE_PRESTIGE_OBS': True,          # HeuristicPushPolicy reads observation['position']
    'USE_VISUALIZATION_OBS': False,
    'OBS': {'NUM_POINTS': 256, 'CROP_MIN': None, 'CROP_MAX': None},
    'ACTION': {
        # reachable part of the table for the Sawyer stand-in; border hints push_reward.py:243-245
        'CSPACE': {'LOW': [0.35, -0.35, 0.0], 'HIGH': [0.85, 0.35, 0.04]},
        'MOTION': {'TRANSLATION_X': 0.2, 'TRANSLATION_Y': 0.2},
        'MIN_DELTA_POSITION': 0.01, 'MIN_DELTA_ANGLE': 0.05,
    },
    'ARM': {
        'FINGER_TIP_OFFSET': 0.125, 'GRIPPER_SAFE_HEIGHT': 0.35,
        'OFFSTAGE_POSITIONS': [-1.5, -1.26, 0.00, 1.98, 0.00, 0.85, 3.3161],
    },
    'TABLE': {'HEIGHT_RANGE': [0.0, 0.0], 'X_RANGE': 0.76, 'Y_RANGE': 1.22},   # layouts.py:30
    'SIM': {
        'TIME_STEP': 1e-3,             # Simulator default (simulator.py:26)
        'ARM': {'CONFIG': 'sawyer'},
        # URDF origin of the table at its top surface (tiles sit at table.z + 0.001 - 0.025, push_env.py:349)
        'GROUND': {'POSE': [[0, 0, -0.9], [0, 0, 0]]},
        'TABLE': {'POSE': [[0.6, 0.0, 0.0], [0, 0, 0]], 'THICKNESS': 0.05, 'FRICTION': 1.0},
        # the reference loads SIM.WALL.PATH (a URDF of its data package, which is not in the repository): here the wall is
        # a static box given by POSE (centre) and SIZE (full extents), or a URDF from the asset pipeline in PATH
        'WALL': {'USE': False, 'POSE': [[1.1, 0.0, 0.25], [0, 0, 0]], 'SIZE': [0.05, 1.6, 1.4], 'PATH': None, 'FRICTION': 1.0},
        'TILE': {'HEIGHT': 0.05, 'COLLIDE': True},
        'STEPS_CHECK': 20, 'MAX_PHASE_STEPS': 3000, 'MAX_MOTION_STEPS': 4000, 'MAX_OFFSTAGE_STEPS': 4000,
    },
    'ROBOT': {                         # sawyer_sim.py config keys
        'LIMB_MAX_VELOCITY_RATIO': 0.5, 'LIMB_TIMEOUT': 15.0, 'LIMB_POSITION_THRESHOLD': 0.008726640,
        'CLAMP_JOINT_VELOCITY': True, 'ARM_POSE': [[0, 0, 0.0], [0, 0, 0]], 'ARM_FRICTION': 1.0,
    },
    'KINECT2': {'DEPTH': {
        'HEIGHT': 424, 'WIDTH': 512,                     # bullet_camera.py:22-23
        'INTRINSICS': [365.0, 0, 256.0, 0, 365.0, 212.0, 0, 0, 1],
        # camera 1.2 m above the table centre looking straight down: x_cam = R x_world + t
        'TRANSLATION': [0.6, 0.0, 1.2], 'ROTATION': [math.pi, 0, 0],
        'INTRINSICS_NOISE': None, 'TRANSLATION_NOISE': None, 'ROTATION_NOISE': None}},
    'RECORDING': {'USE': False},
    'PHYSICS': {                       # Bullet defaults, SURVEY.md 3.4 [upstream-recall]
        'SOLVER_ITERATIONS': 50, 'FRICTION_DIRS': 2, 'ERP2': 0.08, 'LINEAR_SLOP': 1e-5,
        'WARMSTART': 0.85, 'RESIDUAL_THRESHOLD': 1e-7, 'LINEAR_DAMPING': 0.04, 'ANGULAR_DAMPING': 0.04,
        'BREAKING_FACTOR': 0.02, 'GRAVITY': [0, 0, -9.8],
        # torsional friction of every movable URDF (tools/templates/urdf_template.xml:12-14); 0 = rows off
        'ROLLING_FRICTION': 0.001, 'SPINNING_FRICTION': 0.001,
    },
}

DEFAULT_POLICY = {
    'ACTION': {'CSPACE': DEFAULT_PUSH_ENV['ACTION']['CSPACE'], 'MOTION': DEFAULT_PUSH_ENV['ACTION']['MOTION']},
    'HEURISTICS': {'MAX_ATTEMPS': 2000},
}


def default_push_env_config(**bindings):
    """Deep copy of the defaults with `--config_bindings`-style overrides (tools/run_env.py:169-173)."""
    cfg = AttrDict(copy.deepcopy(DEFAULT_PUSH_ENV))
    for k, v in bindings.items():
        cfg[k] = v
    return cfg


def default_policy_config(**bindings):
    cfg = AttrDict(copy.deepcopy(DEFAULT_POLICY))
    for k, v in bindings.items():
        cfg[k] = v
    return cfg


def load_yaml(path, base=None):
    import yaml
    with open(path) as f:
        data = yaml.safe_load(f) or {}
    cfg = AttrDict(copy.deepcopy(base if base is not None else DEFAULT_PUSH_ENV))

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                merge(dst[k], v)
            else:
                dst[k] = v
    merge(cfg, data)
    return cfg


def _range(v):
    return [float(v), float(v)] if isinstance(v, (int, float)) else [float(v[0]), float(v[1])]


def build_scene(config):
    """Flatten a PushEnv config into (AssetLibrary, Scene): the host URDF-loader stand-in."""
    cfg = config
    lib = assets_lib.AssetLibrary()
    table = cfg.SIM.TABLE
    tx, ty, tz = table.POSE[0]
    th = float(table.THICKNESS)
    statics = []
    ground = lib.add_asset('ground', [assets_lib.box_vertices(5.0, 5.0, 0.05, (0, 0, -0.05))], center_on_com=False)
    gpos = cfg.SIM.GROUND.POSE[0]
    statics.append({'name': 'ground', 'asset': ground, 'pose': list(gpos) + [0, 0, 0, 1.0], 'friction': 1.0, 'flags': 0})
    tab = lib.add_asset('table', [assets_lib.box_vertices(0.5 * cfg.TABLE.X_RANGE, 0.5 * cfg.TABLE.Y_RANGE, 0.5 * th,
                                                          (0, 0, -0.5 * th))], center_on_com=False)
    statics.append({'name': 'table', 'asset': tab, 'pose': [tx, ty, tz, 0, 0, 0, 1.0], 'friction': float(table.FRICTION),
                    'flags': _capi.STATIC_ON_TABLE | _capi.STATIC_IS_TABLE})
    if cfg.SIM.WALL.get('USE', False):                         # ArmEnv._reset_scene (arm_env.py:93-98): third static body
        wall = cfg.SIM.WALL
        if wall.get('PATH'):
            path = wall.PATH if os.path.isabs(wall.PATH) else os.path.join(assets_lib.DATA_DIR, wall.PATH)
            hulls, _ = mesh_io.urdf_asset(path)
            wid = lib.add_asset('wall', hulls, center_on_com=False)
        else:
            sx, sy, sz = [0.5 * float(v) for v in wall.SIZE]
            wid = lib.add_asset('wall', [assets_lib.box_vertices(sx, sy, sz)], center_on_com=False)
        wq = assets_lib.quat_from_euler(*[float(v) for v in wall.POSE[1]])
        statics.append({'name': 'wall', 'asset': wid, 'pose': [float(v) for v in wall.POSE[0]] + [float(v) for v in wq],
                        'friction': float(wall.get('FRICTION', 1.0)), 'flags': 0})
    layout = {}
    task = cfg.TASK_NAME
    if task not in (None, 'data_collection'):
        lay = push_layouts.TASK_NAME_TO_LAYOUTS[task][cfg.LAYOUT_ID]
        layout = {'size': lay.size, 'offset': lay.offset, 'region': lay.region, 'goal': lay.goal,
                  'target': lay.target, 'obstacle': lay.obstacle}
        tile_h = float(cfg.SIM.TILE.HEIGHT)
        tile = lib.add_asset('tile', [assets_lib.box_vertices(0.5 * lay.size, 0.5 * lay.size, 0.5 * tile_h)],
                             center_on_com=False)
        flags = _capi.STATIC_ON_TABLE | _capi.STATIC_IS_TILE | (0 if cfg.SIM.TILE.COLLIDE else _capi.STATIC_NO_COLLIDE)
        # PushEnv._load_tiles: region at z_offset 0.001 - 0.025, goal at 0.0015 - 0.025 (push_env.py:343-357)
        for i, c in enumerate(lay.region):
            statics.append({'name': 'tile_%d' % i, 'asset': tile, 'friction': 1.0, 'flags': flags,
                            'pose': [lay.offset[0] + c[0] * lay.size, lay.offset[1] + c[1] * lay.size,
                                     tz + 0.001 - 0.025, 0, 0, 0, 1.0]})
        for i, c in enumerate(lay.goal or []):
            statics.append({'name': 'tile_%d' % i, 'asset': tile, 'friction': 1.0, 'flags': flags,
                            'pose': [lay.offset[0] + c[0] * lay.size, lay.offset[1] + c[1] * lay.size,
                                     tz + 0.0015 - 0.025, 0, 0, 0, 1.0]})
    arm = assets_lib.add_sawyer(lib, finger_length=float(cfg.ARM.FINGER_TIP_OFFSET))
    mname = cfg.MOVABLE_NAME.upper()
    mcfg = cfg.MOVABLE[mname]
    shapes = {}
    shapes.update(assets_lib.convex_movables())
    shapes.update(assets_lib.concave_movables())
    vh = assets_lib.load_vhacd_movables()
    if vh:
        shapes.update({'vhacd_' + k: v for k, v in vh.items()})
    ids = {}
    for name in list(mcfg.PATHS) + list(mcfg.TARGET_PATHS):
        if name in ids:
            continue
        if name.endswith('.urdf'):
            # a URDF from the reference asset pipeline (tools/convert_obj_to_urdf.py): hulls in the inertial frame
            path = name if os.path.isabs(name) else os.path.join(assets_lib.DATA_DIR, name)
            hulls, _ = mesh_io.urdf_asset(path)
            ids[name] = lib.add_asset(name, hulls, center_on_com=False)
        else:
            ids[name] = lib.add_asset(name, shapes[name])
    movable_assets = [ids[n] for n in mcfg.PATHS]
    target_assets = [ids[n] for n in mcfg.TARGET_PATHS]
    pose = mcfg.POSE
    sampling = {
        'scale_range': _range(mcfg.SCALE), 'mass_range': _range(mcfg.MASS), 'friction_range': _range(mcfg.FRICTION),
        'pose_x': _range(pose.X), 'pose_y': _range(pose.Y), 'pose_z': _range(pose.Z),
        'pose_roll': _range(pose.ROLL), 'pose_pitch': _range(pose.PITCH), 'pose_yaw': _range(pose.YAW),
        'placement_margin': float(mcfg.MARGIN), 'min_movables': int(cfg.MIN_MOVABLE_BODIES),
        'table_height_range': _range(cfg.TABLE.HEIGHT_RANGE), 'safe_drop_height': 0.2,
    }
    apos, aeul = cfg.ROBOT.ARM_POSE
    arm_base = list(apos) + list(assets_lib.quat_from_euler(*aeul))
    scene = assets_lib.Scene(lib, statics, arm, movable_assets, target_assets, layout, sampling, arm_base,
                             float(cfg.ROBOT.ARM_FRICTION))
    return scene


def build_params(config, scene, num_envs, env_id_offset=0, lib=None, **overrides):
    """B2SParams from a PushEnv config (+ capacities derived from the scene)."""
    cfg = config
    p = _capi.B2SParams()
    if lib is not None:
        _capi.check(lib, lib.b2s_default_params(p))
    phys = cfg.PHYSICS
    nmax = int(cfg.MAX_MOVABLE_BODIES)
    p.num_envs, p.env_id_offset, p.max_movables = int(num_envs), int(env_id_offset), nmax
    movable_hulls = nmax * scene.max_movable_hulls
    p.max_colliders = scene.fixed_colliders + movable_hulls
    p.max_manifolds = max(32, 6 * movable_hulls)
    p.max_pairs = max(64, 2 * p.max_manifolds)
    p.envs_per_block = int(os.environ.get('B2S_EPB', 0))     # 0 = library default
    p.export_debug = int(os.environ.get('B2S_EXPORT_DEBUG', 0))
    p.num_goal_steps = int(cfg.NUM_GOAL_STEPS or 0)               # push_env.py:259-262
    # <= 32 contact points keeps the solver's Jacobian rows in registers (one contact per lane): enough for movables on
    # the table (four points each + the pusher).  On colliding tiles a movable rests on up to four bodies at once
    # (measured at rest, 3 convex movables: clearing layouts up to 50 points, 30 % of the env-substeps above 32; crossing
    # up to 37), and a capacity that is too small drops contact points (error flag 8): such scenes get the room and with
    # it the record-based solve.
    # (20 per movable BODY: the hulls of a multi-hull movable share its supports -- 8 concave movables of 3 hulls measure
    # up to 142 points against the 192 that 8 per hull provide)
    p.max_contacts = max(32, 8 * movable_hulls)
    if getattr(scene, 'colliding_tiles', 0) > 0:
        p.max_contacts = max(p.max_contacts, 20 * nmax + 4)
    p.solver_iterations, p.friction_dirs = int(phys.SOLVER_ITERATIONS), int(phys.FRICTION_DIRS)
    p.gjk_max_iters, p.epa_max_iters, p.ik_max_iters = 32, 32, 20
    p.ik_interval, p.check_done_interval = 10, 100
    p.steps_check = int(cfg.SIM.STEPS_CHECK)
    p.max_phase_steps, p.max_motion_steps = int(cfg.SIM.MAX_PHASE_STEPS), int(cfg.SIM.MAX_MOTION_STEPS)
    p.max_offstage_steps = int(cfg.SIM.MAX_OFFSTAGE_STEPS)
    p.stable_check_after, p.stable_min_steps, p.stable_max_steps = 100, 100, 2000
    p.clamp_joint_velocity = 1 if cfg.ROBOT.CLAMP_JOINT_VELOCITY else 0
    p.cam_height, p.cam_width = int(cfg.KINECT2.DEPTH.HEIGHT), int(cfg.KINECT2.DEPTH.WIDTH)
    p.num_points = int(cfg.OBS.NUM_POINTS)
    if cfg.OBS.CROP_MIN is not None and cfg.OBS.CROP_MAX is not None:        # camera_obs.py:143-150, 187-193
        p.use_crop = 1
        p.crop_min[:] = [float(x) for x in cfg.OBS.CROP_MIN]
        p.crop_max[:] = [float(x) for x in cfg.OBS.CROP_MAX]
    p.task = _capi.TASK_IDS[cfg.TASK_NAME]
    if os.environ.get('B2S_WARPS'):
        p.warps_per_block = int(os.environ['B2S_WARPS'])           # default: what the library was built for
    p.time_step = float(cfg.SIM.TIME_STEP)
    p.gravity[:] = phys.GRAVITY
    p.erp2, p.linear_slop, p.warmstart = phys.ERP2, phys.LINEAR_SLOP, phys.WARMSTART
    p.residual_threshold = phys.RESIDUAL_THRESHOLD
    p.linear_damping, p.angular_damping = phys.LINEAR_DAMPING, phys.ANGULAR_DAMPING
    p.breaking_factor = phys.BREAKING_FACTOR
    p.rolling_friction = float(phys.get('ROLLING_FRICTION', 0.001))
    p.spinning_friction = float(phys.get('SPINNING_FRICTION', 0.001))
    p.ik_damping, p.ik_residual, p.ik_max_step = 0.1, 1e-4, math.pi / 4
    p.position_gain, p.velocity_gain = 0.05, 1.0
    p.joint_pos_threshold = float(cfg.ROBOT.LIMB_POSITION_THRESHOLD)
    p.joint_vel_threshold = 0.05
    p.limb_timeout = float(cfg.ROBOT.LIMB_TIMEOUT)
    p.limb_velocity_ratio = float(cfg.ROBOT.LIMB_MAX_VELOCITY_RATIO)
    p.stable_lin_threshold = p.stable_ang_threshold = 0.005
    p.cspace_low[:] = cfg.ACTION.CSPACE.LOW
    p.cspace_high[:] = cfg.ACTION.CSPACE.HIGH
    p.translation_x, p.translation_y = cfg.ACTION.MOTION.TRANSLATION_X, cfg.ACTION.MOTION.TRANSLATION_Y
    p.finger_tip_offset, p.gripper_safe_height = cfg.ARM.FINGER_TIP_OFFSET, cfg.ARM.GRIPPER_SAFE_HEIGHT
    p.offstage_positions[:] = cfg.ARM.OFFSTAGE_POSITIONS
    p.min_delta_position, p.min_delta_angle = cfg.ACTION.MIN_DELTA_POSITION, cfg.ACTION.MIN_DELTA_ANGLE
    tx, ty = cfg.SIM.TABLE.POSE[0][0], cfg.SIM.TABLE.POSE[0][1]
    p.table_workspace_low[:] = [tx - 0.5 * cfg.TABLE.X_RANGE, ty - 0.5 * cfg.TABLE.Y_RANGE]
    p.table_workspace_high[:] = [tx + 0.5 * cfg.TABLE.X_RANGE, ty + 0.5 * cfg.TABLE.Y_RANGE]
    p.cam_near, p.cam_far = 0.02, 100.0
    for k, v in overrides.items():
        setattr(p, k, v)
    return p
