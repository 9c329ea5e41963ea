"""ctypes mirror of include/b2s.h and the loader of libb2s.so.

Host code stays Python, as in the reference (robovat is pure Python over the
pybullet C extension, robovat/simulation/physics/bullet_physics.py:28).  The
structures below must match include/b2s.h field for field; ``load()`` checks
their sizes against ``b2s_sizeof`` so a drift fails loudly at import time.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded,
``load()`` raises.
"""
import ctypes as C
import os

NUM_JOINTS = 7
MAX_LINKS = 12
MAX_TILES = 32
CP_FLOATS = 16
CTRL_FLOATS = 40

OK, E_INVALID, E_CUDA, E_STATE, E_CAPACITY, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

(PHASE_INITIAL, PHASE_PRE, PHASE_START, PHASE_MOTION, PHASE_POST, PHASE_OFFSTAGE,
 PHASE_DONE, PHASE_SETTLE, PHASE_IDLE, PHASE_RESET_DROP, PHASE_RESET_WAIT) = range(11)
PHASE_NAMES = ['initial', 'pre', 'start', 'motion', 'post', 'offstage', 'done',
               'settle', 'idle', 'reset_drop', 'reset_wait']

POLICY_HEURISTIC, POLICY_AIMED = 0, 1
TASK_NONE, TASK_CLEARING, TASK_INSERTION, TASK_CROSSING = range(4)
TASK_IDS = {None: TASK_NONE, 'data_collection': TASK_NONE, 'clearing': TASK_CLEARING,
            'insertion': TASK_INSERTION, 'crossing': TASK_CROSSING}

STATIC_ON_TABLE, STATIC_IS_TABLE, STATIC_NO_COLLIDE, STATIC_IS_TILE = 1, 2, 4, 8

(ARR_MANIFOLD_KEYS, ARR_MANIFOLD_NPTS, ARR_MANIFOLD_PTS, ARR_NUM_MANIFOLDS, ARR_PAIR_KEYS,
 ARR_NUM_PAIRS, ARR_PHASE, ARR_NUM_STEPS, ARR_CTRL, ARR_CTRL_FLAGS, ARR_LINK_POSES,
 ARR_MOV_PARAMS, ARR_TABLE_DZ, ARR_ERROR_FLAGS, ARR_WAYPOINTS, ARR_STATUS, ARR_CONTACT_FLAGS,
 ARR_PHASE_STATE, ARR_SOLVER_STATS, ARR_CTRL_TIME, ARR_LINK_VEL, ARR_NUM_COLLIDERS,
 ARR_COL_SLOT, ARR_COL_HULL, ARR_PROF, ARR_NUM_EPISODES, ARR_ROLLOUT_STATE, ARR_RAY_SCENE) = range(28)

f32, i32, u32, u8, f64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint8, C.c_double
P = C.POINTER


class B2SParams(C.Structure):
    _fields_ = [
        ('num_envs', i32), ('env_id_offset', i32), ('max_movables', i32), ('max_pairs', i32),
        ('max_manifolds', i32), ('solver_iterations', i32), ('friction_dirs', i32),
        ('gjk_max_iters', i32), ('epa_max_iters', i32), ('ik_max_iters', i32),
        ('ik_interval', i32), ('check_done_interval', i32), ('steps_check', i32),
        ('max_phase_steps', i32), ('max_motion_steps', i32), ('max_offstage_steps', i32),
        ('stable_check_after', i32), ('stable_min_steps', i32), ('stable_max_steps', i32),
        ('clamp_joint_velocity', i32), ('cam_height', i32), ('cam_width', i32),
        ('num_points', i32), ('task', i32), ('max_contacts', i32), ('max_colliders', i32),
        ('warps_per_block', i32), ('envs_per_block', i32), ('export_debug', i32), ('num_goal_steps', i32), ('use_crop', i32),
        ('time_step', f64), ('gravity', f32 * 3), ('erp2', f32), ('linear_slop', f32),
        ('warmstart', f32), ('residual_threshold', f32), ('linear_damping', f32),
        ('angular_damping', f32), ('breaking_factor', f32), ('ik_damping', f32),
        ('ik_residual', f32), ('ik_max_step', f32), ('position_gain', f32),
        ('velocity_gain', f32), ('joint_pos_threshold', f32), ('joint_vel_threshold', f32),
        ('limb_timeout', f32), ('limb_velocity_ratio', f32), ('stable_lin_threshold', f32),
        ('stable_ang_threshold', f32), ('cspace_low', f32 * 3), ('cspace_high', f32 * 3),
        ('translation_x', f32), ('translation_y', f32), ('finger_tip_offset', f32),
        ('gripper_safe_height', f32), ('offstage_positions', f32 * NUM_JOINTS),
        ('min_delta_position', f32), ('min_delta_angle', f32),
        ('table_workspace_low', f32 * 2), ('table_workspace_high', f32 * 2),
        ('cam_near', f32), ('cam_far', f32), ('crop_min', f32 * 3), ('crop_max', f32 * 3), ('rolling_friction', f32), ('spinning_friction', f32),
    ]


class B2SSceneDesc(C.Structure):
    _fields_ = [
        ('num_verts', i32), ('verts', P(f32)),
        ('num_hulls', i32), ('hull_vert_off', P(i32)), ('hull_vert_cnt', P(i32)),
        ('hull_margin', P(f32)),
        ('num_planes', i32), ('planes', P(f32)), ('hull_plane_off', P(i32)),
        ('hull_plane_cnt', P(i32)),
        ('num_assets', i32), ('asset_hull_off', P(i32)), ('asset_hull_cnt', P(i32)),
        ('num_statics', i32), ('static_asset', P(i32)), ('static_pose', P(f32)),
        ('static_friction', P(f32)), ('static_flags', P(u32)),
        ('num_movable_assets', i32), ('movable_assets', P(i32)),
        ('num_target_assets', i32), ('target_assets', P(i32)),
        ('arm_base_pose', f32 * 7), ('joint_origin', (f32 * 7) * NUM_JOINTS),
        ('joint_axis', (f32 * 3) * NUM_JOINTS), ('joint_lower', f32 * NUM_JOINTS),
        ('joint_upper', f32 * NUM_JOINTS), ('joint_max_velocity', f32 * NUM_JOINTS),
        ('ee_pose', f32 * 7), ('num_links', i32), ('link_joint', i32 * MAX_LINKS),
        ('link_asset', i32 * MAX_LINKS), ('link_pose', (f32 * 7) * MAX_LINKS),
        ('arm_friction', f32),
        ('tile_size', f32), ('tile_offset', f32 * 2),
        ('num_region', i32), ('region', (f32 * 2) * MAX_TILES),
        ('num_goal', i32), ('goal', (f32 * 2) * MAX_TILES),
        ('num_target', i32), ('target', (f32 * 2) * MAX_TILES),
        ('num_obstacle', i32), ('obstacle', (f32 * 2) * MAX_TILES),
        ('scale_range', f32 * 2), ('mass_range', f32 * 2), ('friction_range', f32 * 2),
        ('pose_x', f32 * 2), ('pose_y', f32 * 2), ('pose_z', f32 * 2),
        ('pose_roll', f32 * 2), ('pose_pitch', f32 * 2), ('pose_yaw', f32 * 2),
        ('placement_margin', f32), ('min_movables', i32),
        ('table_height_range', f32 * 2), ('safe_drop_height', f32),
    ]


class B2SBuffers(C.Structure):
    _fields_ = [
        ('body_state', C.c_void_p), ('joint_state', C.c_void_p), ('action', C.c_void_p),
        ('obs_position', C.c_void_p), ('num_movables', C.c_void_p), ('body_mask', C.c_void_p),
        ('depth', C.c_void_p), ('segmask', C.c_void_p), ('point_cloud', C.c_void_p),
        ('reward', C.c_void_p), ('termination', C.c_void_p), ('is_safe', C.c_void_p),
        ('is_effective', C.c_void_p), ('episode_return', C.c_void_p),
    ]


class B2SRollout(C.Structure):
    _fields_ = [
        ('num_actions', i32), ('max_attempts', i32), ('num_episodes', i32), ('max_reset_retries', i32),
        ('seed', C.c_uint64), ('reset_seed', C.c_uint64),
        ('drop_lin_threshold', f32), ('drop_ang_threshold', f32), ('drop_max_steps', i32), ('policy_kind', i32), ('free_running', i32), ('reserved', i32),
        ('first_action', C.c_void_p), ('actions', C.c_void_p), ('rewards', C.c_void_p), ('positions', C.c_void_p),
        ('flags', C.c_void_p), ('substeps', C.c_void_p), ('lengths', C.c_void_p), ('returns', C.c_void_p),
    ]


# every symbol include/b2s.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    'b2s_version': (C.c_int, []),
    'b2s_last_error': (C.c_char_p, []),
    'b2s_default_params': (C.c_int, [P(B2SParams)]),
    'b2s_create': (C.c_int, [P(B2SParams), C.c_int, P(_vp)]),
    'b2s_destroy': (C.c_int, [_vp]),
    'b2s_load_scene': (C.c_int, [_vp, P(B2SSceneDesc)]),
    'b2s_bind_buffers': (C.c_int, [_vp, P(B2SBuffers)]),
    'b2s_get_params': (C.c_int, [_vp, P(B2SParams)]),
    'b2s_reset': (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    'b2s_settle': (C.c_int, [_vp, C.c_float, C.c_float, C.c_int, _vp]),
    'b2s_settle_masked': (C.c_int, [_vp, _vp, C.c_float, C.c_float, C.c_int, _vp]),
    'b2s_begin_episode': (C.c_int, [_vp, _vp, _vp]),
    'b2s_step': (C.c_int, [_vp, C.c_int, _vp]),
    'b2s_step_staged': (C.c_int, [_vp, C.c_int, _vp]),
    'b2s_set_action': (C.c_int, [_vp, _vp]),
    'b2s_env_substeps': (C.c_int, [_vp, C.c_int, P(C.c_int), _vp]),
    'b2s_env_substeps_free': (C.c_int, [_vp, C.c_int, P(C.c_int), _vp]),
    'b2s_env_step': (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    'b2s_rollout_begin': (C.c_int, [_vp, P(B2SRollout), _vp]),
    'b2s_rollout_run': (C.c_int, [_vp, C.c_int, C.c_int, P(C.c_int), _vp]),
    'b2s_env_async_step': (C.c_int, [_vp, _vp, C.c_int, C.c_uint64, _vp, _vp]),
    'b2s_env_async_step_free': (C.c_int, [_vp, _vp, C.c_int, C.c_uint64, _vp, _vp]),
    'b2s_arm_move_to_gripper_pose': (C.c_int, [_vp, _vp, _vp, _vp]),
    'b2s_arm_move_to_joint_positions': (C.c_int, [_vp, _vp, _vp, _vp]),
    'b2s_arm_reset_targets': (C.c_int, [_vp, _vp, _vp]),
    'b2s_arm_is_ready': (C.c_int, [_vp, _vp, _vp]),
    'b2s_set_motor_targets': (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'b2s_rebuild_colliders': (C.c_int, [_vp, _vp]),
    'b2s_inverse_kinematics': (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'b2s_forward_kinematics': (C.c_int, [_vp, _vp]),
    'b2s_query_contacts': (C.c_int, [_vp, _vp, _vp, _vp]),
    'b2s_observe': (C.c_int, [_vp, _vp]),
    'b2s_set_camera': (C.c_int, [_vp, P(f32), P(f32), P(f32), C.c_int]),
    'b2s_render': (C.c_int, [_vp, _vp]),
    'b2s_point_cloud': (C.c_int, [_vp, C.c_uint64, _vp]),
    'b2s_reward': (C.c_int, [_vp, _vp, _vp, _vp]),
    'b2s_allgather_returns': (C.c_int, [_vp, _vp, _vp, _vp]),
    'b2s_array': (C.c_int, [_vp, C.c_int, P(_vp), P(C.c_int64)]),
    'b2s_launch_count': (C.c_int64, [_vp]),
    'b2s_substeps_executed': (C.c_int64, [_vp, _vp]),
    'b2s_sizeof': (C.c_int, [C.c_int]),
    'b2s_se3_quat_from_euler': (C.c_int, [_vp, _vp, C.c_int, _vp]),
    'b2s_se3_euler_from_quat': (C.c_int, [_vp, _vp, C.c_int, _vp]),
    'b2s_se3_matrix_from_quat': (C.c_int, [_vp, _vp, C.c_int, _vp]),
    'b2s_se3_quat_multiply': (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    'b2s_se3_pose_inverse': (C.c_int, [_vp, _vp, C.c_int, _vp]),
    'b2s_se3_pose_transform': (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libb2s.so')
_lib = None


class B2SError(RuntimeError):
    """Raised for a negative return code of the C-ABI (message from b2s_last_error)."""

    def __init__(self, code, message):
        super(B2SError, self).__init__('b2s error %d: %s' % (code, message))
        self.code = code


def load(path=None):
    """Load libb2s.so and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get('B2S_LIB') or LIB_PATH      # B2S_LIB: try another build of the same library
    if not os.path.exists(path):
        raise OSError('%s not found: build it with `python -c "import __graft_entry__ as g; '
                      'g.build()"` (there is no CPU fallback)' % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate((B2SParams, B2SSceneDesc, B2SBuffers, B2SRollout)):
        if lib.b2s_sizeof(which) != C.sizeof(struct):
            raise RuntimeError('ctypes layout of %s (%d bytes) differs from include/b2s.h (%d)' % (
                struct.__name__, C.sizeof(struct), lib.b2s_sizeof(which)))
    _lib = lib
    return lib


def check(lib, code):
    if code != 0:
        raise B2SError(code, (lib.b2s_last_error() or b'').decode('utf-8', 'replace'))
    return code
