"""ctypes wrapper of the CPU oracle (oracle/libb2o.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by robovat_b200.  Mirrors the b2s_* C-ABI
with host pointers; arrays come back as numpy views of the oracle's own storage.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from robovat_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libb2o.so')
LIB64_PATH = os.path.join(HERE, 'libb2o64.so')      # the same sources in double precision (b2o_f64.h)

BUF_IDS = {'body_state': 100, 'joint_state': 101, 'action': 102, 'obs_position': 103,
           'num_movables': 104, 'body_mask': 105, 'depth': 106, 'segmask': 107, 'point_cloud': 108,
           'reward': 109, 'termination': 110, 'is_safe': 111, 'is_effective': 112, 'episode_return': 113}
DTYPES = {
    _capi.ARR_MANIFOLD_KEYS: np.int32, _capi.ARR_MANIFOLD_NPTS: np.int32, _capi.ARR_MANIFOLD_PTS: np.float32,
    _capi.ARR_NUM_MANIFOLDS: np.int32, _capi.ARR_PAIR_KEYS: np.int32, _capi.ARR_NUM_PAIRS: np.int32,
    _capi.ARR_PHASE: np.int32, _capi.ARR_NUM_STEPS: np.int32, _capi.ARR_CTRL: np.float32,
    _capi.ARR_CTRL_FLAGS: np.int32, _capi.ARR_LINK_POSES: np.float32, _capi.ARR_MOV_PARAMS: np.float32,
    _capi.ARR_TABLE_DZ: np.float32, _capi.ARR_ERROR_FLAGS: np.int32, _capi.ARR_WAYPOINTS: np.float32,
    _capi.ARR_STATUS: np.float32, _capi.ARR_CONTACT_FLAGS: np.int32, _capi.ARR_PHASE_STATE: np.int32,
    _capi.ARR_SOLVER_STATS: np.int32, _capi.ARR_CTRL_TIME: np.float64, _capi.ARR_LINK_VEL: np.float32,
    _capi.ARR_NUM_COLLIDERS: np.int32, _capi.ARR_COL_SLOT: np.int32, _capi.ARR_COL_HULL: np.int32,
    _capi.ARR_NUM_EPISODES: np.int32, _capi.ARR_ROLLOUT_STATE: np.int32,
    100: np.float32, 101: np.float32, 102: np.float32, 103: np.float32, 104: np.int32, 105: np.uint8,
    106: np.float32, 107: np.uint8, 108: np.float32, 109: np.float32, 110: np.uint8, 111: np.uint8,
    112: np.uint8, 113: np.float32,
}
_libs = {}


def build():
    subprocess.check_call(['make', '-s', '-C', HERE, 'all'])


def load(f64=False):
    if f64 not in _libs:
        path = LIB64_PATH if f64 else LIB_PATH
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.b2o_last_error.restype = C.c_char_p
        lib.b2o_substeps_executed.restype = C.c_int64
        _libs[f64] = lib
    return _libs[f64]


class OracleWorld(object):
    """f64=True: the double-precision build; every float array / argument of this wrapper is then float64."""

    def __init__(self, params, scene, threads=1, f64=False):
        self.lib = load(f64)
        self.f64 = bool(f64)
        self.real = np.float64 if f64 else np.float32
        self.creal = C.c_double if f64 else C.c_float
        self.params = params
        self.scene = scene
        self.h = C.c_void_p()
        self._chk(self.lib.b2o_create(C.byref(params), C.byref(self.h)))
        self._chk(self.lib.b2o_load_scene(self.h, C.byref(scene.desc)))
        self.lib.b2o_set_threads(int(threads))
        self.B, self.N = params.num_envs, params.max_movables

    def _chk(self, code):
        if code != 0:
            raise RuntimeError('oracle error %d: %s' % (code, self.lib.b2o_last_error()))

    def close(self):
        if self.h:
            self.lib.b2o_destroy(self.h)
            self.h = C.c_void_p()

    def array(self, which):
        if isinstance(which, str):
            which = BUF_IDS[which]
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._chk(self.lib.b2o_array(self.h, which, C.byref(ptr), C.byref(nbytes)))
        dt = np.dtype(DTYPES[which])
        if dt == np.float32:
            dt = np.dtype(self.real)
        n = nbytes.value // dt.itemsize
        if n == 0:
            return np.zeros(0, dtype=dt)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    # views with the documented shapes
    @property
    def body_state(self):
        return self.array('body_state').reshape(13, self.B, self.N)

    @property
    def joint_state(self):
        return self.array('joint_state').reshape(2, 7, self.B)

    def set_threads(self, n):
        self.lib.b2o_set_threads(int(n))

    def reset(self, seed=0, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_reset(self.h, m, C.c_uint64(seed)))

    def settle(self, lin=0.005, ang=0.005, max_steps=2000, mask=None):
        if mask is None:
            self._chk(self.lib.b2o_settle(self.h, self.creal(lin), self.creal(ang), int(max_steps)))
        else:
            m = np.ascontiguousarray(mask, np.uint8)
            self._chk(self.lib.b2o_settle_masked(self.h, m.ctypes.data_as(C.c_void_p), self.creal(lin), self.creal(ang), int(max_steps)))

    def begin_episode(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_begin_episode(self.h, m))

    def step(self, n=1):
        self._chk(self.lib.b2o_step(self.h, int(n)))

    def set_action(self, action):
        self.array('action')[:] = np.asarray(action, np.float32).ravel()
        self._chk(self.lib.b2o_set_action(self.h))

    def env_substeps(self, n):
        u = C.c_int()
        self._chk(self.lib.b2o_env_substeps(self.h, int(n), C.byref(u)))
        return u.value

    # -- episodes without the host (b2s_rollout_*) ----------------------------------------------
    def rollout_begin(self, num_actions, num_episodes=1, policy_seed=0, reset_seed=0, max_attempts=20000,
                      first_action=None, positions=True, record=True, max_reset_retries=8, drop_thresholds=(0.1, 0.1, 500), policy_kind=0):
        """Returns the record: dict of numpy arrays with the layouts of B2SRollout."""
        B, N, EP, A = self.B, self.N, int(num_episodes), int(num_actions)
        if record:
            rec = {'actions': np.zeros((B, EP, A, 4), self.real), 'rewards': np.zeros((B, EP, A), self.real),
                   'positions': np.zeros((B, EP, A + 1, N, 3), self.real) if positions else None,
                   'flags': np.zeros((B, EP, A), np.uint8), 'substeps': np.zeros((B, EP, A), np.int32),
                   'lengths': np.zeros((B, EP), np.int32), 'returns': np.zeros((B, EP), self.real)}
        else:                                   # nothing is recorded (throughput runs over many episodes)
            rec = dict.fromkeys(('actions', 'rewards', 'positions', 'flags', 'substeps', 'lengths', 'returns'))
        r = _capi.B2SRollout()
        r.num_actions, r.num_episodes, r.max_attempts = A, EP, int(min(max_attempts, 65535))
        r.max_reset_retries = int(max_reset_retries)
        r.policy_kind = int(policy_kind)
        r.seed, r.reset_seed = int(policy_seed), int(reset_seed)
        r.drop_lin_threshold, r.drop_ang_threshold, r.drop_max_steps = float(drop_thresholds[0]), float(drop_thresholds[1]), int(drop_thresholds[2])
        for k in ('flags', 'substeps', 'lengths'):
            if rec[k] is not None:
                setattr(r, k, rec[k].ctypes.data)
        fa = None if first_action is None else np.ascontiguousarray(first_action, self.real).reshape(B, 4)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._rollout_keepalive = (rec, fa)
        self._chk(self.lib.b2o_rollout_begin(self.h, C.byref(r), vp(fa), vp(rec['actions']), vp(rec['rewards']),
                                             vp(rec['positions']), vp(rec['returns'])))
        return rec

    def rollout_run(self, n):
        u = C.c_int()
        self._chk(self.lib.b2o_rollout_run(self.h, int(n), C.byref(u)))
        return u.value

    def env_async_step(self, command, n, reset_seed=0):
        cmd = None if command is None else np.ascontiguousarray(command, np.uint8)
        status = np.zeros(self.B, np.uint8)
        self._chk(self.lib.b2o_env_async_step(self.h, None if cmd is None else cmd.ctypes.data_as(C.c_void_p), int(n),
                                              C.c_uint64(int(reset_seed)), status.ctypes.data_as(C.c_void_p)))
        return status

    def policy_sample(self, seed, action_index, num_episodes, max_attempts=20000):
        out = np.zeros((self.B, 4), self.real)
        ne = np.ascontiguousarray(np.broadcast_to(np.asarray(num_episodes, np.int32), (self.B,)))
        self._chk(self.lib.b2o_policy_sample(self.h, C.c_uint64(int(seed)), int(action_index), ne.ctypes.data_as(C.c_void_p),
                                             int(min(max_attempts, 65535)), out.ctypes.data_as(C.c_void_p)))
        return out

    def move_to_gripper_pose(self, pose, mask=None):
        p = np.ascontiguousarray(pose, self.real)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_arm_move_to_gripper_pose(self.h, p.ctypes.data_as(C.c_void_p), m))

    def move_to_joint_positions(self, q, mask=None):
        p = np.ascontiguousarray(q, self.real)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_arm_move_to_joint_positions(self.h, p.ctypes.data_as(C.c_void_p), m))

    def set_motor_targets(self, q, qd=None, mask=None):
        a = np.ascontiguousarray(q, self.real)
        b = None if qd is None else np.ascontiguousarray(qd, self.real)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_set_motor_targets(self.h, a.ctypes.data_as(C.c_void_p),
                                                 None if b is None else b.ctypes.data_as(C.c_void_p), m))

    def rebuild_colliders(self):
        self._chk(self.lib.b2o_rebuild_colliders(self.h))

    def arm_reset_targets(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.c_void_p)
        self._chk(self.lib.b2o_arm_reset_targets(self.h, m))

    # the caller-owned buffers of the CUDA world, under the same attribute names
    num_movables = property(lambda self: self.array('num_movables'))
    body_mask = property(lambda self: self.array('body_mask').reshape(self.B, self.N))

    def arm_is_ready(self):
        out = np.zeros(self.B, np.uint8)
        self._chk(self.lib.b2o_arm_is_ready(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def inverse_kinematics(self, pose, q_start):
        p = np.ascontiguousarray(pose, self.real)
        qs = np.ascontiguousarray(q_start, self.real)
        out = np.zeros((7, self.B), self.real)
        self._chk(self.lib.b2o_inverse_kinematics(self.h, p.ctypes.data_as(C.c_void_p), qs.ctypes.data_as(C.c_void_p),
                                                  out.ctypes.data_as(C.c_void_p)))
        return out

    def forward_kinematics(self):
        self._chk(self.lib.b2o_forward_kinematics(self.h))
        return self.array(_capi.ARR_LINK_POSES).reshape(self.B, -1, 7)

    def observe(self):
        self._chk(self.lib.b2o_observe(self.h))
        return self.array('obs_position').reshape(self.B, self.N, 3)

    def reward(self, prev_xy=None, next_xy=None):
        a = None if prev_xy is None else np.ascontiguousarray(prev_xy, self.real)
        b = None if next_xy is None else np.ascontiguousarray(next_xy, self.real)
        self._chk(self.lib.b2o_reward(self.h, None if a is None else a.ctypes.data_as(C.c_void_p),
                                      None if b is None else b.ctypes.data_as(C.c_void_p)))
        return self.array('reward').copy(), self.array('termination').copy()

    def set_camera(self, K, R, t, per_env=False):
        K, R, t = (np.ascontiguousarray(x, self.real) for x in (K, R, t))
        fp = C.POINTER(self.creal)
        self._chk(self.lib.b2o_set_camera(self.h, K.ctypes.data_as(fp), R.ctypes.data_as(fp), t.ctypes.data_as(fp),
                                          int(bool(per_env))))

    def render(self):
        self._chk(self.lib.b2o_render(self.h))
        H, W = self.params.cam_height, self.params.cam_width
        return self.array('depth').reshape(self.B, H, W), self.array('segmask').reshape(self.B, H, W)

    def point_cloud(self, seed=0):
        self._chk(self.lib.b2o_point_cloud(self.h, C.c_uint64(seed)))
        return self.array('point_cloud').reshape(self.B, self.N, self.params.num_points, 3)

    def substeps_executed(self):
        return self.lib.b2o_substeps_executed(self.h)
