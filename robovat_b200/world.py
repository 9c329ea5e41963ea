"""Batched physics world: Python handle over the C-ABI of libb2s.so.

This is the seam the reference fills with `BulletPhysics`
(robovat/simulation/simulator.py:45-49: `getattr(physics, physics_backend)`), for
B environments at once.  PyTorch is used only to own the device buffers (state,
actions, observations, rewards) that the library borrows by raw pointer, and to
provide the CUDA stream; all arithmetic happens in the hand-written kernels.
"""
import ctypes as C

import numpy as np
import torch

from robovat_b200 import _capi


class _CudaArray(object):
    """Minimal __cuda_array_interface__ holder so torch can view world-owned memory zero-copy."""

    def __init__(self, ptr, nbytes, typestr, owner):
        itemsize = np.dtype(typestr).itemsize
        self.__cuda_array_interface__ = {
            'shape': (nbytes // itemsize,), 'typestr': typestr, 'data': (ptr, False), 'version': 2}
        self._owner = owner


_ARR_TYPES = {
    _capi.ARR_MANIFOLD_KEYS: '<i4', _capi.ARR_MANIFOLD_NPTS: '<i4', _capi.ARR_MANIFOLD_PTS: '<f4',
    _capi.ARR_NUM_MANIFOLDS: '<i4', _capi.ARR_PAIR_KEYS: '<i4', _capi.ARR_NUM_PAIRS: '<i4',
    _capi.ARR_PHASE: '<i4', _capi.ARR_NUM_STEPS: '<i4', _capi.ARR_CTRL: '<f4', _capi.ARR_CTRL_FLAGS: '<i4',
    _capi.ARR_LINK_POSES: '<f4', _capi.ARR_MOV_PARAMS: '<f4', _capi.ARR_TABLE_DZ: '<f4',
    _capi.ARR_ERROR_FLAGS: '<i4', _capi.ARR_WAYPOINTS: '<f4', _capi.ARR_STATUS: '<f4',
    _capi.ARR_CONTACT_FLAGS: '<i4', _capi.ARR_PHASE_STATE: '<i4', _capi.ARR_SOLVER_STATS: '<i4',
    _capi.ARR_CTRL_TIME: '<f8', _capi.ARR_LINK_VEL: '<f4', _capi.ARR_NUM_COLLIDERS: '<i4',
    _capi.ARR_COL_SLOT: '<i4', _capi.ARR_COL_HULL: '<i4', _capi.ARR_PROF: '<i8',
    _capi.ARR_NUM_EPISODES: '<i4', _capi.ARR_ROLLOUT_STATE: '<i4', _capi.ARR_RAY_SCENE: '|u1',
}


class RolloutRecord(object):
    """Device-resident records of a rollout (layouts of B2SRollout, include/b2s.h)."""

    def __init__(self, B, N, num_episodes, num_actions, device, positions=True):
        EP, A = int(num_episodes), int(num_actions)
        f32 = torch.float32
        self.actions = torch.zeros(B, EP, A, 4, dtype=f32, device=device)
        self.rewards = torch.zeros(B, EP, A, dtype=f32, device=device)
        self.positions = torch.zeros(B, EP, A + 1, N, 3, dtype=f32, device=device) if positions else None
        self.flags = torch.zeros(B, EP, A, dtype=torch.uint8, device=device)
        self.substeps = torch.zeros(B, EP, A, dtype=torch.int32, device=device)
        self.lengths = torch.zeros(B, EP, dtype=torch.int32, device=device)
        self.returns = torch.zeros(B, EP, dtype=f32, device=device)

    def tensors(self):
        return {k: v for k, v in vars(self).items() if v is not None}

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors().values())


class World(object):
    """B independent PushEnv scenes on one GPU."""

    def __init__(self, params, scene, device=0, with_camera=False):
        if not torch.cuda.is_available():
            raise RuntimeError('robovat_b200 needs a CUDA device: there is no CPU fallback')
        self.lib = _capi.load()
        self.params = params
        self.scene = scene
        self.device = torch.device('cuda', device)
        self.B, self.N = params.num_envs, params.max_movables
        self.G = max(1, int(params.num_goal_steps))        # goal steps per action (NUM_GOAL_STEPS)
        self.h = C.c_void_p()
        self._chk(self.lib.b2s_create(C.byref(params), int(device), C.byref(self.h)))
        self._chk(self.lib.b2s_load_scene(self.h, C.byref(scene.desc)))
        B, N, dev = self.B, self.N, self.device
        f32, u8, i32 = torch.float32, torch.uint8, torch.int32
        self.body_state = torch.zeros(13, B, N, dtype=f32, device=dev)
        self.joint_state = torch.zeros(2, 7, B, dtype=f32, device=dev)
        self.action = torch.zeros(B, self.G * 4, dtype=f32, device=dev)      # [B][G][4]
        self.obs_position = torch.zeros(B, N, 3, dtype=f32, device=dev)
        self.num_movables = torch.zeros(B, dtype=i32, device=dev)
        self.body_mask = torch.zeros(B, N, dtype=u8, device=dev)
        self.reward_buf = torch.zeros(B, dtype=f32, device=dev)
        self.termination = torch.zeros(B, dtype=u8, device=dev)
        self.is_safe = torch.ones(B, dtype=u8, device=dev)
        self.is_effective = torch.ones(B, dtype=u8, device=dev)
        self.episode_return = torch.zeros(B, dtype=f32, device=dev)
        self.depth = self.segmask = self.point_cloud_buf = None
        if with_camera:
            H, W = params.cam_height, params.cam_width
            self.depth = torch.zeros(B, H, W, dtype=f32, device=dev)
            self.segmask = torch.full((B, H, W), 255, dtype=u8, device=dev)
            self.point_cloud_buf = torch.zeros(B, N, params.num_points, 3, dtype=f32, device=dev)
        bufs = _capi.B2SBuffers()
        for name, t in (('body_state', self.body_state), ('joint_state', self.joint_state),
                        ('action', self.action), ('obs_position', self.obs_position),
                        ('num_movables', self.num_movables), ('body_mask', self.body_mask),
                        ('depth', self.depth), ('segmask', self.segmask), ('point_cloud', self.point_cloud_buf),
                        ('reward', self.reward_buf), ('termination', self.termination),
                        ('is_safe', self.is_safe), ('is_effective', self.is_effective),
                        ('episode_return', self.episode_return)):
            setattr(bufs, name, None if t is None else t.data_ptr())
        self._bufs = bufs
        self._chk(self.lib.b2s_bind_buffers(self.h, C.byref(bufs)))

    # -- plumbing ---------------------------------------------------------------------------
    def _chk(self, code):
        _capi.check(self.lib, code)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if self.h:
            torch.cuda.synchronize(self.device)
            self.lib.b2s_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, which):
        """Zero-copy torch view of a world-owned device array (B2S_ARR_*)."""
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._chk(self.lib.b2s_array(self.h, int(which), C.byref(ptr), C.byref(nbytes)))
        return torch.as_tensor(_CudaArray(ptr.value, nbytes.value, _ARR_TYPES[which], self), device=self.device)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _mask(self, mask):
        if mask is None:
            return None
        return torch.as_tensor(mask, dtype=torch.uint8, device=self.device).contiguous()

    # -- lifecycle --------------------------------------------------------------------------
    def reset(self, seed=0, mask=None):
        m = self._mask(mask)
        self._chk(self.lib.b2s_reset(self.h, self._ptr(m), C.c_uint64(int(seed)), self._stream()))

    def settle(self, lin=0.005, ang=0.005, max_steps=2000, mask=None):
        m = self._mask(mask)
        self._chk(self.lib.b2s_settle_masked(self.h, self._ptr(m), C.c_float(lin), C.c_float(ang), int(max_steps), self._stream()))

    def begin_episode(self, mask=None):
        """End of RobotEnv.reset: the settled movable xy becomes the reward's previous state."""
        m = self._mask(mask)
        self._chk(self.lib.b2s_begin_episode(self.h, self._ptr(m), self._stream()))

    def step(self, n=1):
        self._chk(self.lib.b2s_step(self.h, int(n), self._stream()))

    def step_staged(self, n=1):
        self._chk(self.lib.b2s_step_staged(self.h, int(n), self._stream()))

    def set_action(self, action=None):
        if action is not None:
            self.action.copy_(torch.as_tensor(action, dtype=torch.float32).reshape(self.B, self.G * 4), non_blocking=True)
        self._chk(self.lib.b2s_set_action(self.h, self._stream()))

    def env_substeps(self, n, sync=True, free_running=False):
        u = C.c_int(-1)
        fn = self.lib.b2s_env_substeps_free if free_running else self.lib.b2s_env_substeps
        self._chk(fn(self.h, int(n), C.byref(u) if sync else None, self._stream()))
        return u.value

    def env_step(self, chunk=200, max_substeps=40000):
        self._chk(self.lib.b2s_env_step(self.h, int(chunk), int(max_substeps), self._stream()))

    # -- episodes on the device -----------------------------------------------------------------
    def rollout_begin(self, num_actions, num_episodes=1, policy_seed=0, reset_seed=0, max_attempts=20000,
                      first_action=None, record=None, max_reset_retries=8, drop_thresholds=(0.1, 0.1, 500), policy_kind=0, free_running=False):
        """Start episode 0 of a device-side rollout in every env (from its reset, settled state).  `record` is a
        RolloutRecord (or None: nothing is recorded).  Returns the record."""
        r = _capi.B2SRollout()
        r.num_actions, r.num_episodes, r.max_attempts = int(num_actions), int(num_episodes), int(min(max_attempts, 65535))
        r.max_reset_retries = int(max_reset_retries)
        r.policy_kind = int(policy_kind)
        r.free_running = int(bool(free_running))
        r.seed, r.reset_seed = int(policy_seed), int(reset_seed)
        r.drop_lin_threshold, r.drop_ang_threshold, r.drop_max_steps = float(drop_thresholds[0]), float(drop_thresholds[1]), int(drop_thresholds[2])
        fa = None
        if first_action is not None:
            fa = torch.as_tensor(first_action, dtype=torch.float32, device=self.device).reshape(self.B, 4).contiguous()
            r.first_action = fa.data_ptr()
        if record is not None:
            for name, t in record.tensors().items():
                setattr(r, name, t.data_ptr())
        self._rollout_keepalive = (record, fa)
        self._chk(self.lib.b2s_rollout_begin(self.h, C.byref(r), self._stream()))
        return record

    def rollout_run(self, chunk=250, max_substeps=1 << 30):
        """Advance the rollout until every env finished its last episode or `max_substeps` per env were launched.
        Returns the number of envs that are still running."""
        u = C.c_int(-1)
        self._chk(self.lib.b2s_rollout_run(self.h, int(chunk), int(max_substeps), C.byref(u), self._stream()))
        return u.value

    def env_async_step(self, command, n, reset_seed=0, status=None, free_running=False):
        """b2s_env_async_step(_free): `command` uint8 [B] device tensor (or None), `status` uint8 [B] device tensor out."""
        fn = self.lib.b2s_env_async_step_free if free_running else self.lib.b2s_env_async_step
        self._chk(fn(self.h, self._ptr(command), int(n), C.c_uint64(int(reset_seed)),
                                              self._ptr(status), self._stream()))
        return status

    # -- robot ------------------------------------------------------------------------------
    def move_to_gripper_pose(self, pose, mask=None):
        p = torch.as_tensor(pose, dtype=torch.float32, device=self.device).reshape(self.B, 7).contiguous()
        m = self._mask(mask)
        self._chk(self.lib.b2s_arm_move_to_gripper_pose(self.h, self._ptr(p), self._ptr(m), self._stream()))

    def move_to_joint_positions(self, q, mask=None):
        p = torch.as_tensor(q, dtype=torch.float32, device=self.device).reshape(self.B, 7).contiguous()
        m = self._mask(mask)
        self._chk(self.lib.b2s_arm_move_to_joint_positions(self.h, self._ptr(p), self._ptr(m), self._stream()))

    def arm_reset_targets(self, mask=None):
        m = self._mask(mask)
        self._chk(self.lib.b2s_arm_reset_targets(self.h, self._ptr(m), self._stream()))

    def set_motor_targets(self, q, qd=None, mask=None):
        a = torch.as_tensor(q, dtype=torch.float32, device=self.device).reshape(self.B, 7).contiguous()
        b = None if qd is None else torch.as_tensor(qd, dtype=torch.float32, device=self.device).reshape(self.B, 7).contiguous()
        m = self._mask(mask)
        self._chk(self.lib.b2s_set_motor_targets(self.h, self._ptr(a), self._ptr(b), self._ptr(m), self._stream()))

    def rebuild_colliders(self):
        self._chk(self.lib.b2s_rebuild_colliders(self.h, self._stream()))

    def arm_is_ready(self):
        out = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        self._chk(self.lib.b2s_arm_is_ready(self.h, self._ptr(out), self._stream()))
        return out

    def inverse_kinematics(self, pose, q_start):
        p = torch.as_tensor(pose, dtype=torch.float32, device=self.device).reshape(self.B, 7).contiguous()
        qs = torch.as_tensor(q_start, dtype=torch.float32, device=self.device).reshape(7, self.B).contiguous()
        out = torch.zeros(7, self.B, dtype=torch.float32, device=self.device)
        self._chk(self.lib.b2s_inverse_kinematics(self.h, self._ptr(p), self._ptr(qs), self._ptr(out), self._stream()))
        return out

    def forward_kinematics(self):
        self._chk(self.lib.b2s_forward_kinematics(self.h, self._stream()))
        return self.array(_capi.ARR_LINK_POSES).view(self.B, -1, 7)

    def query_contacts(self):
        a = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        b = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        self._chk(self.lib.b2s_query_contacts(self.h, self._ptr(a), self._ptr(b), self._stream()))
        return a, b

    # -- observations / reward ----------------------------------------------------------------
    def observe(self):
        self._chk(self.lib.b2s_observe(self.h, self._stream()))
        return self.obs_position

    def reward(self, prev_xy=None, next_xy=None):
        a = None if prev_xy is None else torch.as_tensor(prev_xy, dtype=torch.float32, device=self.device).contiguous()
        b = None if next_xy is None else torch.as_tensor(next_xy, dtype=torch.float32, device=self.device).contiguous()
        self._chk(self.lib.b2s_reward(self.h, self._ptr(a), self._ptr(b), self._stream()))
        return self.reward_buf, self.termination

    def set_camera(self, K, R, t, per_env=False):
        K, R, t = (np.ascontiguousarray(x, np.float32) for x in (K, R, t))
        fp = C.POINTER(C.c_float)
        self._chk(self.lib.b2s_set_camera(self.h, K.ctypes.data_as(fp), R.ctypes.data_as(fp), t.ctypes.data_as(fp),
                                          int(bool(per_env))))

    def render(self):
        self._chk(self.lib.b2s_render(self.h, self._stream()))
        return self.depth, self.segmask

    def point_cloud(self, seed=0):
        self._chk(self.lib.b2s_point_cloud(self.h, C.c_uint64(int(seed)), self._stream()))
        return self.point_cloud_buf

    # -- counters -----------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.b2s_launch_count(self.h))

    def substeps_executed(self):
        return int(self.lib.b2s_substeps_executed(self.h, self._stream()))
