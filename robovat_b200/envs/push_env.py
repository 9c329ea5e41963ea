"""PushEnv over the batched CUDA world.

Keeps the call surface of the reference environment
(robovat/envs/push/push_env.py:32-937 on top of robovat/envs/robot_env.py:40-324):
`reset() -> OrderedDict`, `step(action) -> (obs, reward, done, None)`, `observations`,
`reward_fns`, `action_space`, `simulator`, `movable_bodies`, `robot`, `table`,
`attributes`, the statistics counters and the step-after-done ValueError.  With
`num_envs == 1` every observation has the reference's unbatched shape, so
`HeuristicPushPolicy` (robovat/policies/push_policy.py:33-52) runs unchanged; with
`num_envs > 1` every array gains a leading batch axis and `done` is a bool array.

What runs where: the per-substep loop of `_execute_action` (phase machine, arm
control, physics, settle) is `b2s_env_substeps`; `get_observation` is `b2s_observe`
(+ `b2s_render`/`b2s_point_cloud` when the point-cloud observation is on);
`get_reward` is `b2s_reward`.  Python only moves actions in and results out.
"""
import collections

import numpy as np
import torch

from robovat_b200 import _capi, config as config_lib
from robovat_b200.simulation.simulator import Simulator


class Box(object):
    """Stand-in for gym.spaces.Box (gym is not a dependency here)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = dtype

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape[-len(self.shape):] == self.shape and np.all(x >= self.low) and np.all(x <= self.high)


class _Obs(object):
    """One named observation (robovat/observations/observation.py:11-40)."""

    def __init__(self, name, fn):
        self.name = name
        self._fn = fn
        self.env = None

    def initialize(self, env):
        self.env = env

    def on_episode_start(self):
        pass

    def on_episode_end(self):
        pass

    def get_observation(self):
        return self._fn()


class PushReward(object):
    """robovat/reward_fns/push_reward.py:377-405 over the reward kernel."""

    def __init__(self, name, task_name, layout_id, is_planning=False):
        if is_planning:
            raise NotImplementedError('is_planning=True is a model-based-planning path, not the env path')
        self.name = name
        self.task_name = task_name
        self.layout_id = layout_id
        self.env = None

    def initialize(self, env):
        self.env = env

    def on_episode_start(self):
        pass

    def on_episode_end(self):
        pass

    def get_reward(self):
        env = self.env
        assert env.prev_obs_data is not None
        assert env.obs_data is not None
        reward, termination = env.world.reward()          # previous xy is kept on the device
        reward = reward.cpu().numpy()
        termination = termination.cpu().numpy().astype(bool)
        if env.num_envs == 1:
            return float(reward[0]), bool(termination[0])
        return reward, termination


class PushEnv(object):
    """Pushing task environment, B copies stepped together on the GPU."""

    def __init__(self, simulator=None, config=None, debug=False, num_envs=1, seed=0, device=0, env_id_offset=0):
        self._config = config or config_lib.default_push_env_config()
        self._debug = debug
        self.num_envs = int(num_envs)
        self.seed = int(seed)
        cfg = self._config
        self.task_name = cfg.TASK_NAME
        self.layout_id = cfg.LAYOUT_ID
        if self.task_name in (None, 'data_collection'):
            self.layouts, self.num_layouts = None, 1
        else:
            from robovat_b200 import layouts
            self.layouts = layouts.TASK_NAME_TO_LAYOUTS[self.task_name]
            self.num_layouts = len(self.layouts)
        self.num_goal_steps = cfg.NUM_GOAL_STEPS           # None: action [4]; G: action [G, 4], post -> pre G times
        self._action_floats = 4 * max(1, int(self.num_goal_steps or 0))
        self.cspace = Box(cfg.ACTION.CSPACE.LOW, cfg.ACTION.CSPACE.HIGH)
        start_low = np.array(cfg.ACTION.CSPACE.LOW, dtype=np.float32)
        start_high = np.array(cfg.ACTION.CSPACE.HIGH, dtype=np.float32)
        self.start_offset = 0.5 * (start_high + start_low)
        self.start_range = 0.5 * (start_high - start_low)
        self.start_z = cfg.ARM.FINGER_TIP_OFFSET + self.start_offset[2]
        self.min_movable_bodies = cfg.MIN_MOVABLE_BODIES
        self.max_movable_bodies = cfg.MAX_MOVABLE_BODIES
        self.phase_list = ['initial', 'pre', 'start', 'motion', 'post', 'offstage', 'done']

        use_camera = bool(cfg.get('USE_POINT_CLOUD_OBS', False) or cfg.USE_VISUALIZATION_OBS)
        if simulator is None or simulator is True:
            simulator = Simulator(config=cfg, num_envs=self.num_envs, device=device, with_camera=use_camera,
                                  env_id_offset=env_id_offset)
        self._simulator = simulator
        self.world = simulator.world
        self.use_camera = use_camera
        if use_camera:
            self.camera = simulator.create_camera(cfg.KINECT2.DEPTH)
        else:
            self.camera = None
        tx, ty = cfg.SIM.TABLE.POSE[0][0], cfg.SIM.TABLE.POSE[0][1]
        self.table_workspace = Box([tx - 0.5 * cfg.TABLE.X_RANGE, ty - 0.5 * cfg.TABLE.Y_RANGE],
                                   [tx + 0.5 * cfg.TABLE.X_RANGE, ty + 0.5 * cfg.TABLE.Y_RANGE])
        B = self.num_envs
        self._num_episodes = np.zeros(B, np.int64)
        self._num_steps = np.zeros(B, np.int64)
        self._episode_reward = np.zeros(B, np.float64)
        self._total_reward = np.zeros(B, np.float64)
        self._done = np.ones(B, bool)
        self.num_total_steps = 0
        self.num_unsafe = self.num_ineffective = self.num_useful = self.num_successes = 0
        self.num_successes_by_step = [0] * int(cfg.MAX_STEPS + 1)
        self.attributes = None
        self._obs_data = self._prev_obs_data = None
        self.substep_chunk = int(cfg.get('SUBSTEP_CHUNK', 250))
        self.max_action_substeps = int(cfg.get('MAX_ACTION_SUBSTEPS', 60000))
        self._pinned_action = torch.zeros(B, self._action_floats, dtype=torch.float32).pin_memory()
        self._observations = self._create_observations()
        for obs in self._observations:
            obs.initialize(self)
        self._reward_fns = [PushReward('reward', self.task_name, self.layout_id)]
        for fn in self._reward_fns:
            fn.initialize(self)
        shape = [4] if self.num_goal_steps is None else [int(self.num_goal_steps), 4]      # push_env.py:253-262
        self._action_space = Box(-np.ones(shape, np.float32), np.ones(shape, np.float32))
        self._reset_count = 0
        self._calibration = None
        self._async = None

    # -- reference properties ---------------------------------------------------------------
    simulator = property(lambda self: self._simulator)
    config = property(lambda self: self._config)
    debug = property(lambda self: self._debug)
    is_simulation = property(lambda self: True)
    observations = property(lambda self: self._observations)
    reward_fns = property(lambda self: self._reward_fns)
    action_space = property(lambda self: self._action_space)
    obs_data = property(lambda self: self._obs_data)
    prev_obs_data = property(lambda self: self._prev_obs_data)
    robot = property(lambda self: self._simulator.robot)
    table = property(lambda self: self._simulator.table)
    movable_bodies = property(lambda self: self._simulator.movable_bodies)

    def _scalar(self, a):
        return a[0] if self.num_envs == 1 else a

    num_episodes = property(lambda self: self._scalar(self._num_episodes))
    num_steps = property(lambda self: self._scalar(self._num_steps))
    episode_reward = property(lambda self: self._scalar(self._episode_reward))
    total_reward = property(lambda self: self._scalar(self._total_reward))
    done = property(lambda self: self._scalar(self._done))

    # -- observations (push_env.py:163-236) ---------------------------------------------------
    def _batched(self, a):
        a = np.asarray(a)
        return a[0] if self.num_envs == 1 else a

    def _create_observations(self):
        w = self.world
        obs = [
            _Obs('num_episodes', lambda: self._batched(self.attributes['num_episodes'])),
            _Obs('num_steps', lambda: self._batched(self.attributes['num_steps'])),
            _Obs('layout_id', lambda: np.array(self.layout_id, dtype=np.int64) if self.num_envs == 1
                 else np.full(self.num_envs, self.layout_id, np.int64)),
            _Obs('body_mask', lambda: self._batched(self.attributes['movable_body_mask'].astype(np.float32))),
        ]
        if self.use_camera:
            obs.append(_Obs('point_cloud', lambda: self._batched(self._point_cloud())))
        if self._config.USE_PRESTIGE_OBS:
            obs += [
                _Obs('position', lambda: self._batched(w.observe().cpu().numpy())),
                _Obs('is_safe', lambda: self._batched(self.attributes['is_safe'].astype(np.int64))),
                _Obs('is_effective', lambda: self._batched(self.attributes['is_effective'].astype(np.int64))),
            ]
        return obs

    def _point_cloud(self):
        self.world.render()
        self._pc_seed = getattr(self, '_pc_seed', 0) + 1
        return self.world.point_cloud(seed=self.seed * 7919 + self._pc_seed).cpu().numpy()

    def _refresh_attributes(self, counters=True):
        """`attributes` as PushEnv._execute_action keeps them (push_env.py:637-644, 715, 727): the counters are
        snapshotted when the action STARTS (before `_num_steps += 1`, robot_env.py:245-246), the flags when it ends."""
        w = self.world
        old = self.attributes or {}
        self.attributes = {
            'num_episodes': self._num_episodes.copy() if counters else old['num_episodes'],
            'num_steps': self._num_steps.copy() if counters else old['num_steps'],
            'layout_id': self.layout_id,
            'movable_body_mask': w.body_mask.cpu().numpy(),
            'is_safe': w.is_safe.cpu().numpy().astype(bool),
            'is_effective': w.is_effective.cpu().numpy().astype(bool),
        }

    def _reset_camera(self, mask=None):
        """ArmEnv._reset_camera (arm_env.py:109-152) for the envs being reset: calibration = configured values plus
        uniform noise in [-NOISE, NOISE], drawn in the reference's order (intrinsics, translation, rotation).  One env
        draws from the global numpy generator exactly like the reference; a batch draws per env from a generator
        seeded by (seed, reset count, global env id)."""
        from robovat_b200.simulation.camera import draw_calibration
        cam_cfg = self._config.KINECT2.DEPTH
        B = self.num_envs
        noises = (cam_cfg.get('INTRINSICS_NOISE'), cam_cfg.get('TRANSLATION_NOISE'), cam_cfg.get('ROTATION_NOISE'))
        if all(n is None for n in noises):
            return
        K0 = np.array(cam_cfg.INTRINSICS, np.float32).reshape(3, 3)
        if noises[0] is not None:
            noises = (np.array(noises[0], np.float64).reshape(3, 3),) + noises[1:]
        if self._calibration is None:
            self._calibration = [np.tile(K0, (B, 1, 1)), np.tile(np.array(cam_cfg.TRANSLATION, np.float32), (B, 1)),
                                 np.tile(np.array(cam_cfg.ROTATION, np.float32).reshape(1, -1), (B, 1))]
        m = np.ones(B, bool) if mask is None else np.asarray(mask, bool)
        offset = int(self.world.params.env_id_offset)
        for e in np.nonzero(m)[0]:
            rs = np.random if B == 1 else np.random.RandomState((self.seed * 1000003 + self._reset_count * 7919 + offset + int(e)) % (1 << 32))
            vals = draw_calibration(K0, cam_cfg.TRANSLATION, cam_cfg.ROTATION, noises[0], noises[1], noises[2], rs=rs)
            for dst, v in zip(self._calibration, vals):
                dst[e] = v
        K, t, r = self._calibration
        self.camera.set_calibration(K if B > 1 else K[0], t if B > 1 else t[0], r if B > 1 else r[0])

    # -- gym API (robot_env.py:202-310) -------------------------------------------------------
    def reset(self, mask=None):
        """Reset every env (or the envs in `mask`): new scene, movables dropped and settled."""
        B = self.num_envs
        m = np.ones(B, bool) if mask is None else np.asarray(mask, bool).reshape(B)
        self._num_steps[m] = 0
        self._episode_reward[m] = 0.0
        self._done[m] = False
        if self._config.MAX_STEPS is not None and self._config.MAX_STEPS == 0:
            self._done[m] = True
        self._simulator.reset_scene(seed=self.seed, mask=None if mask is None else m)
        self._reset_count += 1
        if self.camera is not None:
            self._reset_camera(None if mask is None else m)
        self._async = None
        self._refresh_attributes()
        self._obs_data = self._prev_obs_data = None
        return self.get_observation(force=True)

    def get_observation(self, force=False):
        if force or self._obs_data is None:
            data = collections.OrderedDict()
            for obs in self._observations:
                data[obs.name] = obs.get_observation()
            self._prev_obs_data, self._obs_data = self._obs_data, data
        return self._obs_data

    def step(self, action):
        if np.all(self._done):
            raise ValueError('The environment is done. Forget to reset?')
        active = ~self._done
        self._refresh_attributes()                     # counters as of the start of the action
        self._execute_action(action)
        self._num_steps[active] += 1
        self._refresh_attributes(counters=False)       # flags as of its end
        observation = self.get_observation(force=True)
        reward, termination = self._reward_fns[0].get_reward()
        reward = np.atleast_1d(np.asarray(reward, np.float64))
        termination = np.atleast_1d(np.asarray(termination, bool))
        reward = np.where(active, reward, 0.0)
        self._episode_reward += reward
        env_done = self.world.array(_capi.ARR_PHASE_STATE).view(self.num_envs, 8)[:, 4].cpu().numpy() != 0
        self._done = self._done | (active & (termination | env_done))
        if self._config.MAX_STEPS is not None:
            self._done |= self._num_steps >= self._config.MAX_STEPS
        finished = active & self._done
        if self.num_envs == 1 and finished[0] and reward[0] >= self._config.SUCCESS_THRESH:
            self.num_successes += 1
            self.num_successes_by_step[int(self._num_steps[0])] += 1
        self._num_episodes[finished] += 1
        self._total_reward[finished] += self._episode_reward[finished]
        if self.num_envs == 1:
            return observation, float(reward[0]), bool(self._done[0]), None
        return observation, reward.astype(np.float32), self._done.copy(), None

    def _execute_action(self, action):
        """push_env.py:631-733: the whole loop runs on the device."""
        a = np.asarray(action, dtype=np.float32).reshape(-1, self._action_floats)
        if a.shape[0] != self.num_envs:
            raise ValueError('action must have shape [%d%s, 4] (or without the batch axis for one env)' % (
                self.num_envs, '' if self.num_goal_steps is None else ', %d' % self.num_goal_steps))
        self._pinned_action.copy_(torch.from_numpy(a))
        self.world.action.copy_(self._pinned_action, non_blocking=True)
        self.world.set_action()
        if np.any(self._done):                     # finished envs do not execute (reference raises instead)
            ph = self.world.array(_capi.ARR_PHASE)
            ph[torch.from_numpy(self._done).to(ph.device)] = _capi.PHASE_IDLE
        done_substeps = 0
        while done_substeps < self.max_action_substeps:
            unfinished = self.world.env_substeps(self.substep_chunk, free_running=True)   # order of stepping is immaterial here
            done_substeps += self.substep_chunk
            if unfinished == 0:
                break
        safe = self.world.is_safe.cpu().numpy().astype(bool)
        eff = self.world.is_effective.cpu().numpy().astype(bool)
        active = ~self._done
        self.num_total_steps += int(active.sum())
        self.num_unsafe += int((~safe & active).sum())
        self.num_ineffective += int((~eff & active).sum())
        self.num_useful += int((safe & eff & active).sum())

    # -- asynchronous stepping: nobody waits for the slowest env -------------------------------------
    def step_async(self, action, substeps=None, free_running=True):
        """One slice of asynchronous stepping with the policy on the host.

        The reference collects data with one process per env (tools/parallel_run.py), each alternating
        `policy.action(obs)` / `env.step(action)` / `env.reset()` at its own pace.  Here every call
          * starts `action[e]` in every env that is ready (idle) and whose episode is not over; the rows of busy
            envs are ignored,
          * re-samples the scene of every ready env whose episode is over (auto-reset: drop + settle on the device),
          * advances every busy env by up to `substeps` substeps (default SUBSTEP_CHUNK); with `free_running` the slice
            is `substeps` x (busy envs) substeps in total and cheap envs get further than expensive ones
            (b2s_env_async_step_free),
        and returns `(obs, reward, done, info)`: `info['finished']` marks the envs whose action completed in this
        call (reward / done / is_safe / is_effective rows are those of that transition, as `step` would return
        them), `info['reset']` those whose reset completed (the row of `obs` is the episode's first observation),
        `info['ready']` those that will take an action (or, if done, be reset) in the next call.  Rows of other envs
        keep their last values.  Call `reset()` once before the first slice."""
        B = self.num_envs
        if self._async is None:
            dev = self.world.device
            self._async = {
                'cmd_host': torch.zeros(B, dtype=torch.uint8).pin_memory(), 'cmd': torch.zeros(B, dtype=torch.uint8, device=dev),
                'status': torch.zeros(B, dtype=torch.uint8, device=dev), 'status_host': torch.zeros(B, dtype=torch.uint8).pin_memory(),
                'reward_host': torch.zeros(B, dtype=torch.float32).pin_memory(), 'term_host': torch.zeros(B, dtype=torch.uint8).pin_memory(),
                'safe_host': torch.zeros(B, dtype=torch.uint8).pin_memory(), 'eff_host': torch.zeros(B, dtype=torch.uint8).pin_memory(),
                'pos_host': torch.zeros(B, self.world.N, 3, dtype=torch.float32).pin_memory(),
                'mask_host': torch.zeros(B, self.world.N, dtype=torch.uint8).pin_memory(),
                'ready': np.ones(B, bool), 'calls': 0,
            }
            self._async['pos_host'].copy_(self.world.obs_position)
            self._async['mask_host'].copy_(self.world.body_mask)
        st = self._async
        w = self.world
        a = np.asarray(action, dtype=np.float32).reshape(-1, self._action_floats)
        if a.shape[0] != B:
            raise ValueError('action must have shape [%d, %d]' % (B, self._action_floats))
        ready = st['ready']
        start = ready & ~self._done
        cmd = np.where(start, 1, np.where(ready & self._done, 2, 0)).astype(np.uint8)
        st['cmd_host'].copy_(torch.from_numpy(cmd))
        self._pinned_action.copy_(torch.from_numpy(a))
        w.action.copy_(self._pinned_action, non_blocking=True)
        st['cmd'].copy_(st['cmd_host'], non_blocking=True)
        st['calls'] += 1
        # counters as of the start of the action, for the envs that start one (robot_env.py:245-246)
        started_steps = self._num_steps.copy()
        w.env_async_step(st['cmd'], int(substeps or self.substep_chunk), reset_seed=self.seed * 1000003 + 7919, status=st['status'],
                         free_running=free_running)
        st['status_host'].copy_(st['status'], non_blocking=True)
        st['reward_host'].copy_(w.reward_buf, non_blocking=True)
        st['term_host'].copy_(w.termination, non_blocking=True)
        st['safe_host'].copy_(w.is_safe, non_blocking=True)
        st['eff_host'].copy_(w.is_effective, non_blocking=True)
        st['pos_host'].copy_(w.obs_position, non_blocking=True)
        st['mask_host'].copy_(w.body_mask, non_blocking=True)
        torch.cuda.current_stream(w.device).synchronize()
        status = st['status_host'].numpy()
        finished = (status & 2) != 0
        was_reset = (status & 4) != 0
        reward = np.where(finished, st['reward_host'].numpy().astype(np.float64), 0.0)
        termination = st['term_host'].numpy().astype(bool)
        env_done = (status & 8) != 0
        self._num_steps[finished] += 1
        self._episode_reward += reward
        done_now = finished & (termination | env_done)
        if self._config.MAX_STEPS is not None:
            done_now |= finished & (self._num_steps >= self._config.MAX_STEPS)
        self._done |= done_now
        self._num_episodes[done_now] += 1
        self._total_reward[done_now] += self._episode_reward[done_now]
        self._num_steps[was_reset] = 0
        self._episode_reward[was_reset] = 0.0
        self._done[was_reset] = False
        safe = st['safe_host'].numpy().astype(bool)
        eff = st['eff_host'].numpy().astype(bool)
        self.num_total_steps += int(finished.sum())
        self.num_unsafe += int((~safe & finished).sum())
        self.num_ineffective += int((~eff & finished).sum())
        self.num_useful += int((safe & eff & finished).sum())
        st['ready'] = (status & 1) != 0
        obs = collections.OrderedDict()
        obs['num_episodes'] = self._num_episodes.copy()
        obs['num_steps'] = self._num_steps.copy()
        obs['layout_id'] = np.full(B, self.layout_id if self.layout_id is not None else 0, np.int64)
        obs['body_mask'] = st['mask_host'].numpy().astype(np.float32)
        obs['position'] = st['pos_host'].numpy().copy()
        obs['is_safe'] = safe.astype(np.int64)
        obs['is_effective'] = eff.astype(np.int64)
        info = {'finished': finished, 'reset': was_reset, 'ready': st['ready'].copy(), 'started_num_steps': started_steps}
        return obs, reward.astype(np.float32), self._done.copy(), info

    async_h2d_bytes = property(lambda self: self.num_envs * (4 * 4 + 1))
    async_d2h_bytes = property(lambda self: self.num_envs * (1 + 4 + 1 + 1 + 1 + self.world.N * 13))

    # -- helpers kept from the reference --------------------------------------------------------
    def _compute_waypoints(self, action):
        """push_env.py:752-786 on the host (for policies/visualisation); the device twin is k_set_action."""
        action = np.reshape(action, [2, 2])
        start, motion = action[0, :], action[1, :]
        x = start[0] * self.start_range[0] + self.start_offset[0]
        y = start[1] * self.start_range[1] + self.start_offset[1]
        z = self.start_z
        x2 = np.clip(x + motion[0] * self._config.ACTION.MOTION.TRANSLATION_X, self.cspace.low[0], self.cspace.high[0])
        y2 = np.clip(y + motion[1] * self._config.ACTION.MOTION.TRANSLATION_Y, self.cspace.low[1], self.cspace.high[1])
        return [np.array([[x, y, z], [np.pi, 0, 0]]), np.array([[x2, y2, z], [np.pi, 0, 0]])]

    def close(self):
        self._simulator.close()
