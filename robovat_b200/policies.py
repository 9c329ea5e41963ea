"""Host-side push policies.

The reference policy (`robovat.policies.HeuristicPushPolicy`, push_policy.py:12-52, over
`HeuristicPushSampler`, heuristic_push_sampler.py:16-168) is plain numpy on the host and
"drops in unchanged": with `num_envs == 1` our PushEnv returns observations of the
reference's shapes, so the reference class can be used as is.  This module provides

  * `HeuristicPushSampler` -- the same rejection sampler restated with the SAME order of
    `np.random` draws, so that for a given global numpy seed it returns exactly what the
    reference returns (checked against tests/golden/sampler.json);
  * `HeuristicPushPolicy` -- the policy over it, accepting batched observations (one
    sampler state per environment);
  * `BatchedHeuristicPolicy` -- a vectorised, distribution-equivalent variant for thousands
    of environments (one numpy pass per attempt round instead of a Python loop per env).
"""
import numpy as np

ANGLE_SEED = 42            # heuristic_push_sampler.py:13


class HeuristicPushSampler(object):
    def __init__(self, cspace_low, cspace_high, translation_x, translation_y, start_margin=0.05,
                 motion_margin=0.01, max_attemps=20000):
        lo, hi = np.array(cspace_low), np.array(cspace_high)
        self.cspace_low, self.cspace_high = lo, hi
        self.cspace_offset, self.cspace_range = 0.5 * (hi + lo), 0.5 * (hi - lo)
        self.translation_x, self.translation_y = translation_x, translation_y
        self.start_margin, self.motion_margin, self.max_attemps = start_margin, motion_margin, max_attemps
        self.last_end = None

    def get_waypoints(self, start, motion):
        """Start point and clipped end point(s) in table coordinates (:125-146)."""
        x = start[0] * self.cspace_range[0] + self.cspace_offset[0]
        y = start[1] * self.cspace_range[1] + self.cspace_offset[1]
        pts = [[x, y]]
        for m in np.reshape(motion, [-1, 2]):
            x = np.clip(x + m[0] * self.translation_x, self.cspace_low[0], self.cspace_high[0])
            y = np.clip(y + m[1] * self.translation_y, self.cspace_low[1], self.cspace_high[1])
            pts.append([x, y])
        return pts

    @staticmethod
    def _dist(position, p):
        d = position[..., :2] - np.asarray(p)
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])

    def is_waypoint_clear(self, waypoint1, waypoint2, position, margin):
        if waypoint2 is None:
            return bool(np.all(self._dist(position, waypoint1) > margin))
        return bool(np.all(np.logical_and(self._dist(position, waypoint1) >= margin,
                                          self._dist(position, waypoint2) >= margin)))

    def _sample(self, position, body_mask, num_episodes, num_steps):
        nb = int(np.sum(body_mask))
        target = int(num_episodes) % nb
        position = position[:nb]
        base_angle = (num_episodes * ANGLE_SEED) % (2 * np.pi)
        if num_steps == 0:
            self.last_end = None
        start = motion = None
        for _ in range(self.max_attemps):
            # draw order matters for parity: start (2), angle jitter (1), motion jitter (2)
            start = np.random.uniform(-1., 1., [2])
            angle = base_angle + np.random.uniform(-0.25 * np.pi, 0.25 * np.pi)
            motion = np.array([np.cos(angle), np.sin(angle)], dtype=np.float32)
            motion += np.random.uniform(-0.3, 0.3, [2])       # in place: stays float32 like the reference (:89-91)
            motion = np.clip(motion, -1.0, 1.0)
            wp = self.get_waypoints(start, motion)
            if not self.is_waypoint_clear(wp[0], None, position, self.start_margin):
                continue                        # start too close to a body
            if self.is_waypoint_clear(wp[0], wp[1], position[target:target + 1], self.motion_margin):
                continue                        # neither end point touches the target body
            self.last_end = (np.asarray(wp[1]) - self.cspace_offset[:2]) / self.cspace_range[:2]
            break
        return np.concatenate([np.array(start, dtype=np.float32), np.array(motion, dtype=np.float32)], axis=-1)

    def sample(self, position, body_mask, num_episodes, num_steps, num_samples=1):
        return np.stack([self._sample(position, body_mask, num_episodes, num_steps) for _ in range(num_samples)], axis=0)


class HeuristicPushPolicy(object):
    """`action(observation)` as robovat.policies.Policy (policy.py:40-49); batched observations allowed."""

    def __init__(self, env, config=None):
        from robovat_b200 import config as config_lib
        self.env = env
        self.config = config or config_lib.default_policy_config()
        c = self.config
        self.num_envs = getattr(env, 'num_envs', 1)
        self._samplers = [HeuristicPushSampler(c.ACTION.CSPACE.LOW, c.ACTION.CSPACE.HIGH, c.ACTION.MOTION.TRANSLATION_X,
                                               c.ACTION.MOTION.TRANSLATION_Y, max_attemps=c.HEURISTICS.MAX_ATTEMPS)
                          for _ in range(self.num_envs)]

    def action(self, observation):
        pos = np.asarray(observation['position'])
        if pos.ndim == 2:                                   # the reference's unbatched layout: returns [1, 4]
            return self._samplers[0].sample(pos, observation['body_mask'], observation['num_episodes'],
                                            observation['num_steps'], num_samples=1)
        out = [self._samplers[e].sample(pos[e], observation['body_mask'][e], observation['num_episodes'][e],
                                        observation['num_steps'][e], num_samples=1)[0] for e in range(pos.shape[0])]
        return np.stack(out, axis=0)


class BatchedHeuristicPolicy(object):
    """Vectorised rejection sampling over all environments at once (same acceptance rule, own RNG)."""

    def __init__(self, config=None, seed=0, rounds=64):
        from robovat_b200 import config as config_lib
        c = config or config_lib.default_policy_config()
        lo, hi = np.array(c.ACTION.CSPACE.LOW[:2]), np.array(c.ACTION.CSPACE.HIGH[:2])
        self.lo, self.hi, self.off, self.rng = lo, hi, 0.5 * (hi + lo), 0.5 * (hi - lo)
        self.t = np.array([c.ACTION.MOTION.TRANSLATION_X, c.ACTION.MOTION.TRANSLATION_Y])
        self.rs = np.random.RandomState(seed)
        self.rounds = rounds

    def action(self, observation):
        pos = np.asarray(observation['position'], np.float64)[..., :2]         # [B, N, 2]
        mask = np.asarray(observation['body_mask']) > 0
        B, N, _ = pos.shape
        nb = np.maximum(mask.sum(axis=1), 1)
        target = np.asarray(observation['num_episodes']).astype(np.int64) % nb
        base = (np.asarray(observation['num_episodes']) * ANGLE_SEED) % (2 * np.pi)
        action = np.zeros((B, 4), np.float32)
        todo = np.ones(B, bool)
        tgt = pos[np.arange(B), target]
        for _ in range(self.rounds):
            idx = np.nonzero(todo)[0]
            if idx.size == 0:
                break
            start = self.rs.uniform(-1, 1, (idx.size, 2))
            ang = base[idx] + self.rs.uniform(-0.25 * np.pi, 0.25 * np.pi, idx.size)
            motion = np.clip(np.stack([np.cos(ang), np.sin(ang)], 1) + self.rs.uniform(-0.3, 0.3, (idx.size, 2)), -1, 1)
            p0 = start * self.rng + self.off
            p1 = np.clip(p0 + motion * self.t, self.lo, self.hi)
            d0 = np.linalg.norm(pos[idx] - p0[:, None, :], axis=-1)
            safe = np.all((d0 > 0.05) | ~mask[idx], axis=1)
            near = (np.linalg.norm(tgt[idx] - p0, axis=-1) < 0.01) | (np.linalg.norm(tgt[idx] - p1, axis=-1) < 0.01)
            action[idx] = np.concatenate([start, motion], axis=1)       # the reference also returns the last try
            todo[idx[safe & near]] = False
        return action
