"""Episode driver and episode files against the reference's own driver run on a scripted
stand-in env (tests/golden/episodes.json, written by oracle/gen_golden.py)."""
import json
import os

import numpy as np

from robovat_b200 import episodes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'episodes.json')


class ScriptedEnv(object):
    def __init__(self, script):
        self.script, self.t, self.resets = script, 0, 0

    def reset(self):
        self.t = 0
        self.resets += 1
        return {'position': [float(self.resets), 0.0]}

    def step(self, action):
        reward, done = self.script[self.t]
        self.t += 1
        return {'position': [float(self.resets), float(self.t)]}, reward + float(action[0]), done, None


class ScriptedPolicy(object):
    def action(self, observation):
        return [observation['position'][1] * 0.5, 1.0]


def test_batched_driver_with_one_env_matches_reference_driver():
    """The reference's generate_episode(s) (golden, run from /root/reference on the scripted env) is the B = 1 case of
    the batched driver: same transitions, same episode lengths over two consecutive episodes."""
    with open(GOLDEN) as f:
        cases = json.load(f)
    assert len(cases) == 3
    for case in cases:
        script = [tuple(x) for x in case['script']]
        env = ScriptedEnv(script)
        batch = episodes.collect(env, ScriptedPolicy(), num_steps=case['num_steps'])
        ep = batch.episodes()[0]
        assert sorted(ep.keys()) == case['keys']
        got = [{'state': {k: np.asarray(v).tolist() for k, v in t['state'].items()}, 'action': np.asarray(t['action'], np.float64).tolist(),
                'reward': t['reward'], 'info': t['info']} for t in ep['transitions']]
        assert got == case['transitions']
        second = episodes.collect(env, ScriptedPolicy(), num_steps=case['num_steps'])
        assert [int(batch.lengths[0]), int(second.lengths[0])] == case['lengths']
        assert np.isclose(batch.returns[0], sum(t['reward'] for t in case['transitions']))


class BatchEnv(object):
    num_envs = 3

    def reset(self):
        self.t = 0
        return {'position': np.zeros((3, 2, 3)), 'body_mask': np.ones((3, 2))}

    def step(self, action):
        self.t += 1
        done = np.array([self.t >= 1, self.t >= 2, self.t >= 3])
        return {'position': np.full((3, 2, 3), float(self.t)), 'body_mask': np.ones((3, 2))}, np.arange(3.0) + self.t, done, None


class BatchPolicy(object):
    def action(self, observation):
        return np.tile(np.arange(4, dtype=np.float32), (3, 1)) + observation['position'][:, 0, :1]


def test_batched_driver_slices_per_environment():
    batch = episodes.collect(BatchEnv(), BatchPolicy())
    assert batch.lengths.tolist() == [1, 2, 3] and batch.actions.shape == (3, 3, 4)
    np.testing.assert_allclose(batch.returns, [1.0, 2.0 + 3.0, 3.0 + 4.0 + 5.0])     # rewards after an env's done do not count
    eps = batch.episodes()
    assert [len(e['transitions']) for e in eps] == [1, 2, 3]
    assert eps[2]['transitions'][1]['state']['position'].shape == (2, 3)
    assert eps[1]['transitions'][1]['reward'] == 3.0 and eps[1]['transitions'][1]['action'].shape == (4,)
    assert episodes.generate_batched_episodes(BatchEnv(), BatchPolicy(), num_steps=2)[2]['transitions'][1]['reward'] == 4.0


def test_shard_round_trip(tmp_path):
    writer = episodes.ShardWriter(str(tmp_path))
    batch = episodes.collect(BatchEnv(), BatchPolicy())
    paths = [writer(batch), writer(batch)]
    assert sorted(os.listdir(str(tmp_path))) == ['episodes_000000.npz', 'episodes_000001.npz']
    back = episodes.read_shard(paths[1])
    assert back.lengths.tolist() == [1, 2, 3] and back.timestamp == batch.timestamp
    np.testing.assert_array_equal(back.actions, batch.actions)
    np.testing.assert_array_equal(back.states['position'], batch.states['position'])
    np.testing.assert_array_equal(back.final['position'], batch.final['position'])


def test_batched_sampler_is_distribution_equivalent_to_the_scalar_one():
    """BatchedHeuristicPolicy (vectorised rejection sampling) vs HeuristicPushSampler (bit-identical to the reference,
    tests/test_golden_cpu.py): every accepted action satisfies the reference's acceptance rule and the accepted
    actions have the same distribution (first two moments, 240 samples each)."""
    from robovat_b200 import config, policies
    pc = config.default_policy_config()
    position = np.array([[0.55, 0.10, 0.03], [0.70, -0.15, 0.03], [0.45, -0.20, 0.03]])
    mask = np.ones(3)
    n = 240
    s = policies.HeuristicPushSampler(pc.ACTION.CSPACE.LOW, pc.ACTION.CSPACE.HIGH, pc.ACTION.MOTION.TRANSLATION_X,
                                      pc.ACTION.MOTION.TRANSLATION_Y, max_attemps=20000)
    np.random.seed(5)
    scalar = np.stack([s.sample(position, mask, 1, 0)[0] for _ in range(n)])
    obs = {'position': np.tile(position, (n, 1, 1)), 'body_mask': np.ones((n, 3)), 'num_episodes': np.ones(n, np.int64),
           'num_steps': np.zeros(n, np.int64)}
    batched = policies.BatchedHeuristicPolicy(pc, seed=7, rounds=20000).action(obs)
    for a in (scalar, batched):
        for row in a[:200]:
            wp = s.get_waypoints(row[:2], row[2:])
            assert s.is_waypoint_clear(wp[0], None, position, s.start_margin)
            assert not s.is_waypoint_clear(wp[0], wp[1], position[1:2], s.motion_margin)     # target = num_episodes % 3 = 1
    se = np.sqrt(scalar.var(axis=0) / n + batched.var(axis=0) / n)
    assert (np.abs(scalar.mean(axis=0) - batched.mean(axis=0)) < 5 * se + 1e-3).all(), (scalar.mean(axis=0), batched.mean(axis=0))
    assert (np.abs(scalar.std(axis=0) - batched.std(axis=0)) < 0.25 * scalar.std(axis=0) + 1e-3).all()


def test_rollout_records_become_episode_batches():
    """batches_from_records on the oracle's device-rollout twin: the per-episode batches carry the recorded actions,
    rewards and observations with `collect`'s conventions (state t = observation before step t, final = after the
    last step, nothing counts after an environment's own end) and survive the shard round trip."""
    from tests import helpers
    cfg, w = helpers.make_oracle(5, threads=4, TASK_NAME='clearing', LAYOUT_ID=0)
    w.reset(seed=2)
    w.settle(0.1, 0.1, 500)
    w.settle()
    w.begin_episode()
    rec = w.rollout_begin(num_actions=3, num_episodes=2, policy_seed=9, reset_seed=4, max_attempts=500)
    while w.rollout_run(2000) > 0:
        pass
    batches = episodes.batches_from_records(rec)
    assert len(batches) == 2
    for ep, b in enumerate(batches):
        L = rec['lengths'][:, ep]
        assert b.lengths.tolist() == L.tolist() and b.actions.shape == (3, 5, 4) and b.states['position'].shape == (3, 5, w.N, 3)
        for e in range(5):
            n = int(L[e])
            np.testing.assert_array_equal(b.actions[:n, e], rec['actions'][e, ep, :n])
            np.testing.assert_array_equal(b.states['position'][:n, e], rec['positions'][e, ep, :n])
            np.testing.assert_array_equal(b.final['position'][e], rec['positions'][e, ep, n])
            assert (b.actions[n:, e] == 0).all() and (b.rewards[n:, e] == 0).all()
            assert b.states['is_safe'][0, e] == 1 and b.final['is_safe'][e] == (rec['flags'][e, ep, n - 1] & 1)
        np.testing.assert_allclose(b.returns, rec['returns'][:, ep], rtol=1e-6)
        eps = b.episodes()
        assert [len(x['transitions']) for x in eps] == L.tolist()
    np.testing.assert_array_equal(batches[1].final['body_mask'], w.body_mask.astype(np.float32))     # the last scene's bodies
    w.close()
