"""Episodes without the host (b2s_rollout_*), CPU side: the oracle's rollout is checked against the oracle's own
lock-step calls (reset / settle / begin_episode / set_action / env_substeps / reward -- the calls PushEnv.step and
Simulator.reset_scene make), and the Philox policy against the acceptance rule of the reference's sampler
(robovat/envs/push/heuristic_push_sampler.py:66-123).  The CUDA rollout is compared with the oracle's in
tests/test_gpu_rollout.py."""
import numpy as np

from robovat_b200 import _capi
from tests import helpers

TASK = dict(TASK_NAME='clearing', LAYOUT_ID=0)


def _settled_oracle(num_envs, seed, **bindings):
    cfg, w = helpers.make_oracle(num_envs, threads=4, **bindings)
    w.reset(seed=seed)
    w.settle(0.1, 0.1, 500)
    w.settle()
    return cfg, w


def _lockstep_episodes(w, num_actions, num_episodes, policy_seed, reset_seed, max_attempts, max_retries=8):
    """The host loop of generate_episodes over the lock-step calls, one env at a time semantics (every env of the
    world goes through the same calls; envs are independent, so masks select who is stepped)."""
    B, N = w.B, w.N
    A, EP = num_actions, num_episodes
    rec = {'actions': np.zeros((B, EP, A, 4), np.float32), 'rewards': np.zeros((B, EP, A), np.float32),
           'positions': np.zeros((B, EP, A + 1, N, 3), np.float32), 'flags': np.zeros((B, EP, A), np.uint8),
           'substeps': np.zeros((B, EP, A), np.int32), 'lengths': np.zeros((B, EP), np.int32),
           'returns': np.zeros((B, EP), np.float32)}
    num_eps = np.zeros(B, np.int32)
    table_z = [s['pose'][2] for s in w.scene.statics if s['flags'] & _capi.STATIC_IS_TABLE][-1]
    for e in range(B):
        only = np.zeros(B, np.uint8)
        only[e] = 1
        for ep in range(EP):
            if ep > 0:
                for attempt in range(max_retries + 1):
                    w.reset(seed=reset_seed, mask=only)
                    w.settle(0.1, 0.1, 500, mask=only)
                    w.settle(mask=only)
                    nm = int(w.num_movables[e])
                    z = w.body_state[2, e, :nm]
                    bad = bool((z < np.float32(table_z) + w.array(_capi.ARR_TABLE_DZ)[e]).any()) or bool(w.array(_capi.ARR_ERROR_FLAGS)[e] & 128)
                    if not bad:
                        break
                else:
                    raise AssertionError('no valid scene')
            w.begin_episode(mask=only)
            rec['positions'][e, ep, 0] = w.observe()[e]
            ret = np.float32(0.0)
            for t in range(A):
                act = w.policy_sample(policy_seed, t, num_eps, max_attempts)[e]
                a = w.array('action').reshape(B, 4).copy()
                a[e] = act
                ph = w.array(_capi.ARR_PHASE).copy()
                w.set_action(a)
                others = np.arange(B) != e
                w.array(_capi.ARR_PHASE)[others] = ph[others]            # only env e executes
                while w.env_substeps(500) > 0 and w.array(_capi.ARR_PHASE)[e] != _capi.PHASE_IDLE:
                    pass
                assert w.array(_capi.ARR_PHASE)[e] == _capi.PHASE_IDLE
                pos = w.observe()[e].copy()
                r, term = w.reward()
                env_done = w.array(_capi.ARR_PHASE_STATE).reshape(B, 8)[e, 4] != 0
                rec['actions'][e, ep, t] = act
                rec['rewards'][e, ep, t] = r[e]
                rec['positions'][e, ep, t + 1] = pos
                rec['flags'][e, ep, t] = (int(w.array('is_safe')[e]) | (int(w.array('is_effective')[e]) << 1)
                                          | (int(term[e]) << 2) | (int(env_done) << 3))
                rec['substeps'][e, ep, t] = w.array(_capi.ARR_NUM_STEPS)[e]
                ret = np.float32(ret + r[e])
                rec['lengths'][e, ep] = t + 1
                if term[e] or env_done:
                    break
            rec['returns'][e, ep] = ret
            num_eps[e] += 1
    return rec


def test_rollout_equals_lockstep_calls():
    """Three episodes of up to three actions in four envs: the rollout's records (actions drawn on its own, rewards,
    observations, flags, Simulator.num_steps, lengths, returns) are bit-identical to the host-driven loop, including
    the scene resets between episodes."""
    A, EP, B = 3, 3, 4
    cfg, w = _settled_oracle(B, seed=3, **TASK)
    w.begin_episode()
    rec = w.rollout_begin(num_actions=A, num_episodes=EP, policy_seed=21, reset_seed=5, max_attempts=2000)
    launched = 0
    while w.rollout_run(1000) > 0:
        launched += 1000
        assert launched < 200000
    assert ((w.array(_capi.ARR_ERROR_FLAGS) & ~8) == 0).all()          # bit3: the tile scene can exceed max_contacts while dropping
    assert (w.array(_capi.ARR_NUM_EPISODES) == EP).all()

    cfg2, w2 = _settled_oracle(B, seed=3, **TASK)
    ref = _lockstep_episodes(w2, A, EP, policy_seed=21, reset_seed=5, max_attempts=2000)
    for k in ('lengths', 'flags', 'substeps'):
        assert np.array_equal(rec[k], ref[k]), k
    valid = np.arange(A)[None, None, :] < rec['lengths'][:, :, None]
    for k in ('actions', 'rewards'):
        a, b = rec[k], ref[k]
        m = valid if a.ndim == 3 else valid[..., None]
        helpers.assert_bits_equal(np.where(m, a, 0), np.where(m, b, 0), k)
    pvalid = (np.arange(A + 1)[None, None, :] <= rec['lengths'][:, :, None])[..., None, None]
    helpers.assert_bits_equal(np.where(pvalid, rec['positions'], 0), np.where(pvalid, ref['positions'], 0), 'positions')
    helpers.assert_bits_equal(rec['returns'], ref['returns'], 'returns')
    assert rec['lengths'].min() >= 1 and rec['lengths'].max() <= A
    w.close()
    w2.close()


def test_rollout_is_independent_of_batch_composition_and_chunking():
    """An env's episodes depend on its global id only: a world holding envs [2, 4) reproduces envs 2..3 of a world
    holding [0, 4), and the size of the run calls does not matter."""
    A, EP = 2, 2
    _, wa = _settled_oracle(4, seed=9)
    wa.begin_episode()
    ra = wa.rollout_begin(num_actions=A, num_episodes=EP, policy_seed=4, reset_seed=8, max_attempts=2000)
    while wa.rollout_run(777) > 0:
        pass
    cfg, scene, params = helpers.make_inputs(2, params={'env_id_offset': 2})
    from oracle import b2o
    wb = b2o.OracleWorld(params, scene, threads=2)
    wb.reset(seed=9)
    wb.settle(0.1, 0.1, 500)
    wb.settle()
    wb.begin_episode()
    rb = wb.rollout_begin(num_actions=A, num_episodes=EP, policy_seed=4, reset_seed=8, max_attempts=2000)
    while wb.rollout_run(250) > 0:
        pass
    for k in ra:
        assert np.array_equal(ra[k][2:], rb[k]), k
    wa.close()
    wb.close()


def test_policy_accepts_what_the_reference_sampler_accepts():
    """Every action the Philox policy returns within its attempt budget satisfies the reference's acceptance rule
    (start more than 5 cm from every body; start or clipped end point within 1 cm of the target body), the target and
    the base angle follow num_episodes as in heuristic_push_sampler.py:66-73, and the draws are inside the sampler's
    ranges.  With one attempt the candidate is returned as is (the reference also returns its last try)."""
    B = 64
    cfg, w = _settled_oracle(B, seed=1)
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
    off, rng = 0.5 * (hi + lo), 0.5 * (hi - lo)
    tr = np.array([cfg.ACTION.MOTION.TRANSLATION_X, cfg.ACTION.MOTION.TRANSLATION_Y])
    pos = w.observe()[..., :2].astype(np.float64)
    nm = np.asarray(w.num_movables)
    accepted = 0
    for num_episodes in (0, 1, 5):
        act = w.policy_sample(seed=77, action_index=2, num_episodes=num_episodes, max_attempts=60000).astype(np.float64)
        assert np.all(np.abs(act) <= 1.0)
        base = (num_episodes * 42) % (2 * np.pi)
        for e in range(B):
            p0 = act[e, :2] * rng + off
            p1 = np.clip(p0 + act[e, 2:] * tr, lo, hi)
            d = np.linalg.norm(pos[e, :nm[e]] - p0, axis=-1)
            tgt = pos[e, num_episodes % nm[e]]
            ok = (d > 0.05 - 1e-6).all() and min(np.linalg.norm(tgt - p0), np.linalg.norm(tgt - p1)) < 0.01 + 1e-6
            accepted += ok
            if ok:
                # direction: base angle +- pi/4, jitter +-0.3 per axis, clipped
                lo_m = np.clip(np.cos(base + np.linspace(-np.pi / 4, np.pi / 4, 41)).min() - 0.3, -1, 1)
                hi_m = np.clip(np.cos(base + np.linspace(-np.pi / 4, np.pi / 4, 41)).max() + 0.3, -1, 1)
                assert lo_m - 1e-3 <= act[e, 2] <= hi_m + 1e-3
    # an accepted sample needs start = target - motion * translation inside the c-space: not every scene has one
    assert accepted >= 0.6 * 3 * B
    one = w.policy_sample(seed=77, action_index=0, num_episodes=0, max_attempts=1)
    again = w.policy_sample(seed=77, action_index=0, num_episodes=0, max_attempts=1)
    other = w.policy_sample(seed=78, action_index=0, num_episodes=0, max_attempts=1)
    assert np.array_equal(one, again) and not np.array_equal(one, other)
    # the start points are uniform over the square
    assert abs(one[:, :2].mean()) < 0.2 and one[:, :2].std() > 0.4
    w.close()


def test_async_stepping_equals_lockstep_per_env():
    """b2s_env_async_step's semantics on the oracle: envs take scripted actions whenever they are ready, are reset after
    three actions, and advance in slices of 300 substeps.  Per env the sequence of finished transitions (reward, PoseObs
    row, flags) and the final state equal the lock-step calls applied to that env alone -- waiting for a slice boundary
    changes nothing."""
    B, K = 5, 5
    rs = np.random.RandomState(3)
    script = rs.uniform(-1, 1, (B, K, 4)).astype(np.float32)
    cfg, w = _settled_oracle(B, seed=6, **TASK)
    w.begin_episode()
    taken = np.zeros(B, int)           # actions started
    since_reset = np.zeros(B, int)
    need_reset = np.zeros(B, bool)
    ready = np.ones(B, bool)
    log = [[] for _ in range(B)]
    for it in range(2000):
        if (taken >= K).all() and ready.all():
            break
        cmd = np.zeros(B, np.uint8)
        act = np.zeros((B, 4), np.float32)
        for e in range(B):
            if not ready[e]:
                continue
            if need_reset[e]:
                cmd[e] = 2
                need_reset[e] = False
                since_reset[e] = 0
            elif taken[e] < K:
                cmd[e] = 1
                act[e] = script[e, taken[e]]
                taken[e] += 1
                since_reset[e] += 1
        w.array('action')[:] = act.ravel()
        status = w.env_async_step(cmd, 300, reset_seed=12)
        for e in range(B):
            if status[e] & 2:
                log[e].append(('action', float(w.array('reward')[e]), w.array('obs_position').reshape(B, -1, 3)[e].copy(),
                               int(w.array('is_safe')[e]), int(w.array('is_effective')[e]), int(w.array('termination')[e])))
                if since_reset[e] == 3:
                    need_reset[e] = True
            if status[e] & 4:
                log[e].append(('reset', w.array('obs_position').reshape(B, -1, 3)[e].copy()))
        ready = (status & 1) != 0
    assert (taken == K).all()
    final = w.body_state.copy()

    cfg2, w2 = _settled_oracle(B, seed=6, **TASK)
    w2.begin_episode()
    table_z = [s['pose'][2] for s in w2.scene.statics if s['flags'] & _capi.STATIC_IS_TABLE][-1]
    for e in range(B):
        only = np.zeros(B, np.uint8)
        only[e] = 1
        others = np.arange(B) != e
        expect = []
        for k in range(K):
            if k == 3:
                for attempt in range(9):
                    w2.reset(seed=12, mask=only)
                    w2.settle(0.1, 0.1, 500, mask=only)
                    w2.settle(mask=only)
                    nm = int(w2.num_movables[e])
                    bad = bool((w2.body_state[2, e, :nm] < np.float32(table_z) + w2.array(_capi.ARR_TABLE_DZ)[e]).any()) or bool(w2.array(_capi.ARR_ERROR_FLAGS)[e] & 128)
                    if not bad:
                        break
                w2.begin_episode(mask=only)
                expect.append(('reset', w2.observe()[e].copy()))
            a = w2.array('action').reshape(B, 4).copy()
            a[e] = script[e, k]
            ph = w2.array(_capi.ARR_PHASE).copy()
            w2.set_action(a)
            w2.array(_capi.ARR_PHASE)[others] = ph[others]
            while w2.array(_capi.ARR_PHASE)[e] != _capi.PHASE_IDLE:
                w2.env_substeps(500)
            pos = w2.observe()[e].copy()
            r, term = w2.reward()
            expect.append(('action', float(r[e]), pos, int(w2.array('is_safe')[e]), int(w2.array('is_effective')[e]), int(term[e])))
        assert len(expect) == len(log[e]), (e, len(expect), len(log[e]))
        for got, want in zip(log[e], expect):
            assert got[0] == want[0]
            if got[0] == 'action':
                assert got[1] == want[1] and got[3:] == want[3:], (e, got, want)
                helpers.assert_bits_equal(got[2], want[2], 'PoseObs row of env %d' % e)
            else:
                helpers.assert_bits_equal(got[1], want[1], 'first observation of env %d' % e)
    helpers.assert_bits_equal(final, w2.body_state, 'final body_state')
    w.close()
    w2.close()
