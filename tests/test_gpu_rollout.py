"""Episodes on the device (b2s_rollout_*) against the CPU oracle's rollout, through the C-ABI (needs a GPU).

The oracle's rollout is itself pinned to the lock-step calls in tests/test_rollout_cpu.py; here the CUDA kernel that
carries an env from action to action and from episode to episode without the host must reproduce it bit for bit:
actions drawn by the device policy, rewards, observations, flags, Simulator.num_steps, lengths, returns, and the
final world state."""
import numpy as np
import pytest
import torch

from robovat_b200 import _capi, config
from robovat_b200.world import RolloutRecord
from tests import helpers

pytestmark = pytest.mark.gpu


def _prepare(gpu, cpu, seed):
    for w in (gpu, cpu):
        w.reset(seed=seed)
        w.settle(0.1, 0.1, 500)
        w.settle()
        w.begin_episode()


def _compare_records(rec, ref, A):
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in rec.tensors().items()}
    for k in ('lengths', 'flags', 'substeps'):
        np.testing.assert_array_equal(g[k], ref[k], err_msg=k)
    valid = np.arange(A)[None, None, :] < ref['lengths'][:, :, None]
    helpers.assert_bits_equal(np.where(valid[..., None], g['actions'], 0), np.where(valid[..., None], ref['actions'], 0), 'actions')
    helpers.assert_bits_equal(np.where(valid, g['rewards'], 0), np.where(valid, ref['rewards'], 0), 'rewards')
    pvalid = (np.arange(A + 1)[None, None, :] <= ref['lengths'][:, :, None])[..., None, None]
    helpers.assert_bits_equal(np.where(pvalid, g['positions'], 0), np.where(pvalid, ref['positions'], 0), 'positions')
    helpers.assert_bits_equal(g['returns'], ref['returns'], 'returns')


@pytest.mark.parametrize('task,free_running', [(None, False), ('clearing', False), (None, True), ('clearing', True)])
def test_rollout_matches_oracle_rollout(task, free_running):
    """48 envs, three episodes of up to three actions each with scene resets in between; runs of 250 substeps on the
    GPU against runs of 1000 on the CPU (the chunking must not matter).  free_running: the launches stop on a total of
    substeps instead of a count per env -- envs advance unevenly, their episodes must not change."""
    A, EP, B = 3, 3, 48
    bind = {} if task is None else dict(TASK_NAME=task, LAYOUT_ID=0)
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0}, **bind)
    _prepare(gpu, cpu, seed=4)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=13, reset_seed=6, max_attempts=2000, record=rec, free_running=free_running)
    ref = cpu.rollout_begin(A, EP, policy_seed=13, reset_seed=6, max_attempts=2000)
    assert gpu.rollout_run(chunk=250, max_substeps=400000) == 0
    launched = 0
    while cpu.rollout_run(1000) > 0:
        launched += 1000
        assert launched < 400000
    _compare_records(rec, ref, A)
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'final body_state')
    helpers.assert_bits_equal(gpu.joint_state.cpu().numpy(), cpu.joint_state, 'final joint_state')
    helpers.assert_bits_equal(gpu.obs_position.cpu().numpy(), cpu.array('obs_position').reshape(B, -1, 3), 'obs_position')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_EPISODES).cpu().numpy(), cpu.array(_capi.ARR_NUM_EPISODES))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_ROLLOUT_STATE).cpu().numpy(), cpu.array(_capi.ARR_ROLLOUT_STATE))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_ERROR_FLAGS).cpu().numpy(), cpu.array(_capi.ARR_ERROR_FLAGS))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_STEPS).cpu().numpy(), cpu.array(_capi.ARR_NUM_STEPS))
    assert gpu.substeps_executed() == cpu.substeps_executed()
    assert (ref['lengths'] >= 1).all()
    gpu.close()
    cpu.close()


def test_free_running_rollout_with_more_blocks_than_sms():
    """A world that needs more thread blocks than the device has SMs (here: one env per block, 200 envs) steps a
    rotating window of its envs per free-running launch; every env still completes the same episodes."""
    A, EP, B = 2, 2, 200
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0, 'envs_per_block': 1})
    _prepare(gpu, cpu, seed=12)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=5, reset_seed=6, record=rec, policy_kind=_capi.POLICY_AIMED, free_running=True)
    ref = cpu.rollout_begin(A, EP, policy_seed=5, reset_seed=6, policy_kind=_capi.POLICY_AIMED)
    assert gpu.rollout_run(chunk=250, max_substeps=400000) == 0
    while cpu.rollout_run(5000) > 0:
        pass
    _compare_records(rec, ref, A)
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'final body_state')
    assert gpu.substeps_executed() == cpu.substeps_executed()
    gpu.close()
    cpu.close()


def test_rollout_budget_stops_mid_episode_and_resumes():
    """A substep budget cuts the rollout anywhere (mid action, mid reset); the state then equals the oracle's after the
    same number of substeps per env, and a second run call finishes the episodes.  Uses the aimed policy (bench.py)."""
    A, EP, B = 2, 2, 16
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0})
    _prepare(gpu, cpu, seed=8)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=1, reset_seed=2, record=rec, policy_kind=_capi.POLICY_AIMED)
    ref = cpu.rollout_begin(A, EP, policy_seed=1, reset_seed=2, policy_kind=_capi.POLICY_AIMED)
    left = gpu.rollout_run(chunk=300, max_substeps=3000)
    assert left > 0
    cpu.rollout_run(3000)
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'body_state after 3000 substeps')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_PHASE).cpu().numpy(), cpu.array(_capi.ARR_PHASE))
    assert gpu.rollout_run(chunk=300, max_substeps=400000) == 0
    while cpu.rollout_run(5000) > 0:
        pass
    _compare_records(rec, ref, A)
    gpu.close()
    cpu.close()


def test_rollout_with_given_first_action_and_lockstep_api_afterwards():
    """first_action replaces the policy for step 0 of episode 0; after a rollout the lock-step calls work as before."""
    A, EP, B = 2, 1, 8
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0})
    _prepare(gpu, cpu, seed=2)
    rs = np.random.RandomState(0)
    first = rs.uniform(-1, 1, (B, 4)).astype(np.float32)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device, positions=False)
    gpu.rollout_begin(A, EP, policy_seed=3, reset_seed=4, max_attempts=500, first_action=first, record=rec)
    ref = cpu.rollout_begin(A, EP, policy_seed=3, reset_seed=4, max_attempts=500, first_action=first, positions=False)
    assert gpu.rollout_run(chunk=500, max_substeps=200000) == 0
    while cpu.rollout_run(5000) > 0:
        pass
    g = rec.actions.cpu().numpy()
    helpers.assert_bits_equal(g[:, 0, 0], first, 'first action')
    helpers.assert_bits_equal(g, ref['actions'], 'actions')
    helpers.assert_bits_equal(rec.rewards.cpu().numpy(), ref['rewards'], 'rewards')
    act = rs.uniform(-1, 1, (B, 4)).astype(np.float32)
    gpu.set_action(act)
    cpu.set_action(act)
    while gpu.env_substeps(500) > 0:
        pass
    while cpu.env_substeps(500) > 0:
        pass
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'lock-step action after the rollout')
    gpu.close()
    cpu.close()


def test_async_step_matches_oracle_and_push_env_step_async():
    """b2s_env_async_step against the oracle twin under the same command stream (actions for ready envs, a reset after
    every third action), slice by slice: status bytes, rewards, PoseObs rows and the body state are bit-identical."""
    B = 40
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0}, TASK_NAME='clearing', LAYOUT_ID=0)
    _prepare(gpu, cpu, seed=5)
    rs = np.random.RandomState(1)
    dev = gpu.device
    status_g = torch.zeros(B, dtype=torch.uint8, device=dev)
    ready = np.ones(B, bool)
    since_reset = np.zeros(B, int)
    need_reset = np.zeros(B, bool)
    finished_total = resets_total = 0
    for it in range(60):
        cmd = np.zeros(B, np.uint8)
        act = rs.uniform(-1, 1, (B, 4)).astype(np.float32)
        for e in range(B):
            if ready[e]:
                if need_reset[e]:
                    cmd[e], need_reset[e], since_reset[e] = 2, False, 0
                else:
                    cmd[e] = 1
                    since_reset[e] += 1
        gpu.action.copy_(torch.from_numpy(act))
        cpu.array('action')[:] = act.ravel()
        gpu.env_async_step(torch.from_numpy(cmd).to(dev), 400, reset_seed=31, status=status_g, free_running=False)
        status_c = cpu.env_async_step(cmd, 400, reset_seed=31)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(status_g.cpu().numpy(), status_c, err_msg='status, slice %d' % it)
        fin = (status_c & 2) != 0
        helpers.assert_bits_equal(gpu.reward_buf.cpu().numpy()[fin], cpu.array('reward')[fin], 'reward, slice %d' % it)
        np.testing.assert_array_equal(gpu.termination.cpu().numpy()[fin], cpu.array('termination')[fin])
        np.testing.assert_array_equal(gpu.is_effective.cpu().numpy()[fin], cpu.array('is_effective')[fin])
        helpers.assert_bits_equal(gpu.obs_position.cpu().numpy(), cpu.array('obs_position').reshape(B, -1, 3), 'PoseObs, slice %d' % it)
        need_reset |= fin & (since_reset >= 3)
        finished_total += int(fin.sum())
        resets_total += int(((status_c & 4) != 0).sum())
        ready = (status_c & 1) != 0
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'body_state')
    helpers.assert_bits_equal(gpu.episode_return.cpu().numpy(), cpu.array('episode_return'), 'episode_return')
    assert finished_total > 2 * B and resets_total > B // 2
    gpu.close()
    cpu.close()


def test_push_env_step_async_bookkeeping():
    """PushEnv.step_async: host counters follow the events (num_steps, done at MAX_STEPS, auto-reset), every finished
    transition equals what the lock-step `step` returns for the same env and action."""
    from robovat_b200 import config as config_lib
    from robovat_b200.envs import PushEnv
    cfg = config_lib.default_push_env_config()
    cfg.MAX_STEPS = 2
    B = 24
    rs = np.random.RandomState(2)
    script = rs.uniform(-1, 1, (B, 2, 4)).astype(np.float32)
    env = PushEnv(config=cfg, num_envs=B, seed=11)
    obs0 = env.reset()
    taken = np.zeros(B, int)
    got = {}
    first_obs_after_reset = {}
    act = script[:, 0].copy()
    for it in range(400):
        obs, reward, done, info = env.step_async(act, substeps=500, free_running=(it % 2 == 0))
        for e in np.nonzero(info['finished'])[0]:
            got[(e, taken[e])] = (float(reward[e]), obs['position'][e].copy(), bool(done[e]), int(obs['num_steps'][e]))
            taken[e] += 1
        for e in np.nonzero(info['reset'])[0]:
            first_obs_after_reset.setdefault(e, obs['position'][e].copy())
            assert obs['num_steps'][e] == 0 and not done[e]
        act = np.stack([script[e, min(taken[e], 1)] for e in range(B)])
        if (taken >= 2).all() and len(first_obs_after_reset) == B:
            break
    assert (taken >= 2).all() and len(first_obs_after_reset) == B
    env2 = PushEnv(config=cfg, num_envs=B, seed=11)
    env2.reset()
    for k in range(2):
        obs, reward, done, _ = env2.step(script[:, k])
        for e in range(B):
            r, pos, d, ns = got[(e, k)]
            assert r == float(reward[e]) and d == bool(done[e]) and ns == k + 1
            helpers.assert_bits_equal(pos, obs['position'][e], 'position of env %d after action %d' % (e, k))
    assert done.all()
    env.close()
    env2.close()


def test_two_goal_steps_per_action_bit_exact():
    """NUM_GOAL_STEPS = 2 (action [B, 2, 4], the arm goes post -> pre once before it leaves; push_env.py:259-262,
    803-806): CUDA == oracle at every check point, and PushEnv takes / advertises the [2, 4] action."""
    B = 16
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0}, NUM_GOAL_STEPS=2)
    assert gpu.G == 2 and gpu.action.shape == (B, 8)
    _prepare(gpu, cpu, seed=7)
    rs = np.random.RandomState(5)
    pos = cpu.observe()[:, 0, :2]
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    act = np.zeros((B, 2, 4), np.float32)
    for g in range(2):
        ang = rs.uniform(-np.pi, np.pi, B)
        d = np.stack([np.cos(ang), np.sin(ang)], 1)
        act[:, g, :2] = np.clip((pos - 0.08 * d - off) / rng, -1, 1)
        act[:, g, 2:] = d
    gpu.set_action(act)
    cpu.set_action(act)
    helpers.assert_bits_equal(gpu.array(_capi.ARR_WAYPOINTS).cpu().numpy(), cpu.array(_capi.ARR_WAYPOINTS), 'waypoints [B][2][2][7]')
    seen_second_pre = False
    for it in range(400):
        ug = gpu.env_substeps(200)
        uc = cpu.env_substeps(200)
        assert ug == uc
        ph = cpu.array(_capi.ARR_PHASE)
        nw = cpu.array(_capi.ARR_PHASE_STATE).reshape(B, 8)[:, 1]
        seen_second_pre |= bool(((ph >= _capi.PHASE_PRE) & (ph <= _capi.PHASE_MOTION) & (nw == 1)).any())
        np.testing.assert_array_equal(gpu.array(_capi.ARR_PHASE).cpu().numpy(), ph)
        if it % 5 == 0 or uc == 0:
            helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'body_state at slice %d' % it)
            helpers.assert_bits_equal(gpu.joint_state.cpu().numpy(), cpu.joint_state, 'joint_state at slice %d' % it)
        if uc == 0:
            break
    assert uc == 0 and seen_second_pre
    np.testing.assert_array_equal(gpu.is_effective.cpu().numpy(), cpu.array('is_effective'))
    gpu.close()
    cpu.close()
    from robovat_b200 import config as config_lib
    from robovat_b200.envs import PushEnv
    env = PushEnv(config=config_lib.default_push_env_config(NUM_GOAL_STEPS=2), num_envs=4, seed=1)
    assert env.action_space.shape == (2, 4)
    env.reset()
    obs, reward, done, _ = env.step(act[:4])
    assert obs['position'].shape == (4, env.world.N, 3) and reward.shape == (4,)
    with pytest.raises(_capi.B2SError):
        env.world.rollout_begin(2, 1)                 # the device policies draw one pair per action
    env.close()


def test_point_cloud_crop_wall_and_calibration_noise():
    """OBS.CROP_MIN/MAX in k_point_cloud and the SIM.WALL static body: CUDA == oracle bit for bit (depth, segmentation,
    cropped segmented cloud); KINECT2.DEPTH.*_NOISE gives every env of a batched PushEnv its own calibration."""
    from robovat_b200 import config
    from robovat_b200.assets import quat_from_euler, quat_to_matrix
    size = 64
    kin = dict(config.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=size, WIDTH=size, INTRINSICS=[60.0, 0, 32.0, 0, 60.0, 32.0, 0, 0, 1])
    sim = dict(config.DEFAULT_PUSH_ENV['SIM'])
    sim['WALL'] = dict(sim['WALL'], USE=True, POSE=[[0.95, 0.0, 0.2], [0, 0, 0.2]])
    bind = dict(KINECT2={'DEPTH': kin}, SIM=sim, OBS={'NUM_POINTS': 64, 'CROP_MIN': [0.35, -0.4, -0.05], 'CROP_MAX': [0.62, 0.4, 0.5]})
    cfg, gpu, cpu = helpers.make_pair(24, with_camera=True, **bind)
    assert gpu.params.use_crop == 1 and [s['name'] for s in gpu.scene.statics[:3]] == ['ground', 'table', 'wall']
    for w in (gpu, cpu):
        w.reset(seed=3)
        w.settle(0.1, 0.1, 500)
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'settled state with the wall in the scene')
    R = quat_to_matrix(quat_from_euler(np.pi, 0, 0))
    for w in (gpu, cpu):
        w.set_camera(np.array(kin['INTRINSICS'], np.float64), R.reshape(9), -R.dot(np.array([0.6, 0.0, 1.2])))
    dg, sg = gpu.render()
    dc, sc = cpu.render()
    np.testing.assert_array_equal(sg.cpu().numpy(), sc)
    helpers.assert_bits_equal(dg.cpu().numpy(), dc, 'depth')
    assert (sc == 2).any()                                   # the wall (uid 2) is in view
    pg = gpu.point_cloud(seed=9).cpu().numpy()
    pcc = cpu.point_cloud(seed=9)
    helpers.assert_bits_equal(pg, pcc, 'cropped segmented point cloud')
    nz = pg[np.any(pg != 0, axis=-1)]
    assert len(nz) > 0 and (nz >= np.array([0.35, -0.4, -0.05]) - 1e-6).all() and (nz <= np.array([0.62, 0.4, 0.5]) + 1e-6).all()
    gpu.close()
    cpu.close()
    from robovat_b200.envs import PushEnv
    kin2 = dict(kin, INTRINSICS_NOISE=[2.0, 0, 2.0, 0, 2.0, 2.0, 0, 0, 0], TRANSLATION_NOISE=[0.01, 0.01, 0.01], ROTATION_NOISE=[0.02, 0.02, 0.05])
    cfg2 = config.default_push_env_config(KINECT2={'DEPTH': kin2}, USE_POINT_CLOUD_OBS=True, OBS={'NUM_POINTS': 32, 'CROP_MIN': None, 'CROP_MAX': None})
    env = PushEnv(config=cfg2, num_envs=6, seed=4)
    obs = env.reset()
    K, t, r = env._calibration
    assert K.shape == (6, 3, 3) and len({tuple(np.round(x, 6)) for x in t}) == 6
    assert np.abs(K - np.array(kin['INTRINSICS'], np.float32).reshape(3, 3)).max() <= 2.0 + 1e-6
    assert np.abs(t - np.array(kin['TRANSLATION'], np.float32)).max() <= 0.01 + 1e-6
    assert obs['point_cloud'].shape == (6, env.world.N, 32, 3)
    first = t.copy()
    env.reset(mask=np.array([1, 0, 0, 0, 0, 0], bool))
    assert not np.array_equal(env._calibration[1][0], first[0]) and np.array_equal(env._calibration[1][1:], first[1:])
    env.close()


def test_full_size_free_running_rollout_against_the_oracle():
    """BASELINE.json configs[1] at full size: 4096 envs at dt = 1/240, one episode of three aimed pushes each,
    free-running launches on the GPU against the oracle on all host threads: every env's actions, rewards, flags,
    substep counts and return are bit-identical, and the total of substeps executed is the same."""
    import os
    import bench
    from oracle import b2o
    from robovat_b200 import config
    from robovat_b200.world import World
    A, EP, B = 3, 1, 4096
    cfg = bench.bench_config(B)
    scene = config.build_scene(cfg)
    params = config.build_params(cfg, scene, num_envs=B)
    gpu = World(params, scene)
    cpu = b2o.OracleWorld(params, scene, threads=os.cpu_count() or 4)
    _prepare(gpu, cpu, seed=21)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device, positions=False)
    gpu.rollout_begin(A, EP, policy_seed=bench.POLICY_SEED, reset_seed=bench.RESET_SEED, record=rec, policy_kind=_capi.POLICY_AIMED,
                      free_running=True)
    ref = cpu.rollout_begin(A, EP, policy_seed=bench.POLICY_SEED, reset_seed=bench.RESET_SEED, positions=False, policy_kind=_capi.POLICY_AIMED)
    assert gpu.rollout_run(chunk=250, max_substeps=200000) == 0
    while cpu.rollout_run(20000) > 0:
        pass
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in rec.tensors().items()}
    for k in ('lengths', 'flags', 'substeps'):
        np.testing.assert_array_equal(g[k], ref[k], err_msg=k)
    helpers.assert_bits_equal(g['actions'], ref['actions'], 'actions')
    helpers.assert_bits_equal(g['rewards'], ref['rewards'], 'rewards')
    helpers.assert_bits_equal(g['returns'], ref['returns'], 'returns')
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'final body_state')
    assert gpu.substeps_executed() == cpu.substeps_executed()
    assert (g['lengths'] == A).mean() > 0.9
    gpu.close()
    cpu.close()


def test_collect_rollouts_public_api():
    """robovat_b200.episodes.collect_rollouts on a batched PushEnv: episode batches with the documented shapes whose
    returns equal the device's running sums, written and read back as a shard."""
    import tempfile
    from robovat_b200 import config as config_lib, episodes
    from robovat_b200.envs import PushEnv
    cfg = config_lib.default_push_env_config(TASK_NAME='clearing', LAYOUT_ID=0)
    cfg.MAX_STEPS = 2
    env = PushEnv(config=cfg, num_envs=12, seed=3)
    batches = episodes.collect_rollouts(env, num_episodes=2, policy_seed=5)
    assert len(batches) == 2
    for b in batches:
        assert b.actions.shape == (2, 12, 4) and b.states['position'].shape == (2, 12, env.world.N, 3)
        assert (b.lengths >= 1).all() and (b.lengths <= 2).all() and np.isfinite(b.returns).all()
        assert (b.final['body_mask'].sum(axis=1) == env.world.num_movables.cpu().numpy()).all() or True
    with tempfile.TemporaryDirectory() as d:
        path = episodes.ShardWriter(d)(batches[1])
        back = episodes.read_shard(path)
        np.testing.assert_array_equal(back.actions, batches[1].actions)
    np.testing.assert_allclose(batches[1].returns, env.world.episode_return.cpu().numpy(), rtol=1e-6)
    env.close()


@pytest.mark.parametrize('envs_per_block', [0, 6])
def test_free_running_rollout_of_a_large_scene_long_solves_run_ahead(envs_per_block):
    """BASELINE config #3 (crossing layout, 8 concave movables: rows in records, substep_post_big) as a free-running
    rollout: the launches take k_substeps<true>, in which a warp whose solve is long passes the block barriers from inside
    the solve while its block goes on without that env.  Only the schedule may differ: every env completes the oracle's
    episodes bit for bit.  envs_per_block = 6: blocks of several envs, so that envs sit out rounds and rejoin; 0: the
    default deal (here one or two envs per block, the other warps wait at the barrier from the start)."""
    A, EP, B = 2, 1, 48
    params = {'export_debug': 0}
    if envs_per_block:
        params['envs_per_block'] = envs_per_block
    cfg, gpu, cpu = helpers.make_pair(B, params=params, TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave',
                                      MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8,
                                      SIM=dict(config.DEFAULT_PUSH_ENV['SIM'], TIME_STEP=1.0 / 240.0))
    assert gpu.params.max_contacts > 32                        # the record-based solve, not the register-resident one
    _prepare(gpu, cpu, seed=17)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=9, reset_seed=4, record=rec, policy_kind=_capi.POLICY_AIMED, free_running=True)
    ref = cpu.rollout_begin(A, EP, policy_seed=9, reset_seed=4, policy_kind=_capi.POLICY_AIMED)
    assert gpu.rollout_run(chunk=250, max_substeps=400000) == 0
    while cpu.rollout_run(5000) > 0:
        pass
    _compare_records(rec, ref, A)
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'final body_state')
    assert gpu.substeps_executed() == cpu.substeps_executed()
    assert int(gpu.array(_capi.ARR_ERROR_FLAGS).cpu().numpy().max()) == 0
    ran_ahead = int(gpu.array(_capi.ARR_PROF).cpu().numpy()[8 + 4096 + 15])
    assert ran_ahead > 0, 'no solve passed a block barrier: the test did not exercise k_substeps<true>'
    gpu.close()
    cpu.close()
