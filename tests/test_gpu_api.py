"""Every remaining C-ABI entry point on the GPU against the oracle and the reference goldens (needs a GPU):
the small kernels (observe, SE(3), contact queries, reward on every task, env_step, all-gather of returns), the
product configuration without the inspection write-backs, full-size batches against the oracle, the reset / reward
semantics of PushEnv, and two worlds sharing one device."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from robovat_b200 import _capi
from tests import helpers
from tests.test_golden_cpu import quat_close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def gpu_se3(op, a, b, width):
    lib = _capi.load()
    fn = [lib.b2s_se3_quat_from_euler, lib.b2s_se3_euler_from_quat, lib.b2s_se3_matrix_from_quat,
          lib.b2s_se3_quat_multiply, lib.b2s_se3_pose_inverse, lib.b2s_se3_pose_transform][op]
    ta = torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()
    out = torch.zeros(ta.shape[0], width, dtype=torch.float32, device='cuda')
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if op in (3, 5):
        tb = torch.as_tensor(np.ascontiguousarray(b, np.float32)).cuda()
        _capi.check(lib, fn(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()), ta.shape[0], st))
    else:
        _capi.check(lib, fn(C.c_void_p(ta.data_ptr()), C.c_void_p(out.data_ptr()), ta.shape[0], st))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def cpu_se3(op, a, b, width):
    from oracle import b2o
    lib = b2o.load()
    a = np.ascontiguousarray(a, np.float32)
    out = np.zeros((a.shape[0], width), np.float32)
    bp = None if b is None else np.ascontiguousarray(b, np.float32).ctypes.data_as(C.c_void_p)
    assert lib.b2o_se3(op, a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), a.shape[0]) == 0
    return out


def test_se3_kernels_match_reference_goldens_and_oracle():
    """All six b2s_se3_* exports: against third_party/transformations.py and robovat.math.Pose (goldens, fp32
    tolerances of tests/test_golden_cpu.py) and bit for bit against the same leaf math on the CPU."""
    g = load('transformations.json')
    p = load('pose.json')
    a, b = np.array(p['a']), np.array(p['b'])
    cases = [(0, g['euler'], None, 4), (1, g['quaternion_from_euler'], None, 3), (2, g['quaternion_from_euler'], None, 9),
             (3, g['qa'], g['qb'], 4), (4, a, None, 7), (5, a, b, 7)]
    outs = []
    for op, x, y, width in cases:
        got = gpu_se3(op, x, y, width)
        helpers.assert_bits_equal(got, cpu_se3(op, x, y, width), 'se3 op %d' % op)
        outs.append(got)
    np.testing.assert_allclose(outs[0], g['quaternion_from_euler'], atol=2e-6)
    np.testing.assert_allclose(outs[1], g['euler_from_quaternion'], atol=2e-5)
    np.testing.assert_allclose(outs[2].reshape(-1, 3, 3), g['matrix3_from_quaternion'], atol=2e-6)
    np.testing.assert_allclose(outs[3], g['quaternion_multiply'], atol=2e-6)
    np.testing.assert_allclose(outs[4][:, :3], np.array(p['inverse_a'])[:, :3], atol=3e-6)
    assert quat_close(outs[4][:, 3:], np.array(p['inverse_a'])[:, 3:], 3e-6)
    np.testing.assert_allclose(outs[5][:, :3], np.array(p['a_transform_b'])[:, :3], atol=3e-6)
    assert quat_close(outs[5][:, 3:], np.array(p['a_transform_b'])[:, 3:], 2e-4)
    rel = gpu_se3(5, gpu_se3(4, b, None, 7), a, 7)                    # get_transform(source=a, target=b)
    ref = np.array(p['get_transform_source_a_target_b'])
    np.testing.assert_allclose(rel[:, :3], ref[:, :3], atol=5e-6)
    assert quat_close(rel[:, 3:], ref[:, 3:], 2e-4)
    d = g['doctests']
    np.testing.assert_allclose(gpu_se3(3, [[1, -2, 3, 4]], [[-5, 6, 7, 8]], 4)[0], d['quaternion_multiply([1,-2,3,4],[-5,6,7,8])'], atol=1e-5)


def test_reward_kernel_on_every_task_matches_reference_goldens():
    """k_reward for clearing / insertion / crossing layouts: termination exact and reward within 2e-5 of
    push_reward.get_reward_fn (goldens), bit-identical to the oracle."""
    from robovat_b200.world import World
    from oracle import b2o
    for case in load('reward.json'):
        s0, s1 = np.array(case['state'], np.float32), np.array(case['next_state'], np.float32)
        B, n_max, _ = s0.shape
        cfg, scene, params = helpers.make_inputs(B, TASK_NAME=case['task'], LAYOUT_ID=case['layout_id'],
                                                 MIN_MOVABLE_BODIES=n_max, MAX_MOVABLE_BODIES=n_max)
        gpu, cpu = World(params, scene), b2o.OracleWorld(params, scene)
        rg, tg = gpu.reward(s0, s1)
        rc, tc = cpu.reward(s0, s1)
        torch.cuda.synchronize()
        helpers.assert_bits_equal(rg.cpu().numpy(), rc, 'reward %s/%d' % (case['task'], case['layout_id']))
        np.testing.assert_array_equal(tg.cpu().numpy(), tc)
        np.testing.assert_array_equal(tg.cpu().numpy().astype(bool), np.array(case['termination']))
        np.testing.assert_allclose(rg.cpu().numpy(), case['reward'], atol=2e-5)
        helpers.assert_bits_equal(gpu.episode_return.cpu().numpy(), cpu.array('episode_return'), 'episode_return')
        gpu.close(); cpu.close()


def test_observe_query_contacts_and_env_step_match_oracle():
    """b2s_observe (ragged movable counts), b2s_query_contacts while the arm pushes, b2s_env_step (set_action + the
    chunked loop inside the library) against the oracle."""
    cfg, gpu, cpu = helpers.make_pair(24, MIN_MOVABLE_BODIES=1, MAX_MOVABLE_BODIES=4)
    gpu.reset(seed=4); cpu.reset(seed=4)
    gpu.settle(0.1, 0.1, 500); cpu.settle(0.1, 0.1, 500)
    helpers.assert_bits_equal(gpu.observe().cpu().numpy(), cpu.observe(), 'observe after settle')
    np.testing.assert_array_equal(gpu.body_mask.cpu().numpy(), cpu.body_mask)
    pos0 = cpu.observe().copy()
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    act = np.zeros((24, 4), np.float32)
    act[:, :2] = np.clip((pos0[:, 0, :2] - [0.08, 0.0] - off) / rng, -1, 1)
    act[:, 2] = 1.0
    gpu.set_action(act); cpu.set_action(act)
    seen_table = seen_movable = 0
    for _ in range(60):
        ug, uc = gpu.env_substeps(100), cpu.env_substeps(100)
        assert ug == uc
        at, am = gpu.query_contacts()
        cf = cpu.array(_capi.ARR_CONTACT_FLAGS)
        np.testing.assert_array_equal(at.cpu().numpy(), cf & 1)
        np.testing.assert_array_equal(am.cpu().numpy(), (cf >> 1) & 1)
        seen_table += int((cf & 1).sum()); seen_movable += int(((cf >> 1) & 1).sum())
        helpers.assert_bits_equal(gpu.observe().cpu().numpy(), cpu.observe(), 'observe mid action')
        if uc == 0:
            break
    assert seen_movable > 0
    # a second action through b2s_env_step; the oracle runs the same loop from Python
    act2 = np.tile(np.array([0.2, -0.1, -0.6, 0.8], np.float32), (24, 1))
    gpu.action.copy_(torch.from_numpy(act2))
    gpu.env_step(chunk=200, max_substeps=40000)
    cpu.set_action(act2)
    while cpu.env_substeps(200) > 0:
        pass
    torch.cuda.synchronize()
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'env_step: body_state')
    helpers.assert_bits_equal(gpu.joint_state.cpu().numpy(), cpu.joint_state, 'env_step: joint_state')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_STEPS).cpu().numpy(), cpu.array(_capi.ARR_NUM_STEPS))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_PHASE).cpu().numpy(), cpu.array(_capi.ARR_PHASE))


def test_product_configuration_without_inspection_writebacks_is_bit_exact():
    """params.export_debug = 0 (the default of the product): pair keys / link poses / link velocities are not written by
    the substep kernel; everything that defines the simulation still matches the oracle bit for bit."""
    cfg, gpu, cpu = helpers.make_pair(32, params={'export_debug': 0})
    assert gpu.params.export_debug == 0
    gpu.reset(seed=12); cpu.reset(seed=12)
    M, B = gpu.params.max_manifolds, 32
    for k in range(6):
        gpu.step(100); cpu.step(100)
        torch.cuda.synchronize()
        helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'body_state %d' % k)
        gk = helpers.manifold_view(gpu.array(_capi.ARR_MANIFOLD_KEYS).cpu().numpy(), gpu.array(_capi.ARR_MANIFOLD_NPTS).cpu().numpy(),
                                   gpu.array(_capi.ARR_MANIFOLD_PTS).cpu().numpy(), B, M)
        ck = helpers.manifold_view(cpu.array(_capi.ARR_MANIFOLD_KEYS), cpu.array(_capi.ARR_MANIFOLD_NPTS),
                                   cpu.array(_capi.ARR_MANIFOLD_PTS), B, M)
        for a, b, what in zip(gk, ck, ('keys', 'npts', 'points', 'gjk cache')):
            helpers.assert_bits_equal(a, b, 'manifold ' + what)
        np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_PAIRS).cpu().numpy(), cpu.array(_capi.ARR_NUM_PAIRS))
        np.testing.assert_array_equal(gpu.array(_capi.ARR_SOLVER_STATS).cpu().numpy(), cpu.array(_capi.ARR_SOLVER_STATS))
    # the calls that need link poses refresh them themselves
    helpers.assert_bits_equal(gpu.forward_kinematics().cpu().numpy(), cpu.forward_kinematics(), 'fk')
    act = np.tile(np.array([0.0, 0.0, 1.0, 0.0], np.float32), (32, 1))
    gpu.set_action(act); cpu.set_action(act)
    while True:
        ug, uc = gpu.env_substeps(400), cpu.env_substeps(400)
        assert ug == uc
        if uc == 0:
            break
    helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'after action')
    helpers.assert_bits_equal(gpu.joint_state.cpu().numpy(), cpu.joint_state, 'joints after action')


def test_full_size_batch_against_the_oracle():
    """BASELINE's 4096 envs vs the oracle: an oracle world that owns the global ids 1024..1183 (env_id_offset) is
    compared bit for bit with rows 1024..1183 of the 4096-env CUDA world after reset, settle and a push."""
    from oracle import b2o
    from robovat_b200.world import World
    cfg, scene, params = helpers.make_inputs(4096, params={'export_debug': 0})
    big = World(params, scene)
    cfg2, scene2, params2 = helpers.make_inputs(160, params={'export_debug': 0})
    params2.env_id_offset = 1024
    cpu = b2o.OracleWorld(params2, scene2, threads=8)
    act = np.random.RandomState(0).uniform(-1, 1, (4096, 4)).astype(np.float32)
    for w, a in ((big, act), (cpu, act[1024:1184])):
        w.reset(seed=21)
        w.settle(0.1, 0.1, 500)
        w.settle()
        w.set_action(a)
        for _ in range(4):
            w.env_substeps(400)
    torch.cuda.synchronize()
    helpers.assert_bits_equal(big.body_state.cpu().numpy()[:, 1024:1184], cpu.body_state, 'body_state')
    helpers.assert_bits_equal(big.joint_state.cpu().numpy()[:, :, 1024:1184], cpu.joint_state, 'joint_state')
    np.testing.assert_array_equal(big.array(_capi.ARR_PHASE).cpu().numpy()[1024:1184], cpu.array(_capi.ARR_PHASE))
    np.testing.assert_array_equal(big.array(_capi.ARR_NUM_STEPS).cpu().numpy()[1024:1184], cpu.array(_capi.ARR_NUM_STEPS))
    helpers.assert_bits_equal(big.observe().cpu().numpy()[1024:1184], cpu.observe(), 'observe')


def test_render_at_2048_envs_against_the_oracle_on_a_slice():
    """BASELINE config #4 at full size (2048 envs, 128 x 128): the oracle renders global envs 1000..1015."""
    from oracle import b2o
    from robovat_b200 import config as config_lib
    from robovat_b200.world import World
    kin = dict(config_lib.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=128, WIDTH=128,
               INTRINSICS=[120.0, 0.0, 64.0, 0, 120.0, 64.0, 0, 0, 1], TRANSLATION=[0.6, 0.0, 1.1])
    cfg, scene, params = helpers.make_inputs(2048, KINECT2={'DEPTH': kin}, TASK_NAME='crossing', LAYOUT_ID=1)
    gpu = World(params, scene, with_camera=True)
    cfg2, scene2, params2 = helpers.make_inputs(16, KINECT2={'DEPTH': kin}, TASK_NAME='crossing', LAYOUT_ID=1)
    params2.env_id_offset = 1000
    cpu = b2o.OracleWorld(params2, scene2, threads=8)
    from robovat_b200.assets import quat_from_euler, quat_to_matrix
    R = quat_to_matrix(quat_from_euler(np.pi, 0.05, 0.1))
    K = np.array([[120.0, 0, 64.0], [0, 120.0, 64.0], [0, 0, 1.0]])
    t = -R.dot(np.array([0.6, 0.0, 1.1]))
    for w in (gpu, cpu):
        w.reset(seed=8)
        w.step(300)
        w.set_camera(K, R, t)
    dg, sg = gpu.render()
    dc, sc = cpu.render()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(sg.cpu().numpy()[1000:1016], sc)
    helpers.assert_bits_equal(dg.cpu().numpy()[1000:1016], dc, 'depth')
    helpers.assert_bits_equal(gpu.point_cloud(seed=4).cpu().numpy()[1000:1016], cpu.point_cloud(seed=4), 'point cloud')
    assert len(np.unique(sc)) >= 5


def _nccl():
    lib = None
    for name in ('libnccl.so.2',):
        try:
            lib = C.CDLL(name, mode=C.RTLD_GLOBAL)
            break
        except OSError:
            pass
    if lib is None:
        import glob
        import site
        for root in site.getsitepackages():
            for path in glob.glob(os.path.join(root, 'nvidia', 'nccl', 'lib', 'libnccl.so*')):
                lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
                break
    return lib


@pytest.mark.parametrize('ranks', [1, 2])
def test_allgather_returns_through_a_raw_nccl_communicator(ranks):
    """b2s_allgather_returns (replaces tools/parallel_run.py's result collection): one ncclAllGather of the episode
    returns.  One process drives `ranks` devices through ncclCommInitAll; with a single GPU only ranks = 1 runs."""
    if torch.cuda.device_count() < ranks:
        pytest.skip('needs %d GPUs' % ranks)
    nccl = _nccl()
    if nccl is None:
        pytest.skip('libnccl not found')
    from robovat_b200.world import World
    comms = (C.c_void_p * ranks)()
    devs = (C.c_int * ranks)(*range(ranks))
    assert nccl.ncclCommInitAll(comms, ranks, devs) == 0
    B = 64
    worlds, outs, want = [], [], []
    for r in range(ranks):
        cfg, scene, params = helpers.make_inputs(B)
        params.env_id_offset = r * B
        with torch.cuda.device(r):
            w = World(params, scene, device=r)
            w.episode_return.copy_(torch.arange(B, dtype=torch.float32, device=w.device) + 1000.0 * r)
            worlds.append(w)
            outs.append(torch.zeros(ranks * B, dtype=torch.float32, device=w.device))
            want.append(np.arange(B, dtype=np.float32) + 1000.0 * r)
    for r in range(ranks):
        torch.cuda.synchronize(r)
    lib = _capi.load()
    assert nccl.ncclGroupStart() == 0
    for r in range(ranks):
        with torch.cuda.device(r):
            st = C.c_void_p(torch.cuda.current_stream(r).cuda_stream)
            _capi.check(lib, lib.b2s_allgather_returns(worlds[r].h, comms[r], C.c_void_p(outs[r].data_ptr()), st))
    assert nccl.ncclGroupEnd() == 0
    for r in range(ranks):
        torch.cuda.synchronize(r)
        np.testing.assert_array_equal(outs[r].cpu().numpy(), np.concatenate(want))
    for r in range(ranks):
        nccl.ncclCommDestroy(C.c_void_p(comms[r]))


def test_first_step_reward_uses_the_settled_reset_observation():
    """PushReward.get_reward compares prev_obs_data with obs_data (push_reward.py:396-405); at the first step of an
    episode prev_obs_data is the observation reset() returned, i.e. the SETTLED scene, not the drop poses.  Also
    obs['num_steps'] counts as the reference does: 0 at reset, 0 after the first step, 1 after the second."""
    from oracle import b2o
    from robovat_b200 import config
    from robovat_b200.envs import PushEnv
    cfg = config.default_push_env_config(TASK_NAME='crossing', LAYOUT_ID=0)
    env = PushEnv(config=cfg, num_envs=16, seed=5)
    obs0 = env.reset()
    assert (np.asarray(obs0['num_steps']) == 0).all()
    act = np.tile(np.array([0.0, 0.0, 0.5, 0.5], np.float32), (16, 1))
    obs1, r1, done1, _ = env.step(act)
    scene = config.build_scene(cfg)
    params = config.build_params(cfg, scene, num_envs=16)
    cpu = b2o.OracleWorld(params, scene)
    rc, tc = cpu.reward(np.asarray(obs0['position'])[..., :2], np.asarray(obs1['position'])[..., :2])
    helpers.assert_bits_equal(np.asarray(r1, np.float32), rc, 'first-step reward')
    assert (np.asarray(obs1['num_steps']) == 0).all()
    env._done[:] = False
    obs2, r2, _, _ = env.step(act)
    assert (np.asarray(obs2['num_steps']) == 1).all()
    rc2, _ = cpu.reward(np.asarray(obs1['position'])[..., :2], np.asarray(obs2['position'])[..., :2])
    helpers.assert_bits_equal(np.asarray(r2, np.float32), rc2, 'second-step reward')
    env.close()


def test_masked_reset_leaves_the_other_envs_untouched():
    """PushEnv.reset(mask): only the masked envs are re-sampled and settled; the others keep their state and their
    substep counters (the reference has one world per env)."""
    from robovat_b200.envs import PushEnv
    env = PushEnv(num_envs=12, seed=2)
    env.reset()
    w = env.world
    before = w.body_state.clone()
    steps = w.array(_capi.ARR_NUM_STEPS).clone()
    mask = np.arange(12) % 3 == 0
    env.reset(mask=mask)
    torch.cuda.synchronize()
    keep = torch.from_numpy(~mask).to(before.device)
    assert torch.equal(w.body_state[:, keep], before[:, keep])
    assert torch.equal(w.array(_capi.ARR_NUM_STEPS)[keep], steps[keep])
    assert not torch.equal(w.body_state[:, ~keep], before[:, ~keep])
    env.close()


def test_two_worlds_on_one_device_on_different_streams():
    """The world description of a launch sits in one __constant__ symbol per device: launches of two worlds from two
    streams must not see each other's description."""
    from robovat_b200.world import World
    cfg_a, scene_a, params_a = helpers.make_inputs(40)
    cfg_b, scene_b, params_b = helpers.make_inputs(24, MIN_MOVABLE_BODIES=1, MAX_MOVABLE_BODIES=5)

    def run(interleaved):
        wa, wb = World(params_a, scene_a), World(params_b, scene_b)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        with torch.cuda.stream(sa):
            wa.reset(seed=1)
        with torch.cuda.stream(sb):
            wb.reset(seed=2)
        if interleaved:
            for _ in range(40):
                with torch.cuda.stream(sa):
                    wa.step(5)
                with torch.cuda.stream(sb):
                    wb.step(5)
        else:
            with torch.cuda.stream(sa):
                wa.step(200)
            torch.cuda.synchronize()
            with torch.cuda.stream(sb):
                wb.step(200)
        torch.cuda.synchronize()
        out = wa.body_state.cpu().numpy().copy(), wb.body_state.cpu().numpy().copy()
        wa.close(); wb.close()
        return out
    a1, b1 = run(True)
    a2, b2 = run(False)
    helpers.assert_bits_equal(a1, a2, 'world A')
    helpers.assert_bits_equal(b1, b2, 'world B')
