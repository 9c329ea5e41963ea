"""CUDA path vs CPU oracle on the same seeded inputs, through the C-ABI (needs a GPU)."""
import numpy as np
import pytest
import torch

from robovat_b200 import _capi
from tests import helpers

pytestmark = pytest.mark.gpu


def _compare_state(gpu, cpu, what, exact=True):
    torch.cuda.synchronize()
    g = gpu.body_state.cpu().numpy()
    c = cpu.body_state
    if exact:
        helpers.assert_bits_equal(g, c, what + ': body_state')
        helpers.assert_bits_equal(gpu.joint_state.cpu().numpy(), cpu.joint_state, what + ': joint_state')
    else:
        np.testing.assert_allclose(g, c, atol=1e-5, err_msg=what)


def _compare_contacts(gpu, cpu, what):
    B, M = gpu.B, gpu.params.max_manifolds
    gk = helpers.manifold_view(gpu.array(_capi.ARR_MANIFOLD_KEYS).cpu().numpy(), gpu.array(_capi.ARR_MANIFOLD_NPTS).cpu().numpy(),
                               gpu.array(_capi.ARR_MANIFOLD_PTS).cpu().numpy(), B, M)
    ck = helpers.manifold_view(cpu.array(_capi.ARR_MANIFOLD_KEYS), cpu.array(_capi.ARR_MANIFOLD_NPTS),
                               cpu.array(_capi.ARR_MANIFOLD_PTS), B, M)
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_PAIRS).cpu().numpy(), cpu.array(_capi.ARR_NUM_PAIRS), err_msg=what + ': num_pairs')
    npairs = cpu.array(_capi.ARR_NUM_PAIRS)
    gp = gpu.array(_capi.ARR_PAIR_KEYS).cpu().numpy().reshape(B, -1)
    cp = cpu.array(_capi.ARR_PAIR_KEYS).reshape(B, -1)
    for e in range(B):
        np.testing.assert_array_equal(gp[e, :npairs[e]], cp[e, :npairs[e]], err_msg='%s: pair keys env %d' % (what, e))
    np.testing.assert_array_equal(gk[0], ck[0], err_msg=what + ': manifold keys')
    np.testing.assert_array_equal(gk[1], ck[1], err_msg=what + ': manifold npts')
    helpers.assert_bits_equal(gk[2], ck[2], what + ': manifold points')
    np.testing.assert_array_equal(gk[3], ck[3], err_msg=what + ': GJK simplex cache')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_CONTACT_FLAGS).cpu().numpy(), cpu.array(_capi.ARR_CONTACT_FLAGS))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_SOLVER_STATS).cpu().numpy(), cpu.array(_capi.ARR_SOLVER_STATS))
    np.testing.assert_array_equal(gpu.array(_capi.ARR_ERROR_FLAGS).cpu().numpy(), cpu.array(_capi.ARR_ERROR_FLAGS))


def test_reset_is_bit_exact():
    _, gpu, cpu = helpers.make_pair(64)
    gpu.reset(seed=3); cpu.reset(seed=3)
    _compare_state(gpu, cpu, 'reset')
    helpers.assert_bits_equal(gpu.array(_capi.ARR_MOV_PARAMS).cpu().numpy(), cpu.array(_capi.ARR_MOV_PARAMS), 'mov_params')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_COL_HULL).cpu().numpy(), cpu.array(_capi.ARR_COL_HULL))
    helpers.assert_bits_equal(gpu.array(_capi.ARR_LINK_POSES).cpu().numpy(), cpu.array(_capi.ARR_LINK_POSES), 'link poses')


@pytest.mark.parametrize('dt', [1e-3, 1.0 / 240.0])
def test_drop_and_settle_substeps_bit_exact(dt):
    """Free fall, first impacts (EPA), resting contact: 600 raw substeps, checked every 50."""
    cfg, gpu, cpu = helpers.make_pair(32, SIM={'TIME_STEP': dt, 'ARM': {'CONFIG': 'sawyer'},
                                                'GROUND': {'POSE': [[0, 0, -0.9], [0, 0, 0]]},
                                                'TABLE': {'POSE': [[0.6, 0.0, 0.0], [0, 0, 0]], 'THICKNESS': 0.05, 'FRICTION': 1.0},
                                                'WALL': {'USE': False}, 'TILE': {'HEIGHT': 0.05, 'COLLIDE': True},
                                                'STEPS_CHECK': 20, 'MAX_PHASE_STEPS': 3000, 'MAX_MOTION_STEPS': 4000,
                                                'MAX_OFFSTAGE_STEPS': 4000})
    gpu.reset(seed=11); cpu.reset(seed=11)
    for k in range(12):
        gpu.step(50); cpu.step(50)
        _compare_state(gpu, cpu, 'dt=%g after %d substeps' % (dt, 50 * (k + 1)))
        _compare_contacts(gpu, cpu, 'dt=%g after %d substeps' % (dt, 50 * (k + 1)))
    assert int(cpu.array(_capi.ARR_NUM_MANIFOLDS).min()) >= 1


def test_settle_matches():
    _, gpu, cpu = helpers.make_pair(48)
    gpu.reset(seed=5); cpu.reset(seed=5)
    gpu.settle(0.1, 0.1, 500); cpu.settle(0.1, 0.1, 500)
    gpu.settle(); cpu.settle()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_STEPS).cpu().numpy(), cpu.array(_capi.ARR_NUM_STEPS))
    _compare_state(gpu, cpu, 'settle')


def test_push_action_phase_machine_bit_exact():
    """A full PushEnv action (7 phases + wait_until_stable) with the arm pushing body 0."""
    cfg, gpu, cpu = helpers.make_pair(32)
    gpu.reset(seed=1); cpu.reset(seed=1)
    gpu.settle(0.1, 0.1, 500); cpu.settle(0.1, 0.1, 500)
    gpu.settle(); cpu.settle()
    pos0 = cpu.observe().copy()
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    act = np.zeros((32, 4), np.float32)
    for e in range(32):
        act[e, :2] = np.clip((pos0[e, 0, :2] - [0.08, 0.0] - off) / rng, -1, 1)
        act[e, 2:] = [1.0, 0.0]
    gpu.set_action(act); cpu.set_action(act)
    total = 0
    while total < 40000:
        ug = gpu.env_substeps(250); uc = cpu.env_substeps(250)
        total += 250
        np.testing.assert_array_equal(gpu.array(_capi.ARR_PHASE).cpu().numpy(), cpu.array(_capi.ARR_PHASE), err_msg='phase after %d' % total)
        assert ug == uc
        if uc == 0:
            break
    assert uc == 0
    _compare_state(gpu, cpu, 'after action')
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_STEPS).cpu().numpy(), cpu.array(_capi.ARR_NUM_STEPS))
    np.testing.assert_array_equal(gpu.is_safe.cpu().numpy(), cpu.array('is_safe'))
    np.testing.assert_array_equal(gpu.is_effective.cpu().numpy(), cpu.array('is_effective'))
    helpers.assert_bits_equal(gpu.array(_capi.ARR_STATUS).cpu().numpy(), cpu.array(_capi.ARR_STATUS), 'status')
    moved = np.linalg.norm(cpu.observe()[:, 0, :2] - pos0[:, 0, :2], axis=1)
    assert (moved > 0.02).sum() >= 8, moved


def test_stacked_movables_couple_dynamic_bodies_bit_exact():
    """Movables piled on each other: contacts between two DYNAMIC bodies (both lanes of the body-centric
    solve exchange velocities), checked every 25 substeps while the pile collapses and settles."""
    _, gpu, cpu = helpers.make_pair(32)
    gpu.reset(seed=7); cpu.reset(seed=7)
    gpu.settle(0.1, 0.1, 500); cpu.settle(0.1, 0.1, 500)
    st = np.array(cpu.body_state)                      # [13][B][Nmax]
    nm = cpu.array('num_movables')
    rs = np.random.RandomState(3)
    for e in range(32):
        for i in range(1, int(nm[e])):                 # body i goes 6 cm above body i-1, slightly off centre
            st[0:2, e, i] = st[0:2, e, 0] + rs.uniform(-0.015, 0.015, size=2)
            st[2, e, i] = st[2, e, 0] + 0.06 * i
            st[7:13, e, i] = 0.0
    cpu.body_state[...] = st
    gpu.body_state.copy_(torch.from_numpy(st))
    coupled = 0
    for k in range(16):
        gpu.step(25); cpu.step(25)
        _compare_state(gpu, cpu, 'pile after %d substeps' % (25 * (k + 1)))
        _compare_contacts(gpu, cpu, 'pile after %d substeps' % (25 * (k + 1)))
        keys = cpu.array(_capi.ARR_MANIFOLD_KEYS).reshape(32, -1)
        slots = cpu.array(_capi.ARR_COL_SLOT).reshape(32, -1)
        first_movable = gpu.params.max_colliders - gpu.params.max_movables     # one hull per movable in this config
        for e in range(32):
            for key in keys[e][keys[e] >= 0]:
                if (key >> 16) >= first_movable and (key & 0xffff) >= first_movable:
                    coupled += 1
        del slots
    assert coupled > 100, coupled


def test_ik_fk_bit_exact():
    cfg, gpu, cpu = helpers.make_pair(16)
    gpu.reset(seed=0); cpu.reset(seed=0)
    from robovat_b200.assets import quat_from_euler
    rs = np.random.RandomState(0)
    pose = np.zeros((16, 7), np.float32)
    pose[:, 0] = rs.uniform(0.4, 0.8, 16); pose[:, 1] = rs.uniform(-0.3, 0.3, 16); pose[:, 2] = rs.uniform(0.14, 0.4, 16)
    pose[:, 3:] = quat_from_euler(np.pi, 0, 0)
    q0 = np.tile(np.array([0.0, -1.18, 0.0, 2.18, 0.0, 0.57, 3.3161], np.float32)[:, None], (1, 16))
    qg = gpu.inverse_kinematics(pose, q0).cpu().numpy()
    qc = cpu.inverse_kinematics(pose, q0)
    helpers.assert_bits_equal(qg, qc, 'ik')
    gpu.joint_state[0].copy_(torch.from_numpy(qc)); cpu.joint_state[0] = qc
    helpers.assert_bits_equal(gpu.forward_kinematics().cpu().numpy(), cpu.forward_kinematics(), 'fk')


def test_crossing_task_concave_movables_bit_exact():
    """BASELINE config #3: TASK_NAME='crossing' LAYOUT_ID=0, 8 concave (multi-hull) movables, tiles as static bodies.
    More than 32 contact points per env: exercises the shared-memory-row solver path."""
    cfg, gpu, cpu = helpers.make_pair(12, TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave',
                                      MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8)
    assert gpu.params.max_contacts > 32
    gpu.reset(seed=2); cpu.reset(seed=2)
    for k in range(6):
        gpu.step(100); cpu.step(100)
        _compare_state(gpu, cpu, 'crossing after %d substeps' % (100 * (k + 1)))
        _compare_contacts(gpu, cpu, 'crossing after %d substeps' % (100 * (k + 1)))
    assert int(cpu.array(_capi.ARR_SOLVER_STATS).reshape(12, 4)[:, 3].max()) > 32
    act = np.tile(np.array([0.0, 0.1, 0.8, -0.6], np.float32), (12, 1))
    gpu.set_action(act); cpu.set_action(act)
    for _ in range(40):
        ug, uc = gpu.env_substeps(500), cpu.env_substeps(500)
        assert ug == uc
        if uc == 0:
            break
    _compare_state(gpu, cpu, 'crossing after action')
    rg, tg = gpu.reward(); rc, tc = cpu.reward()
    helpers.assert_bits_equal(rg.cpu().numpy(), rc, 'reward')
    np.testing.assert_array_equal(tg.cpu().numpy(), tc)


def test_vhacd_urdf_movables_bit_exact():
    """Movables loaded from the URDF files of the reference asset pipeline (V-HACD hulls, inertial frame = the
    reference's area-weighted centroid): drop, settle and one push, CUDA == oracle."""
    cfg, gpu, cpu = helpers.make_pair(16, MOVABLE_NAME='vhacd', MIN_MOVABLE_BODIES=3, MAX_MOVABLE_BODIES=3)
    gpu.reset(seed=6); cpu.reset(seed=6)
    for k in range(5):
        gpu.step(100); cpu.step(100)
        _compare_state(gpu, cpu, 'vhacd after %d substeps' % (100 * (k + 1)))
        _compare_contacts(gpu, cpu, 'vhacd after %d substeps' % (100 * (k + 1)))
    gpu.settle(); cpu.settle()
    act = np.tile(np.array([0.0, -0.1, 0.7, 0.7], np.float32), (16, 1))
    gpu.set_action(act); cpu.set_action(act)
    for _ in range(40):
        ug, uc = gpu.env_substeps(500), cpu.env_substeps(500)
        assert ug == uc
        if uc == 0:
            break
    _compare_state(gpu, cpu, 'vhacd after action')


def test_render_and_point_cloud_bit_exact():
    """BASELINE config #4 shape: 128x128 depth + segmentation, then the segmented point cloud."""
    from robovat_b200 import config as config_lib
    from robovat_b200.assets import quat_from_euler, quat_to_matrix
    kin = dict(config_lib.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=128, WIDTH=128,
               INTRINSICS=[120.0, 0.5, 64.0, 0, 118.0, 63.0, 0, 0, 1])
    cfg, gpu, cpu = helpers.make_pair(6, with_camera=True, KINECT2={'DEPTH': kin}, TASK_NAME='crossing', LAYOUT_ID=1)
    gpu.reset(seed=8); cpu.reset(seed=8)
    gpu.step(400); cpu.step(400)
    rs = np.random.RandomState(0)
    K = np.tile(np.array([120.0, 0.5, 64.0, 0, 118.0, 63.0, 0, 0, 1.0]), (6, 1)) + rs.uniform(-1, 1, (6, 9)) * [1, 0, 1, 0, 1, 1, 0, 0, 0]
    R = np.stack([quat_to_matrix(quat_from_euler(np.pi + rs.uniform(-0.1, 0.1), rs.uniform(-0.1, 0.1), rs.uniform(-0.2, 0.2))) for _ in range(6)])
    t = np.stack([-R[i].dot(np.array([0.6, 0.0, 1.1]) + rs.uniform(-0.05, 0.05, 3)) for i in range(6)])
    gpu.set_camera(K, R.reshape(6, 9), t, per_env=True); cpu.set_camera(K, R.reshape(6, 9), t, per_env=True)
    dg, sg = gpu.render(); dc, sc = cpu.render()
    np.testing.assert_array_equal(sg.cpu().numpy(), sc)
    helpers.assert_bits_equal(dg.cpu().numpy(), dc, 'depth')
    assert len(np.unique(sc)) >= 6
    helpers.assert_bits_equal(gpu.point_cloud(seed=4).cpu().numpy(), cpu.point_cloud(seed=4), 'point cloud')


def test_physics_plugin_calls_match():
    """The BulletPhysics-compatible seam (robovat_b200.physics.CudaPhysics) over the CUDA world and over the oracle:
    the same scripted sequence of add_body / set_body_dynamics / step / position_control_array / IK / contact queries."""
    from robovat_b200.physics import CudaPhysics, EE_LINK_INDEX
    from robovat_b200.assets import quat_from_euler
    cfg, gpu, cpu = helpers.make_pair(1)
    outs = []
    for world in (gpu, cpu):
        ph = CudaPhysics(world=world, scene=world.scene, time_step=cfg.SIM.TIME_STEP)
        ph.reset(); ph.set_gravity([0, 0, -9.8]); ph.start()
        ground = ph.add_body('/assets/scene/ground.urdf', [[0, 0, -0.9], [0, 0, 0]], is_static=True)
        table = ph.add_body('/assets/scene/table.urdf', [[0.6, 0, 0.0], [0, 0, 0]], is_static=True)
        box = ph.add_body('/assets/movables/box.urdf', [[0.6, 0.0, 0.1], [0.2, 0.1, 0.3]], scale=1.1)
        ph.set_body_dynamics(box, mass=0.2, lateral_friction=0.7)
        arm = ph.add_body('/assets/robots/sawyer_arm.urdf', [[0, 0, 0], [0, 0, 0]], is_static=True)
        for j, q in enumerate([0.0, -1.18, 0.0, 2.18, 0.0, 0.57, 3.3161]):
            ph.set_joint_position((arm, j), q)
        pose = [[0.6, 0.0, 0.3], list(quat_from_euler(np.pi, 0, 0))]
        trace = []
        for s in range(1200):
            if s % 10 == 0:
                q = ph.compute_inverse_kinematics((arm, EE_LINK_INDEX), pose)[:7]
            ph.position_control_array(arm, range(7), q, [0.0] * 7)
            ph.step()
            trace.append([float(ph.get_joint_position((arm, j))) for j in range(7)] + list(ph.get_body_position(box)))
        assert ph.num_steps == 1200 and len(ph.get_contact_points(box, table)) >= 1 and not ph.get_contact_points(arm, table)
        ee = ph.get_link_pose((arm, EE_LINK_INDEX))
        assert np.abs(np.asarray(ee.position) - [0.6, 0.0, 0.3]).max() < 5e-3
        outs.append(np.array(trace, np.float32))
    helpers.assert_bits_equal(outs[0], outs[1], 'plugin trace')


def test_episode_driver_on_the_cuda_env():
    """robovat_b200.episodes over the CUDA PushEnv: a single env with the reference-identical HeuristicPushPolicy and a
    batch with the vectorised policy."""
    from robovat_b200 import config, episodes, policies
    from robovat_b200.envs import PushEnv
    cfg = config.default_push_env_config()
    env = PushEnv(config=cfg, num_envs=1, seed=3)
    ep = episodes.collect(env, policies.HeuristicPushPolicy(env), num_steps=2).episodes()[0]
    assert 1 <= len(ep['transitions']) <= 2
    t0 = ep['transitions'][0]
    assert np.asarray(t0['action']).shape[-1] == 4 and np.isfinite(t0['reward'])
    assert np.asarray(t0['state']['position']).shape == (int(cfg.MAX_MOVABLE_BODIES), 3)
    env.close()
    env = PushEnv(config=cfg, num_envs=8, seed=3)
    batch = episodes.collect(env, policies.BatchedHeuristicPolicy(seed=1), num_steps=2)
    eps = batch.episodes()
    assert len(eps) == 8 and all(1 <= len(e['transitions']) <= 2 for e in eps)
    assert np.isfinite(batch.returns).all()
    assert eps[0]['transitions'][0]['state']['position'].shape == (int(cfg.MAX_MOVABLE_BODIES), 3)
    env.close()


def test_ragged_movable_counts_and_partial_reset_bit_exact():
    """Edge cases of the batch: environments with 1..5 movables (zero padded slots), a masked reset of some
    environments mid-run, and environments whose action never started (idle) next to running ones."""
    cfg, gpu, cpu = helpers.make_pair(24, MIN_MOVABLE_BODIES=1, MAX_MOVABLE_BODIES=5)
    gpu.reset(seed=9); cpu.reset(seed=9)
    nm = cpu.array('num_movables')
    assert nm.min() < nm.max() and nm.min() >= 1 and nm.max() <= 5
    np.testing.assert_array_equal(gpu.num_movables.cpu().numpy(), nm)
    gpu.settle(0.1, 0.1, 500); cpu.settle(0.1, 0.1, 500)
    _compare_state(gpu, cpu, 'ragged settle')
    mask = (np.arange(24) % 3 == 0)
    gpu.reset(seed=10, mask=mask); cpu.reset(seed=10, mask=mask)
    gpu.step(300); cpu.step(300)
    _compare_state(gpu, cpu, 'ragged after masked reset')
    _compare_contacts(gpu, cpu, 'ragged after masked reset')
    act = np.tile(np.array([0.1, 0.0, -0.8, 0.5], np.float32), (24, 1))
    gpu.set_action(act); cpu.set_action(act)
    idle = np.arange(24) % 4 == 0                      # these environments do not execute the action
    ph_g = gpu.array(_capi.ARR_PHASE); ph_c = cpu.array(_capi.ARR_PHASE)
    ph_g[torch.from_numpy(idle).to(ph_g.device)] = _capi.PHASE_IDLE
    ph_c[idle] = _capi.PHASE_IDLE
    before = np.array(cpu.array(_capi.ARR_NUM_STEPS))
    for _ in range(4):
        ug, uc = gpu.env_substeps(300), cpu.env_substeps(300)
        assert ug == uc
    _compare_state(gpu, cpu, 'ragged mid action')
    after = cpu.array(_capi.ARR_NUM_STEPS)
    assert (after[idle] == before[idle]).all() and (after[~idle] > before[~idle]).all()
    np.testing.assert_array_equal(gpu.array(_capi.ARR_NUM_STEPS).cpu().numpy(), after)


def test_capacity_overflow_is_flagged_identically():
    """Maximum sizes: a pair / manifold capacity that is too small raises the same error flags on both sides and the
    surviving contacts still agree."""
    cfg, gpu, cpu = helpers.make_pair(8, TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave',
                                      MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8, params={'max_pairs': 12, 'max_manifolds': 8})
    gpu.reset(seed=2); cpu.reset(seed=2)
    for k in range(3):
        gpu.step(100); cpu.step(100)
        np.testing.assert_array_equal(gpu.array(_capi.ARR_ERROR_FLAGS).cpu().numpy(), cpu.array(_capi.ARR_ERROR_FLAGS))
        _compare_contacts(gpu, cpu, 'overflow after %d substeps' % (100 * (k + 1)))
    assert int(cpu.array(_capi.ARR_ERROR_FLAGS).max()) & 3


def test_full_size_batch_is_shard_and_schedule_independent():
    """BASELINE's full size (4096 envs) through size-independent properties: the state of global environment i does
    not depend on which rank / block / warp steps it.  A 4096-env world and a 160-env world that owns the global
    ids 1024..1183 (env_id_offset) must agree bit for bit on those environments after reset, settle and a push --
    the two worlds deal their environments to blocks completely differently -- and a second 4096-env run must
    reproduce the first one exactly (determinism)."""
    from robovat_b200.world import World
    cfg, scene, params = helpers.make_inputs(4096)
    big = World(params, scene)
    cfg2, scene2, params2 = helpers.make_inputs(160)
    params2.env_id_offset = 1024
    small = World(params2, scene2)
    rs = np.random.RandomState(0)
    act = rs.uniform(-1, 1, (4096, 4)).astype(np.float32)

    def run(w, a):
        w.reset(seed=21)
        w.settle(0.1, 0.1, 500)
        w.settle()
        w.set_action(a)
        for _ in range(3):
            w.env_substeps(400)
        torch.cuda.synchronize()
        return (w.body_state.cpu().numpy().copy(), w.joint_state.cpu().numpy().copy(),
                w.array(_capi.ARR_PHASE).cpu().numpy().copy(), w.array(_capi.ARR_NUM_STEPS).cpu().numpy().copy())
    b1 = run(big, act)
    s1 = run(small, act[1024:1184])
    helpers.assert_bits_equal(b1[0][:, 1024:1184], s1[0], 'shard: body_state')
    helpers.assert_bits_equal(b1[1][:, :, 1024:1184], s1[1], 'shard: joint_state')
    np.testing.assert_array_equal(b1[2][1024:1184], s1[2])
    np.testing.assert_array_equal(b1[3][1024:1184], s1[3])
    big2 = World(params, scene)                          # a fresh world: the reset count is part of the RNG key
    b2 = run(big2, act)
    helpers.assert_bits_equal(b1[0], b2[0], 'determinism: body_state')
    np.testing.assert_array_equal(b1[3], b2[3])
    assert np.isfinite(b1[0]).all() and int(big.array(_capi.ARR_ERROR_FLAGS).max().item()) == 0
    assert (b1[0][2] > -0.95).all()                      # a pushed body may leave the table, nothing falls through the ground (z = -0.9)
    assert len(np.unique(b1[2])) >= 3                    # environments are spread over several phases


@pytest.mark.parametrize('roll,spin', [(0.0, 0.001), (0.002, 0.005)])
@pytest.mark.parametrize('big', [False, True])
def test_torsional_friction_options_bit_exact(roll, spin, big):
    """Rolling / spinning friction rows (B2SParams.rolling_friction / spinning_friction): rows off (rolling = 0, Bullet's
    gate) and a spinning coefficient that differs from the rolling one, on the register-resident solve (3 convex
    movables) and on the record-based solve of large scenes (8 concave movables): drop, impacts, rest."""
    from robovat_b200 import config
    phys = dict(config.DEFAULT_PUSH_ENV['PHYSICS'], ROLLING_FRICTION=roll, SPINNING_FRICTION=spin)
    kw = dict(PHYSICS=phys)
    if big:
        kw.update(TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave', MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8)
    cfg, gpu, cpu = helpers.make_pair(12 if big else 24, **kw)
    assert abs(gpu.params.rolling_friction - roll) < 1e-9 and abs(gpu.params.spinning_friction - spin) < 1e-9
    gpu.reset(seed=21); cpu.reset(seed=21)
    for k in range(6):
        gpu.step(60); cpu.step(60)
        _compare_state(gpu, cpu, 'roll=%g spin=%g big=%d after %d substeps' % (roll, spin, big, 60 * (k + 1)))
        _compare_contacts(gpu, cpu, 'roll=%g spin=%g big=%d after %d substeps' % (roll, spin, big, 60 * (k + 1)))
