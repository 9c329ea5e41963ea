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
