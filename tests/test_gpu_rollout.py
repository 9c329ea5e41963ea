"""Episodes on the device (b2s_rollout_*) against the CPU oracle's rollout, through the C-ABI (needs a GPU).

The oracle's rollout is itself pinned to the lock-step calls in tests/test_rollout_cpu.py; here the CUDA kernel that
carries an env from action to action and from episode to episode without the host must reproduce it bit for bit:
actions drawn by the device policy, rewards, observations, flags, Simulator.num_steps, lengths, returns, and the
final world state."""
import numpy as np
import pytest
import torch

from robovat_b200 import _capi
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


@pytest.mark.parametrize('task', [None, 'clearing'])
def test_rollout_matches_oracle_rollout(task):
    """48 envs, three episodes of up to three actions each with scene resets in between; runs of 250 substeps on the
    GPU against runs of 1000 on the CPU (the chunking must not matter)."""
    A, EP, B = 3, 3, 48
    bind = {} if task is None else dict(TASK_NAME=task, LAYOUT_ID=0)
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0}, **bind)
    _prepare(gpu, cpu, seed=4)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=13, reset_seed=6, max_attempts=2000, record=rec)
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


def test_rollout_budget_stops_mid_episode_and_resumes():
    """A substep budget cuts the rollout anywhere (mid action, mid reset); the state then equals the oracle's after the
    same number of substeps per env, and a second run call finishes the episodes."""
    A, EP, B = 2, 2, 16
    cfg, gpu, cpu = helpers.make_pair(B, params={'export_debug': 0})
    _prepare(gpu, cpu, seed=8)
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=1, reset_seed=2, max_attempts=2000, record=rec)
    ref = cpu.rollout_begin(A, EP, policy_seed=1, reset_seed=2, max_attempts=2000)
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
