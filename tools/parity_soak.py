"""Randomised parity soak (needs a GPU): device-side rollouts of the CUDA path against the oracle's rollout over seeds,
tasks, time steps, scene sizes, policies and schedules that the pytest suite does not enumerate.  Every env must
complete the oracle's episodes bit for bit (actions, rewards, flags, substep counts, returns, final state).

usage: python tools/parity_soak.py [cases] [first_seed]        prints one line per case and a summary; exit 1 on a mismatch
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import b2o
from robovat_b200 import _capi, config
from robovat_b200.world import RolloutRecord, World
from tests import helpers

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
threads = os.cpu_count() or 4
bad = 0
for c in range(cases):
    rs = np.random.RandomState(seed0 + c)
    big = rs.rand() < 0.3
    dt = [1e-3, 1.0 / 240.0][int(rs.rand() < 0.7)]
    task = [None, 'clearing', 'insertion', 'crossing'][rs.randint(4)]
    sim = dict(config.DEFAULT_PUSH_ENV['SIM'], TIME_STEP=dt)
    if rs.rand() < 0.3:
        sim['WALL'] = dict(sim['WALL'], USE=True)
    kw = dict(SIM=sim, TASK_NAME=task, LAYOUT_ID=int(rs.randint(3)) if task else 0)
    if big:
        kw.update(TASK_NAME='crossing', LAYOUT_ID=int(rs.randint(3)), MOVABLE_NAME='concave', MIN_MOVABLE_BODIES=6, MAX_MOVABLE_BODIES=8)
    else:
        kw.update(MIN_MOVABLE_BODIES=int(rs.randint(1, 4)), MAX_MOVABLE_BODIES=3)
    phys = dict(config.DEFAULT_PUSH_ENV['PHYSICS'])
    if rs.rand() < 0.3:
        phys.update(ROLLING_FRICTION=float(rs.choice([0.0, 0.003])), SPINNING_FRICTION=float(rs.choice([0.001, 0.004])))
    kw['PHYSICS'] = phys
    B = int(rs.choice([40, 96]) if big else rs.choice([64, 300, 777]))
    params = {'export_debug': 0}
    if rs.rand() < 0.4:
        params['envs_per_block'] = int(rs.choice([3, 5, 9]))
    A, EP = int(rs.randint(1, 4)), int(rs.randint(1, 3))
    free = bool(rs.rand() < 0.75)
    policy = int(rs.choice([_capi.POLICY_AIMED, _capi.POLICY_HEURISTIC])) if hasattr(_capi, 'POLICY_HEURISTIC') else _capi.POLICY_AIMED
    cfg, scene, p = helpers.make_inputs(B, params=params, **kw)
    gpu = World(p, scene)
    cpu = b2o.OracleWorld(p, scene, threads=threads)
    t0 = time.time()
    for w in (gpu, cpu):
        w.reset(seed=seed0 + c); w.settle(0.1, 0.1, 500); w.settle(); w.begin_episode()
    rec = RolloutRecord(B, gpu.N, EP, A, gpu.device)
    gpu.rollout_begin(A, EP, policy_seed=7 + c, reset_seed=11 + c, record=rec, policy_kind=policy, free_running=free)
    ref = cpu.rollout_begin(A, EP, policy_seed=7 + c, reset_seed=11 + c, policy_kind=policy)
    left = gpu.rollout_run(chunk=int(rs.choice([100, 250])), max_substeps=600000)
    while cpu.rollout_run(20000) > 0:
        pass
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in rec.tensors().items()}
    what = 'case %d: %s B=%d dt=%.4g task=%s/%d wall=%d A=%d EP=%d free=%d policy=%d epb=%s roll=%g' % (
        c, 'big' if big else 'small', B, dt, kw['TASK_NAME'], kw['LAYOUT_ID'], int(sim['WALL']['USE']), A, EP, free, policy, params.get('envs_per_block'), phys['ROLLING_FRICTION'])
    try:
        assert left == 0, 'rollout not finished (%d envs left)' % left
        for k in ('lengths', 'flags', 'substeps'):
            np.testing.assert_array_equal(g[k], ref[k], err_msg=k)
        valid = np.arange(A)[None, None, :] < ref['lengths'][:, :, None]
        helpers.assert_bits_equal(np.where(valid[..., None], g['actions'], 0), np.where(valid[..., None], ref['actions'], 0), 'actions')
        helpers.assert_bits_equal(np.where(valid, g['rewards'], 0), np.where(valid, ref['rewards'], 0), 'rewards')
        helpers.assert_bits_equal(g['returns'], ref['returns'], 'returns')
        helpers.assert_bits_equal(gpu.body_state.cpu().numpy(), cpu.body_state, 'final body_state')
        assert gpu.substeps_executed() == cpu.substeps_executed()
        flags = int(gpu.array(_capi.ARR_ERROR_FLAGS).cpu().numpy().max())
        print('%s: OK, %d env-substeps, error flags %d, %.1f s' % (what, cpu.substeps_executed(), flags, time.time() - t0), flush=True)
    except AssertionError as e:
        bad += 1
        print('%s: MISMATCH %s' % (what, str(e)[:300]), flush=True)
    gpu.close(); cpu.close()
print('soak: %d cases, %d mismatches' % (cases, bad))
sys.exit(1 if bad else 0)
