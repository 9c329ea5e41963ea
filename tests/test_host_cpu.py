"""CPU suite: C-ABI surface, host logic, physical invariants of the oracle, env sharding over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from robovat_b200 import _capi, assets, config
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- C-ABI ----------------------------

def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'b2s.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b2s_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    """libb2s.so loads without a GPU and exports exactly what include/b2s.h declares (no compute calls here)."""
    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_capi.SYMBOLS), set(declared) ^ set(_capi.SYMBOLS)
    assert lib.b2s_version() == 100
    for which, struct in enumerate((_capi.B2SParams, _capi.B2SSceneDesc, _capi.B2SBuffers)):
        assert lib.b2s_sizeof(which) == C.sizeof(struct)


def test_error_convention_without_gpu():
    """Bad arguments give negative codes + a message; without a CUDA device b2s_create fails loudly (no CPU fallback)."""
    lib = _capi.load()
    p = _capi.B2SParams()
    assert lib.b2s_default_params(p) == 0
    assert p.solver_iterations == 50 and abs(p.time_step - 1e-3) < 1e-12 and p.ik_interval == 10
    h = C.c_void_p()
    assert lib.b2s_create(None, 0, C.byref(h)) == _capi.E_INVALID
    assert b'NULL' in lib.b2s_last_error()
    assert lib.b2s_create(C.byref(p), 0, C.byref(h)) == _capi.E_INVALID          # sizes are still zero
    import torch
    if not torch.cuda.is_available():
        p.num_envs, p.max_movables, p.max_pairs, p.max_manifolds, p.max_contacts, p.max_colliders = 4, 3, 64, 32, 32, 16
        assert lib.b2s_create(C.byref(p), 0, C.byref(h)) == _capi.E_CUDA
        assert b'no CPU fallback' in lib.b2s_last_error()
        from robovat_b200.world import World
        with pytest.raises(RuntimeError):
            World(p, None)
    assert lib.b2s_step(None, 1, None) == _capi.E_INVALID


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'robovat_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                # comments may cite oracle files; code may not include, import, load or call them
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert not re.search(r'#include\s+"[^"]*oracle', text) and 'libb2o' not in text, f
                assert not re.search(r'\bb2o_[a-z0-9_]+\s*\(', text), f


# ---------------------------------------------------------------- host logic -------------------------

def test_scene_flattening_and_capacities():
    cfg = config.default_push_env_config()
    scene = config.build_scene(cfg)
    lib = scene.lib
    assert max(lib.hull_vert_cnt) <= assets.MAX_HULL_VERTS
    names = [s['name'] for s in scene.statics]
    assert names[:2] == ['ground', 'table']                      # reference uid order (arm_env.py:84-93)
    p = config.build_params(cfg, scene, num_envs=8)
    assert p.max_colliders == scene.fixed_colliders + 3 and p.max_contacts == 32
    # crossing layout 0: 12 region tiles + 1 goal tile as static bodies (push_env.py:343-357)
    cfg3 = config.default_push_env_config(TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave',
                                          MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8)
    scene3 = config.build_scene(cfg3)
    assert sum(1 for s in scene3.statics if s['name'].startswith('tile')) == 13
    assert scene3.desc.num_region == 12 and scene3.desc.num_goal == 1 and scene3.desc.num_obstacle == 18
    tile = [s for s in scene3.statics if s['name'].startswith('tile')][0]
    assert abs(tile['pose'][2] - (0.001 - 0.025)) < 1e-9
    assert scene3.max_movable_hulls == 3


def test_concave_assets_are_centred_on_their_com():
    lib = assets.AssetLibrary()
    for name, hulls in assets.concave_movables().items():
        aid = lib.add_asset(name, hulls)
        vols, cens = [], []
        for h in range(lib.asset_hull_off[aid], lib.asset_hull_off[aid] + lib.asset_hull_cnt[aid]):
            pts = np.array(lib.verts[lib.hull_vert_off[h]:lib.hull_vert_off[h] + lib.hull_vert_cnt[h]])
            v, c = assets._hull_volume_centroid(pts)
            vols.append(v); cens.append(c)
        com = np.average(np.array(cens), axis=0, weights=np.array(vols))
        assert np.abs(com).max() < 1e-9, name


# ---------------------------------------------------------------- oracle physics invariants -----------

def test_resting_box_stays_at_rest():
    cfg, w = helpers.make_oracle(16)
    w.reset(seed=4)
    w.settle(0.1, 0.1, 500)
    w.settle()
    z0 = w.body_state[2].copy()
    xy0 = w.body_state[0:2].copy()
    w.step(1000)
    # wait_until_stable counts stable checks cumulatively (simulator.py:366-372), so a prism lying on an edge may
    # still rock slightly; it must stay tiny and must not drift
    assert np.abs(w.body_state[7:10]).max() < 2e-3 and np.abs(w.body_state[10:13]).max() < 3e-2
    assert np.abs(w.body_state[2] - z0).max() < 5e-4 and np.abs(w.body_state[0:2] - xy0).max() < 2e-3
    assert (w.body_state[2] > 0.015).all()                     # nothing sank into the table (top at z = 0)
    assert (w.array(_capi.ARR_ERROR_FLAGS) == 0).all()


def test_free_fall_matches_closed_form():
    cfg, w = helpers.make_oracle(4, PHYSICS=dict(config.DEFAULT_PUSH_ENV['PHYSICS'], LINEAR_DAMPING=0.0, ANGULAR_DAMPING=0.0))
    w.reset(seed=1)
    z0 = w.body_state[2].copy()
    n, dt = 100, 1e-3
    w.step(n)
    # semi-implicit Euler: z_n = z_0 + g dt^2 n (n + 1) / 2
    np.testing.assert_allclose(w.body_state[2], z0 - 9.8 * dt * dt * n * (n + 1) / 2, atol=2e-6)
    np.testing.assert_allclose(w.body_state[9], -9.8 * dt * n, rtol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(w.body_state[3:7], axis=0), 1.0, atol=1e-6)


def test_friction_cone_and_nonnegative_normal_impulses():
    cfg, w = helpers.make_oracle(32)
    w.reset(seed=9)
    w.step(400)                                                  # impacts + sliding
    M = w.params.max_manifolds
    pts = w.array(_capi.ARR_MANIFOLD_PTS).reshape(32, M, 4, _capi.CP_FLOATS)
    npts = w.array(_capi.ARR_MANIFOLD_NPTS).reshape(32, M)
    live = np.arange(4)[None, None, :] < npts[:, :, None]
    lam_n, lam_t = pts[..., 10][live], pts[..., 11:13][live]
    assert live.sum() > 50 and (lam_n >= 0).all()
    # pyramid friction with mu <= 1.0 * 1.0: |lambda_t| <= mu lambda_n per direction
    assert (np.abs(lam_t) <= lam_n[:, None] * 1.0 + 1e-9).all()


def test_push_moves_the_target_and_reports_flags():
    cfg, w = helpers.make_oracle(16, threads=4)
    w.reset(seed=1); w.settle(0.1, 0.1, 500); w.settle()
    pos0 = w.observe().copy()
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    act = np.zeros((16, 4), np.float32)
    act[:, :2] = np.clip((pos0[:, 0, :2] - [0.08, 0.0] - off) / rng, -1, 1)
    act[:, 2] = 1.0
    w.set_action(act)
    total = 0
    while w.env_substeps(500) > 0 and total < 40000:
        total += 500
    assert (w.array(_capi.ARR_PHASE) == _capi.PHASE_IDLE).all()
    moved = np.linalg.norm(w.observe()[:, 0, :2] - pos0[:, 0, :2], axis=1)
    eff = w.array('is_effective').astype(bool)
    assert (moved > 0.05).sum() >= 8
    assert (eff[moved > 0.02]).all()                             # _check_effectiveness (push_env.py:900-923)
    assert (w.array(_capi.ARR_ERROR_FLAGS) == 0).all()


def test_reset_is_independent_of_sharding():
    """Philox is keyed by the global env id: two half-size shards give the rows of the full world."""
    cfg, full = helpers.make_oracle(8)
    full.reset(seed=21)
    halves = []
    for r in range(2):
        _, w = helpers.make_oracle(4, params={'env_id_offset': 4 * r})
        w.reset(seed=21)
        halves.append(w.body_state.copy())
    helpers.assert_bits_equal(np.concatenate(halves, axis=1), full.body_state, 'sharded reset')


# ---------------------------------------------------------------- multi-process path (gloo) ------------

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from tests import helpers
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%%s' %% os.environ['PORT'], rank=int(os.environ['RANK']), world_size=2)
rank = dist.get_rank()
cfg, w = helpers.make_oracle(4, params={'env_id_offset': 4 * rank}, TASK_NAME='crossing', LAYOUT_ID=0)
w.reset(seed=5); w.settle(0.1, 0.1, 300)
act = np.tile(np.array([0.1, -0.2, 0.5, 0.5], np.float32), (4, 1))
w.set_action(act)
while w.env_substeps(500) > 0:
    pass
w.reward()
local = torch.from_numpy(w.array('episode_return').copy())
out = [torch.zeros(4), torch.zeros(4)]
dist.all_gather(out, local)          # the one collective of the path: episode returns (replaces tools/parallel_run.py)
if rank == 0:
    np.save(os.environ['OUT'], torch.cat(out).numpy())
dist.destroy_process_group()
'''


def test_two_rank_sharding_allgathers_the_same_returns(tmp_path):
    out = str(tmp_path / 'returns.npy')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, PORT=str(29500 + os.getpid() % 2000), OUT=out)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    gathered = np.load(out)
    cfg, w = helpers.make_oracle(8, TASK_NAME='crossing', LAYOUT_ID=0)
    w.reset(seed=5); w.settle(0.1, 0.1, 300)
    w.set_action(np.tile(np.array([0.1, -0.2, 0.5, 0.5], np.float32), (8, 1)))
    while w.env_substeps(500) > 0:
        pass
    w.reward()
    helpers.assert_bits_equal(gathered, w.array('episode_return'), 'all-gathered returns')


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the CUDA path): one JSON line with the contract
    keys on rank 0, silence and exit 0 on every other rank."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--envs', '256', '--substeps', '500']
    env = dict(os.environ, RANK='0', WORLD_SIZE='1', LOCAL_RANK='0')
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'PushEnv substeps/sec at 4096 envs' and line['unit'] == 'substeps/s'
    assert line['higher_is_better'] is True and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'substeps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('PushEnv 4096 batched envs') and line['config']['policy'] == 'B2S_POLICY_AIMED'
    assert line['ms_per_step'] > 0 and line['cpu_baseline']['reference_cpu'].startswith('pybullet ')
    env['RANK'] = '1'
    env['WORLD_SIZE'] = '2'
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env=env, cwd=root)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_bench_cuda_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: the CUDA arm of bench.py refuses to run when there is no device (skipped on a GPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode != 0 and 'no CPU fallback' in out.stderr and out.stdout.strip() == ''
