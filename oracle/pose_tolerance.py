"""The stated fp32 pose tolerance: fp32 path (CUDA world or fp32 oracle) against the double-precision oracle.

TEST INFRASTRUCTURE ONLY (used by tests/ and by bench.py's parity line).

BASELINE.json asks for "body poses within a stated fp32 tolerance after 240 substeps" against PyBullet, which computes
in double.  PyBullet cannot run here, so the tolerance is stated against the same algorithm carried out
in double precision (oracle/libb2o64.so, b2o_f64.h): both worlds start from the SAME state, run 240 substeps at
dt = 1/240 s, and the Euclidean distance of every movable's position is reported.  Three scenarios:

  drop    the reset state: bodies fall 0.2 m with random orientation, hit the table, tumble and come to rest
          (impacts amplify rounding differences: the hard case)
  rest    the settled scene simply stays at rest (rounding-level differences only)
  slide   bodies at rest on the table get a horizontal velocity of 0.4 m/s and a spin of 2 rad/s and slide to rest
          under Coulomb friction (the regime of a push)
"""
import numpy as np

from oracle import b2o
from robovat_b200 import _capi


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)


def _assign(dst, src):
    if hasattr(dst, 'copy_'):
        import torch
        dst.copy_(torch.as_tensor(np.ascontiguousarray(src, np.float32)).reshape(dst.shape))
    else:
        dst[...] = np.asarray(src).reshape(dst.shape)


def _sync(world):
    if hasattr(world.body_state, 'detach'):
        import torch
        torch.cuda.synchronize()


def _copy_state(src32, dst64):
    """fp32 state -> fp64 world: poses, velocities, joints, sampled scale / mass / friction (the asset ids already
    agree: they are integer draws of the same Philox stream)."""
    dst64.body_state[...] = _np(src32.body_state).astype(np.float64)
    dst64.joint_state[...] = _np(src32.joint_state).astype(np.float64)
    mp32 = _np(src32.array(_capi.ARR_MOV_PARAMS)).reshape(4, src32.B, src32.N)
    mp64 = dst64.array(_capi.ARR_MOV_PARAMS).reshape(4, dst64.B, dst64.N)
    mp64[1:] = mp32[1:].astype(np.float64)


def _copy_contacts(src32, dst64):
    """persistent manifolds (contact points + warm-start impulses) fp32 -> fp64; the GJK simplex cache words (integers
    stored in float slots, point 0 words 13..15) are cleared: a cold GJK start finds the same closest features"""
    B, M = src32.B, src32.params.max_manifolds
    keys = _np(src32.array(_capi.ARR_MANIFOLD_KEYS)).reshape(B, M)
    npts = _np(src32.array(_capi.ARR_MANIFOLD_NPTS)).reshape(B, M)
    pts = _np(src32.array(_capi.ARR_MANIFOLD_PTS)).reshape(B, M, 4, _capi.CP_FLOATS).astype(np.float64)
    pts[:, :, :, 13:] = 0.0
    live = np.arange(M)[None, :] < _np(src32.array(_capi.ARR_NUM_MANIFOLDS)).reshape(B, 1)
    dst64.array(_capi.ARR_MANIFOLD_KEYS).reshape(B, M)[...] = np.where(live, keys, -1)
    dst64.array(_capi.ARR_MANIFOLD_NPTS).reshape(B, M)[...] = np.where(live, npts, 0)
    dst64.array(_capi.ARR_MANIFOLD_PTS).reshape(B, M, 4, _capi.CP_FLOATS)[...] = pts * live[:, :, None, None]
    dst64.array(_capi.ARR_NUM_MANIFOLDS)[...] = _np(src32.array(_capi.ARR_NUM_MANIFOLDS)).ravel()


def _l2(w32, w64):
    _sync(w32)
    p32 = _np(w32.body_state)[0:3].astype(np.float64)
    p64 = np.asarray(w64.body_state)[0:3]
    live = np.arange(w64.N)[None, :] < np.asarray(w64.array('num_movables'))[:, None]
    d = np.sqrt(((p32 - p64) ** 2).sum(axis=0))
    return d[live]


def _summary(d, scenario, substeps):
    return {'scenario': scenario, 'substeps': substeps, 'bodies': int(d.size), 'median_m': float(np.median(d)),
            'p90_m': float(np.percentile(d, 90)), 'p99_m': float(np.percentile(d, 99)), 'max_m': float(d.max()),
            'frac_le_1e-3': float((d <= 1e-3).mean()), 'frac_le_1e-4': float((d <= 1e-4).mean())}


def measure(world32, params, scene, seed=0, substeps=240, threads=4):
    """world32: a freshly created fp32 world (robovat_b200.world.World on the GPU, or b2o.OracleWorld) built from
    `params` / `scene`.  Returns the two scenario summaries."""
    w64 = b2o.OracleWorld(params, scene, threads=threads, f64=True)
    out = []
    # ---- drop
    world32.reset(seed=seed)
    w64.reset(seed=seed)
    _copy_state(world32, w64)
    world32.step(substeps)
    w64.step(substeps)
    out.append(_summary(_l2(world32, w64), 'drop', substeps))
    # ---- slide: settle in fp32, hand the rest state and the contact caches over, kick
    world32.settle(0.1, 0.1, 500)
    world32.settle()
    _sync(world32)
    # ---- rest: the settled scene simply stays (rounding-level differences only)
    _copy_state(world32, w64)
    _copy_contacts(world32, w64)
    world32.step(substeps)
    w64.step(substeps)
    out.append(_summary(_l2(world32, w64), 'rest', substeps))
    st = _np(world32.body_state).copy()
    rs = np.random.RandomState(seed)
    B, N = params.num_envs, params.max_movables
    ang = rs.uniform(-np.pi, np.pi, (B, N))
    st[7], st[8], st[9] = 0.4 * np.cos(ang), 0.4 * np.sin(ang), 0.0
    st[10], st[11], st[12] = 0.0, 0.0, rs.choice([-2.0, 2.0], (B, N))
    live = np.arange(N)[None, :] < _np(world32.num_movables)[:, None]
    st[7:13] *= live[None]
    _assign(world32.body_state, st)
    _copy_state(world32, w64)
    _copy_contacts(world32, w64)
    world32.step(substeps)
    w64.step(substeps)
    out.append(_summary(_l2(world32, w64), 'slide', substeps))
    w64.close()
    return out
