"""Shared helpers of the parity tests: build the CUDA world and the CPU oracle from one config."""
import numpy as np

from robovat_b200 import _capi, config


def make_inputs(num_envs, **bindings):
    overrides = dict(bindings.pop('params', {}))
    # the parity tests look at the inspection arrays too (pair keys, link poses / velocities of every substep); the
    # product default leaves them out of the substep kernel (params.export_debug = 0, covered by its own test)
    overrides.setdefault('export_debug', 1)
    cfg = config.default_push_env_config(**bindings)
    scene = config.build_scene(cfg)
    params = config.build_params(cfg, scene, num_envs=num_envs, **overrides)
    return cfg, scene, params


def make_oracle(num_envs, threads=1, **bindings):
    from oracle import b2o
    cfg, scene, params = make_inputs(num_envs, **bindings)
    return cfg, b2o.OracleWorld(params, scene, threads=threads)


def make_pair(num_envs, with_camera=False, **bindings):
    from oracle import b2o
    from robovat_b200.world import World
    cfg, scene, params = make_inputs(num_envs, **bindings)
    gpu = World(params, scene, with_camera=with_camera)
    cpu = b2o.OracleWorld(params, scene, threads=4)
    return cfg, gpu, cpu


def manifold_view(keys, npts, pts, B, M):
    """Canonical comparable form: (keys, npts, points masked to the live entries, GJK simplex cache words)."""
    keys = np.asarray(keys).reshape(B, M).copy()
    npts = np.asarray(npts).reshape(B, M).copy()
    pts = np.asarray(pts).reshape(B, M, 4, _capi.CP_FLOATS).copy()
    live = np.arange(4)[None, None, :] < npts[:, :, None]
    pts[~live] = 0
    cache = np.ascontiguousarray(pts[:, :, 0, 13:16]).view(np.uint32).copy()     # n, (ia | ib << 8) x 4 (include/b2s.h)
    cache[npts == 0] = 0
    pts[..., 13:] = 0
    return keys, npts, pts, cache


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, '%s: shape %s vs %s' % (what, a.shape, b.shape)
    if a.dtype.kind == 'f':
        neq = bits(a) != bits(b)
        # +0/-0 and NaN payloads do not matter
        neq &= ~((a == 0) & (b == 0))
    else:
        neq = a != b
    if neq.any():
        idx = np.argwhere(neq)[:5]
        raise AssertionError('%s: %d of %d entries differ, first at %s: %s vs %s' % (
            what, int(neq.sum()), neq.size, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])]))
