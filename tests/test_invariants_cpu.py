"""Physics invariants that do NOT share code with the implementation ("what's weak").

The CUDA kernels and the CPU oracle compile the same leaf headers (include/b2s_math.h, b2s_geom.h), so a wrong
Voronoi test or Jacobian there would pass every CUDA == oracle parity test.  The checks below pin the oracle (and with
it, bit for bit, the CUDA path) against closed forms and against independent numpy / scipy computations:

  * Coulomb sliding distance v^2 / (2 mu g)
  * linear momentum across the impact of two free-floating movables
  * no tunnelling at dt = 1/240 s: a fast drop, and a body hit by the finger at the arm's maximum joint speed
  * GJK distance / EPA depth of near-touching hull pairs against a scipy QP (separated) and a SAT sweep (overlapping)
  * the stated fp32 pose tolerance against the double-precision build of the oracle
"""
import copy
import itertools

import numpy as np
import pytest

from robovat_b200 import _capi, assets, config
from tests import helpers

G = 9.8
TABLE_MU = 1.0


def _cfg(paths=('box',), n=1, mu=0.5, mass=0.2, table_mu=TABLE_MU, dt=1.0 / 240.0, damping=0.0, gravity=(0, 0, -G), **extra):
    mov = copy.deepcopy(config.DEFAULT_PUSH_ENV['MOVABLE'])
    mov['CONVEX'].update(PATHS=list(paths), TARGET_PATHS=[paths[0]], SCALE=[1.0, 1.0], MASS=[mass, mass], FRICTION=[mu, mu])
    sim = copy.deepcopy(config.DEFAULT_PUSH_ENV['SIM'])
    sim['TIME_STEP'] = dt
    sim['TABLE']['FRICTION'] = table_mu
    phys = dict(config.DEFAULT_PUSH_ENV['PHYSICS'], LINEAR_DAMPING=damping, ANGULAR_DAMPING=damping, GRAVITY=list(gravity))
    return dict(MOVABLE=mov, SIM=sim, PHYSICS=phys, MIN_MOVABLE_BODIES=n, MAX_MOVABLE_BODIES=n, **extra)


def _place(w, e, i, pos, quat=(0, 0, 0, 1), vel=(0, 0, 0), ang=(0, 0, 0)):
    w.body_state[:, e, i] = list(pos) + list(quat) + list(vel) + list(ang)


def _rest_boxes(B, n=1, **kw):
    """B envs with n boxes (0.08 x 0.08 x 0.05) lying flat on the table top (z = 0), contact caches built."""
    cfg, w = helpers.make_oracle(B, threads=4, **_cfg(n=n, **kw))
    w.reset(seed=0)
    for e in range(B):
        for i in range(n):
            _place(w, e, i, (0.55 + 0.15 * i, 0.0, 0.025 + 0.001))
    w.step(240)
    assert np.abs(w.body_state[7:10]).max() < 2e-3 and np.abs(w.body_state[10:13]).max() < 3e-2
    return cfg, w


@pytest.mark.parametrize('mu', [0.3, 0.6])
def test_coulomb_sliding_distance(mu):
    """A flat box launched at v0 along x (a direction of the friction pyramid) stops after v0^2 / (2 mu g).
    Speeds stay below (contact breaking threshold) / dt = 0.02 * 0.063 m * 240 / s = 0.30 m/s: a persistent manifold
    (Bullet's and this one) loses its cached points when the body travels further than the threshold in one substep
    and then rebuilds them one GJK point per substep, during which a small box rocks instead of sliding."""
    v0s = np.array([0.1, 0.15, 0.2, 0.28])
    cfg, w = _rest_boxes(len(v0s), mu=mu)
    x0 = w.body_state[0, :, 0].copy()
    w.body_state[7, :, 0] = v0s
    w.step(480)
    assert np.abs(w.body_state[7:10]).max() < 1e-3                      # at rest again
    slid = w.body_state[0, :, 0] - x0
    expect = v0s ** 2 / (2 * mu * TABLE_MU * G)
    # semi-implicit Euler at dt = 1/240 stops within one step's travel of the closed form
    np.testing.assert_allclose(slid, expect, rtol=0.04, atol=1.5e-3)
    assert np.abs(w.body_state[1, :, 0]).max() < 2e-3                   # no sideways drift
    w.close()


def test_impact_of_two_movables_conserves_linear_momentum():
    """Two bodies floating without gravity (so that only the movable-movable contact acts), the first flying into the
    second off-centre: the contact impulses are equal and opposite, so the total linear momentum is kept to rounding
    through the impact, whatever the mass ratio and the shapes; there is no restitution, so they do not separate
    faster than they met."""
    B = 6
    cfg, w = helpers.make_oracle(B, **_cfg(paths=('box', 'hex', 'wedge'), n=2, mu=0.5, gravity=(0, 0, 0)))
    w.reset(seed=2)
    m = w.array(_capi.ARR_MOV_PARAMS).reshape(4, B, 2)
    m[2, :, 0] = [0.1, 0.2, 0.3, 0.1, 0.05, 0.3]
    m[2, :, 1] = [0.1, 0.1, 0.1, 0.3, 0.30, 0.3]
    rs = np.random.RandomState(0)
    for e in range(B):
        qa = rs.normal(size=4); qa /= np.linalg.norm(qa)
        qb = rs.normal(size=4); qb /= np.linalg.norm(qb)
        _place(w, e, 0, (0.5, 0.0, 0.4), qa, vel=(0.25, 0.0, 0.0))
        _place(w, e, 1, (0.62, 0.02 * (e % 3 - 1), 0.4 + 0.01 * (e % 2)), qb)
    mass = m[2].copy()                                                   # [B, 2]
    p0 = (mass[None] * w.body_state[7:10]).sum(axis=2)                   # [3, B]
    hit = np.zeros(B, bool)
    for _ in range(30):
        w.step(8)
        p = (mass[None] * w.body_state[7:10]).sum(axis=2)
        np.testing.assert_allclose(p, p0, atol=2e-6)
        hit |= np.abs(w.body_state[7, :, 1]) > 0.01
    assert hit.all()                                                     # every second body was set in motion
    ke0 = 0.5 * mass[:, 0] * 0.25 ** 2
    ke = 0.5 * (mass[None] * w.body_state[7:10] ** 2).sum(axis=(0, 2))
    assert (ke <= ke0 * (1 + 1e-4)).all()                                # translational energy never grows (rotation takes some)
    w.close()


def test_no_tunnelling_of_a_fast_drop():
    """dt = 1/240: a box arriving at 6 m/s moves 2.5 cm per substep, half the 5 cm table thickness; it must end on top."""
    cfg, w = helpers.make_oracle(6, **_cfg(n=1))
    w.reset(seed=0)
    for e, vz in enumerate([-1.0, -2.0, -3.0, -4.0, -5.0, -6.0]):
        _place(w, e, 0, (0.6, 0.0, 0.3), vel=(0, 0, vz))
    zmin = np.full(6, 1.0)
    for _ in range(60):
        w.step(4)
        zmin = np.minimum(zmin, w.body_state[2, :, 0])
    # discrete collision detection: the box may enter by up to one substep of travel (2.5 cm at 6 m/s) before the
    # contact pushes it back, but its centre never reaches the middle of the 5 cm table
    assert (zmin > -0.02).all() and (zmin[:3] > 0.005).all(), zmin
    w.step(240)
    np.testing.assert_allclose(w.body_state[2, :, 0], 0.025, atol=2e-3)
    w.close()


def test_finger_at_max_joint_speed_does_not_pass_through_a_body():
    """The arm sweeps its finger through the place of a box as fast as the controller allows (LIMB_MAX_VELOCITY_RATIO
    x max joint velocity): the box is pushed ahead of the finger, it is never left behind inside or across it."""
    cfg, w = helpers.make_oracle(4, threads=4, **_cfg(n=1, mu=0.5))
    w.reset(seed=0)
    w.settle(0.1, 0.1, 500)
    q_down = assets.quat_from_euler(np.pi, 0, 0)
    z = cfg.ARM.FINGER_TIP_OFFSET + 0.5 * (cfg.ACTION.CSPACE.LOW[2] + cfg.ACTION.CSPACE.HIGH[2])
    start = np.tile(np.array([0.45, 0.0, z] + list(q_down), np.float32), (4, 1))
    w.move_to_gripper_pose(start)
    for _ in range(40):
        w.step(100)
        if w.arm_is_ready().all():
            break
    for e in range(4):
        _place(w, e, 0, (0.55 + 0.02 * e, 0.0, 0.026))
    w.step(60)
    goal = start.copy()
    goal[:, 0] = 0.80
    w.move_to_gripper_pose(goal)
    x_body, x_tip, speed = [], [], 0.0
    for _ in range(240):
        w.step(5)
        tip = w.forward_kinematics()[:, -1, :3].copy()
        x_tip.append(tip[:, 0]); x_body.append(w.body_state[0, :, 0].copy())
        if len(x_tip) > 1:
            speed = max(speed, float(np.abs(x_tip[-1] - x_tip[-2]).max() / (5 * w.params.time_step)))
    x_body, x_tip = np.array(x_body), np.array(x_tip)
    assert speed > 0.3                                                   # the sweep was fast (m/s at the finger tip)
    assert (x_tip[-1] > 0.7).all()                                       # the finger went all the way
    ahead = x_body - x_tip                                               # body centre relative to the finger axis
    touched = ahead < 0.06
    assert touched.any(axis=0).all()
    # once in contact the body centre stays ahead of the finger axis (half box = 4 cm, finger radius ~1 cm)
    assert (ahead[touched] > 0.02).all(), ahead[touched].min()
    assert (w.body_state[2, :, 0] > 0.0).all() and (w.array(_capi.ARR_ERROR_FLAGS) == 0).all()
    w.close()


# ---------------------------------------------------------------- independent narrow phase -------------

def _hull_faces(verts):
    from scipy.spatial import ConvexHull
    h = ConvexHull(verts)
    n = h.equations[:, :3]
    keep = []
    for v in n:                                                           # unique face normals
        if not any(np.allclose(v, k, atol=1e-9) for k in keep):
            keep.append(v)
    edges = set()
    for s in h.simplices:
        for a, b in itertools.combinations(sorted(s), 2):
            edges.add((a, b))
    dirs = []
    for a, b in edges:                                                    # unique edge directions
        d = verts[b] - verts[a]
        d = d / np.linalg.norm(d)
        if not any(abs(abs(d.dot(k)) - 1) < 1e-9 for k in dirs):
            dirs.append(d)
    return np.array(keep), np.array(dirs)


def _sat_signed_distance(A, B):
    """max over separating-axis candidates (face normals, edge x edge) of the gap; < 0 = penetration depth"""
    nA, eA = _hull_faces(A)
    nB, eB = _hull_faces(B)
    axes = list(nA) + list(nB)
    for a in eA:
        for b in eB:
            c = np.cross(a, b)
            if np.linalg.norm(c) > 1e-9:
                axes.append(c / np.linalg.norm(c))
    best = -np.inf
    for n in axes:
        for s in (n, -n):
            best = max(best, (B @ s).min() - (A @ s).max())
    return best


def _qp_distance(A, B):
    from scipy.optimize import minimize
    na, nb = len(A), len(B)

    def f(x):
        d = x[:na] @ A - x[na:] @ B
        return d @ d

    def g(x):
        d = x[:na] @ A - x[na:] @ B
        return np.concatenate([2 * A @ d, -2 * B @ d])
    cons = [{'type': 'eq', 'fun': lambda x: x[:na].sum() - 1}, {'type': 'eq', 'fun': lambda x: x[na:].sum() - 1}]
    best = np.inf
    for trial in range(3):
        rs = np.random.RandomState(trial)
        x0 = np.concatenate([rs.dirichlet(np.ones(na)), rs.dirichlet(np.ones(nb))])
        r = minimize(f, x0, jac=g, bounds=[(0, 1)] * (na + nb), constraints=cons, method='SLSQP',
                     options={'ftol': 1e-16, 'maxiter': 500})
        best = min(best, np.sqrt(max(r.fun, 0.0)))
    return best


def _quat_rot(q, v):
    return v @ assets.quat_to_matrix(q).T


def test_narrow_phase_distance_against_independent_solvers():
    """Two movables floating in zero gravity, nearly touching or slightly overlapping: the distance the narrow phase
    stores in the manifold (core distance minus the two hull margins) equals an independent scipy QP distance
    (separated pairs) / SAT penetration depth (overlapping pairs) of the same world-space vertex sets."""
    shapes = assets.convex_movables()
    names = ['box', 'hex', 'wedge']
    B = 48
    cfg, w = helpers.make_oracle(B, **_cfg(paths=names, n=2, gravity=(0, 0, 0)))
    w.reset(seed=3)
    rs = np.random.RandomState(5)
    mp = w.array(_capi.ARR_MOV_PARAMS).reshape(4, B, 2)
    margin = assets.HULL_MARGIN
    lib = w.scene.lib
    verts = [[(int(mp[0, e, i:i + 1].view(np.int32)[0]), None) for i in range(2)] for e in range(B)]
    hull_v = {}
    for aid in set(a for p in verts for a, _ in p):
        h = lib.asset_hull_off[aid]
        off, cnt = lib.hull_vert_off[h], lib.hull_vert_cnt[h]
        hull_v[aid] = np.array(lib.verts[off:off + cnt], np.float32).astype(np.float64)
    want = []
    for e in range(B):
        qa = rs.normal(size=4); qa /= np.linalg.norm(qa)
        qb = rs.normal(size=4); qb /= np.linalg.norm(qb)
        A = _quat_rot(qa, hull_v[verts[e][0][0]])
        Bv = _quat_rot(qb, hull_v[verts[e][1][0]])
        # slide B along a random direction until the SAT gap is the wanted signed distance
        u = rs.normal(size=3); u /= np.linalg.norm(u)
        gap = rs.uniform(-0.004, 0.0025)
        lo, hi = 0.0, 0.5
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if _sat_signed_distance(A, Bv + mid * u) < gap:
                lo = mid
            else:
                hi = mid
        t = 0.5 * (lo + hi)
        pa, pb = np.array([0.6, 0.0, 0.5]), np.array([0.6, 0.0, 0.5]) + t * u
        _place(w, e, 0, pa, qa)
        _place(w, e, 1, pb, qb)
        want.append((A + pa, Bv + pb))
    w.step(1)
    M = w.params.max_manifolds
    keys = w.array(_capi.ARR_MANIFOLD_KEYS).reshape(B, M)
    npts = w.array(_capi.ARR_MANIFOLD_NPTS).reshape(B, M)
    pts = w.array(_capi.ARR_MANIFOLD_PTS).reshape(B, M, 4, _capi.CP_FLOATS)
    first_mov = w.params.max_colliders - 2
    checked_sep = checked_pen = 0
    for e in range(B):
        # float32 state: recompute the reference from the float32 poses the oracle actually used
        A, Bv = want[e]
        sd = _sat_signed_distance(A, Bv)
        found = [k for k in range(M) if keys[e, k] == ((first_mov + 1) << 16 | first_mov) and npts[e, k] > 0]
        threshold = 0.02 * (np.linalg.norm(hull_v[verts[e][0][0]], axis=1).max() + margin)
        if sd > 0:
            ref = _qp_distance(A, Bv) - 2 * margin
        else:
            ref = sd - 2 * margin
        if ref >= threshold - 2e-4:
            continue                                                      # too far for a contact (or borderline)
        assert found, (e, sd, ref)
        got = float(pts[e, found[0], 0, 9])
        assert abs(got - ref) < 3e-5, (e, got, ref, sd)
        n = pts[e, found[0], 0, 6:9]
        assert abs(np.linalg.norm(n) - 1) < 1e-5
        if sd > 0:
            checked_sep += 1
        else:
            checked_pen += 1
    assert checked_sep >= 8 and checked_pen >= 8, (checked_sep, checked_pen)
    w.close()


def test_stated_fp32_pose_tolerance_against_double_precision():
    """The tolerance DESIGN.md states: fp32 path vs the double-precision build of the same algorithm, 240 substeps at
    dt = 1/240 from identical states (oracle/pose_tolerance.py).  Rounding differences grow through contact events,
    so the statement is distributional: median well under BASELINE's 1 mm, a tail of diverged tumbling bodies."""
    from oracle import b2o, pose_tolerance
    sim = copy.deepcopy(config.DEFAULT_PUSH_ENV['SIM'])
    sim['TIME_STEP'] = 1.0 / 240.0
    cfg, scene, params = helpers.make_inputs(96, SIM=sim)
    w32 = b2o.OracleWorld(params, scene, threads=4)
    out = {o['scenario']: o for o in pose_tolerance.measure(w32, params, scene, seed=0, threads=4)}
    w32.close()
    assert out['rest']['median_m'] < 3e-5 and out['rest']['max_m'] < 1e-3
    assert out['slide']['median_m'] < 1e-3 and out['slide']['p90_m'] < 2e-3 and out['slide']['max_m'] < 1e-2
    assert out['drop']['median_m'] < 2e-4 and out['drop']['frac_le_1e-3'] > 0.8


@pytest.mark.parametrize('spin', [0.001, 0.002])
def test_spinning_friction_stops_a_frictionless_box(spin):
    """Torsional rows (urdf_template.xml:12-14, body.py:225-230).  A box WITHOUT lateral friction spinning flat on the
    table about the vertical: the pyramid rows cannot touch it, only the spinning rows of its contact points do, each
    limited to mu_c * (its normal impulse) with mu_c = spinning * table friction.  The normal impulses sum to m g dt, so
    the spin decays at alpha = mu_c m g / I_zz (I_zz of the 0.08 x 0.08 box = m (0.04^2 + 0.04^2) / 3, the AABB inertia)
    and the box turns by w0^2 / (2 alpha) before it stops.  With rolling_friction = 0 the rows are off (Bullet's gate)
    and the spin is kept."""
    w0s = np.array([1.0, 1.5, 2.0])
    on = dict(config.DEFAULT_PUSH_ENV['PHYSICS'], ROLLING_FRICTION=0.001, SPINNING_FRICTION=spin)
    off = dict(on, ROLLING_FRICTION=0.0)
    izz = (0.04 ** 2 + 0.04 ** 2) / 3.0
    alpha = spin * TABLE_MU * G / izz
    for phys, expect_stop in ((on, True), (off, False)):
        kw = _cfg(mu=0.0)
        kw['PHYSICS'] = dict(phys, LINEAR_DAMPING=0.0, ANGULAR_DAMPING=0.0)
        cfg, w = helpers.make_oracle(len(w0s), threads=4, **kw)
        w.reset(seed=0)
        for e in range(len(w0s)):
            _place(w, e, 0, (0.55, 0.0, 0.025 + 0.001))
        w.step(240)
        yaw0 = 2 * np.arctan2(w.body_state[5, :, 0], w.body_state[6, :, 0])
        w.body_state[12, :, 0] = w0s
        if expect_stop:
            w.step(240)
            yaw = 2 * np.arctan2(w.body_state[5, :, 0], w.body_state[6, :, 0]) - yaw0
            assert np.abs(w.body_state[10:13]).max() < 2e-2
            # a substep of travel at the start (semi-implicit Euler) is the discretisation error of the closed form
            expect = w0s ** 2 / (2 * alpha)
            assert (np.abs(yaw - expect) <= 0.05 * expect + w0s / 240.0).all(), (yaw, expect)
        else:
            # (a tenth of a second, 2 %: a frictionless box that keeps turning loses and rebuilds its cached contact points
            # and starts to rock, which is the manifold's business, not the solver's)
            w.step(24)
            np.testing.assert_allclose(w.body_state[12, :, 0], w0s, rtol=2e-2)
        w.close()


@pytest.mark.parametrize('task,layout', [('clearing', 0), ('clearing', 2), ('crossing', 0), ('insertion', 0)])
def test_default_capacities_hold_the_contacts_of_tiled_tasks(task, layout):
    """Capacities are part of the physics: a contact list that overflows drops contact points (error flag 8).  On the
    colliding tiles of the task layouts a movable rests on up to four static bodies at once (up to 50 points for three
    convex movables on the clearing layouts), which config.build_params has to provide: no capacity flag after reset,
    settle and 240 substeps, and the most crowded env stays below the capacity."""
    sim = copy.deepcopy(config.DEFAULT_PUSH_ENV['SIM'])
    sim['TIME_STEP'] = 1.0 / 240.0
    cfg, w = helpers.make_oracle(192, threads=4, TASK_NAME=task, LAYOUT_ID=layout, SIM=sim)
    w.reset(seed=3)
    w.settle(0.1, 0.1, 500)
    w.settle()
    most = 0
    for _ in range(12):
        w.step(20)
        most = max(most, int(w.array(_capi.ARR_SOLVER_STATS).reshape(-1, 4)[:, 3].max()))
    assert int((w.array(_capi.ARR_ERROR_FLAGS) & (1 | 2 | 8 | 16)).max()) == 0
    assert most < w.params.max_contacts, (most, w.params.max_contacts)
    w.close()
