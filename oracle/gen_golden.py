"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the authoring container (needs /root/reference):

    cd /tmp && PYTHONPATH=/root/repo python -B -m oracle.gen_golden

Everything written here is an OUTPUT OF THE REFERENCE'S OWN CODE:
  transformations.json  third_party/transformations.py functions + its doctest vectors
  pose.json             robovat.math.Pose.inverse / transform / get_transform
  layouts.json          robovat.envs.push.layouts.TASK_NAME_TO_LAYOUTS
  reward.json           robovat.reward_fns.push_reward.get_reward_fn(task, layout)(state, next_state)
  sampler.json          robovat.envs.push.heuristic_push_sampler.HeuristicPushSampler.sample
  camera.json           robovat.perception.camera.camera.Camera + bullet_camera.intrinsic_to_projection_matrix
  episodes.json         robovat.io.episode_generation.generate_episode(s) on a scripted stand-in env
  mesh.json             robovat.utils.mesh_utils (OBJ reader, volume / area / centroid), the URDF text written from
                        tools/templates/*.xml and tools/convert_obj_to_urdf.split_wrl_file
  waypoints.json        robovat.envs.push.push_env.PushEnv._compute_waypoints
  push_step_trace.npz   the reference PushEnv.step() (its own _execute_action, SawyerSim, ControllableBody,
                        Simulator.wait_until_stable) driving OUR physics backend substep by substep
                        (oracle/ref_cosim.py): start snapshot, action, phase log and final state.
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')

SNAP_IDS = ['body_state', 'joint_state', 'num_movables', 'body_mask', 'is_safe', 'is_effective']


def _tolist(x):
    return np.asarray(x, dtype=np.float64).tolist()


def snapshot(world):
    from robovat_b200 import _capi
    snap = {k: np.array(world.array(k)) for k in SNAP_IDS}
    for which in (_capi.ARR_MANIFOLD_KEYS, _capi.ARR_MANIFOLD_NPTS, _capi.ARR_MANIFOLD_PTS, _capi.ARR_NUM_MANIFOLDS,
                  _capi.ARR_NUM_STEPS, _capi.ARR_CTRL, _capi.ARR_CTRL_FLAGS, _capi.ARR_CTRL_TIME, _capi.ARR_MOV_PARAMS,
                  _capi.ARR_TABLE_DZ, _capi.ARR_NUM_COLLIDERS, _capi.ARR_COL_SLOT, _capi.ARR_COL_HULL,
                  _capi.ARR_PHASE_STATE, _capi.ARR_PHASE):
        snap['arr%d' % which] = np.array(world.array(which))
    return snap


def gen_transformations(rs):
    import third_party.transformations as T
    eul = rs.uniform(-np.pi, np.pi, (64, 3))
    eul[:, 1] = rs.uniform(-1.5, 1.5, 64)
    quat = np.array([T.quaternion_from_euler(*e) for e in eul])
    qa, qb = rs.normal(size=(64, 4)), rs.normal(size=(64, 4))
    qa /= np.linalg.norm(qa, axis=1, keepdims=True)
    qb /= np.linalg.norm(qb, axis=1, keepdims=True)
    return {
        'euler': _tolist(eul),
        'quaternion_from_euler': _tolist(quat),
        'euler_from_quaternion': _tolist([T.euler_from_quaternion(q) for q in quat]),
        'matrix3_from_quaternion': _tolist([T.matrix3_from_quaternion(q) for q in quat]),
        'qa': _tolist(qa), 'qb': _tolist(qb),
        'quaternion_multiply': _tolist([T.quaternion_multiply(a, b) for a, b in zip(qa, qb)]),
        # doctest vectors quoted in SURVEY.md section 4 (transformations.py:1186-1188, 1254-1256, 1365-1367)
        'doctests': {
            'euler_from_quaternion([0.06146124,0,0,0.99810947])': _tolist(T.euler_from_quaternion([0.06146124, 0, 0, 0.99810947])),
            'quaternion_about_axis(0.123,(1,0,0))': _tolist(T.quaternion_about_axis(0.123, (1, 0, 0))),
            'quaternion_multiply([1,-2,3,4],[-5,6,7,8])': _tolist(T.quaternion_multiply([1, -2, 3, 4], [-5, 6, 7, 8])),
        },
    }


def gen_pose(rs):
    from robovat.math import Pose, get_transform
    a = [[rs.uniform(-1, 1, 3), rs.uniform(-1.4, 1.4, 3)] for _ in range(32)]
    b = [[rs.uniform(-1, 1, 3), rs.uniform(-1.4, 1.4, 3)] for _ in range(32)]

    def flat(p):
        return list(np.asarray(p.position, np.float64)) + list(np.asarray(p.quaternion, np.float64))
    return {
        'a': [flat(Pose(x)) for x in a], 'b': [flat(Pose(x)) for x in b],
        'inverse_a': [flat(Pose(x).inverse()) for x in a],
        'a_transform_b': [flat(Pose(x).transform(Pose(y))) for x, y in zip(a, b)],
        'get_transform_source_a_target_b': [flat(get_transform(source=Pose(x), target=Pose(y))) for x, y in zip(a, b)],
    }


def gen_layouts():
    from robovat.envs.push import layouts
    return {task: [dict(l._asdict()) for l in ls] for task, ls in layouts.TASK_NAME_TO_LAYOUTS.items()}


def gen_reward(rs):
    from robovat.reward_fns import push_reward
    out = []
    for task in ('clearing', 'insertion', 'crossing'):
        for layout_id in range(3):
            fn = push_reward.get_reward_fn(task, layout_id)
            for n_bodies, n_max in ((3, 3), (2, 4), (8, 8)):
                B = 24
                s0 = np.zeros((B, n_max, 3))
                s0[:, :n_bodies, 0] = rs.uniform(0.2, 1.0, (B, n_bodies))
                s0[:, :n_bodies, 1] = rs.uniform(-0.6, 0.7, (B, n_bodies))
                s1 = s0.copy()
                s1[:, :n_bodies, :2] += rs.uniform(-0.15, 0.15, (B, n_bodies, 2))
                rew, term = [], []
                for b in range(B):          # the env path calls it with unbatched dict observations
                    r, t = fn({'position': s0[b]}, {'position': s1[b]})
                    rew.append(float(r[0])); term.append(bool(t[0]))
                out.append({'task': task, 'layout_id': layout_id, 'state': s0[:, :, :2].tolist(),
                            'next_state': s1[:, :, :2].tolist(), 'reward': rew, 'termination': term})
    return out


def gen_sampler():
    from robovat.envs.push.heuristic_push_sampler import HeuristicPushSampler
    cases = []
    for seed in range(6):
        rs = np.random.RandomState(100 + seed)
        n = 3
        position = np.c_[rs.uniform(0.4, 0.8, n), rs.uniform(-0.3, 0.3, n), np.full(n, 0.03)]
        sampler = HeuristicPushSampler([0.35, -0.35, 0.0], [0.85, 0.35, 0.04], 0.2, 0.2, max_attemps=20000)
        np.random.seed(seed)
        act = sampler.sample(position, np.ones(n), num_episodes=seed, num_steps=0, num_samples=2)
        cases.append({'seed': seed, 'position': position.tolist(), 'num_episodes': seed, 'action': np.asarray(act, np.float64).tolist()})
    return cases


def gen_camera(rs):
    import robovat.perception.camera.camera as cam_mod
    from robovat.simulation.camera import bullet_camera

    class Cam(cam_mod.Camera):
        def _frames(self):
            return {}
    K = np.array([[365.0, 0, 256.0], [0, 365.0, 212.0], [0, 0, 1.0]])
    t = np.array([0.1, -0.2, 1.0])
    eul = np.array([np.pi, 0.0, 0.0])
    cam = Cam(height=424, width=512, intrinsics=K, translation=t, rotation=eul)
    depth = rs.uniform(0.5, 1.2, (6, 7))
    pc = cam.deproject_depth_image(depth)
    pts = rs.uniform(-0.3, 0.3, (16, 3)) + [0.1, 0.2, 0.0]
    return {
        'K': K.tolist(), 't': t.tolist(), 'euler': eul.tolist(), 'rotation_matrix': np.asarray(cam.rotation, np.float64).tolist(),
        'depth': depth.tolist(), 'deproject_depth_image': np.asarray(pc, np.float64).tolist(),
        'points': pts.tolist(), 'project_point': [np.asarray(cam.project_point(p)).tolist() for p in pts],
        'projection_matrix_424x512_near0.02_far100': _tolist(bullet_camera.intrinsic_to_projection_matrix(K, 424, 512, 0.02, 100)),
    }


def gen_calibration():
    """ArmEnv._reset_camera (arm_env.py:109-152) of the unmodified reference with a recording camera stub."""
    from robovat.envs import arm_env

    class Cam(object):
        def set_calibration(self, intrinsics, translation, rotation):
            self.got = (intrinsics, translation, rotation)

    class Self(object):
        is_simulation = True
    out = []
    K = [365.0, 0, 256.0, 0, 365.0, 212.0, 0, 0, 1]
    for seed, (kn, tn, rn) in enumerate([(None, None, None), ([2.0, 0, 2.0, 0, 2.0, 2.0, 0, 0, 0], [0.01, 0.02, 0.03], [0.05, 0.05, 0.1]),
                                         (None, [0.02, 0.02, 0.02], None), ([1.0] * 9, None, [0.01, 0.0, 0.02])]):
        np.random.seed(100 + seed)
        cam = Cam()
        arm_env.ArmEnv._reset_camera(Self(), cam, np.array(K).reshape(3, 3), [0.6, 0.0, 1.2], [np.pi, 0, 0],
                                     None if kn is None else np.array(kn).reshape(3, 3), tn, rn)
        out.append({'seed': 100 + seed, 'intrinsics': K, 'translation': [0.6, 0.0, 1.2], 'rotation': [float(np.pi), 0, 0],
                    'intrinsics_noise': kn, 'translation_noise': tn, 'rotation_noise': rn,
                    'got_intrinsics': _tolist(np.asarray(cam.got[0], np.float64)), 'got_translation': _tolist(np.asarray(cam.got[1], np.float64)),
                    'got_rotation': _tolist(np.asarray(cam.got[2], np.float64))})
    return out


def gen_cosim(cfg_bindings, seed, action_fn, name):
    """Reference-driven PushEnv.step on our backend; returns the arrays of one trace."""
    from oracle import ref_cosim
    from robovat_b200 import _capi, config
    cfg = config.default_push_env_config(**cfg_bindings)
    random.seed(seed)
    np.random.seed(seed)
    env, world, scene = ref_cosim.make_reference_env(cfg)
    obs = env.reset()
    start = snapshot(world)
    arm = env.robot.arm
    jt = arm._joint_target
    ctrl = {
        'joint_target_active': int(not jt.is_ready()),
        'joint_target_positions': np.asarray(jt.positions if jt.positions is not None else np.zeros(7), np.float32),
        'joint_start_time': float(jt.start_time or 0.0), 'joint_stop_time': float(jt.stop_time or 0.0),
        'gripper_ready_time': float(env.robot._gripper_ready_time),
    }
    action = action_fn(np.asarray(obs['position']), cfg)
    log = []
    orig_next = env._get_next_phase

    def logged_next():
        ph = orig_next()
        log.append((int(env.simulator.num_steps), env.phase_list.index(ph)))
        return ph
    env._get_next_phase = logged_next
    waypoints = env._compute_waypoints(action if np.ndim(action) == 1 else action[0])     # of the first goal step
    obs2, reward, done, _ = env.step(action)
    final = snapshot(world)
    out = {'action': np.asarray(action, np.float32), 'phase_log': np.asarray(log, np.int32),
           'reward': np.float32(reward), 'done': np.bool_(done),
           'waypoint_start': np.asarray(list(waypoints[0].position) + list(waypoints[0].quaternion), np.float64),
           'waypoint_end': np.asarray(list(waypoints[1].position) + list(waypoints[1].quaternion), np.float64),
           'obs_position': np.asarray(obs2['position'], np.float32),
           'is_safe': np.int32(obs2['is_safe']), 'is_effective': np.int32(obs2['is_effective']),
           'config_bindings': np.array(json.dumps(cfg_bindings))}
    for k, v in ctrl.items():
        out['ctrl_' + k] = np.asarray(v)
    for k, v in start.items():
        out['start_' + k] = v
    for k, v in final.items():
        out['final_' + k] = v
    print('  cosim %s: %d substeps, phases %s, safe %d effective %d' % (
        name, int(final['arr%d' % _capi.ARR_NUM_STEPS][0] - start['arr%d' % _capi.ARR_NUM_STEPS][0]),
        [p for _, p in log], out['is_safe'], out['is_effective']))
    return out


def push_twice(first, second):
    """NUM_GOAL_STEPS = 2: action [2, 4], both pushes aimed at body 0's position at the start of the action."""
    def fn(position, cfg):
        return np.stack([first(position, cfg), second(position, cfg)])
    return fn


def push_body0(dx, dy):
    def fn(position, cfg):
        lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2]), np.array(cfg.ACTION.CSPACE.HIGH[:2])
        off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
        d = np.array([dx, dy], np.float64)
        start = position[0, :2] - 0.08 * d / np.linalg.norm(d)
        return np.r_[np.clip((start - off) / rng, -1, 1), d].astype(np.float32)
    return fn


OBJ_SAMPLE = '''# comment line
v 0 0 0
v 1 0 0
v 0 1 0

v 0 0 1
vt 0.5 0.5
vn 0 0 1
f 1/1/1 3/1/1 2/1/1
f 1//1 2//1 4//1
f 2 3 4
f 1 4 3
bogus line
'''

WRL_SAMPLE = '''#VRML V2.0 utf8
Group {
 children [
  Shape { geometry IndexedFaceSet { coord Coordinate { point [
   0 0 0,
   1 0 0,
   0 1 0,
   0 0 1,
  ] } coordIndex [ 0, 2, 1, -1, 0, 1, 3, -1, 1, 2, 3, -1, 0, 3, 2, -1, ] } }
 ]
}
#VRML V2.0 utf8
Group {
 children [
  Shape { geometry IndexedFaceSet { coord Coordinate { point [
   2 0 0,
   3 0 0,
   2 1 0,
   2 0 1,
  ] } coordIndex [ 0, 2, 1, -1, 0, 1, 3, -1, 1, 2, 3, -1, 0, 3, 2, -1, ] } }
 ]
}
'''


def gen_mesh():
    """Reference mesh_utils on the committed OBJ sources + a hand-written OBJ with v/vt/vn faces; URDF template text."""
    import importlib.util
    import tempfile
    from robovat.utils import mesh_utils
    out = {'objs': {}}
    data = os.path.join(ROOT, 'robovat_b200', 'data', 'urdf')
    paths = {name: os.path.join(data, name, name + '.obj') for name in sorted(os.listdir(data))}
    tmp = tempfile.mkdtemp()
    sample = os.path.join(tmp, 'sample.obj')
    with open(sample, 'w') as f:
        f.write(OBJ_SAMPLE)
    paths['sample'] = sample
    for name, path in paths.items():
        v, t = mesh_utils.read_from_obj(path)
        out['objs'][name] = {'text': open(path).read(), 'vertices': _tolist(v), 'triangles': np.asarray(t).tolist(),
                             'volume': float(mesh_utils.compute_volume(v, t)),
                             'surface_area': float(mesh_utils.compute_surface_area(v, t)),
                             'centroid': _tolist(mesh_utils.compute_centroid(v, t))}
    tdir = os.path.join('/root/reference', 'tools', 'templates')
    visual = open(os.path.join(tdir, 'visual_template.xml')).read()
    collision = open(os.path.join(tdir, 'collision_template.xml')).read()
    urdf = open(os.path.join(tdir, 'urdf_template.xml')).read()
    cases = []
    for body_name, files, mass, c, scale, rgba in (
            ('L', ['L_vhacd_0_of_2.obj', 'L_vhacd_1_of_2.obj'], 0.1, [-0.0165432, -0.00654321, 1.2e-9], 1.0, '0.50 0.50 0.50 1.00'),
            ('thing', ['thing_vhacd_0_of_1.obj'], 0.25, [0.0, 1e-7, -123.456], 0.75, '0.12 0.34 0.56 1.00')):
        vt = ''.join(visual.format(filename=fn, scale=scale) for fn in files)
        ct = ''.join(collision.format(filename=fn, scale=scale) for fn in files)
        text = urdf.format(body_name=body_name, mass=mass, ixx=1, iyy=1, izz=1, ixy=0, ixz=0, iyz=0, cx=c[0], cy=c[1], cz=c[2],
                           visual=vt, collision=ct, rgba=rgba)
        cases.append({'body_name': body_name, 'files': files, 'mass': mass, 'centroid': c, 'scale': scale, 'rgba': rgba, 'text': text})
    out['urdf'] = cases
    # split_wrl_file of the reference tool (its module imports only os/argparse/numpy + robovat.utils)
    import types
    sys.modules.setdefault('_init_paths', types.ModuleType('_init_paths'))    # tools/_init_paths.py only edits sys.path
    spec = importlib.util.spec_from_file_location('ref_convert', '/root/reference/tools/convert_obj_to_urdf.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    wrl = os.path.join(tmp, 'output.wrl')
    with open(wrl, 'w') as f:
        f.write(WRL_SAMPLE)
    pieces = [open(p).read() for p in mod.split_wrl_file(wrl, tmp_dir=tmp)]
    out['wrl'] = {'text': WRL_SAMPLE, 'pieces': pieces, 'groups': mod.count_output_groups(wrl)}
    return out


class ScriptedEnv(object):
    """Deterministic stand-in env for the episode driver: reward / done follow a script."""

    def __init__(self, script):
        self.script, self.t, self.resets = script, 0, 0

    def reset(self):
        self.t = 0
        self.resets += 1
        return {'position': [float(self.resets), 0.0]}

    def step(self, action):
        reward, done = self.script[self.t]
        self.t += 1
        return {'position': [float(self.resets), float(self.t)]}, reward + float(action[0]), done, None


class ScriptedPolicy(object):
    def action(self, observation):
        return [observation['position'][1] * 0.5, 1.0]


def gen_episodes():
    """robovat.io.episode_generation.generate_episode(s) on the scripted env: transitions per episode."""
    from robovat.io import episode_generation
    out = []
    for script, num_steps in (([(1.0, False), (2.0, False), (3.0, True), (4.0, False)], None),
                              ([(1.0, False), (2.0, False), (3.0, False), (4.0, True)], 2),
                              ([(5.0, True)], None)):
        env = ScriptedEnv(script)
        ep = episode_generation.generate_episode(env, ScriptedPolicy(), num_steps=num_steps)
        gen = episode_generation.generate_episodes(ScriptedEnv(script), ScriptedPolicy(), num_steps=num_steps, debug=True)
        two = [next(gen), next(gen)]
        out.append({'script': script, 'num_steps': num_steps, 'keys': sorted(ep.keys()),
                    'transitions': [{'state': t['state'], 'action': t['action'], 'reward': t['reward'], 'info': t['info']}
                                    for t in ep['transitions']],
                    'indices': [i for i, _ in two], 'lengths': [len(e['transitions']) for _, e in two]})
    return out


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_cosim, ref_shim
    ref_shim.install(ref_cosim._PybulletStub())
    os.makedirs(OUT, exist_ok=True)
    rs = np.random.RandomState(2024)

    def dump(name, obj):
        with open(os.path.join(OUT, name), 'w') as f:
            json.dump(obj, f)
        print('wrote', name)
    dump('mesh.json', gen_mesh())
    dump('episodes.json', gen_episodes())
    if '--only-mesh' in sys.argv:
        return
    dump('transformations.json', gen_transformations(rs))
    dump('pose.json', gen_pose(rs))
    # robovat.envs.__init__ imports gym-based envs: stub the package objects
    import importlib
    import types
    import robovat
    for name, sub in (('robovat.envs', 'envs'), ('robovat.envs.push', 'envs/push')):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(ref_shim.REFERENCE, 'robovat', sub)]
        sys.modules[name] = m
        parent, _, child = name.rpartition('.')
        setattr(importlib.import_module(parent), child, m)
    dump('layouts.json', gen_layouts())
    dump('reward.json', gen_reward(rs))
    dump('sampler.json', gen_sampler())
    traces = {}
    for name, bindings, seed, fn in (
            ('push_x', {}, 3, push_body0(1.0, 0.0)),
            ('push_diag', {}, 5, push_body0(-0.7, 0.7)),
            ('push_240hz', {'SIM_TIME_STEP': 1.0 / 240.0}, 7, push_body0(0.0, -1.0)),
            ('push_two_goals', {'NUM_GOAL_STEPS': 2}, 11, push_twice(push_body0(1.0, 0.0), push_body0(0.0, 1.0)))):
        b = dict(bindings)
        if 'SIM_TIME_STEP' in b:
            from robovat_b200 import config
            sim = dict(config.DEFAULT_PUSH_ENV['SIM'])
            sim['TIME_STEP'] = b.pop('SIM_TIME_STEP')
            b['SIM'] = sim
        tr = gen_cosim(b, seed, fn, name)
        for k, v in tr.items():
            traces['%s/%s' % (name, k)] = v
    np.savez_compressed(os.path.join(OUT, 'push_step_trace.npz'), **traces)
    print('wrote push_step_trace.npz (%d arrays)' % len(traces))
    dump('camera.json', gen_camera(rs))
    dump('calibration.json', gen_calibration())


if __name__ == '__main__':
    main()
