"""`CudaPhysics`: the physics-backend plug-in the reference resolves by name.

`robovat.simulation.Simulator(physics_backend=...)` looks its backend up with
`getattr(robovat.simulation.physics, physics_backend)` and then talks to it through
the methods of `BulletPhysics` (robovat/simulation/physics/bullet_physics.py:28-1304).
This class implements the subset PushEnv, SawyerSim, ControllableBody, Body/Link/Joint
and the observations actually reach, same names, argument meaning and
error behaviour, over one environment (`env` index, default 0) of a batched world:

    lifecycle   reset start step time time_step num_steps uid set_gravity
    bodies      add_body remove_body get_body_pose get_body_position get_body_linear_velocity
                get_body_angular_velocity get_body_mass set_body_dynamics set_body_color
                get_body_link_indices get_body_joint_indices set_body_pose
    links       get_link_name get_link_pose
    joints      get_joint_name get_joint_limit get_joint_dynamics get_joint_position
                get_joint_velocity set_joint_position enable_joint_sensor
    control     position_control_array compute_inverse_kinematics
    contacts    get_contact_points

With it the reference's own Python (ControllableBody.update, PushEnv._execute_action,
wait_until_stable ...) drives the device substep by substep -- the literal drop-in.  The
fast path (robovat_b200.envs.PushEnv) runs the same control flow on the device instead.

`world` is duck-typed: robovat_b200.world.World (CUDA, torch tensors) in the product;
the tests also hand it the CPU oracle's world (numpy views) to generate golden traces.
`pose_cls` lets the caller pass robovat.math.Pose so the reference receives its own type.
"""
import os

import numpy as np

from robovat_b200 import _capi

LIMB_JOINT_NAMES = ['right_j0', 'right_j1', 'right_j2', 'right_j3', 'right_j4', 'right_j5', 'right_j6']
# pybullet numbers a link like the joint that attaches it
ARM_JOINTS = LIMB_JOINT_NAMES + ['right_hand', 'right_gripper_l_finger_joint', 'right_gripper_r_finger_joint',
                                 'right_gripper_l_finger_tip_joint', 'right_gripper_r_finger_tip_joint']
ARM_LINKS = ['right_l0', 'right_l1', 'right_l2', 'right_l3', 'right_l4', 'right_l5', 'right_l6', 'right_hand',
             'right_gripper_l_finger', 'right_gripper_r_finger', 'right_gripper_l_finger_tip',
             'right_gripper_r_finger_tip']
EE_LINK_INDEX = 7
FINGER_LIMITS = {8: (0.0, 0.020833), 9: (-0.020833, 0.0)}


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


class _SimplePose(object):
    def __init__(self, value):
        self.position = np.asarray(value[0], dtype=np.float64)
        self.quaternion = np.asarray(value[1], dtype=np.float32)


class CudaPhysics(object):
    """One environment of a batched world behind the BulletPhysics interface."""

    def __init__(self, world=None, scene=None, env=0, time_step=1e-3, use_visualizer=False, worker_id=0,
                 pose_cls=None, quat_from_euler=None):
        if use_visualizer:
            raise NotImplementedError('the debug visualizer is a pybullet GUI feature')
        if world is None:
            raise ValueError('CudaPhysics needs a world (robovat_b200.world.World)')
        self.world, self.scene, self.env = world, scene, env
        self._uid = worker_id
        self._pose_cls = pose_cls or _SimplePose
        if quat_from_euler is None:
            from robovat_b200.assets import quat_from_euler
        self._quat_from_euler = quat_from_euler
        if abs(world.params.time_step - time_step) > 1e-12:
            raise ValueError('time_step %r differs from the world (%r)' % (time_step, world.params.time_step))
        self._bodies = []          # uid -> dict(kind, index)
        self._static_names = {s['name']: i for i, s in enumerate(scene.statics)}
        self._asset_ids = scene.lib.asset_names
        self._gravity = None

    # ---- helpers -------------------------------------------------------------------------------
    def _arr(self, which):
        return _np(self.world.array(which))

    def _set(self, tensor, index, value):
        if hasattr(tensor, 'detach'):
            import torch
            tensor[index] = torch.as_tensor(value, dtype=tensor.dtype, device=tensor.device)
        else:
            tensor[index] = value

    def _quat(self, pose):
        """quaternion [x,y,z,w] of a robovat Pose / (position, orientation) pair"""
        if hasattr(pose, 'quaternion'):
            return np.asarray(pose.quaternion, dtype=np.float32)
        o = np.asarray(pose[1], dtype=np.float64)
        return np.asarray(self._quat_from_euler(*o) if o.size == 3 else o, dtype=np.float32)

    def _pos(self, pose):
        return np.asarray(pose.position if hasattr(pose, 'position') else pose[0], dtype=np.float32)

    # ---- lifecycle (bullet_physics.py:70-137) --------------------------------------------------
    uid = property(lambda self: self._uid)
    time_step = property(lambda self: self.world.params.time_step)
    gravity = property(lambda self: self._gravity)

    @property
    def num_steps(self):
        return int(self._arr(_capi.ARR_NUM_STEPS)[self.env])

    def time(self):
        return self.time_step * self.num_steps

    def reset(self):
        """resetSimulation: empty scene, step counter 0."""
        w, e = self.world, self.env
        self._bodies = []
        self._set(w.num_movables, e, 0)
        self._set(w.body_mask, e, 0)
        mask = np.zeros(w.B, np.uint8)
        mask[e] = 1
        w.reset(seed=0, mask=mask)                 # clears contacts, controller, counters of this env
        w.arm_reset_targets(mask)                  # the reference's own ControllableBody drives the arm
        self._set(w.num_movables, e, 0)
        self._set(w.body_mask, e, 0)
        w.rebuild_colliders()
        self._removed = set()

    def start(self):
        pass

    def set_gravity(self, gravity):
        g = [float(x) for x in gravity]
        if any(abs(a - b) > 1e-6 for a, b in zip(g, self.world.params.gravity)):
            raise ValueError('gravity %r differs from the world (%r)' % (g, list(self.world.params.gravity)))
        self._gravity = g

    def step(self):
        self.world.step(1)

    # ---- bodies --------------------------------------------------------------------------------
    def add_body(self, filename, pose, scale=1.0, is_static=False):
        name, ext = os.path.splitext(os.path.basename(filename))
        if ext != '.urdf':
            raise ValueError('Unrecognized extension %s.' % ext)
        w, e = self.world, self.env
        uid = len(self._bodies)
        if name in self._static_names:
            idx = self._static_names[name]
            if self.scene.statics[idx]['flags'] & _capi.STATIC_IS_TABLE:
                dz = float(self._pos(pose)[2]) - float(self.scene.statics[idx]['pose'][2])
                self._set(w.array(_capi.ARR_TABLE_DZ), e, dz)
            self._bodies.append({'kind': 'static', 'index': idx})
        elif name.startswith('tile'):
            tiles = [i for i, s in enumerate(self.scene.statics) if s['name'].startswith('tile')]
            used = sum(1 for b in self._bodies if b['kind'] == 'static' and b['index'] in tiles)
            self._bodies.append({'kind': 'static', 'index': tiles[used]})
        elif name.startswith('sawyer_arm') or name == 'sawyer':
            self._bodies.append({'kind': 'arm', 'index': 0})
        elif name.startswith('sawyer_'):
            self._bodies.append({'kind': 'dummy', 'index': 0, 'pose': pose})
        elif name in self._asset_ids:
            n = int(_np(w.num_movables)[e])
            if n >= w.N:
                raise ValueError('more than MAX_MOVABLE_BODIES movables')
            state = np.zeros(13, np.float32)
            state[0:3] = self._pos(pose)
            state[3:7] = self._quat(pose)
            for c in range(13):
                self._set(w.body_state, (c, e, n), state[c])
            mp = w.array(_capi.ARR_MOV_PARAMS)
            mpv = mp.view(4, w.B, w.N) if hasattr(mp, 'view') and hasattr(mp, 'detach') else mp.reshape(4, w.B, w.N)
            asset_bits = np.array([self._asset_ids[name]], np.int32).view(np.float32)[0]
            # URDF defaults of tools/templates/urdf_template.xml: mass 0.1, lateral friction 1.0
            for c, v in enumerate((asset_bits, np.float32(scale), np.float32(0.1), np.float32(1.0))):
                self._set(mpv, (c, e, n), v)
            self._set(w.num_movables, e, n + 1)
            self._set(w.body_mask, (e, n), 1)
            w.rebuild_colliders()
            self._bodies.append({'kind': 'movable', 'index': n})
        else:
            raise AssertionError('File %s does not exist.' % filename)
        return uid

    def remove_body(self, body_uid):
        b = self._bodies[body_uid]
        if b['kind'] != 'movable':
            b['kind'] = 'removed'
            return
        self._removed.add(b['index'])
        b['kind'] = 'removed'
        n = int(_np(self.world.num_movables)[self.env])
        if self._removed >= set(range(n)):           # every movable gone: empty the slots
            self._set(self.world.num_movables, self.env, 0)
            self._set(self.world.body_mask, self.env, 0)
            self.world.rebuild_colliders()
            self._removed = set()

    def _movable_state(self, index):
        return _np(self.world.body_state)[:, self.env, index]

    def get_body_pose(self, body_uid):
        b = self._bodies[body_uid]
        if b['kind'] == 'movable':
            s = self._movable_state(b['index'])
            return self._pose_cls([s[0:3].astype(np.float64), s[3:7]])
        if b['kind'] == 'static':
            p = np.array(self.scene.statics[b['index']]['pose'], np.float64)
            if self.scene.statics[b['index']]['flags'] & _capi.STATIC_ON_TABLE:
                p[2] += float(self._arr(_capi.ARR_TABLE_DZ)[self.env])
            return self._pose_cls([p[0:3], p[3:7].astype(np.float32)])
        base = np.array(self.scene.desc.arm_base_pose[:], np.float64)
        return self._pose_cls([base[0:3], base[3:7].astype(np.float32)])

    def get_body_position(self, body_uid):
        return np.array(self.get_body_pose(body_uid).position, dtype=np.float32)

    def get_body_linear_velocity(self, body_uid):
        b = self._bodies[body_uid]
        if b['kind'] == 'movable':
            return np.array(self._movable_state(b['index'])[7:10], dtype=np.float32)
        return np.zeros(3, np.float32)

    def get_body_angular_velocity(self, body_uid):
        b = self._bodies[body_uid]
        if b['kind'] == 'movable':
            return np.array(self._movable_state(b['index'])[10:13], dtype=np.float32)
        return np.zeros(3, np.float32)

    def get_body_mass(self, body_uid):
        b = self._bodies[body_uid]
        if b['kind'] != 'movable':
            return 0.0
        return float(self._arr(_capi.ARR_MOV_PARAMS).reshape(4, self.world.B, self.world.N)[2, self.env, b['index']])

    def set_body_pose(self, body_uid, pose):
        b = self._bodies[body_uid]
        if b['kind'] != 'movable':
            raise NotImplementedError('only movables can be re-posed')
        for c, v in enumerate(np.r_[self._pos(pose), self._quat(pose)]):
            self._set(self.world.body_state, (c, self.env, b['index']), np.float32(v))

    def set_body_dynamics(self, body_uid, mass=None, lateral_friction=None, rolling_friction=None,
                          spinning_friction=None):
        b = self._bodies[body_uid]
        if b['kind'] != 'movable':
            return
        mp = self.world.array(_capi.ARR_MOV_PARAMS)
        mpv = mp.view(4, self.world.B, self.world.N) if hasattr(mp, 'detach') else mp.reshape(4, self.world.B, self.world.N)
        if mass is not None:
            self._set(mpv, (2, self.env, b['index']), np.float32(mass))
        if lateral_friction is not None:
            self._set(mpv, (3, self.env, b['index']), np.float32(lateral_friction))
        # rolling / spinning friction are one value per world (B2SParams.rolling_friction / spinning_friction = the
        # 0.001 / 0.001 every URDF of tools/templates/urdf_template.xml carries; PushEnv passes None, push_env.py:456-457)
        P = self.world.params
        for name, given, have in (('rolling_friction', rolling_friction, P.rolling_friction),
                                  ('spinning_friction', spinning_friction, P.spinning_friction)):
            if given is not None and abs(float(given) - float(have)) > 1e-9:
                raise NotImplementedError('%s per body (%g) differs from the world value %g: set PHYSICS.%s'
                                          % (name, given, have, name.upper()))

    def set_body_color(self, body_uid, rgba, specular):
        pass

    def get_body_link_indices(self, body_uid):
        return list(range(len(ARM_LINKS))) if self._bodies[body_uid]['kind'] == 'arm' else []

    def get_body_joint_indices(self, body_uid):
        return list(range(len(ARM_JOINTS))) if self._bodies[body_uid]['kind'] == 'arm' else []

    # ---- links / joints ------------------------------------------------------------------------
    def get_link_name(self, link_uid):
        return ARM_LINKS[link_uid[1]]

    def get_link_pose(self, link_uid):
        """World link frame (fields 4,5 of getLinkState, bullet_physics.py:460-473); only the end effector is posed."""
        if link_uid[1] != EE_LINK_INDEX:
            raise NotImplementedError('only the end-effector link pose is on the path')
        lp = _np(self.world.forward_kinematics())[self.env, -1]
        return self._pose_cls([lp[0:3].astype(np.float64), lp[3:7]])

    def get_joint_name(self, joint_uid):
        return ARM_JOINTS[joint_uid[1]]

    def get_joint_limit(self, joint_uid):
        j = joint_uid[1]
        d = self.scene.desc
        if j < 7:
            return {'lower': d.joint_lower[j], 'upper': d.joint_upper[j], 'effort': 80.0, 'velocity': d.joint_max_velocity[j]}
        lo, hi = FINGER_LIMITS.get(j, (0.0, 0.0))
        return {'lower': lo, 'upper': hi, 'effort': 20.0, 'velocity': 5.0}

    def get_joint_dynamics(self, joint_uid):
        return {'damping': 0.0, 'friction': 0.0}

    def get_joint_position(self, joint_uid):
        j = joint_uid[1]
        if j < 7:
            return np.float32(_np(self.world.joint_state)[0, j, self.env])
        return np.float32(FINGER_LIMITS.get(j, (0.0, 0.0))[1 if j == 8 else 0])

    def get_joint_velocity(self, joint_uid):
        j = joint_uid[1]
        return np.float32(_np(self.world.joint_state)[1, j, self.env]) if j < 7 else np.float32(0.0)

    def set_joint_position(self, joint_uid, position):
        j = joint_uid[1]
        if j < 7:                                                # resetJointState: velocity 0
            self._set(self.world.joint_state, (0, j, self.env), np.float32(position))
            self._set(self.world.joint_state, (1, j, self.env), np.float32(0.0))

    def enable_joint_sensor(self, joint_uid):
        pass

    # ---- control -------------------------------------------------------------------------------
    def position_control_array(self, body_uid, joint_inds, target_positions, target_velocities=None,
                               position_gains=None, velocity_gains=None, max_velocities=None):
        if max_velocities is not None:
            raise NotImplementedError('This is not implemented in pybullet')     # bullet_physics.py:1092-1093
        inds = list(joint_inds)
        if not inds or max(inds) >= 7:
            return                                   # finger motors: the stand-in gripper is rigid
        w = self.world
        for gains, ref in ((position_gains, w.params.position_gain), (velocity_gains, w.params.velocity_gain)):
            if gains is not None and any(abs(g - ref) > 1e-9 for g in gains):
                raise NotImplementedError('per-call motor gains differ from the world parameters')
        q = np.array(_np(w.array(_capi.ARR_CTRL)).reshape(w.B, _capi.CTRL_FLOATS)[:, 18:25], np.float32)
        qd = np.zeros((w.B, 7), np.float32)
        for k, j in enumerate(inds):
            q[self.env, j] = np.float32(target_positions[k])
            if target_velocities is not None:
                qd[self.env, j] = np.float32(target_velocities[k])
        mask = np.zeros(w.B, np.uint8)
        mask[self.env] = 1
        w.set_motor_targets(q, qd, mask)

    def compute_inverse_kinematics(self, link_uid, link_pose, upper_limits=None, lower_limits=None, ranges=None,
                                   damping=None, neutral_positions=None):
        """bullet_physics.py:1203-1262: DLS from the current joint state; 7 limb + 2 finger positions."""
        if link_uid[1] != EE_LINK_INDEX:
            raise NotImplementedError('IK is implemented for the end-effector link')
        w = self.world
        pose = np.zeros((w.B, 7), np.float32)
        pose[:, 3:] = [0, 0, 0, 1]
        pose[self.env, 0:3] = self._pos(link_pose)
        pose[self.env, 3:7] = self._quat(link_pose)
        q0 = np.array(_np(w.joint_state)[0], np.float32)
        q = _np(w.inverse_kinematics(pose, q0))[:, self.env]
        return [np.float32(v) for v in q] + [self.get_joint_position((body_uid_of(link_uid), 8)),
                                             self.get_joint_position((body_uid_of(link_uid), 9))]

    # ---- contacts ------------------------------------------------------------------------------
    def get_contact_points(self, a_uid, b_uid=None):
        """Contact distances between two bodies (bullet_physics.py:1268-1304 keeps the last tuple field)."""
        if not isinstance(a_uid, int):
            raise ValueError
        if b_uid is not None and not isinstance(b_uid, int):
            raise ValueError
        w, e = self.world, self.env
        M = w.params.max_manifolds
        keys = self._arr(_capi.ARR_MANIFOLD_KEYS).reshape(w.B, M)[e]
        npts = self._arr(_capi.ARR_MANIFOLD_NPTS).reshape(w.B, M)[e]
        pts = self._arr(_capi.ARR_MANIFOLD_PTS).reshape(w.B, M, 4, _capi.CP_FLOATS)[e]
        n = int(self._arr(_capi.ARR_NUM_MANIFOLDS)[e])
        slots = self._arr(_capi.ARR_COL_SLOT).reshape(w.B, -1)[e]
        ns, nl = self.scene.num_statics, self.scene.num_links

        def slot_set(uid):
            b = self._bodies[uid]
            if b['kind'] == 'static':
                return {b['index']}
            if b['kind'] == 'arm':
                return set(range(ns, ns + nl))
            if b['kind'] == 'movable':
                return {ns + nl + b['index']}
            return set()
        sa = slot_set(a_uid)
        sb = None if b_uid is None else slot_set(b_uid)
        out = []
        for m in range(n):
            ca, cb = int(keys[m]) >> 16, int(keys[m]) & 0xffff
            s1, s2 = int(slots[ca]), int(slots[cb])
            hit = (s1 in sa and (sb is None or s2 in sb)) or (s2 in sa and (sb is None or s1 in sb))
            if hit:
                out += [float(pts[m, k, 9]) for k in range(int(npts[m]))]
        return out


def body_uid_of(link_uid):
    return link_uid[0]
