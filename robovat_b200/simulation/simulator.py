"""Simulator facade over the batched CUDA world.

Mirrors the parts of robovat/simulation/simulator.py:20-376 that PushEnv, its
observations and the Sawyer wrapper reach (`reset/start/step/num_steps/time_step/
check_contact/check_stable/wait_until_stable`, `bodies`), with every quantity
batched over `num_envs`.  `physics_backend` is fixed: the seam the reference fills
with `getattr(physics, 'BulletPhysics')` (:45-49) is filled by `robovat_b200.world.World`.
"""
import numpy as np
import torch

from robovat_b200 import _capi, config as config_lib
from robovat_b200.world import World


class BodyView(object):
    """Batch view of one body (robovat/simulation/body.py:14-240): arrays of shape [B, ...]."""

    def __init__(self, simulator, name, kind, index):
        self.simulator, self.name, self.kind, self.index = simulator, name, kind, index

    @property
    def uid(self):
        return (self.kind, self.index)

    def _state(self, lo, hi):
        return self.simulator.world.body_state[lo:hi, :, self.index].t().cpu().numpy()

    @property
    def position(self):
        if self.kind == 'movable':
            return self._state(0, 3)
        pose = np.asarray(self.simulator.static_pose(self.index))
        return np.tile(pose[None, :3], (self.simulator.num_envs, 1)) + self.simulator.table_offset(self.index)

    @property
    def quaternion(self):
        if self.kind == 'movable':
            return self._state(3, 7)
        return np.tile(np.asarray(self.simulator.static_pose(self.index))[None, 3:], (self.simulator.num_envs, 1))

    @property
    def linear_velocity(self):
        return self._state(7, 10) if self.kind == 'movable' else np.zeros((self.simulator.num_envs, 3), np.float32)

    @property
    def angular_velocity(self):
        return self._state(10, 13) if self.kind == 'movable' else np.zeros((self.simulator.num_envs, 3), np.float32)


class LinkView(object):
    """End-effector link (robovat/simulation/link.py): `.pose` -> [B, 7], `.position` -> [B, 3]."""

    def __init__(self, simulator):
        self.simulator = simulator

    @property
    def pose(self):
        return self.simulator.world.forward_kinematics()[:, -1, :].cpu().numpy()

    @property
    def position(self):
        return self.pose[:, :3]


class SawyerView(object):
    """The `Sawyer` methods PushEnv uses (robovat/robots/sawyer/sawyer.py:27-120, sawyer_sim.py:186-416)."""

    def __init__(self, simulator):
        self.simulator = simulator
        self.end_effector = LinkView(simulator)
        self.arm = self

    uid = ('arm', 0)

    @property
    def joint_positions(self):
        return self.simulator.world.joint_state[0].t().cpu().numpy()

    def move_to_joint_positions(self, positions, mask=None, **_):
        q = np.broadcast_to(np.asarray(positions, np.float32), (self.simulator.num_envs, 7))
        self.simulator.world.move_to_joint_positions(np.ascontiguousarray(q), mask)

    def move_to_gripper_pose(self, pose, mask=None, straight_line=False, **_):
        if straight_line:
            raise NotImplementedError('straight_line=True (move_along_gripper_path) is not on the PushEnv path')
        p = np.broadcast_to(np.asarray(pose, np.float32), (self.simulator.num_envs, 7))
        self.simulator.world.move_to_gripper_pose(np.ascontiguousarray(p), mask)

    def reset_targets(self, mask=None):
        self.simulator.world.arm_reset_targets(mask)

    def is_limb_ready(self):
        return self.simulator.world.arm_is_ready().cpu().numpy().astype(bool)

    def is_gripper_ready(self):
        w = self.simulator.world
        t = w.params.time_step * w.array(_capi.ARR_NUM_STEPS).cpu().numpy()
        return t >= w.array(_capi.ARR_CTRL_TIME).view(-1, 5)[:, 4].cpu().numpy()


class Simulator(object):
    def __init__(self, config=None, num_envs=1, device=0, with_camera=False, env_id_offset=0,
                 assets_dir=None, physics_backend='CudaPhysics', time_step=None, gravity=None,
                 worker_id=0, use_visualizer=False):
        if physics_backend not in ('CudaPhysics', 'BulletPhysics'):
            raise ValueError('Unrecognized physics backend: %r' % physics_backend)
        if use_visualizer:
            raise NotImplementedError('the debug visualizer is a pybullet GUI feature')
        cfg = config or config_lib.default_push_env_config()
        if time_step is not None:
            cfg.SIM.TIME_STEP = time_step
        if gravity is not None:
            cfg.PHYSICS.GRAVITY = list(gravity)
        self.config = cfg
        self.num_envs = int(num_envs)
        self.scene = config_lib.build_scene(cfg)
        lib = _capi.load()
        self.params = config_lib.build_params(cfg, self.scene, num_envs, env_id_offset=env_id_offset, lib=lib)
        self.world = World(self.params, self.scene, device=device, with_camera=with_camera)
        self.robot = SawyerView(self)
        statics = self.scene.statics
        self._static_index = {s['name']: i for i, s in enumerate(statics)}
        self.ground = BodyView(self, 'ground', 'static', self._static_index['ground'])
        self.table = BodyView(self, 'table', 'static', self._static_index['table'])
        self.movable_bodies = [BodyView(self, 'movable_%d' % i, 'movable', i) for i in range(self.params.max_movables)]
        self._reset_calls = 0

    physics = property(lambda self: self.world)
    time_step = property(lambda self: self.params.time_step)

    @property
    def num_steps(self):
        n = self.world.array(_capi.ARR_NUM_STEPS).cpu().numpy()
        return int(n[0]) if self.num_envs == 1 else n

    def static_pose(self, index):
        return self.scene.statics[index]['pose']

    def table_offset(self, index):
        if self.scene.statics[index]['flags'] & _capi.STATIC_ON_TABLE:
            dz = self.world.array(_capi.ARR_TABLE_DZ).cpu().numpy()
            return np.stack([np.zeros_like(dz), np.zeros_like(dz), dz], axis=1)
        return np.zeros((self.num_envs, 3), np.float32)

    def create_camera(self, cam_cfg):
        from robovat_b200.simulation.camera import CudaCamera
        cam = CudaCamera(self, height=cam_cfg.HEIGHT, width=cam_cfg.WIDTH)
        cam.set_calibration(np.array(cam_cfg.INTRINSICS, np.float32).reshape(3, 3), np.array(cam_cfg.TRANSLATION, np.float32),
                            np.array(cam_cfg.ROTATION, np.float32))
        return cam

    # -- reference lifecycle ------------------------------------------------------------------
    def reset(self):
        pass

    def start(self):
        pass

    def step(self, n=1):
        """Simulator.step (simulator.py:94-103) x n for every env."""
        self.world.step(n)

    def reset_scene(self, seed=0, mask=None, max_retries=8):
        """RobotEnv.reset's scene part: sample, drop, settle; re-sample envs whose bodies fell off
        (`body.position.z < table.position.z`, push_env.py:460-468) or that found no placement with the MARGIN
        clearance.  Only the envs being reset are stepped.  Raises if an env is still invalid after
        `max_retries` re-samples (the reference would loop forever)."""
        w = self.world
        self._reset_calls += 1
        todo = None if mask is None else torch.as_tensor(np.asarray(mask, bool), device=w.device)
        table_z = torch.as_tensor(self.scene.statics[self._static_index['table']]['pose'][2], device=w.device)
        for attempt in range(max_retries + 1):
            m8 = None if todo is None else todo.to(torch.uint8)
            w.reset(seed=seed * 1000003 + self._reset_calls, mask=m8)
            # drop settle (0.1 / 0.1 thresholds, <=500 substeps, push_env.py:443-447) then the final wait
            w.settle(0.1, 0.1, 500, mask=m8)
            w.settle(mask=m8)
            z = w.body_state[2]                                   # [B, N]
            live = w.body_mask.bool()
            bad = ((z < (table_z + w.array(_capi.ARR_TABLE_DZ))[:, None]) & live).any(dim=1)
            bad |= (w.array(_capi.ARR_ERROR_FLAGS) & 128) != 0
            if todo is not None:
                bad &= todo
            if not bool(bad.any()):
                break
            todo = bad
        else:
            raise RuntimeError('reset_scene: %d environments have no valid arrangement after %d re-samples'
                               % (int(bad.sum()), max_retries))
        # capacities are part of the physics: a pair / manifold / contact / colour list that overflowed has dropped
        # contacts (b2s.h error flags 1, 2, 8, 16).  A settled scene that already needs more than the world holds will
        # need it during every push too: say so here, once per reset, instead of stepping wrong physics silently.
        over = int(((w.array(_capi.ARR_ERROR_FLAGS) & (1 | 2 | 8 | 16)) != 0).sum().item())
        if over:
            import warnings
            warnings.warn('reset_scene: %d of %d environments exceeded a capacity of the world (max_pairs / max_manifolds / '
                          'max_contacts, B2SParams); contacts were dropped' % (over, w.B), RuntimeWarning)
        w.begin_episode(mask=None if mask is None else np.asarray(mask, bool))
        w.observe()

    def check_contact(self, entity_a, entity_b=None):
        """simulator.py:246-287 for the two queries PushEnv makes: (arm, table) and (arm, movables)."""
        arm_table, arm_movable = self.world.query_contacts()
        kind_b = entity_b[0].kind if isinstance(entity_b, (list, tuple)) else getattr(entity_b, 'kind', None)
        if getattr(entity_a, 'uid', None) != ('arm', 0) or kind_b not in ('static', 'movable'):
            raise NotImplementedError('only check_contact(arm, table) and check_contact(arm, movables) are on the path')
        out = (arm_table if kind_b == 'static' else arm_movable).cpu().numpy().astype(bool)
        return bool(out[0]) if self.num_envs == 1 else out

    def check_stable(self, body, linear_velocity_threshold, angular_velocity_threshold):
        lin = np.linalg.norm(body.linear_velocity, axis=-1)
        ang = np.linalg.norm(body.angular_velocity, axis=-1)
        out = (lin < linear_velocity_threshold) & (ang < angular_velocity_threshold)
        return bool(out[0]) if self.num_envs == 1 else out

    def wait_until_stable(self, body=None, linear_velocity_threshold=0.005, angular_velocity_threshold=0.005,
                          check_after_steps=100, min_stable_steps=100, max_steps=2000):
        if check_after_steps != 100 or min_stable_steps != 100:
            raise NotImplementedError('check_after_steps/min_stable_steps are world parameters (B2SParams)')
        self.world.settle(linear_velocity_threshold, angular_velocity_threshold, max_steps)

    def close(self):
        self.world.close()
