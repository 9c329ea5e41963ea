"""Run the reference's OWN PushEnv / SawyerSim / Simulator / ControllableBody classes on top of
our physics backend (the `Physics` plug-in seam, robovat/simulation/simulator.py:45-49).

TEST INFRASTRUCTURE ONLY (authoring container: needs /root/reference).  The reference code is
imported unmodified through oracle/ref_shim.py; `physics_backend='OraclePhysics'` resolves to
robovat_b200.physics.CudaPhysics bound to the CPU oracle world, so every `physics.step()`,
`position_control_array`, `compute_inverse_kinematics`, `get_contact_points` ... the reference
issues lands in our implementation.  oracle/gen_golden.py records traces of this co-simulation;
tests compare the fused device-side control flow (b2o/b2s set_action + env_substeps) against them.
"""
import os
import tempfile
import types

import numpy as np

from oracle import b2o, ref_shim
from robovat_b200 import config as config_lib
from robovat_b200 import physics as physics_lib


class _PybulletStub(object):
    COV_ENABLE_RENDERING = COV_ENABLE_GUI = COV_ENABLE_SHADOWS = 0
    URDF_USE_SELF_COLLISION_EXCLUDE_PARENT = URDF_USE_SELF_COLLISION = 0
    DIRECT, GUI = 0, 1
    JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FIXED, JOINT_POINT2POINT = 0, 1, 4, 5
    LINK_FRAME, POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 1, 2, 0, 1

    def __init__(self):
        self.image_shape = (424, 512)

    def computeViewMatrix(self, cameraEyePosition, cameraTargetPosition, cameraUpVector):
        return [0.0] * 16

    def getCameraImage(self, height, width, viewMatrix=None, projectionMatrix=None, physicsClientId=0):
        return (width, height, np.zeros((height, width, 4), np.uint8), np.full((height, width), 0.5, np.float32),
                np.full((height, width), -1, np.int32))


def reference_robot_config(cfg):
    return ref_shim.attrdict({
        'ARM_URDF': 'robots/sawyer_arm.urdf', 'BASE_URDF': 'robots/sawyer_base.urdf', 'HEAD_URDF': 'robots/sawyer_head.urdf',
        'LIMB_JOINT_NAMES': list(physics_lib.LIMB_JOINT_NAMES),
        'LIMB_NEUTRAL_POSITIONS': [0.0, -1.18, 0.0, 2.18, 0.0, 0.57, 3.3161],
        'END_EFFCTOR_NAME': 'right_hand',
        'L_FINGER_NAME': 'right_gripper_l_finger_joint', 'R_FINGER_NAME': 'right_gripper_r_finger_joint',
        'L_FINGER_TIP_NAME': 'right_gripper_l_finger_tip', 'R_FINGER_TIP_NAME': 'right_gripper_r_finger_tip',
        'OPEN_GRIPPER_WHEN_RESET': True,
        'LIMB_MAX_VELOCITY_RATIO': cfg.ROBOT.LIMB_MAX_VELOCITY_RATIO, 'LIMB_TIMEOUT': cfg.ROBOT.LIMB_TIMEOUT,
        'LIMB_POSITION_THRESHOLD': cfg.ROBOT.LIMB_POSITION_THRESHOLD, 'END_EFFECTOR_STEP': 0.01,
    })


def reference_env_config(cfg):
    """Our config -> the key layout the reference reads."""
    mname = cfg.MOVABLE_NAME.upper()
    m = cfg.MOVABLE[mname]
    d = {
        'DEBUG': False, 'MAX_STEPS': cfg.MAX_STEPS, 'SUCCESS_THRESH': cfg.SUCCESS_THRESH, 'TASK_NAME': cfg.TASK_NAME,
        'LAYOUT_ID': cfg.LAYOUT_ID, 'NUM_GOAL_STEPS': cfg.NUM_GOAL_STEPS,
        'MIN_MOVABLE_BODIES': cfg.MIN_MOVABLE_BODIES, 'MAX_MOVABLE_BODIES': cfg.MAX_MOVABLE_BODIES,
        'MOVABLE_NAME': cfg.MOVABLE_NAME,
        'MOVABLE': {mname: {
            'PATHS': ['movables/%s.urdf' % p for p in m.PATHS], 'TARGET_PATHS': ['movables/%s.urdf' % p for p in m.TARGET_PATHS],
            'SCALE': list(m.SCALE), 'MASS': m.MASS, 'FRICTION': m.FRICTION, 'MARGIN': m.MARGIN,
            'POSE': {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in m.POSE.items()}}},
        'USE_RANDOM_RGBA': False, 'USE_PRESTIGE_OBS': True, 'USE_VISUALIZATION_OBS': False,
        'OBS': {'NUM_POINTS': 8, 'CROP_MIN': None, 'CROP_MAX': None},
        'ACTION': {'CSPACE': {'LOW': list(cfg.ACTION.CSPACE.LOW), 'HIGH': list(cfg.ACTION.CSPACE.HIGH)},
                   'MOTION': dict(cfg.ACTION.MOTION), 'MIN_DELTA_POSITION': cfg.ACTION.MIN_DELTA_POSITION,
                   'MIN_DELTA_ANGLE': cfg.ACTION.MIN_DELTA_ANGLE},
        'ARM': {'FINGER_TIP_OFFSET': cfg.ARM.FINGER_TIP_OFFSET, 'GRIPPER_SAFE_HEIGHT': cfg.ARM.GRIPPER_SAFE_HEIGHT,
                'OFFSTAGE_POSITIONS': list(cfg.ARM.OFFSTAGE_POSITIONS)},
        'TABLE': {'HEIGHT_RANGE': list(cfg.TABLE.HEIGHT_RANGE), 'X_RANGE': cfg.TABLE.X_RANGE, 'Y_RANGE': cfg.TABLE.Y_RANGE},
        'SIM': {'ARM': {'CONFIG': reference_robot_config(cfg)},
                'GROUND': {'PATH': 'scene/ground.urdf', 'POSE': cfg.SIM.GROUND.POSE},
                'TABLE': {'PATH': 'scene/table.urdf', 'POSE': cfg.SIM.TABLE.POSE},
                'WALL': {'USE': False}, 'TILE': {'PATH': 'scene/tile.urdf'},
                'STEPS_CHECK': cfg.SIM.STEPS_CHECK, 'MAX_PHASE_STEPS': cfg.SIM.MAX_PHASE_STEPS,
                'MAX_MOTION_STEPS': cfg.SIM.MAX_MOTION_STEPS, 'MAX_OFFSTAGE_STEPS': cfg.SIM.MAX_OFFSTAGE_STEPS},
        'KINECT2': {'DEPTH': {'HEIGHT': 8, 'WIDTH': 8, 'INTRINSICS': list(cfg.KINECT2.DEPTH.INTRINSICS),
                              'TRANSLATION': list(cfg.KINECT2.DEPTH.TRANSLATION), 'ROTATION': list(cfg.KINECT2.DEPTH.ROTATION),
                              'INTRINSICS_NOISE': None, 'TRANSLATION_NOISE': None, 'ROTATION_NOISE': None}},
        'RECORDING': {'USE': False},
    }
    return ref_shim.attrdict(d)


def make_assets_dir(cfg):
    root = tempfile.mkdtemp(prefix='b2s_ref_assets_')
    names = ['scene/ground.urdf', 'scene/table.urdf', 'scene/tile.urdf', 'robots/sawyer_arm.urdf',
             'robots/sawyer_base.urdf', 'robots/sawyer_head.urdf']
    for mc in cfg.MOVABLE.values():
        names += ['movables/%s.urdf' % p for p in list(mc.PATHS) + list(mc.TARGET_PATHS)]
    for n in names:
        path = os.path.join(root, n)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, 'w').close()
    return root


def make_reference_env(cfg=None, world_factory=None):
    """Returns (reference PushEnv instance, our world, scene).  The world has one environment."""
    cfg = cfg or config_lib.default_push_env_config()
    ref_shim.install(_PybulletStub())
    import sys
    # skip package __init__ files that pull gym-only / hardware-only modules
    import importlib
    import robovat
    import robovat.perception
    for name, sub in (('robovat.envs', 'envs'), ('robovat.envs.push', 'envs/push'),
                      ('robovat.perception.camera', 'perception/camera')):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(ref_shim.REFERENCE, 'robovat', sub)]
            sys.modules[name] = m
            parent, _, child = name.rpartition('.')
            setattr(importlib.import_module(parent), child, m)
    import robovat.perception.camera.camera as ref_camera
    sys.modules['robovat.perception.camera'].Camera = ref_camera.Camera
    sys.modules['robovat.perception.camera'].Kinect2 = type('Kinect2', (ref_camera.Camera,), {})
    from robovat.math import Pose
    import robovat.simulation.physics as ref_physics
    from third_party import transformations as T

    scene = config_lib.build_scene(cfg)
    params = config_lib.build_params(cfg, scene, num_envs=1)
    world = world_factory(params, scene) if world_factory else b2o.OracleWorld(params, scene)

    class OraclePhysics(physics_lib.CudaPhysics):
        def __init__(self, time_step=1e-3, use_visualizer=False, worker_id=0):
            physics_lib.CudaPhysics.__init__(self, world=world, scene=scene, env=0, time_step=time_step,
                                             use_visualizer=use_visualizer, worker_id=worker_id, pose_cls=Pose,
                                             quat_from_euler=lambda r, p, y: T.quaternion_from_euler(r, p, y))

    ref_physics.OraclePhysics = OraclePhysics
    from robovat.simulation.simulator import Simulator
    from robovat.envs.push.push_env import PushEnv
    ref_shim.patch_point_cloud_utils()
    sim = Simulator(assets_dir=make_assets_dir(cfg), physics_backend='OraclePhysics', time_step=cfg.SIM.TIME_STEP,
                    gravity=list(cfg.PHYSICS.GRAVITY))
    env = PushEnv(simulator=sim, config=reference_env_config(cfg), debug=False)
    return env, world, scene
