"""`CudaCamera`: the camera class ArmEnv._create_camera would build (envs/arm_env.py:38-71).

Same surface as the reference's `BulletCamera` over `Camera`
(robovat/simulation/camera/bullet_camera.py:117-258, robovat/perception/camera/camera.py:17-244):
`height/width/intrinsics/translation/rotation/cx/cy/pose`, `set_calibration(K, t, R)`,
`frames() -> {'rgb', 'depth', 'segmask'}`, `project_point`, `deproject_pixel`,
`deproject_depth_image`.  `frames()` runs the ray-cast kernel (b2s_render) for every
environment; the projection helpers are small host-side numpy like the reference's.
RGB is not produced by the device path (north_star asks for depth + segmentation); `frames()`
returns a zero image under 'rgb' so callers that only forward it keep working.
"""
import numpy as np

from robovat_b200.assets import quat_from_euler, quat_to_matrix

NEAR_PLANE, FAR_PLANE = 0.02, 100          # bullet_camera.py:18-19
DEPTH_HEIGHT, DEPTH_WIDTH = 424, 512       # bullet_camera.py:22-23


def _rotation_matrix(rotation):
    r = np.asarray(rotation, dtype=np.float64)
    if r.size == 9:
        return r.reshape(3, 3)
    if r.size == 4:
        return quat_to_matrix(r)
    return quat_to_matrix(quat_from_euler(*r))


def draw_calibration(intrinsics, translation, rotation, intrinsics_noise=None, translation_noise=None, rotation_noise=None,
                     rs=np.random):
    """The calibration ArmEnv._reset_camera hands to camera.set_calibration in simulation (arm_env.py:109-152): float32
    copies of the configured values plus uniform noise in [-noise, noise], drawn from `rs` in the reference's order
    (intrinsics, translation, rotation)."""
    intrinsics = np.copy(intrinsics).astype(np.float32)
    translation = np.copy(translation).astype(np.float32)
    rotation = np.copy(rotation).astype(np.float32)
    if intrinsics_noise is not None:
        intrinsics += rs.uniform(-np.array(intrinsics_noise), np.array(intrinsics_noise))
    if translation_noise is not None:
        translation += rs.uniform(-np.array(translation_noise), np.array(translation_noise))
    if rotation_noise is not None:
        rotation += rs.uniform(-np.array(rotation_noise), np.array(rotation_noise))
    return intrinsics, translation, rotation


class CudaCamera(object):
    def __init__(self, simulator, height=DEPTH_HEIGHT, width=DEPTH_WIDTH, intrinsics=None, translation=None,
                 rotation=None, crop=None, near=NEAR_PLANE, far=FAR_PLANE, distance=1.0, upside_down=True):
        if crop is not None:
            raise NotImplementedError('crop is a real-camera (Kinect2) feature')
        self._simulator = simulator
        self._height, self._width = int(height), int(width)
        p = simulator.world.params
        if (p.cam_height, p.cam_width) != (self._height, self._width):
            raise ValueError('image size %dx%d differs from the world (%dx%d)' % (
                self._height, self._width, p.cam_height, p.cam_width))
        if abs(p.cam_near - near) > 1e-9 or abs(p.cam_far - far) > 1e-6:
            raise ValueError('near/far differ from the world parameters')
        self._intrinsics = self._translation = self._rotation = None
        self.set_calibration(intrinsics, translation, rotation)

    simulator = property(lambda self: self._simulator)
    height = property(lambda self: self._height)
    width = property(lambda self: self._width)
    intrinsics = property(lambda self: self._intrinsics)
    translation = property(lambda self: self._translation)
    rotation = property(lambda self: self._rotation)
    cx = property(lambda self: self._intrinsics[0, 2])
    cy = property(lambda self: self._intrinsics[1, 2])

    @property
    def pose(self):
        """Camera pose in the world: inverse of (t, R) (camera.py:78-81, pose.py:161-172)."""
        return -self._translation.dot(self._rotation), self._rotation.T

    def start(self):
        pass

    def stop(self):
        return True

    def reset(self):
        pass

    def set_calibration(self, intrinsics, translation, rotation):
        """x_cam = R x_world + t; one calibration for every environment, or [B, ...] arrays for per-env noise."""
        if intrinsics is not None:
            self._intrinsics_all = np.asarray(intrinsics, dtype=np.float64).reshape(-1, 3, 3)
            self._intrinsics = self._intrinsics_all[0]
        if translation is not None:
            self._translation_all = np.asarray(translation, dtype=np.float64).reshape(-1, 3)
            self._translation = self._translation_all[0]
        if rotation is not None:
            r = np.asarray(rotation, dtype=np.float64)
            if r.shape in ((3,), (4,), (9,), (3, 3)):
                rs = [r]                                   # Euler / quaternion / matrix shared by all envs
            elif r.ndim == 3:
                rs = list(r)                               # [B, 3, 3]
            else:
                rs = list(r.reshape(r.shape[0], -1))       # [B, 3] Euler, [B, 4] quaternion or [B, 9]
            self._rotation_all = np.stack([_rotation_matrix(x) for x in rs])
            self._rotation = self._rotation_all[0]
        if self._intrinsics is not None and self._translation is not None and self._rotation is not None:
            B = self._simulator.num_envs
            n = max(len(self._intrinsics_all), len(self._translation_all), len(self._rotation_all))
            if n not in (1, B):
                raise ValueError('calibration must be shared or given per environment')
            per_env = n > 1

            def full(a):
                return np.repeat(a, n, axis=0) if len(a) == 1 and n > 1 else a
            self._simulator.world.set_camera(full(self._intrinsics_all).reshape(n, 9), full(self._rotation_all).reshape(n, 9),
                                             full(self._translation_all), per_env=per_env)

    def frames(self):
        depth, seg = self._simulator.world.render()
        depth, seg = depth.cpu().numpy(), seg.cpu().numpy()
        rgb = np.zeros(depth.shape + (3,), np.uint8)
        if self._simulator.num_envs == 1:
            return {'rgb': rgb[0], 'depth': depth[0], 'segmask': seg[0]}
        return {'rgb': rgb, 'depth': depth, 'segmask': seg}

    def project_point(self, point, is_world_frame=True):
        """camera.py:170-192"""
        point = np.array(point, dtype=np.float64)
        if is_world_frame:
            pos, mat = self.pose
            point = np.dot(point - pos, mat)
        proj = np.dot(point, self._intrinsics.T)
        proj = np.round(proj / proj[..., 2:3])
        return np.array(proj[..., :2]).astype(np.int16)

    def deproject_pixel(self, pixel, depth, is_world_frame=True):
        """camera.py:194-211"""
        point = depth * np.linalg.inv(self._intrinsics).dot(np.r_[pixel, 1.0])
        if is_world_frame:
            pos, mat = self.pose
            point = pos + np.dot(point, mat.T)
        return point

    def deproject_depth_image(self, image, crop=None, is_world_frame=True):
        """camera.py:213-244: [H*W, 3] points, pixel (row v, col u) -> depth * K^-1 [u, v, 1]"""
        h, w = image.shape[0], image.shape[1]
        v, u = np.indices((h, w)).reshape(2, -1)
        pix = np.stack([u, v, np.ones_like(u)], axis=0) * image.reshape(1, -1)
        pc = np.linalg.inv(self._intrinsics).dot(pix)
        if crop is not None:
            raise NotImplementedError('crop')
        if is_world_frame:
            pos, mat = self.pose
            pc = pos.reshape(3, 1) + mat.dot(pc)
        return np.array(pc.T)
