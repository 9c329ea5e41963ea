"""What the raster's tile classification decided (tuning): class histogram and list lengths of env 0..k.
usage: python tools/render_lists.py [envs] [size]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from robovat_b200 import _capi, config as config_lib
from robovat_b200.assets import quat_from_euler, quat_to_matrix
from robovat_b200.world import World

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kin = dict(config_lib.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=size, WIDTH=size,
           INTRINSICS=[120.0 * size / 128, 0.0, size / 2.0, 0, 120.0 * size / 128, size / 2.0, 0, 0, 1])
cfg = config_lib.default_push_env_config(KINECT2={'DEPTH': kin})
scene = config_lib.build_scene(cfg)
params = config_lib.build_params(cfg, scene, num_envs=envs)
w = World(params, scene, with_camera=True)
w.reset(seed=1)
w.settle(0.1, 0.1, 500)
R = quat_to_matrix(quat_from_euler(np.pi, 0, 0))
w.set_camera(np.array(kin['INTRINSICS'], np.float64), R.reshape(9), -R.dot(np.array([0.6, 0.0, 1.1])), per_env=False)
w.render()
torch.cuda.synchronize()
raw = w.array(_capi.ARR_RAY_SCENE).cpu().numpy()
per_env = raw.size // envs
tiles = ((size + 31) // 32) * ((size + 7) // 8)
lib = _capi.load()
# layout: planes[max_planes] float4, cols[max_cols] 28 B, cnt[tiles] int, list[tiles][max_cols] u16 -- sizes from the counts
found = None
for max_cols in range(1, 257):
    rest = per_env - max_cols * 28 - tiles * 4 - tiles * max_cols * 2
    for pad in range(16):
        if rest - pad > 0 and (rest - pad) % 16 == 0:
            max_planes = (rest - pad) // 16
            e = raw[:per_env]
            off = max_planes * 16
            cols = e[off:off + max_cols * 28].view(np.int32).reshape(max_cols, 7)
            cnt = e[off + max_cols * 28:off + max_cols * 28 + tiles * 4].view(np.int32)
            if cols[0, 0] == 0 and 0 < cols[0, 1] <= 64 and cnt.min() >= 1 and cnt.max() <= max_cols and (np.diff(cols[:3, 0]) > 0).all():
                found = (max_planes, max_cols)
    if found:
        break
max_planes, max_cols = found
print('per env bytes', per_env, 'max_planes', max_planes, 'max_cols', max_cols, 'tiles', tiles)
e = raw[:per_env]
off = max_planes * 16
cols = e[off:off + max_cols * 28].view(np.int32).reshape(max_cols, 7)
cnt = e[off + max_cols * 28:off + max_cols * 28 + tiles * 4].view(np.int32)
lst = e[off + max_cols * 28 + tiles * 4:off + max_cols * 28 + tiles * 4 + tiles * max_cols * 2].view(np.uint16).reshape(tiles, max_cols)
print('planes per hull', (cols[:, 1] - cols[:, 0]).tolist())
print('list length per tile: mean %.2f max %d' % (cnt.mean(), cnt.max()))
full = fast = 0
per_hull = {}
for t in range(tiles):
    for k in range(cnt[t]):
        c, cl = int(lst[t, k]) & 255, int(lst[t, k]) >> 8
        per_hull.setdefault(c, [0, 0])[0 if cl == 255 else 1] += 1
        full += cl == 255
        fast += cl != 255
print('entries: full %d fast %d' % (full, fast))
print('per hull (full, fast):', {k: tuple(v) for k, v in sorted(per_hull.items())})
