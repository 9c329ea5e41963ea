"""BASELINE config #4: 128x128 depth + segmentation raster and segmented point cloud at 2048 envs, timed with
CUDA events against the HBM roofline (81 920 B written per env and frame).

usage: python tools/bench_render.py [envs] [size] [repeats]
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from robovat_b200 import config as config_lib
from robovat_b200.assets import quat_from_euler, quat_to_matrix
from robovat_b200.world import World

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kin = dict(config_lib.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=size, WIDTH=size,
           INTRINSICS=[120.0 * size / 128, 0.0, size / 2.0, 0, 120.0 * size / 128, size / 2.0, 0, 0, 1])
cfg = config_lib.default_push_env_config(KINECT2={'DEPTH': kin})
scene = config_lib.build_scene(cfg)
params = config_lib.build_params(cfg, scene, num_envs=envs)
w = World(params, scene, with_camera=True)
w.reset(seed=1)
w.settle(0.1, 0.1, 500)
R = quat_to_matrix(quat_from_euler(np.pi, 0, 0))
t = -R.dot(np.array([0.6, 0.0, 1.1]))
w.set_camera(np.array(kin['INTRINSICS'], np.float64), R.reshape(9), t, per_env=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=w.device)


def timed(fn):
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


w.render(); w.point_cloud(seed=0)
torch.cuda.synchronize()
t_render = timed(w.render)
t_pc = timed(lambda: w.point_cloud(seed=0))
peak = 6542.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
out_bytes = envs * size * size * 5
seg = w.segmask.cpu().numpy()
print(json.dumps({
    'workload': 'PushEnv + %dx%d depth/segmentation camera observation, %d envs' % (size, size, envs),
    'render_ms': t_render, 'render_frames_per_s': envs / (t_render * 1e-3),
    'render_GBps_algorithmic': out_bytes / (t_render * 1e-3) / 1e9, 'render_frac_of_hbm_peak': out_bytes / (t_render * 1e-3) / 1e9 / peak,
    'point_cloud_ms': t_pc, 'point_cloud_read_GBps': out_bytes / (t_pc * 1e-3) / 1e9, 'hbm_peak_GBps': peak,
    'visible_body_pixels_frac': float((seg != 255).mean()), 'l2': 'flushed (256 MB memset) before every timed call'}))
