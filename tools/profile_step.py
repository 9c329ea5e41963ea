"""Short driver for ncu: 4096 envs mid-push, a few launches of the substep kernel.

usage: python tools/profile_step.py [envs] [substeps_per_launch] [launches] [skip_substeps]
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from robovat_b200.envs import PushEnv

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
per = int(sys.argv[2]) if len(sys.argv) > 2 else 50
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 4
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 600
cfg = bench.bench_config(envs)
if os.environ.get('B2S_CFG') == 'crossing':          # BASELINE configs[2]: crossing layout 0, 8 concave (multi-hull) movables
    from robovat_b200 import config as config_lib
    cfg = config_lib.default_push_env_config(TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME=os.environ.get('B2S_MOVABLE', 'concave'),
                                             MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8)
    cfg.SIM.TIME_STEP = 1.0 / 240.0
env = PushEnv(config=cfg, num_envs=envs, seed=0)
env.reset()
w = env.world
gen = torch.Generator(device=w.device); gen.manual_seed(0)
w.action.copy_(bench.heuristic_actions_torch(w.obs_position, w.body_mask, cfg, gen))
w.set_action()
if skip:
    w.env_substeps(skip)          # get every arm into the pre/start/motion phases
torch.cuda.synchronize()
w.array(24).zero_()
torch.cuda.synchronize()
t = []
for i in range(launches):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if i == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()      # ncu --profile-from-start off captures this launch
    a.record(); w.env_substeps(per, sync=False); b.record()
    if i == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
    t.append((a, b))
torch.cuda.synchronize()
ph = w.array(6).cpu().numpy()
print('phase histogram', np.bincount(ph, minlength=9).tolist())
print('ms per launch', [round(a.elapsed_time(b), 3) for a, b in t])
st = w.array(18).view(envs, 4).cpu().numpy()
print('mean rows %.1f colours %.2f iters %.1f contacts %.1f' % tuple(st.mean(axis=0)))
print('pairs mean %.2f manifolds mean %.2f' % (w.array(5).float().mean().item(), w.array(3).float().mean().item()))
prof = w.array(24).cpu().numpy().astype(float)
if prof[6] > 0:
    blocks = (envs + w.params.reserved_i[0] - 1) // w.params.reserved_i[0] if w.params.reserved_i[0] else 147
    r = prof[6]
    print('stage ns per round (block mean): A %.0f B %.0f C %.0f | warp-busy fraction A %.2f B %.2f C %.2f | longest env in C %.0f' % (
        prof[0] / r, prof[1] / r, prof[2] / r, prof[3] / (16 * prof[0]), prof[4] / (16 * prof[1]), prof[5] / (16 * prof[2]), prof[7] / r))
    pb = prof[8:8 + 4 * 1024].reshape(1024, 4)
    pb = pb[pb[:, 3] > 0]
    tot = pb[:, :3].sum(axis=1) / 1e6
    order = np.argsort(tot)
    print('blocks %d: total ms min %.1f median %.1f max %.1f' % (len(pb), tot.min(), np.median(tot), tot.max()))
    for i in list(order[:2]) + list(order[-4:]):
        print('  block %4d  A %.1f B %.1f C %.1f ms  rounds %d' % (i, pb[i, 0] / 1e6, pb[i, 1] / 1e6, pb[i, 2] / 1e6, pb[i, 3]))
    sec = prof[8 + 4096:8 + 4096 + 12] / 1e6
    names = ['C rows', 'C colour', 'C sweeps', 'C integrate', 'B lookup/refresh', 'B collide', 'B manifold', 'B tail', 'A arm+fk', 'A bodies', 'A colliders', 'A broad']
    print('warp-ms by section: ' + ', '.join('%s %.0f' % (n, v) for n, v in zip(names, sec)))
