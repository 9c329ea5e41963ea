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
env = PushEnv(config=cfg, num_envs=envs, seed=0)
env.reset()
w = env.world
gen = torch.Generator(device=w.device); gen.manual_seed(0)
w.action.copy_(bench.heuristic_actions_torch(w.obs_position, w.body_mask, cfg, gen))
w.set_action()
if skip:
    w.env_substeps(skip)          # get every arm into the pre/start/motion phases
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
