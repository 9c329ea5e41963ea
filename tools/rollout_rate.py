"""Throughput of the device-side rollout: B envs, every launch advances every env by `chunk` substeps.

usage: python tools/rollout_rate.py [envs] [chunk] [launches] [num_actions] [free_running] [policy_kind]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from robovat_b200.envs import PushEnv

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 40
A = int(sys.argv[4]) if len(sys.argv) > 4 else 8
free = int(sys.argv[5]) if len(sys.argv) > 5 else 0
kind = int(sys.argv[6]) if len(sys.argv) > 6 else 1
cfg = bench.bench_config(envs)
if os.environ.get('B2S_TASK'):                      # e.g. B2S_TASK=clearing: the task layouts (colliding tiles) with 3 convex movables
    cfg.TASK_NAME, cfg.LAYOUT_ID = os.environ['B2S_TASK'], int(os.environ.get('B2S_LAYOUT', 0))
env = PushEnv(config=cfg, num_envs=envs, seed=0)
env.reset()
w = env.world
w.rollout_begin(A, num_episodes=1 << 20, policy_seed=1, reset_seed=2, max_attempts=2000, free_running=bool(free), policy_kind=kind)
torch.cuda.synchronize()
rates = []
for i in range(launches):
    s0 = w.substeps_executed()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    left = w.rollout_run(chunk=chunk, max_substeps=chunk)
    b.record()
    torch.cuda.synchronize()
    n = w.substeps_executed() - s0
    ms = a.elapsed_time(b)
    rates.append(n / ms / 1e3)
    ph = np.bincount(w.array(6).cpu().numpy(), minlength=11).tolist()
    print('launch %2d: %.2f ms, %d substeps, %.2f M substeps/s, left %d, phases %s' % (i, ms, n, n / ms / 1e3, left, ph))
print('mean of the second half: %.2f M substeps/s' % np.mean(rates[len(rates) // 2:]))
print('episodes per env so far: mean %.2f' % w.array(25).float().mean().item(), 'errors', np.bincount(w.array(13).cpu().numpy()).tolist()[:3])
