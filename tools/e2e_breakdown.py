#!/usr/bin/env python
"""Where the time of one batched PushEnv.step goes on the host side (GPU needed). usage: python tools/e2e_breakdown.py [envs] [steps]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from robovat_b200.envs import PushEnv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.bench_config(B)
env = PushEnv(config=cfg, num_envs=B, seed=17, device=0)
acc = {}


def timed(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, wrap)


timed(env, '_execute_action'); timed(env, '_refresh_attributes'); timed(env, 'get_observation')
timed(env._reward_fns[0], 'get_reward'); timed(env.world, 'env_substeps'); timed(env.world, 'set_action')
rs = np.random.RandomState(0)
obs = env.reset()
for k in range(3 + steps):
    if k == 3:
        acc.clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
    env._done[:] = False
    act = bench.heuristic_actions_np(np.asarray(obs['position']).reshape(B, -1, 3), np.asarray(obs['body_mask']).reshape(B, -1), cfg, rs)
    obs, rew, done, _ = env.step(act)
torch.cuda.synchronize(); total = time.perf_counter() - t0
print('total ms/step %.1f' % (1e3 * total / steps))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print('  %-22s %.1f ms/step' % (k, 1e3 * v / steps))
