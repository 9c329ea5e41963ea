#!/usr/bin/env python
"""Why is PushEnv.step slower than the device-resident loop? Four drivers of the same env, same workload statistics."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from robovat_b200.envs import PushEnv

B, steps = 4096, 3
cfg = bench.bench_config(B)


def run(mode):
    mode, _, sd = mode.partition(':')
    sd = int(sd or 0)
    env = PushEnv(config=cfg, num_envs=B, seed=17, device=0)
    w = env.world
    obs = env.reset()
    gen = torch.Generator(device=w.device); gen.manual_seed(sd)
    rs = np.random.RandomState(sd)
    for k in range(3 + steps):
        if k == 3:
            torch.cuda.synchronize(); t0 = time.perf_counter(); s0 = w.substeps_executed()
        if mode in ('dev_torch', 'step_torch'):
            act = bench.heuristic_actions_torch(w.obs_position, w.body_mask, cfg, gen)
        else:
            act = bench.heuristic_actions_np(w.obs_position.cpu().numpy(), w.body_mask.cpu().numpy(), cfg, rs)
        if mode.startswith('dev'):
            w.action.copy_(torch.as_tensor(act, device=w.device))
            w.set_action()
            done = 0
            while done < env.max_action_substeps:
                u = w.env_substeps(env.substep_chunk)
                done += env.substep_chunk
                if u == 0:
                    break
            w.observe(); w.reward()
        else:
            env._done[:] = False
            env.step(act.cpu().numpy() if torch.is_tensor(act) else act)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('%-11s seed %d: %.1f ms/step, %.0f substeps/step, act mean %s' % (mode, sd, 1e3 * dt / steps, (w.substeps_executed() - s0) / steps,
                                                                   np.round(np.asarray(act.cpu() if torch.is_tensor(act) else act).mean(0), 3)), flush=True)


for m in sys.argv[1:] or ['dev_torch', 'step_np', 'step_torch', 'dev_np']:
    run(m)
