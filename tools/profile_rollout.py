"""Short driver for ncu: 4096 envs in the steady state of the device-side rollout, a few launches of the substep kernel.

usage: python tools/profile_rollout.py [envs] [substeps_per_launch] [launches] [warm_launches_of_250] [free_running]
The launch between cudaProfilerStart/Stop (the second timed one) is what `ncu --profile-from-start off` captures.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from robovat_b200 import _capi
from robovat_b200.envs import PushEnv

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
per = int(sys.argv[2]) if len(sys.argv) > 2 else 50
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 4
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 24
free = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cfg = bench.crossing_config() if os.environ.get('B2S_CFG') == 'crossing' else bench.bench_config(envs)
env = PushEnv(config=cfg, num_envs=envs, seed=0)
env.reset()
w = env.world
w.rollout_begin(cfg.MAX_STEPS, num_episodes=1 << 20, policy_seed=bench.POLICY_SEED, reset_seed=bench.RESET_SEED,
                policy_kind=_capi.POLICY_AIMED, free_running=bool(free))
for _ in range(warm):
    w.rollout_run(chunk=250, max_substeps=250)
torch.cuda.synchronize()
w.array(24).zero_()
torch.cuda.synchronize()
t = []
for i in range(launches):
    s0 = w.substeps_executed()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if i == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    a.record(); w.rollout_run(chunk=per, max_substeps=per); b.record()
    torch.cuda.synchronize()
    if i == 1:
        torch.cuda.profiler.stop()
    t.append((a.elapsed_time(b), w.substeps_executed() - s0))
ph = w.array(6).cpu().numpy()
print('phase histogram', np.bincount(ph, minlength=11).tolist())
print('ms per launch', [round(x[0], 3) for x in t], 'env-substeps per launch', [x[1] for x in t])
print('M substeps/s', [round(x[1] / x[0] / 1e3, 2) for x in t])
st = w.array(18).view(envs, 4).cpu().numpy()
print('mean rows %.1f colours %.2f iters %.1f contacts %.1f' % tuple(st.mean(axis=0)))
print('pairs mean %.2f manifolds mean %.2f' % (w.array(5).float().mean().item(), w.array(3).float().mean().item()))
prof = w.array(24).cpu().numpy().astype(float)
if prof[6] > 0:
    r = prof[6]
    print('stage ns per round (block mean): A %.0f B %.0f C %.0f | warp-busy fraction A %.2f B %.2f C %.2f | longest env in C %.0f' % (
        prof[0] / r, prof[1] / r, prof[2] / r, prof[3] / (16 * prof[0]), prof[4] / (16 * prof[1]), prof[5] / (16 * prof[2]), prof[7] / r))
    pb = prof[8:8 + 4 * 1024].reshape(1024, 4)
    pb = pb[pb[:, 3] > 0]
    tot = pb[:, :3].sum(axis=1) / 1e6
    print('blocks %d: total ms min %.1f median %.1f max %.1f; rounds min %d median %d max %d' % (
        len(pb), tot.min(), np.median(tot), tot.max(), pb[:, 3].min(), np.median(pb[:, 3]), pb[:, 3].max()))
    sec = prof[8 + 4096:8 + 4096 + 16]
    n = max(1.0, float(sum(x[1] for x in t)))
    print('sections, ns per env-substep (warp time): ' + ' '.join('%d:%.0f' % (i, v / n) for i, v in enumerate(sec) if v > 0))
    if len(prof) >= 8 + 4096 + 16 + 128:
        h = prof[8 + 4096 + 16:8 + 4096 + 16 + 128]
        print('stage-C time of an env, 2 us bins:', h[:64].astype(int).tolist())
        print('start of an env within stage C, 4 us bins:', h[64:96].astype(int).tolist())
        print('stage-C time of the env handed out first, 4 us bins:', h[96:128].astype(int).tolist())
