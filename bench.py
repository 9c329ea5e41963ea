#!/usr/bin/env python
"""Headline benchmark: PushEnv substeps/s at 4096 envs per GPU (BASELINE.json configs[1]).

One bench "step" = one `PushEnv.step` for every env of the rank: set the actions,
run the device-side phase machine + physics until every env finished its push and
settled, gather PoseObs and the reward.  Reported:
  value   substeps/s with the actions already on the device (heuristic push actions
          are derived from the device-resident observation with a few torch ops)
  e2e     the same metric through the public API `robovat_b200.envs.PushEnv.step`
          with HOST actions produced by the host policy from the HOST observation
          (pinned H2D of the actions, D2H of observation + reward inside the timing)
  roofline / cpu_baseline as the task contract describes (Measurement).
`--impl reference` times the CPU oracle port (PyBullet is not installable here,
BASELINE.md section 3) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'PushEnv substeps/sec at 4096 envs'
UNIT = 'substeps/s'
WORKLOAD = 'PushEnv 4096 batched envs, 3 convex movables, 240 Hz substeps'
# whole-substep algorithmic traffic B_step = 104*Nb + 24*C + 84 bytes with
# Nb = 3 movables + 1 pusher, C = 16 contact points
ALGO_BYTES_PER_SUBSTEP = 104 * 4 + 24 * 16 + 84


def bench_config(num_envs):
    from robovat_b200 import config
    cfg = config.default_push_env_config()
    cfg.SIM.TIME_STEP = 1.0 / 240.0
    cfg.MIN_MOVABLE_BODIES = cfg.MAX_MOVABLE_BODIES = 3
    return cfg


def heuristic_actions_np(position, mask, cfg, rs):
    """Vectorised stand-in for HeuristicPushPolicy: start 8 cm behind a random body, push through it."""
    B, N, _ = position.shape
    lo, hi = np.array(cfg.ACTION.CSPACE.LOW[:2], np.float32), np.array(cfg.ACTION.CSPACE.HIGH[:2], np.float32)
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    nb = np.maximum(mask.sum(axis=1).astype(np.int64), 1)
    body = rs.randint(0, 1 << 30, size=B) % nb
    ang = rs.uniform(-np.pi, np.pi, size=B).astype(np.float32)
    d = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    tgt = position[np.arange(B), body, :2] - 0.08 * d
    a = np.concatenate([np.clip((tgt - off) / rng, -1, 1), d], axis=1).astype(np.float32)
    return a


def heuristic_actions_torch(position, mask, cfg, gen):
    import torch
    B = position.shape[0]
    dev = position.device
    lo = torch.tensor(cfg.ACTION.CSPACE.LOW[:2], device=dev)
    hi = torch.tensor(cfg.ACTION.CSPACE.HIGH[:2], device=dev)
    off, rng = 0.5 * (lo + hi), 0.5 * (hi - lo)
    nb = mask.sum(dim=1).clamp(min=1).long()
    body = torch.randint(0, 1 << 30, (B,), device=dev, generator=gen) % nb
    ang = (torch.rand(B, device=dev, generator=gen) * 2 - 1) * np.pi
    d = torch.stack([torch.cos(ang), torch.sin(ang)], dim=1)
    tgt = position[torch.arange(B, device=dev), body, :2] - 0.08 * d
    return torch.cat([((tgt - off) / rng).clamp(-1, 1), d], dim=1).float()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic():
    """DRAM bytes per env-substep of k_substeps from the committed ncu --set full capture (None when absent)."""
    path = os.path.join(ROOT, 'profiles', 'r01_k_substeps_dram_traffic.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['dram_bytes_per_env_substep'])
    return None


def run_cpu_sample(cfg, threads, target_seconds, seed):
    """Oracle port on the host cores over a bounded sample of the same workload."""
    from oracle import b2o
    from robovat_b200 import config
    b2o.build()
    scene = config.build_scene(cfg)

    def one(num_envs, steps):
        params = config.build_params(cfg, scene, num_envs=num_envs)
        w = b2o.OracleWorld(params, scene, threads=threads)
        w.reset(seed=seed)
        w.settle(0.1, 0.1, 500)
        w.settle()
        rs = np.random.RandomState(seed)
        t0 = time.perf_counter()
        s0 = w.substeps_executed()
        for _ in range(steps):
            pos = w.observe().copy()
            mask = w.array('body_mask').reshape(num_envs, -1)
            w.set_action(heuristic_actions_np(pos, mask, cfg, rs))
            while w.env_substeps(500) > 0:
                pass
            w.observe()
            w.reward()
        dt = time.perf_counter() - t0
        n = w.substeps_executed() - s0
        w.close()
        return n, dt
    n, dt = one(max(threads * 4, 16), 1)                      # calibration
    rate = n / max(dt, 1e-9)
    per_env_step = n / float(max(threads * 4, 16))
    envs = int(min(4096, max(threads * 8, target_seconds * rate / max(per_env_step, 1.0))))
    n, dt = one(envs, 1)
    return n / dt, '%d envs x 1 PushEnv.step (%d substeps, %.1f s) on %d threads' % (envs, n, dt, threads)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='envs per GPU')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = bench_config(args.envs)
    threads = os.cpu_count() or 1

    if args.impl == 'reference':
        if rank != 0:
            return 0
        values = []
        sample = ''
        for _ in range(max(1, min(args.steps, 3))):
            v, sample = run_cpu_sample(cfg, threads, max(3.0, args.cpu_seconds / max(1, min(args.steps, 3))), args.seed)
            values.append(v)
        v = float(np.mean(values))
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'note': 'CPU oracle port of the same path; pybullet==2.6.5 is not installable here'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        }))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port')
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from robovat_b200 import _capi
    from robovat_b200.envs import PushEnv

    B = args.envs
    env = PushEnv(config=cfg, num_envs=B, seed=args.seed + 17 * rank, device=local_rank, env_id_offset=rank * B)
    w = env.world
    env.reset()
    dev = w.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    # Returns of all ranks (tools/parallel_run.py:54-90 collects the workers' results; its workers never wait for one
    # another).  The all-gather of rollout k is asynchronous and double buffered: it travels on NCCL's stream while
    # rollout k+1 runs, so a rank only waits for a peer that is two rollouts behind; the last ones are waited for
    # inside the timed region.
    returns_all = [torch.zeros(world_size * B, dtype=torch.float32, device=dev) for _ in range(2)]
    returns_mine = [torch.zeros(B, dtype=torch.float32, device=dev) for _ in range(2)]
    gathers = [None, None]
    gather_count = [0]

    def gather_returns(world, last=False):
        if world_size == 1:
            return
        i = gather_count[0] & 1
        gather_count[0] += 1
        if gathers[i] is not None:
            gathers[i].wait()
        returns_mine[i].copy_(world.episode_return.reshape(-1))
        gathers[i] = dist.all_gather_into_tensor(returns_all[i], returns_mine[i], async_op=True)
        if last:
            for g in gathers:
                if g is not None:
                    g.wait()

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_events = []

    def host_actions(world, rs):
        """The policy stand-in: host arithmetic on the host copy of the observation (identical in both legs)."""
        return heuristic_actions_np(world.obs_position.cpu().numpy(), world.body_mask.cpu().numpy(), cfg, rs)

    def device_step(timed, act, last=False):
        w.action.copy_(act)
        w.set_action()
        done = 0
        while done < env.max_action_substeps:
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            u = w.env_substeps(env.substep_chunk)
            if timed:
                e1.record()
                kernel_events.append((e0, e1))
            done += env.substep_chunk
            if u == 0:
                break
        w.observe()
        w.reward()
        gather_returns(w, last)

    # ---- device-resident leg -------------------------------------------------------------
    # Both legs run the SAME workload: same env seed, same action stream (the duration of a batched step is set by its
    # slowest env, and two random action streams differ by +-10% in that).  Here the actions are uploaded before the
    # timed region starts (inputs resident in HBM); the end-to-end leg below computes and uploads them inside it.
    rs = np.random.RandomState(args.seed + rank)
    for _ in range(args.warmup):
        device_step(False, torch.from_numpy(host_actions(w, rs)).to(dev))
    barrier()
    s0, l0 = w.substeps_executed(), w.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_events = []
    barrier()
    for k in range(args.steps):
        act = torch.from_numpy(host_actions(w, rs)).to(dev)
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        device_step(True, act, last=(k == args.steps - 1))
        b.record()
        step_events.append((a, b))
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in step_events)
    substeps = w.substeps_executed() - s0
    launches = w.launch_count() - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events)
    n_kernel_launches = len(kernel_events)

    # ---- end-to-end leg through PushEnv.step with host actions -------------------------------
    e2e = None
    if not args.no_e2e:
        env2 = PushEnv(config=cfg, num_envs=B, seed=args.seed + 17 * rank, device=local_rank, env_id_offset=rank * B)
        w2 = env2.world
        rs = np.random.RandomState(args.seed + rank)
        obs = env2.reset()

        def policy(obs):
            return heuristic_actions_np(np.asarray(obs['position']).reshape(B, -1, 3), np.asarray(obs['body_mask']).reshape(B, -1), cfg, rs)
        for _ in range(args.warmup):
            env2._done[:] = False
            obs, _, _, _ = env2.step(policy(obs))
        barrier()
        s1 = w2.substeps_executed()
        t0 = time.perf_counter()
        for k in range(args.steps):
            env2._done[:] = False
            obs, rew, done, _ = env2.step(policy(obs))
            gather_returns(w2, last=(k == args.steps - 1))
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_sub = w2.substeps_executed() - s1
        nmax = w2.N
        e2e = {'substeps': e2e_sub, 'seconds': e2e_s, 'h2d': B * 4 * 4,
               'd2h': B * nmax * 3 * 4 + B * nmax + 2 * B + B * 4 + B + B * 32}

    # ---- max over ranks ------------------------------------------------------------------------
    stats = torch.tensor([ms, float(substeps), kernel_ms, float(e2e['seconds'] if e2e else 0), float(e2e['substeps'] if e2e else 0)],
                         dtype=torch.float64, device=dev)
    if world_size > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, kernel_ms_max = float(mx[0]), float(mx[2])
        total_substeps, e2e_seconds, e2e_substeps = float(sm[1]), float(mx[3]), float(sm[4])
    else:
        total_substeps, e2e_seconds, e2e_substeps, kernel_ms_max = float(substeps), float(stats[3]), float(stats[4]), kernel_ms
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return 0

    value = total_substeps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    achieved = (substeps * ALGO_BYTES_PER_SUBSTEP) / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    per_launch = substeps / float(max(1, n_kernel_launches))          # env-substeps one launch processes on this rank
    traffic = measured_traffic()
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world_size, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'envs_per_gpu': B, 'time_step': cfg.SIM.TIME_STEP, 'solver_iterations': 50,
                   'step': 'one PushEnv.step per env (7-phase push + wait_until_stable)', 'substeps_per_step': total_substeps / args.steps,
                   'l2': 'flushed between timed steps (256 MB memset, untimed)', 'parallelism': 'env-sharded x%d' % world_size},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic * per_launch if traffic else None, 'traffic_unit': 'bytes per launch (ncu dram read+write)',
                     'algorithmic_bytes_per_launch': ALGO_BYTES_PER_SUBSTEP * per_launch, 'env_substeps_per_launch': per_launch,
                     'kernel': 'k_substeps', 'launches': n_kernel_launches,
                     'avg_launch_ms': kernel_ms / max(1, n_kernel_launches), 'kernel_share_of_step': kernel_ms / ms if ms > 0 else None,
                     'algorithmic_bytes_per_substep': ALGO_BYTES_PER_SUBSTEP, 'peak_source': peak_src,
                     'note': 'state is shared-memory/L2 resident by design: the kernel is latency/issue bound, not HBM bound (DESIGN.md)'},
    }
    if e2e:
        out['e2e'] = {'value': e2e_substeps / e2e_seconds, 'unit': UNIT, 'h2d_bytes_per_step': e2e['h2d'], 'd2h_bytes_per_step': e2e['d2h'],
                      'ms_per_step': 1e3 * e2e_seconds / args.steps, 'substeps_per_step': e2e_substeps / args.steps,
                      'api': 'robovat_b200.envs.PushEnv.step(host actions) -> host obs, reward, done'}
    if not args.no_cpu_baseline and world_size == 1:
        v, sample = run_cpu_sample(cfg, threads, args.cpu_seconds, args.seed)
        out['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample}
    print(json.dumps(out))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
