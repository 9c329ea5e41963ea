#!/usr/bin/env python
"""Headline benchmark: PushEnv substeps/s at 4096 envs per GPU (BASELINE.json configs[1]).

The workload is data collection as tools/parallel_run.py runs it: every env works through its own stream of
episodes (reset -> drop -> settle -> [policy -> 7-phase push -> wait_until_stable -> reward] x MAX_STEPS -> reset ...)
at its own pace; nobody waits for the slowest env.  One bench "step" = every env of the rank advances by
`--substeps` (2000) substeps of that stream, in launches of 250.  Reported:
  value     substeps/s of the device-side rollout (b2s_rollout_*: policy, reward, reset on the device; nothing
            crosses PCIe inside the timed region)
  e2e       the same metric through the public API `robovat_b200.envs.PushEnv.step_async` with the policy on the
            HOST: every slice uploads the host policy's actions from pinned memory and downloads observation, reward
            and status (b2s_env_async_step)
  lockstep  the reference-shaped `PushEnv.step` (the whole batch waits for its slowest env), for comparison
  roofline / cpu_baseline / pose_tolerance / other_configs as the task contract and VERDICT ask.
`--impl reference` times the CPU oracle port of the same rollout (same seeds, hence the same episodes) on all host
cores; PyBullet itself is probed for and reported as unavailable.
The actions come from the "aimed" synthetic policy (B2S_POLICY_AIMED: start 8 cm behind a random body, push through
it), so every push makes contact; the reference's HeuristicPushSampler is also on the device (policy_kind 0).
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


CHUNK = 250                    # substeps per launch
POLICY_SEED, RESET_SEED = 1234, 4321


def crossing_config():
    """BASELINE.json configs[2]: TASK_NAME='crossing' LAYOUT_ID=0, 8 V-HACD concave movables."""
    from robovat_b200 import config
    cfg = config.default_push_env_config(TASK_NAME='crossing', LAYOUT_ID=0, MOVABLE_NAME='concave',
                                         MIN_MOVABLE_BODIES=8, MAX_MOVABLE_BODIES=8)
    cfg.SIM.TIME_STEP = 1.0 / 240.0
    return cfg


def config_dict(cfg, B, substeps, world_size):
    return {'workload': WORKLOAD, 'envs_per_gpu': B, 'time_step': cfg.SIM.TIME_STEP, 'solver_iterations': 50,
            'step': 'every env advances %d substeps of its own episode stream (reset, drop, settle, MAX_STEPS=%d x '
                    '[aimed push policy, 7-phase push, wait_until_stable, reward]); launches of %d substeps' % (substeps, cfg.MAX_STEPS, CHUNK),
            'substeps_per_env_per_step': substeps, 'policy': 'B2S_POLICY_AIMED', 'seeds': [POLICY_SEED, RESET_SEED],
            'l2': 'flushed between timed steps (256 MB memset, untimed)', 'parallelism': 'env-sharded x%d' % world_size}


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
    for name in ('r02_k_substeps_dram_traffic.json', 'r01_k_substeps_dram_traffic.json'):
        path = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(path):
            with open(path) as f:
                return float(json.load(f)['dram_bytes_per_env_substep'])
    return None


def probe_pybullet():
    """the reference's own CPU path needs pybullet==2.6.5 (requirements.txt:9)."""
    try:
        import pybullet  # noqa: F401
        return 'importable (not used: the reference arm is the oracle port so that both arms run the same algorithm)'
    except Exception as e:            # ImportError here: no wheel, no source, no network
        return 'unavailable: %s' % (str(e).splitlines()[0] if str(e) else type(e).__name__)


def cpu_rollout(cfg, num_envs, threads, seed, env_id_offset=0):
    """Oracle world with the same rollout as the device leg (same seeds => the same episodes)."""
    from oracle import b2o
    from robovat_b200 import _capi, config
    b2o.build()
    scene = config.build_scene(cfg)
    params = config.build_params(cfg, scene, num_envs=num_envs, env_id_offset=env_id_offset)
    w = b2o.OracleWorld(params, scene, threads=threads)
    table_z = [st['pose'][2] for st in scene.statics if st['flags'] & _capi.STATIC_IS_TABLE][-1]
    todo = None
    for _ in range(9):                  # Simulator.reset_scene: re-sample envs whose bodies fell off the table
        w.reset(seed=seed * 1000003 + 1, mask=todo)
        w.settle(0.1, 0.1, 500, mask=todo)
        w.settle(mask=todo)
        live = w.body_mask > 0
        bad = ((w.body_state[2] < (np.float32(table_z) + w.array(_capi.ARR_TABLE_DZ))[:, None]) & live).any(axis=1)
        bad |= (w.array(_capi.ARR_ERROR_FLAGS) & 128) != 0
        if todo is not None:
            bad &= todo.astype(bool)
        if not bad.any():
            break
        todo = bad.astype(np.uint8)
    w.begin_episode()
    w.rollout_begin(cfg.MAX_STEPS, num_episodes=1 << 20, policy_seed=POLICY_SEED, reset_seed=RESET_SEED, record=False,
                    policy_kind=_capi.POLICY_AIMED)
    return w


def run_cpu_steps(cfg, num_envs, threads, seed, substeps, steps, warmup):
    """`steps` bench steps of the oracle rollout after `warmup` untimed ones; returns (substeps/s, seconds, substeps)."""
    w = cpu_rollout(cfg, num_envs, threads, seed)
    for _ in range(warmup):
        w.rollout_run(substeps)
    s0 = w.substeps_executed()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.rollout_run(substeps)
    dt = time.perf_counter() - t0
    n = w.substeps_executed() - s0
    w.close()
    return n / dt, dt, n


def cpu_baseline(cfg, num_envs, threads, seed, substeps, target_seconds):
    """Bounded sample: whole bench steps on all host threads for about target_seconds, plus a 1-thread number."""
    rate, dt, n = run_cpu_steps(cfg, num_envs, threads, seed, substeps, 1, 0)            # one step to size the sample
    steps = int(max(1, min(8, round(target_seconds / max(dt, 1e-3)))))
    rate, dt, n = run_cpu_steps(cfg, num_envs, threads, seed, substeps, steps, 1)
    sample = '%d envs x %d steps x %d substeps (%d substeps, %.1f s) on %d threads, after 1 warm-up step' % (num_envs, steps, substeps, n, dt, threads)
    one_envs = max(16, min(num_envs, int(rate / threads * 3.0 / substeps)))              # about 3 s of one thread
    r1, d1, n1 = run_cpu_steps(cfg, one_envs, 1, seed, substeps, 1, 0)
    return {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
            'one_thread': {'value': r1, 'unit': UNIT, 'sample': '%d envs x 1 step x %d substeps (%.1f s)' % (one_envs, substeps, d1)},
            'reference_cpu': 'pybullet ' + probe_pybullet()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=4096, help='envs per GPU')
    ap.add_argument('--substeps', type=int, default=2000, help='substeps every env advances per bench step')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--exact-schedule', action='store_true', help='every launch advances every env by exactly 250 substeps (default: free-running launches)')
    ap.add_argument('--no-extras', action='store_true', help='skip lockstep / pose_tolerance / other_configs')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = bench_config(args.envs)
    threads = os.cpu_count() or 1
    B, S = args.envs, args.substeps
    slices = (S + CHUNK - 1) // CHUNK

    if args.impl == 'reference':
        if rank != 0:
            return 0
        v, dt, n = run_cpu_steps(cfg, B, threads, args.seed, S, args.steps, args.warmup)
        sample = '%d envs x %d steps x %d substeps (%d substeps, %.1f s) on %d threads' % (B, args.steps, S, n, dt, threads)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config_dict(cfg, B, S, max(1, args.gpus)),
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
                             'reference_cpu': 'pybullet ' + probe_pybullet()},
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

    env = PushEnv(config=cfg, num_envs=B, seed=args.seed, device=local_rank, env_id_offset=rank * B)
    w = env.world
    env.reset()
    dev = w.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    # Returns of all ranks (tools/parallel_run.py:54-90 collects the workers' results; its workers never wait for one
    # another).  The all-gather after step k is asynchronous and double buffered: it travels on NCCL's stream while
    # step k+1 runs; the last ones are waited for inside the timed region.
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

    # ---- device-resident leg: the rollout runs entirely on the device ---------------------------
    w.rollout_begin(cfg.MAX_STEPS, num_episodes=1 << 20, policy_seed=POLICY_SEED, reset_seed=RESET_SEED, record=None,
                    policy_kind=_capi.POLICY_AIMED, free_running=not args.exact_schedule)
    kernel_events = []

    def device_step(timed, last=False):
        left = S
        while left > 0:
            n = min(CHUNK, left)
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            w.rollout_run(chunk=n, max_substeps=n)
            if timed:
                e1.record()
                kernel_events.append((e0, e1))
            left -= n
        gather_returns(w, last)

    for _ in range(args.warmup):
        device_step(False)
    barrier()
    s0, l0 = w.substeps_executed(), w.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_events = []
    barrier()
    for k in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        device_step(True, last=(k == args.steps - 1))
        b.record()
        step_events.append((a, b))
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in step_events)
    substeps = w.substeps_executed() - s0
    launches = w.launch_count() - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events)
    n_kernel_launches = len(kernel_events)
    episodes_done = float(w.array(_capi.ARR_NUM_EPISODES).float().mean().item())
    errors = int((w.array(_capi.ARR_ERROR_FLAGS) != 0).sum().item())

    # ---- end-to-end leg: PushEnv.step_async, policy on the host --------------------------------
    e2e = None
    if not args.no_e2e:
        env2 = PushEnv(config=cfg, num_envs=B, seed=args.seed, device=local_rank, env_id_offset=rank * B)
        w2 = env2.world
        rs = np.random.RandomState(args.seed + rank)
        obs = env2.reset()

        def e2e_step(obs, last=False):
            for _ in range(slices):
                act = heuristic_actions_np(obs['position'], obs['body_mask'], cfg, rs)       # host policy on the host observation
                obs, rew, done, info = env2.step_async(act, substeps=CHUNK, free_running=not args.exact_schedule)
            gather_returns(w2, last)
            return obs
        for _ in range(args.warmup):
            obs = e2e_step(obs)
        barrier()
        s1 = w2.substeps_executed()
        t0 = time.perf_counter()
        for k in range(args.steps):
            obs = e2e_step(obs, last=(k == args.steps - 1))
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_sub = w2.substeps_executed() - s1
        e2e = {'substeps': e2e_sub, 'seconds': e2e_s, 'h2d': slices * env2.async_h2d_bytes, 'd2h': slices * env2.async_d2h_bytes,
               'episodes': float(np.mean(env2._num_episodes))}
        env2.close()

    # ---- max over ranks ------------------------------------------------------------------------
    stats = torch.tensor([ms, float(substeps), kernel_ms, float(e2e['seconds'] if e2e else 0), float(e2e['substeps'] if e2e else 0)],
                         dtype=torch.float64, device=dev)
    if world_size > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
        total_substeps, e2e_seconds, e2e_substeps = float(sm[1]), float(mx[3]), float(sm[4])
    else:
        total_substeps, e2e_seconds, e2e_substeps = float(substeps), float(stats[3]), float(stats[4])
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
        'dtype': 'f32', 'data': 'synthetic', 'config': config_dict(cfg, B, S, world_size),
        'gpu_launches': int(launches),
        'clocks': clocks,
        'episodes_per_env_so_far': episodes_done, 'envs_with_error_flags': errors,
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
                      'slices_per_step': slices, 'episodes_per_env_so_far': e2e['episodes'],
                      'api': 'robovat_b200.envs.PushEnv.step_async(host actions) -> host obs, reward, done, info (b2s_env_async_step)'}
    if world_size == 1 and not args.no_extras:
        out.update(extras(args, cfg, env, local_rank, threads))
    if not args.no_cpu_baseline and world_size == 1:
        out['cpu_baseline'] = cpu_baseline(cfg, B, threads, args.seed, S, args.cpu_seconds)
    print(json.dumps(out))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def extras(args, cfg, env, device, threads):
    """Numbers next to the headline (N = 1 only): the lock-step API, the fp32 pose tolerance, BASELINE configs[2], [3]."""
    import torch
    from robovat_b200 import _capi, config
    from robovat_b200.envs import PushEnv
    from robovat_b200.world import World
    out = {}
    B = args.envs
    peak, _ = measured_peak()
    # -- the reference-shaped PushEnv.step: the batch waits for its slowest env
    w = env.world
    rs = np.random.RandomState(args.seed)
    env.reset()
    obs = env.get_observation(force=True)
    t_sub = 0
    for k in range(3):
        if k == 1:
            torch.cuda.synchronize()
            t0, s0 = time.perf_counter(), w.substeps_executed()
        env._done[:] = False
        act = heuristic_actions_np(np.asarray(obs['position']).reshape(B, -1, 3), np.asarray(obs['body_mask']).reshape(B, -1), cfg, rs)
        obs, _, _, _ = env.step(act)
    torch.cuda.synchronize()
    dt, t_sub = time.perf_counter() - t0, w.substeps_executed() - s0
    out['lockstep'] = {'value': t_sub / dt, 'unit': UNIT, 'ms_per_step': 1e3 * dt / 2, 'api': 'robovat_b200.envs.PushEnv.step(host actions), 2 steps after 1 warm-up',
                       'h2d_bytes_per_step': B * 16, 'd2h_bytes_per_step': B * w.N * 13 + B * 39}
    env.close()
    # -- stated fp32 pose tolerance: CUDA fp32 vs the double-precision oracle, 240 substeps at dt = 1/240 from identical states
    try:
        from oracle import b2o, pose_tolerance
        b2o.build()
        scene = config.build_scene(cfg)
        params = config.build_params(cfg, scene, num_envs=256)
        w32 = World(params, scene, device=device)
        tol = {o['scenario']: {k: o[k] for k in ('bodies', 'median_m', 'p99_m', 'max_m', 'frac_le_1e-3')}
               for o in pose_tolerance.measure(w32, params, scene, seed=args.seed, threads=threads)}
        w32.close()
        out['pose_tolerance'] = {'against': 'double-precision build of the oracle (pybullet is unavailable)', 'substeps': 240, 'envs': 256, **tol}
    except Exception as e:             # the checker must not take the bench line down
        out['pose_tolerance'] = {'error': str(e)[:200]}
    other = {}
    # -- configs[2]: crossing, 8 concave movables, 4096 envs: the same rollout
    try:
        cfg3 = crossing_config()
        env3 = PushEnv(config=cfg3, num_envs=B, seed=args.seed, device=device)
        env3.reset()
        w3 = env3.world
        w3.rollout_begin(cfg3.MAX_STEPS, num_episodes=1 << 20, policy_seed=POLICY_SEED, reset_seed=RESET_SEED, record=None, policy_kind=_capi.POLICY_AIMED,
                         free_running=not args.exact_schedule)
        w3.rollout_run(chunk=CHUNK, max_substeps=500)
        torch.cuda.synchronize()
        s0 = w3.substeps_executed()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        w3.rollout_run(chunk=CHUNK, max_substeps=500)
        b.record()
        torch.cuda.synchronize()
        n3, ms3 = w3.substeps_executed() - s0, a.elapsed_time(b)
        nb3, c3 = 9, float(w3.array(_capi.ARR_SOLVER_STATS).view(B, 4)[:, 3].float().mean().item())
        bytes3 = 104 * nb3 + 24 * c3 + 84
        other['crossing_4096'] = {'workload': 'PushEnv TASK_NAME=crossing LAYOUT_ID=0, 8 V-HACD concave movables, 4096 envs', 'value': n3 / ms3 * 1e3, 'unit': UNIT,
                                  'ms': ms3, 'substeps': n3,
                                  'roofline': {'bound': 'hbm', 'achieved': n3 * bytes3 / ms3 / 1e6, 'peak': peak, 'unit': 'GB/s',
                                               'frac': n3 * bytes3 / ms3 / 1e6 / peak, 'algorithmic_bytes_per_substep': bytes3, 'mean_contacts': c3}}
        env3.close()
        if not args.no_cpu_baseline:
            v, dt, n = run_cpu_steps(cfg3, 512, threads, args.seed, 250, 1, 0)
            other['crossing_4096']['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': '512 envs x 250 substeps (%.1f s)' % dt}
    except Exception as e:
        other['crossing_4096'] = {'error': str(e)[:200]}
    # -- configs[3]: 128 x 128 depth + segmentation, 2048 envs
    try:
        from robovat_b200.assets import quat_from_euler, quat_to_matrix
        size = 128
        kin = dict(config.DEFAULT_PUSH_ENV['KINECT2']['DEPTH'], HEIGHT=size, WIDTH=size,
                   INTRINSICS=[120.0, 0.0, size / 2.0, 0, 120.0, size / 2.0, 0, 0, 1])
        cfg4 = config.default_push_env_config(KINECT2={'DEPTH': kin}, MIN_MOVABLE_BODIES=3, MAX_MOVABLE_BODIES=3)
        scene4 = config.build_scene(cfg4)
        w4 = World(config.build_params(cfg4, scene4, num_envs=2048), scene4, device=device, with_camera=True)
        w4.reset(seed=1)
        w4.settle(0.1, 0.1, 500)
        R = quat_to_matrix(quat_from_euler(np.pi, 0, 0))
        w4.set_camera(np.array(kin['INTRINSICS'], np.float64), R.reshape(9), -R.dot(np.array([0.6, 0.0, 1.1])), per_env=False)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=w4.device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn, reps=10):
            t = []
            for _ in range(reps):
                flush.zero_()
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                t.append(a.elapsed_time(b))
            return float(np.median(t))
        w4.render()
        w4.point_cloud(seed=0)
        ms4 = timed(w4.render)
        ms4pc = timed(lambda: w4.point_cloud(seed=0))
        H, Wd = size, size
        bytes4 = 2048 * H * Wd * 5
        other['render_2048_128'] = {'workload': 'PushEnv + %dx%d depth/segmentation camera observation, 2048 envs' % (H, Wd), 'value': 2048 / ms4 * 1e3, 'unit': 'frames/s',
                                    'ms_per_batch': ms4, 'point_cloud_ms_per_batch': ms4pc, 'l2': 'flushed before every timed call',
                                    'roofline': {'bound': 'hbm', 'achieved': bytes4 / ms4 / 1e6, 'peak': peak, 'unit': 'GB/s', 'frac': bytes4 / ms4 / 1e6 / peak,
                                                 'algorithmic_bytes_per_env': H * Wd * 5, 'kernel': 'k_render'}}
        w4.close()
    except Exception as e:
        other['render_2048_128'] = {'error': str(e)[:200]}
    out['other_configs'] = other
    return out


if __name__ == '__main__':
    sys.exit(main())
