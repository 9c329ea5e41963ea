#!/usr/bin/env python
"""Offline model of how environments are dealt to the blocks of k_substeps (CPU only; analysis, not product).

Records, from the CPU oracle, the solver work (iterations x colours) and candidate pairs of every environment for the
250 substeps of one mid-push launch of the bench workload, then replays that launch under a cost model of a block round
calibrated on the -DB2S_PROF timers:

    scene stage   ceil(m / 16) x 18 us                        m environments on 16 warps
    narrow phase  ceil(pairs / 64) x 35 us                    four pairs per warp and grab
    solve stage   greedy list schedule of c_e = 8 + 0.75 x iterations x colours us on 16 warps

and compares (a) the static two-class deal per launch (k_assign_envs) with (b) a dynamic hand-out per round from two
ready rings (expensive / cheap), which an environment re-enters after every substep.  usage:
    python tests/analysis/deal_model.py [envs] [skip] [substeps]
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402
from oracle import b2o                            # noqa: E402
from robovat_b200 import _capi, config            # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SKIP = int(sys.argv[2]) if len(sys.argv) > 2 else 600
N = int(sys.argv[3]) if len(sys.argv) > 3 else 250
WN, E, NBLOCKS, HEAVY_KEY = 16, 32, 147 if B == 4096 else max(1, (B + 27) // 28), 60


def record():
    cfg = bench.bench_config(B)
    scene = config.build_scene(cfg)
    params = config.build_params(cfg, scene, num_envs=B)
    b2o.build()
    w = b2o.OracleWorld(params, scene, threads=os.cpu_count() or 1)
    w.reset(seed=0)
    w.settle(0.1, 0.1, 500)
    w.settle()
    rs = np.random.RandomState(0)
    pos = w.observe().copy()
    mask = w.array('body_mask').reshape(B, -1)
    w.set_action(bench.heuristic_actions_np(pos, mask, cfg, rs))
    w.env_substeps(SKIP)
    key = np.zeros((N, B), np.int32)
    pairs = np.zeros((N, B), np.int32)
    active = np.zeros((N, B), bool)
    for s in range(N):
        active[s] = w.array(_capi.ARR_PHASE).reshape(B) != _capi.PHASE_IDLE
        w.env_substeps(1)
        st = w.array(_capi.ARR_SOLVER_STATS).reshape(B, 4)
        key[s] = st[:, 1] * st[:, 2]
        pairs[s] = w.array(_capi.ARR_NUM_PAIRS).reshape(B)
    w.close()
    return key, pairs, active


def round_time(keys, prs):
    """One block round over environments with solver keys `keys` (sorted by the caller as the kernel would grab them)."""
    m = len(keys)
    if m == 0:
        return 0.0
    a = -(-m // WN) * 18.0
    b = -(-int(prs.sum()) // (4 * WN)) * 35.0
    free = [0.0] * WN
    heapq.heapify(free)
    end = 0.0
    for k in keys:
        t = heapq.heappop(free) + 8.0 + 0.75 * k
        end = max(end, t)
        heapq.heappush(free, t)
    return a + b + end


def static_deal(key, pairs, active, first=0, count=None):
    """k_assign_envs once per launch (ranked by the previous substep), then `count` lock-step rounds per block."""
    count = N if count is None else count
    prev = key[max(first - 1, 0)]
    order = np.argsort(-prev, kind='stable')
    heavy = int((prev >= HEAVY_KEY).sum())
    hb = min((heavy + WN - 1) // WN, (NBLOCKS * E - B) // (E - WN), NBLOCKS - 1)
    blocks = [[] for _ in range(NBLOCKS)]
    for p, e in enumerate(order):
        if p < hb * WN:
            blocks[p % hb].append(e)
        else:
            blocks[hb + (p - hb * WN) % (NBLOCKS - hb)].append(e)
    tot = np.zeros(NBLOCKS)
    for bi, envs in enumerate(blocks):
        envs = np.asarray(envs, int)
        for s in range(first, min(N, first + count)):
            live = envs[active[s, envs]]
            tot[bi] += round_time(key[s, live], pairs[s, live])
    return tot


def dynamic_rings(key, pairs, active):
    """Every block draws 16 expensive or 32 cheap ready environments per round; an environment re-enters a ring
    (classified by the substep it just did) when its round ends.  Event-driven over block finish times."""
    step = np.zeros(B, int)                       # next substep of every environment
    ready_h, ready_l = [], list(np.argsort(-key[0], kind='stable'))
    ready_h = [e for e in ready_l if key[0, e] >= HEAVY_KEY]
    ready_l = [e for e in ready_l if key[0, e] < HEAVY_KEY]
    events = [(0.0, b, ()) for b in range(NBLOCKS)]          # (time, block, envs it returns)
    heapq.heapify(events)
    busy = np.zeros(NBLOCKS)
    left = B * N
    tmax = 0.0
    while events:
        t, b, back = heapq.heappop(events)
        for e in back:
            if step[e] < N:
                (ready_h if key[step[e] - 1, e] >= HEAVY_KEY else ready_l).append(e)
        if ready_h:
            take, ready_h = ready_h[:WN], ready_h[WN:]
        elif ready_l:
            take, ready_l = ready_l[:E], ready_l[E:]
        else:
            if left > 0 and events:               # nothing ready: wait for the next block to return its environments
                heapq.heappush(events, (events[0][0] + 1e-3, b, ()))
            continue
        take = np.asarray(take, int)
        s = step[take]
        live = active[s, take]
        order = np.argsort(-key[s, take], kind='stable')
        dt = round_time(key[s, take][order][live[order]], pairs[s, take][live])
        step[take] += 1
        left -= len(take)
        busy[b] += dt
        tmax = max(tmax, t + dt)
        heapq.heappush(events, (t + dt, b, tuple(take)))
    return tmax, busy


if __name__ == '__main__':
    cache = '/tmp/deal_model_%d_%d_%d.npz' % (B, SKIP, N)
    if os.path.exists(cache):
        z = np.load(cache)
        key, pairs, active = z['key'], z['pairs'], z['active']
    else:
        key, pairs, active = record()
        np.savez(cache, key=key, pairs=pairs, active=active)
    print('envs %d, substeps %d..%d: mean key %.1f, expensive (key >= %d) %.1f%%, pairs/env %.2f' % (
        B, SKIP, SKIP + N, key.mean(), HEAVY_KEY, 100.0 * (key >= HEAVY_KEY).mean(), pairs.mean()))
    tot = static_deal(key, pairs, active)
    print('static two-class deal : launch %.1f ms (slowest block), mean block %.1f ms, slowest / mean %.2f' % (
        tot.max() / 1e3, tot.mean() / 1e3, tot.max() / tot.mean()))
    for chunk in (125, 50, 25, 10):
        t = sum(static_deal(key, pairs, active, f, chunk).max() for f in range(0, N, chunk))
        print('re-dealt every %3d substeps: %.1f ms (+ %d launches x ~0.03 ms)' % (chunk, t / 1e3, N // chunk))
    tmax, busy = dynamic_rings(key, pairs, active)
    print('dynamic ready rings   : launch %.1f ms, mean block busy %.1f ms, launch / mean busy %.2f' % (
        tmax / 1e3, busy.mean() / 1e3, tmax / busy.mean()))
