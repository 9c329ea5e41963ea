"""Batched episode driver and columnar episode shards.

The reference collects data one environment per process: `generate_episodes`
(robovat/io/episode_generation.py:72-119) resets the env, alternates
`policy.action` / `env.step` until `done` or `num_steps`, and hands one dict per
episode (`hostname`, `timestamp`, `transitions[i] = {state, action, reward,
info}`) to a writer.  Here B environments advance together, so the driver is
organised the other way round: one pass over time, every quantity an array with
a leading [T, B] axis, and a per-environment episode length.  Nothing in this
module loops over environments while stepping.

  collect(env, policy, num_steps)   -> EpisodeBatch (columnar, [T, B, ...])
  collect_rollouts(env, ...)        -> list of EpisodeBatch, one per episode index: the episodes run on the device
                                       (b2s_rollout_*: policy, reward, reset without the host), nobody waits for the
                                       slowest environment; `batches_from_records` is the pure conversion
  EpisodeBatch.episodes()           -> the reference's per-episode dict layout, for
                                       consumers that expect it (one dict per env)
  ShardWriter / read_shard          -> .npz shards of whole batches (robovat/io's
                                       HDF5 and pickle writers are out of scope,
                                       SURVEY.md section 2 #13; h5py is absent here)

A single environment is the B = 1 case of the same code.
"""
import os
import socket
from datetime import datetime

import numpy as np


def _stamp():
    return datetime.now().strftime('%Y-%m-%d-%H-%M-%S-%f')


class EpisodeBatch(object):
    """T time slices of B environments.

    states[k]   [T, B, ...]  observation the action of slice t was computed from
    actions     [T, B, A]
    rewards     [T, B]
    lengths     [B]          number of valid slices of environment b (its episode stops at its own `done`)
    final[k]    [B, ...]     observation after the last step of every environment
    """

    def __init__(self, states, actions, rewards, lengths, final, hostname=None, timestamp=None):
        self.states, self.actions, self.rewards = states, actions, rewards
        self.lengths, self.final = lengths, final
        self.hostname = hostname or socket.gethostname()
        self.timestamp = timestamp or _stamp()

    @property
    def num_envs(self):
        return int(self.lengths.shape[0])

    @property
    def returns(self):
        """Undiscounted return of every environment's episode, [B]."""
        t = np.arange(self.rewards.shape[0])[:, None]
        return np.where(t < self.lengths[None, :], self.rewards, 0.0).sum(axis=0)

    def episodes(self):
        """One dict per environment in the layout the reference's writers receive
        (episode_generation.py:63-69): hostname, timestamp, transitions[{state, action, reward, info}]."""
        out = []
        for b in range(self.num_envs):
            tr = [{'state': {k: v[t, b] for k, v in self.states.items()}, 'action': self.actions[t, b],
                   'reward': float(self.rewards[t, b]), 'info': None} for t in range(int(self.lengths[b]))]
            out.append({'hostname': self.hostname, 'timestamp': self.timestamp, 'transitions': tr})
        return out


def _batched(x, B):
    """Observation / reward / done of a B = 1 reference-style env gain the batch axis."""
    a = np.asarray(x)
    if a.ndim == 0 or a.shape[0] != B:
        a = a[None]
    return a


def collect(env, policy, num_steps=None):
    """Run one episode in every environment of `env` (batched or single) and return an EpisodeBatch.

    Semantics per environment are the reference's (episode_generation.py:41-61): the transition of slice t is
    recorded before `done` is looked at, an episode ends at its own `done` or after `num_steps` steps, and
    environments that are done keep being stepped as no-ops by the batched env until the slowest one finishes.
    """
    B = int(getattr(env, 'num_envs', 1))
    obs = env.reset()
    alive = np.ones(B, bool)
    lengths = np.zeros(B, np.int64)
    states, actions, rewards = {}, [], []
    t = 0
    while alive.any():
        action = _batched(np.asarray(policy.action(obs), np.float32).reshape(B, -1), B)
        for k, v in obs.items():
            states.setdefault(k, []).append(_batched(v, B).copy())
        new_obs, reward, done, _ = env.step(action if B > 1 else action[0])
        actions.append(action)
        rewards.append(np.where(alive, _batched(reward, B).astype(np.float64), 0.0))
        lengths += alive
        alive &= ~_batched(done, B).astype(bool)
        obs = new_obs
        t += 1
        if num_steps is not None and t >= num_steps:
            break
    final = {k: _batched(v, B).copy() for k, v in obs.items()}
    return EpisodeBatch({k: np.stack(v) for k, v in states.items()}, np.stack(actions), np.stack(rewards), lengths, final)


def batches_from_records(records, body_mask=None, hostname=None, timestamp=None):
    """Device-side rollout records (the arrays of B2SRollout: actions [B, EP, A, 4], rewards [B, EP, A], positions
    [B, EP, A+1, N, 3], flags [B, EP, A], substeps [B, EP, A], lengths [B, EP], returns [B, EP]) -> one EpisodeBatch per
    episode index, in the layout `collect` produces: states['position'][t] is the observation action t was computed
    from, final['position'] the one after the episode's last step; is_safe / is_effective / termination of the
    transition and Simulator.num_steps after it ride along as states of the NEXT slice (as the reference's
    observations report them) and in `final`."""
    rec = {k: np.asarray(v) for k, v in records.items() if v is not None}
    B, EP, A = rec['rewards'].shape
    t = np.arange(A)
    out = []
    for ep in range(EP):
        lengths = rec['lengths'][:, ep].astype(np.int64)
        valid = t[None, :] < lengths[:, None]                                   # [B, A]
        actions = np.where(valid[..., None], rec['actions'][:, ep], 0).transpose(1, 0, 2)
        rewards = np.where(valid, rec['rewards'][:, ep], 0.0).transpose(1, 0).astype(np.float64)
        states, final = {}, {}
        if 'positions' in rec:
            pos = rec['positions'][:, ep]                                         # [B, A+1, N, 3]
            states['position'] = np.where(valid[..., None, None], pos[:, :A], 0).transpose(1, 0, 2, 3)
            final['position'] = pos[np.arange(B), lengths]
            # the scene (and with ragged counts the number of bodies) changes from episode to episode: a body is present
            # iff its first recorded position is not the zero padding (no body rests at the world origin)
            mask = (pos[:, 0] != 0).any(axis=-1).astype(np.float32) if body_mask is None or ep > 0 else np.asarray(body_mask, np.float32)
            states['body_mask'] = np.broadcast_to(mask[None], (A,) + mask.shape).copy()
            final['body_mask'] = mask
        flags = rec['flags'][:, ep]
        prev = np.concatenate([np.full((B, 1), 3, flags.dtype), flags[:, :A - 1]], axis=1)     # before step 0: safe, effective
        for name, bit in (('is_safe', 1), ('is_effective', 2)):
            states[name] = np.where(valid, (prev & bit) != 0, False).transpose(1, 0).astype(np.int64)
            final[name] = ((flags[np.arange(B), np.maximum(lengths - 1, 0)] & bit) != 0).astype(np.int64)
        states['num_steps'] = np.broadcast_to(t[:, None], (A, B)).copy()
        final['num_steps'] = lengths.copy()
        final['termination'] = (flags[np.arange(B), np.maximum(lengths - 1, 0)] & 4) != 0
        if 'substeps' in rec:
            final['simulator_num_steps'] = rec['substeps'][np.arange(B), ep, np.maximum(lengths - 1, 0)]
        out.append(EpisodeBatch(states, actions, rewards, lengths, final, hostname, timestamp))
    return out


def collect_rollouts(env, num_episodes=1, num_steps=None, policy_seed=0, reset_seed=None, policy_kind=0, max_attempts=None,
                     chunk=250, free_running=True, max_substeps=1 << 30):
    """`num_episodes` episodes in every environment of a batched PushEnv, run on the device: HeuristicPushPolicy
    (policy_kind 0) or the aimed synthetic policy (1) draws the actions, rewards and resets happen without the host
    (World.rollout_begin / rollout_run).  Episode 0 starts from a fresh `env.reset()`.  Returns one EpisodeBatch per
    episode index (batches_from_records)."""
    from robovat_b200.world import RolloutRecord
    cfg = env.config
    A = int(num_steps if num_steps is not None else cfg.MAX_STEPS)
    env.reset()
    w = env.world
    rec = RolloutRecord(w.B, w.N, num_episodes, A, w.device)
    attempts = int(max_attempts if max_attempts is not None else cfg.get('HEURISTICS', {}).get('MAX_ATTEMPS', 2000))
    w.rollout_begin(A, num_episodes, policy_seed=policy_seed, reset_seed=env.seed * 1000003 + 7919 if reset_seed is None else reset_seed,
                    max_attempts=attempts, record=rec, policy_kind=policy_kind, free_running=free_running)
    left = w.rollout_run(chunk=chunk, max_substeps=max_substeps)
    if left:
        raise RuntimeError('collect_rollouts: %d environments did not finish within %d substeps' % (left, max_substeps))
    records = {k: v.cpu().numpy() for k, v in rec.tensors().items()}
    return batches_from_records(records)


def generate_batched_episodes(env, policy, num_steps=None):
    """Per-environment episode dicts of one batched pass (kept for callers of the round-1 name)."""
    return collect(env, policy, num_steps).episodes()


class ShardWriter(object):
    """Writes every EpisodeBatch as one compressed .npz shard: `<dir>/episodes_<index>.npz` with the arrays
    `state/<key>`, `final/<key>`, `action`, `reward`, `length` and the strings `hostname`, `timestamp`."""

    def __init__(self, output_dir):
        self.output_dir = output_dir
        self.num_shards = 0
        os.makedirs(output_dir, exist_ok=True)

    def __call__(self, batch):
        path = os.path.join(self.output_dir, 'episodes_%06d.npz' % self.num_shards)
        arrays = {'action': batch.actions, 'reward': batch.rewards, 'length': batch.lengths,
                  'hostname': np.array(batch.hostname), 'timestamp': np.array(batch.timestamp)}
        arrays.update({'state/' + k: v for k, v in batch.states.items()})
        arrays.update({'final/' + k: v for k, v in batch.final.items()})
        np.savez_compressed(path, **arrays)
        self.num_shards += 1
        return path


def read_shard(path):
    with np.load(path) as z:
        states = {k[6:]: z[k] for k in z.files if k.startswith('state/')}
        final = {k[6:]: z[k] for k in z.files if k.startswith('final/')}
        return EpisodeBatch(states, z['action'], z['reward'], z['length'], final, str(z['hostname']), str(z['timestamp']))
