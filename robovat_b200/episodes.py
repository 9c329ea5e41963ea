"""Episode driver and episode files: the callers either side of the path.

  generate_episode / generate_episodes    robovat/io/episode_generation.py:21-119
  PickleWriter / read / read_all          robovat/io/pickle_utils.py:20-131 (same stream-of-pickles file format)
  generate_batched_episodes               the same loop over a batched PushEnv: one episode dict per environment

A single-env `robovat_b200.envs.PushEnv` (num_envs=1) with `HeuristicPushPolicy` runs through `generate_episode`
exactly as the reference env does (tools/run_env.py:217-253).  The reference's quirks are kept where a caller can
observe them: `generate_episodes` ignores `num_episodes` unless `strict=True` is passed (episode_generation.py:88-119
loops forever), and an exception discards the episode and continues.  HDF5 output (io/hdf5_utils.py) needs h5py, which
this image does not have; the pickle writer is the reference's other episode format.
"""
import os
import pickle
import signal
import socket
import time
import traceback
import uuid
from datetime import datetime

import numpy as np


def get_timestamp_as_string():
    """robovat/utils/time_utils.py: date and time down to microseconds."""
    return datetime.now().strftime('%Y-%m-%d-%H-%M-%S-%f')


class Timeout(object):
    """SIGALRM guard of one episode (robovat/utils/time_utils.py:19-34); seconds are rounded up to an int, which
    `signal.alarm` needs on Python 3 (the reference passes its float default through and would raise)."""

    def __init__(self, sec):
        self.sec = int(np.ceil(sec)) if sec else 0

    def __enter__(self):
        if self.sec:
            signal.signal(signal.SIGALRM, self.raise_timeout)
            signal.alarm(self.sec)

    def __exit__(self, *args):
        if self.sec:
            signal.alarm(0)

    def raise_timeout(self, *args):
        raise TimeoutError('episode timed out')


def generate_episode(env, policy, num_steps=None, debug=False):
    """One episode of a single environment (episode_generation.py:21-69)."""
    t = 0
    transitions = []
    observation = env.reset()
    while 1:
        action = policy.action(observation)
        new_observation, reward, done, info = env.step(action)
        transitions.append({'state': observation, 'action': action, 'reward': reward, 'info': info})
        observation = new_observation
        if done:
            break
        t += 1
        if (num_steps is not None) and (t >= num_steps):
            break
    return {'hostname': socket.gethostname(), 'timestamp': get_timestamp_as_string(), 'transitions': transitions}


def generate_episodes(env, policy, num_steps=None, num_episodes=None, timeout=30, debug=False, strict=False):
    """Generator of (episode_index, episode) (episode_generation.py:72-119).  `strict=True` honours `num_episodes`
    (the reference never checks it); failed episodes are discarded and the loop goes on, as in the reference."""
    episode_index = 0
    total_time = 0.0
    while 1:
        if strict and num_episodes is not None and episode_index >= num_episodes:
            return
        try:
            tic = time.time()
            if debug:
                episode = generate_episode(env, policy, num_steps, debug)
            else:
                with Timeout(timeout):
                    episode = generate_episode(env, policy, num_steps, debug)
            total_time += time.time() - tic
            yield episode_index, episode
            episode_index += 1
        except (KeyboardInterrupt, GeneratorExit):
            raise
        except Exception:            # the reference prints the traceback and discards the episode
            traceback.print_exc()


def generate_batched_episodes(env, policy, num_steps=None):
    """One episode per environment of a batched PushEnv, stepped together.

    Returns a list of `env.num_envs` episode dicts with the reference's layout; environment e's transitions stop at
    its own `done`.  Observations are sliced per environment so every transition looks like the single-env one.
    """
    B = env.num_envs
    observation = env.reset()
    done = np.zeros(B, bool)
    stamp = get_timestamp_as_string()
    transitions = [[] for _ in range(B)]

    def take(obs, e):
        return {k: np.asarray(v)[e] for k, v in obs.items()}
    t = 0
    while not done.all():
        action = np.asarray(policy.action(observation), np.float32).reshape(B, -1)
        new_observation, reward, new_done, info = env.step(action)
        reward = np.atleast_1d(reward)
        new_done = np.atleast_1d(new_done)
        for e in np.nonzero(~done)[0]:
            transitions[e].append({'state': take(observation, e), 'action': action[e], 'reward': float(reward[e]), 'info': info})
        done = done | new_done
        observation = new_observation
        t += 1
        if (num_steps is not None) and (t >= num_steps):
            break
    host = socket.gethostname()
    return [{'hostname': host, 'timestamp': stamp, 'transitions': transitions[e]} for e in range(B)]


class PickleWriter(object):
    """Episode files as a stream of pickles, `num_entries_per_file` per file (pickle_utils.py:20-94)."""

    def __init__(self, output_dir, num_entries_per_file, use_random_name=True):
        self._output_dir = output_dir
        self._num_entries_per_file = num_entries_per_file
        self._use_random_name = use_random_name
        self._file = None
        self._output_path = None
        self._num_files = 0
        self._num_entries_this_file = 0
        if not os.path.isdir(output_dir):
            os.makedirs(output_dir)

    def __call__(self, data):
        if self._num_entries_this_file == 0:
            if self._use_random_name:
                timestamp = get_timestamp_as_string() + '-' + uuid.uuid4().hex[:6]     # batched episodes share a microsecond
            else:
                timestamp = '%06d' % (self._num_files)
            self._output_path = os.path.join(self._output_dir, 'data_%s.pickle' % (timestamp))
            self._num_files += 1
            if self._file:
                self._file.close()
            self._file = open(self._output_path, 'wb')
        num_entries = self.write(data)
        self._num_entries_this_file += num_entries
        self._num_entries_this_file %= self._num_entries_per_file

    def write(self, data):
        pickle.dump(data, self._file, protocol=pickle.HIGHEST_PROTOCOL)
        return 1

    def close(self):
        if self._file is not None:
            self._file.close()
            self._file = None


def read(filename):
    """Yields the entries of an episode file (pickle_utils.py:97-112)."""
    with open(filename, 'rb') as f:
        while True:
            try:
                yield pickle.load(f)
            except EOFError:
                break


def read_all(filename):
    """pickle_utils.py:115-131."""
    return list(read(filename))
