"""Small training-glue helpers the example scripts expect under ``magent.utility``.

Out of the hot path; kept minimal so ``examples/train_*.py`` import and
run: an episode buffer keyed by agent id, the epsilon schedules, a logger and recursive rounding
(reference surface: python/magent/utility.py:14-150).
"""
import logging
import numbers

import numpy as np


class EpisodesBufferEntry:
    """Trajectory of one agent: parallel lists of view/feature/action/reward + terminal flag."""

    __slots__ = ("views", "features", "actions", "rewards", "terminal")

    def __init__(self):
        self.views, self.features, self.actions, self.rewards = [], [], [], []
        self.terminal = False

    def append(self, view, feature, action, reward, alive):
        self.views.append(np.array(view, copy=True))
        self.features.append(np.array(feature, copy=True))
        self.actions.append(action)
        self.rewards.append(reward)
        self.terminal = self.terminal or not alive


class EpisodesBuffer:
    """Whole-episode replay store: at most ``capacity`` agents, one entry per agent id."""

    def __init__(self, capacity):
        self.buffer = {}
        self.capacity = capacity
        self.is_full = False

    def record_step(self, ids, obs, acts, rewards, alives):
        views, features = obs
        order = range(len(ids)) if self.is_full else np.random.permutation(len(ids))
        for i in order:
            entry = self.buffer.get(ids[i])
            if entry is None:
                if self.is_full:
                    continue
                entry = self.buffer[ids[i]] = EpisodesBufferEntry()
                self.is_full = len(self.buffer) >= self.capacity
            entry.append(views[i], features[i], acts[i], rewards[i], alives[i])

    def reset(self):
        self.buffer = {}
        self.is_full = False

    def episodes(self):
        return self.buffer.values()


def exponential_decay(now_step, total_step, final_value, rate):
    """1 -> final_value, exponentially over total_step."""
    decay = final_value ** (1.0 / (total_step / rate))
    return max(final_value, decay ** (now_step / rate))


def linear_decay(now_step, total_step, final_value):
    """1 -> final_value, linearly over total_step."""
    if now_step >= total_step:
        return final_value
    return 1.0 - (1.0 - final_value) * now_step / total_step


def piecewise_decay(now_step, anchor, anchor_value):
    """piecewise-linear interpolation through (anchor[i], anchor_value[i])."""
    return float(np.interp(now_step, anchor, anchor_value))


def init_logger(filename):
    """log INFO+ to <filename>.log and to the console."""
    root = logging.getLogger()
    root.setLevel(logging.INFO)
    for h in list(root.handlers):
        root.removeHandler(h)
    root.addHandler(logging.FileHandler(filename + ".log", mode="w"))
    root.addHandler(logging.StreamHandler())


def rec_round(x, ndigits=2):
    """round numbers nested in lists/tuples."""
    if isinstance(x, (list, tuple)):
        return [rec_round(item, ndigits) for item in x]
    if isinstance(x, numbers.Number):
        return round(x, ndigits)
    return x
