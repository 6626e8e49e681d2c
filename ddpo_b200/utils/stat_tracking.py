"""Per-prompt reward normalisation (host, NumPy) -- mirrors ``ddpo/utils/stat_tracking.py:6-35``
(``PerPromptStatTracker``: ring buffer of the last ``buffer_size`` rewards per prompt; advantages use the
prompt's own mean/std once ``min_count`` rewards are buffered, else the batch statistics; std + 1e-6)."""
from collections import deque

import numpy as np


class PerPromptStatTracker:
    def __init__(self, buffer_size, min_count):
        self.buffer_size = buffer_size
        self.min_count = min_count
        self.stats = {}

    def update(self, prompts, rewards):
        prompts = np.asarray(prompts)
        rewards = np.asarray(rewards)
        advantages = np.empty_like(rewards)
        batch_mean, batch_std = np.mean(rewards), np.std(rewards) + 1e-6
        for prompt in np.unique(prompts):
            sel = prompts == prompt
            buf = self.stats.setdefault(prompt, deque(maxlen=self.buffer_size))
            buf.extend(rewards[sel])
            if len(buf) < self.min_count:
                mean, std = batch_mean, batch_std
            else:
                mean, std = np.mean(buf), np.std(buf) + 1e-6
            advantages[sel] = (rewards[sel] - mean) / std
        return advantages

    def get_stats(self):
        return {k: {"mean": np.mean(v), "std": np.std(v), "count": len(v)} for k, v in self.stats.items()}


def global_advantages(rewards):
    """``pipeline/policy_gradient.py:347``: plain z-score, no epsilon."""
    rewards = np.asarray(rewards)
    return (rewards - np.mean(rewards)) / np.std(rewards)
