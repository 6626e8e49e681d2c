"""Reward filters for the RWR data path and a running mean.

Call surface of the reference's ``ddpo/utils/logger.py:32-94`` (used by ``pipeline/sample.py`` and
``pipeline/save_sizes.py``): ``make_masker(mode, param)`` returns a callable that takes a batch of rewards (``[N]`` or
``[N, 1]``) and returns the boolean keep-mask ``reward >= p``; the current cut-off is exposed as ``.p`` and printed by
``repr``.  ``StreamingAverage()(x)`` folds ``x`` into ``.avg`` (``.n`` values seen).
"""
import numpy as np


def _flat_rewards(rewards):
    r = np.asarray(rewards)
    return r[:, 0] if r.ndim == 2 else r


class Masker:
    """keep-mask against a cut-off ``p``; subclasses decide how ``p`` follows the rewards they are shown"""
    label = "filter"
    p = None

    def _update_cutoff(self, rewards):
        """called with the flattened batch before the mask is taken"""

    def __call__(self, rewards):
        r = _flat_rewards(rewards)
        self._update_cutoff(r)
        return r >= self.p

    def __repr__(self):
        return f"[ {self.label} | {self.p} ]"


class Threshold(Masker):
    """fixed cut-off (``mask_mode='threshold'``)"""

    def __init__(self, threshold=0.95):
        self.p = threshold
        self.label = f"threshold: {threshold}"

    def __call__(self, rewards):           # a threshold never reshapes: [N, 1] rewards give an [N, 1] mask
        return np.asarray(rewards) >= self.p


class Percentile(Masker):
    """cut-off = q-th percentile of the batch at hand"""

    def __init__(self, q=90, maxsize=None):
        self.q = q
        self.label = f"percentile: {q}"

    def _update_cutoff(self, rewards):
        self.p = np.percentile(rewards, self.q)


class StreamingPercentile(Masker):
    """cut-off = q-th percentile of every reward seen so far (the RWR-sparse default); history is kept in a
    geometrically grown buffer instead of the reference's 5M-entry preallocation"""

    def __init__(self, q=90, maxsize=5e6):
        self.q = q
        self.capacity = int(maxsize)
        self._hist = np.empty(min(self.capacity, 4096), np.float64)
        self.size = 0
        self.label = f"streaming_percentile: {q}"

    def _update_cutoff(self, rewards):
        end = self.size + len(rewards)
        if end > self.capacity:
            raise ValueError(f"StreamingPercentile: more than maxsize={self.capacity} rewards")
        if end > len(self._hist):
            grown = np.empty(min(self.capacity, max(end, 2 * len(self._hist))), np.float64)
            grown[: self.size] = self._hist[: self.size]
            self._hist = grown
        self._hist[self.size:end] = rewards
        self.size = end
        self.p = np.percentile(self._hist[:end], self.q)

    @property
    def xs(self):
        return self._hist


class StreamingAverage:
    def __init__(self):
        self.n, self.avg = 0, 0.0

    def __call__(self, x):
        self.n += 1
        self.avg += (x - self.avg) / self.n


_MASKERS = {"percentile": Percentile, "streaming_percentile": StreamingPercentile, "threshold": Threshold}


def make_masker(mode, param):
    if mode not in _MASKERS:
        raise KeyError(mode)
    return _MASKERS[mode](param)
