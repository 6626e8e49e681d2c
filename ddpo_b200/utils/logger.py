"""Sample-filter maskers and the streaming average -- same names and behaviour as the reference's
``ddpo/utils/logger.py:32-94`` (``make_masker(mode, param)`` -> ``masker(rewards) -> bool mask``, ``xs >= p``)."""
import numpy as np


class Masker:
    def __repr__(self):
        return f"[ {self._name} | {self.p} ]"

    def mask(self, xs):
        return xs >= self.p


class StreamingAverage:
    def __init__(self):
        self.n = 0
        self.avg = 0
        self._name = "streaming_average"

    def __call__(self, x):
        self.n += 1
        self.avg = self.avg * (self.n - 1) / self.n + x / self.n


class StreamingPercentile(Masker):
    """percentile over every reward seen so far (the RWR-sparse filter, ``mask_mode`` default)"""

    def __init__(self, q=90, maxsize=5e6):
        self.q = q
        self.xs = np.zeros(int(maxsize))
        self.size = 0
        self._name = f"streaming_percentile: {q}"

    def __call__(self, xs):
        xs = np.asarray(xs)
        if xs.ndim == 2:
            xs = xs.squeeze(axis=-1)
        n = len(xs)
        self.xs[self.size: self.size + n] = xs[:]
        self.size += n
        self.p = np.percentile(self.xs[: self.size], self.q)
        return super().mask(xs)


class Percentile(Masker):
    def __init__(self, q=90, maxsize=5e6):
        self.q = q
        self._name = f"percentile: {q}"

    def __call__(self, xs):
        xs = np.asarray(xs)
        if xs.ndim == 2:
            xs = xs.squeeze(axis=-1)
        self.p = np.percentile(xs, self.q)
        return super().mask(xs)


class Threshold(Masker):
    def __init__(self, threshold=0.95):
        self.p = threshold
        self._name = f"threshold: {threshold}"

    def __call__(self, xs):
        return super().mask(np.asarray(xs))


def make_masker(mode, param):
    return {"percentile": Percentile, "streaming_percentile": StreamingPercentile, "threshold": Threshold}[mode](param)
