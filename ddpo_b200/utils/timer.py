"""``utils.Timer``: a lap timer with the reference's call convention (``ddpo/utils/timer.py``) -- ``timer()`` gives the
seconds since the previous lap and starts a new one; ``timer(reset=False)`` only reads."""
import time


class Timer:
    def __init__(self):
        self.lap_start = time.perf_counter()

    def __call__(self, reset=True):
        t = time.perf_counter()
        elapsed = t - self.lap_start
        if reset:
            self.lap_start = t
        return elapsed
