"""Drop-in mirror of the reference's modules/one_euro_filter.py (One-Euro temporal smoothing of one scalar track,
used by modules/pose.py:108-116 when tracking with smooth=True).  Per-track sequential host arithmetic in Python
floats: the same operations in the same order as the reference, so the same bits."""
import math


def get_alpha(rate=30, cutoff=1):
    """Smoothing factor of a first-order low-pass at `cutoff` Hz sampled at `rate` Hz (reference :4-7)."""
    time_constant = 1 / (2 * math.pi * cutoff)
    period = 1 / rate
    return 1 / (1 + time_constant / period)


class LowPassFilter:
    """Exponential smoothing with a caller-supplied factor; the first sample passes through (reference :10-20)."""

    def __init__(self):
        self.x_previous = None

    def __call__(self, x, alpha=0.5):
        if self.x_previous is not None:
            x = alpha * x + (1 - alpha) * self.x_previous
        self.x_previous = x
        return x


class OneEuroFilter:
    """Low-pass whose cut-off grows with the (smoothed) speed of the signal (reference :23-43)."""

    def __init__(self, freq=15, mincutoff=1, beta=0.05, dcutoff=1):
        self.freq, self.mincutoff, self.beta, self.dcutoff = freq, mincutoff, beta, dcutoff
        self.filter_x = LowPassFilter()
        self.filter_dx = LowPassFilter()
        self.x_previous = None
        self.dx = None

    def __call__(self, x):
        self.dx = 0 if self.dx is None else (x - self.x_previous) * self.freq
        speed = self.filter_dx(self.dx, get_alpha(self.freq, self.dcutoff))
        cutoff = self.mincutoff + self.beta * abs(speed)
        y = self.filter_x(x, get_alpha(self.freq, cutoff))
        self.x_previous = x
        return y
