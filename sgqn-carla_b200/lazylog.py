"""Deferred scalars for `L.log(key, value, step)` (logger.py:108-113).  The reference hands the logger a CUDA
tensor and pays a `.item()` device sync per key (5 per even step).  A `LazyScalar` is a plain-float stand-in:
the logger's `AverageMeter` arithmetic (`_sum += value`, `_sum / count`, logger.py:37-43) stays lazy and the
device->host read happens once, when a number is really needed (`Logger.dump`, `float()`, formatting)."""
import torch


class _Ring:
    """Pinned host ring of loss vectors, one row per update; rows become readable after their event."""

    def __init__(self, width=8, depth=4096):
        self.host = torch.zeros(depth, width).pin_memory()
        self.events = [None] * depth
        self.depth, self.n = depth, 0

    def push(self, dev_vec):
        slot = self.n % self.depth
        self.host[slot].copy_(dev_vec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[slot] = ev
        self.n += 1
        return slot, self.n - 1

    def read(self, slot, serial, col):
        if self.n - serial > self.depth:
            raise RuntimeError("LazyScalar read after its ring slot was recycled")
        self.events[slot].synchronize()
        return float(self.host[slot, col])


class LazyScalar(object):
    __slots__ = ("_terms", "_scale")

    def __init__(self, terms, scale=1.0):
        self._terms, self._scale = terms, scale          # terms: list of (ring, slot, serial, col, weight)

    def __float__(self):
        return self._scale * sum(w * r.read(s, n, c) for r, s, n, c, w in self._terms)

    item = __float__

    def __add__(self, o):
        if isinstance(o, LazyScalar):
            return LazyScalar([(r, s, n, c, w * self._scale) for r, s, n, c, w in self._terms] +
                              [(r, s, n, c, w * o._scale) for r, s, n, c, w in o._terms])
        if o == 0:
            return self
        return float(self) + o

    __radd__ = __add__

    def __mul__(self, k):
        return LazyScalar(self._terms, self._scale * float(k))

    __rmul__ = __mul__

    def __truediv__(self, k):
        return float(self) / k

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))
