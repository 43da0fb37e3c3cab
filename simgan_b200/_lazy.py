"""Loss scalars that are read back from the device only when somebody looks at them.

The reference's ``update_gail_dyn`` returns three Python floats (gail.py:193) that the caller merely stores until its
log line formats them (main_gail_dyn_ppo.py:255-256, 325-337).  Returning real floats costs one blocking device->host
read per call -- five idle gaps of ~0.2 ms per outer iteration between the discriminator epochs.  ``LazyFloat`` is what
the update returns instead: the per-step loss trace is copied to pinned host memory asynchronously and an event is
recorded; the first ``float()`` / ``format()`` / arithmetic / comparison on any of the three values waits for that
event, sums the trace in the reference's order and from then on behaves like the float it stands for."""
import math
import operator

import torch


class PendingTrace:
    """One update call's (n_steps, k) loss trace on its way to the host."""

    def __init__(self, trace_dev, n_cols, check):
        self._dev = trace_dev
        self._host = torch.empty(trace_dev.shape, dtype=trace_dev.dtype).pin_memory()
        self._host.copy_(trace_dev, non_blocking=True)
        self._event = torch.cuda.current_stream(trace_dev.device).record_event()
        self._n_cols = n_cols
        self._check = check
        self._means = None

    def done(self):
        return self._event is None or self._event.query()

    def trace(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
            self._dev = None
        return self._host

    def means(self):
        if self._means is None:
            rows = self.trace().tolist()
            sums = [0.0] * self._n_cols
            for row in rows:                      # sequential Python-float sums, as the reference accumulates .item()s
                for c in range(self._n_cols):
                    sums[c] += row[c]
            if not all(math.isfinite(s) for s in sums):
                raise self._check("update produced non-finite losses (grid barrier timeout or diverged update)")
            n = max(len(rows), 1)
            self._means = tuple(s / n for s in sums)
        return self._means


class HostTrace:
    """A trace that already lives on the host (assigned by hand)."""

    def __init__(self, value):
        self._value = value

    def done(self):
        return True

    def trace(self):
        return self._value

    def means(self):
        return tuple(self._value.double().mean(dim=0).tolist())


def _binary(op, reflected=False):
    if reflected:
        return lambda self, other: op(other, float(self))
    return lambda self, other: op(float(self), other)


class LazyFloat:
    """Stands for ``pending.means()[col]``; materialises on first use."""
    __slots__ = ("_pending", "_col")

    def __init__(self, pending, col):
        self._pending = pending
        self._col = col

    def __float__(self):
        return self._pending.means()[self._col]

    def item(self):
        return float(self)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))

    __str__ = __repr__

    def __bool__(self):
        return bool(float(self))

    def __int__(self):
        return int(float(self))

    def __hash__(self):
        return hash(float(self))

    def __abs__(self):
        return abs(float(self))

    def __neg__(self):
        return -float(self)

    def __pos__(self):
        return float(self)

    def __round__(self, n=None):
        return round(float(self), n)

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        return np.array(float(self), dtype=dtype or np.float64)

    __add__ = _binary(operator.add)
    __radd__ = _binary(operator.add, True)
    __sub__ = _binary(operator.sub)
    __rsub__ = _binary(operator.sub, True)
    __mul__ = _binary(operator.mul)
    __rmul__ = _binary(operator.mul, True)
    __truediv__ = _binary(operator.truediv)
    __rtruediv__ = _binary(operator.truediv, True)
    __floordiv__ = _binary(operator.floordiv)
    __rfloordiv__ = _binary(operator.floordiv, True)
    __mod__ = _binary(operator.mod)
    __rmod__ = _binary(operator.mod, True)
    __pow__ = _binary(operator.pow)
    __rpow__ = _binary(operator.pow, True)
    __lt__ = _binary(operator.lt)
    __le__ = _binary(operator.le)
    __gt__ = _binary(operator.gt)
    __ge__ = _binary(operator.ge)
    __eq__ = _binary(operator.eq)
    __ne__ = _binary(operator.ne)
