"""Scalar meters for a training / evaluation loop.

Contract kept for callers written against the reference trainer (``lib/engine/train.py:42-89`` reads
``meters['loss'].median``, ``.global_avg`` and prints ``str(meters)``; interface at
``lib/utils/metric_logger.py:6-51``):

* ``MetricLogger(delimiter).update(name=value, ...)`` accepts Python numbers and one-element tensors, anything
  else is an ``AssertionError``; ``meters[name]`` is the meter of that name (created on first use);
* a meter exposes ``median`` and ``global_avg`` over the most recent ``window_size`` values (the reference's
  ``global_avg`` is windowed too, despite its name) and the all-time ``total`` / ``count``;
* ``str(meters)`` is ``"name: median"`` with four decimals, joined by the delimiter.

The implementation is this package's own: a fixed ring of floats per meter, no numpy on the update path.
"""
import statistics


class SmoothedValue:
    """Last ``window_size`` samples of one scalar, plus all-time sum and count."""

    __slots__ = ('_ring', '_next', '_filled', 'total', 'count')

    def __init__(self, window_size=20):
        self._ring = [0.0] * int(window_size)
        self._next = 0
        self._filled = 0
        self.total = 0
        self.count = 0

    def update(self, value):
        self._ring[self._next] = value
        self._next = (self._next + 1) % len(self._ring)
        self._filled = min(self._filled + 1, len(self._ring))
        self.total += value
        self.count += 1

    @property
    def values(self):
        """window contents, oldest first"""
        n = len(self._ring)
        if self._filled < n:
            return self._ring[:self._filled]
        return self._ring[self._next:] + self._ring[:self._next]

    @property
    def median(self):
        return statistics.median(self.values) if self._filled else float('nan')

    @property
    def global_avg(self):
        w = self.values
        return sum(w) / len(w) if w else float('nan')


class MetricLogger:
    def __init__(self, delimiter='\t'):
        self.meters = {}
        self.delimiter = delimiter

    def __getitem__(self, key):
        if key not in self.meters:
            self.meters[key] = SmoothedValue()
        return self.meters[key]

    def update(self, **scalars):
        for name, v in scalars.items():
            if hasattr(v, 'item') and not isinstance(v, (float, int)):
                v = v.item()
            assert isinstance(v, (float, int)), 'meter %r needs a number, got %r' % (name, type(v))
            self[name].update(v)

    def __str__(self):
        return self.delimiter.join('%s: %.4f' % (name, m.median) for name, m in self.meters.items())
