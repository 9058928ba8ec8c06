"""``SmoothedValue`` / ``MetricLogger`` of the reference's training loop (``lib/utils/metric_logger.py:6-51``,
used by ``lib/engine/train.py:42-89``): windowed medians / means of scalar meters.

``global_avg`` is, as in the reference, the mean of the WINDOW (not of everything seen); ``total`` / ``count`` keep the
all-time sums.  ``__str__`` formats ``name: median`` with four decimals -- the reference's format string
(``'{}: {.4f}'``) raises when called, this is the evident intent.
"""
from collections import defaultdict, deque

import numpy as np
import torch


class SmoothedValue:
    def __init__(self, window_size=20):
        self.values = deque(maxlen=window_size)
        self.total = 0
        self.count = 0

    def update(self, value):
        self.values.append(value)
        self.total += value
        self.count += 1

    @property
    def median(self):
        return np.median(np.array(self.values))

    @property
    def global_avg(self):
        return np.mean(np.array(self.values))


class MetricLogger:
    def __init__(self, delimiter='\t'):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kargs):
        for k, v in kargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getitem__(self, key):
        return self.meters[key]

    def __str__(self):
        return self.delimiter.join('{}: {:.4f}'.format(n, m.median) for n, m in self.meters.items())
