"""CPU: observability mirrors: MetricLogger / SmoothedValue against the live reference,
the logger getter, and the TensorBoard wrapper writing event files through torch.utils.tensorboard."""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import ref_loader as R

from iodine_b200.utils.metric_logger import MetricLogger, SmoothedValue
from iodine_b200.utils.vis_logger import Logger, VAEGetter, logger, make_getter


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
def test_meters_match_live_reference():
    if R.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, R.REFERENCE_ROOT)
    from lib.utils.metric_logger import MetricLogger as RefLogger
    a, b = MetricLogger(', '), RefLogger(', ')
    rng = np.random.RandomState(0)
    for i in range(57):                                   # beyond the 20-value window
        v = float(rng.rand())
        for m in (a, b):
            m.update(loss=v, batch_time=torch.tensor(v * 2))
        assert a['loss'].median == b['loss'].median and abs(a['loss'].global_avg - b['loss'].global_avg) < 1e-12
        assert a['batch_time'].total == b['batch_time'].total and a['loss'].count == b['loss'].count == i + 1
    assert a.delimiter == b.delimiter


def test_meters_window_and_str():
    s = SmoothedValue(window_size=3)
    for v in (1.0, 2.0, 30.0, 4.0):
        s.update(v)
    assert s.median == 4.0 and s.global_avg == 12.0 and s.total == 37.0 and s.count == 4
    m = MetricLogger(' | ')
    m.update(loss=torch.tensor(0.5), n=3)
    assert str(m) == 'loss: 0.5000 | n: 3.0000'
    with pytest.raises(AssertionError):
        m.update(bad='x')


def test_getter_detaches_and_moves_to_host():
    lg = Logger()
    w = torch.ones(2, requires_grad=True)
    lg.update(pred=w * 2, kl=1.5)
    data = VAEGetter(lg).get_tensorboard_data()
    assert not data['pred'].requires_grad and data['kl'] == 1.5 and data is lg.things
    assert isinstance(make_getter(NS(GETTER='VAE')), VAEGetter) and make_getter(NS(GETTER='other')) is None
    assert VAEGetter().logger is logger


def test_tensorboard_wrapper_writes_named_events(tmp_path):
    pytest.importorskip('torch.utils.tensorboard')
    from iodine_b200.utils.tensorboard import TensorBoard
    d = str(tmp_path / 'tb')
    tb = TensorBoard(d, scalars=['loss', 'var'], images=['pred'], resume=False)
    tb.update(loss=0.25, pred=torch.rand(3, 8, 8), ignored=7)
    tb.add('train', 10)                                    # 'var' is declared but not staged: skipped
    tb.close()
    files = [f for f in os.listdir(d) if 'tfevents' in f]
    assert len(files) == 1
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(d, size_guidance={'images': 0, 'scalars': 0})
    acc.Reload()
    tags = acc.Tags()
    assert tags['scalars'] == ['train/loss'] and tags['images'] == ['train/pred']
    assert acc.Scalars('train/loss')[0].step == 10 and abs(acc.Scalars('train/loss')[0].value - 0.25) < 1e-7
    # resume=False wipes the directory, resume=True keeps it
    TensorBoard(d, [], [], resume=True).close()
    assert any(f == files[0] for f in os.listdir(d))
    TensorBoard(d, [], [], resume=False).close()
    assert files[0] not in os.listdir(d)
