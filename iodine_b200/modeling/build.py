"""``make_model(cfg)`` -- same call as the reference's ``lib/modeling/build.py:5-19``.

The reference wraps the model in ``torch.nn.DataParallel`` when ``cfg.MODEL.PARALLEL`` is
set (build.py:11-12).  Here data parallelism is one process per GPU sharding whole images
(``iodine_b200.parallel``), so PARALLEL only adds a transparent ``.module`` attribute that
``lib/engine/train.py:97`` / ``lib/engine/eval.py:15-16`` dereference.
"""
import torch

from .iodine import IODINE


class _ModuleAlias(torch.nn.Module):
    """Gives ``model.module`` (and ``module.``-prefixed state_dict keys) like DataParallel."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **kw):
        return self.module(*a, **kw)


def make_model(cfg):
    device = torch.device(cfg.MODEL.DEVICE)
    if cfg.MODEL.NAME != 'IODINE':
        raise ValueError('only MODEL.NAME == "IODINE" is provided (got %r)' % cfg.MODEL.NAME)
    precision = str(getattr(cfg.MODEL, 'PRECISION', 'fp32'))
    model = IODINE(cfg.ARCH, precision=precision).to(device)
    if cfg.MODEL.PARALLEL:
        model = _ModuleAlias(model)
    return model
