"""Process-global side-channel dictionary the model writes visualisation tensors into.

Same contract as the reference's ``lib/utils/vis_logger.py:30-50`` (``logger.update(**kw)``
/ ``logger[key]``): the hot path keeps populating ``image, pred, kl, likelihood, mask_i,
pred_i`` (reference lib/modeling/iodine.py:226-239) so TensorBoard-style consumers keep
working.
"""


class Logger:
    def __init__(self):
        self.things = dict()

    def __getitem__(self, key):
        return self.things[key]

    def __contains__(self, key):
        return key in self.things

    def update(self, **kargs):
        self.things.update(kargs)


logger = Logger()


class VAEGetter:
    """``getter.get_tensorboard_data()`` of the reference (vis_logger.py:61-94): everything in the logger, tensors
    detached and moved to the host, ready for ``TensorBoard.update(**data)``."""

    def __init__(self, logger=logger):
        self.logger = logger

    def get_tensorboard_data(self):
        import torch
        things = self.logger.things
        for k, v in things.items():
            if isinstance(v, torch.Tensor):
                things[k] = v.detach().cpu()
        return things


def make_getter(cfg):
    """vis_logger.py:53-57: ``cfg.GETTER == 'VAE'`` is the only getter the reference defines (IODINE uses it too)."""
    if cfg.GETTER == 'VAE':
        return VAEGetter()
    return None
