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
