"""``TensorBoard`` wrapper with the reference's contract (``lib/utils/tensorboard.py:6-37``): a named set of scalars and
images, ``update(**kw)`` to stage values, ``add(prefix, global_step)`` to write the staged ones that are named as
``<prefix>/<name>``.  Backed by ``torch.utils.tensorboard.SummaryWriter`` (the reference's ``tensorboardX`` and
``easydict`` are not needed); a non-resumed run removes the old log directory with ``shutil`` instead of ``rm -r``.
"""
import os
import shutil


class TensorBoard:
    def __init__(self, logdir, scalars, images, resume):
        if os.path.exists(logdir) and not resume:
            shutil.rmtree(logdir)
        from torch.utils.tensorboard import SummaryWriter
        self.logdir, self.scalars, self.images = logdir, scalars, images
        self.writer = SummaryWriter(log_dir=logdir)
        self.data = {}

    def update(self, **kargs):
        self.data.update(kargs)

    def add(self, prefix, global_step):
        for name in self.scalars:
            if name in self.data:
                self.writer.add_scalar('{}/{}'.format(prefix, name), self.data[name], global_step)
        for name in self.images:
            if name in self.data:
                self.writer.add_image('{}/{}'.format(prefix, name), self.data[name], global_step)

    def close(self):
        self.writer.close()
