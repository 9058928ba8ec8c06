"""``Checkpointer`` with the reference's contract and ON-DISK FORMAT (``lib/utils/checkpoint.py:6-121``), so that a
directory written by either implementation resumes in the other:

* ``<save_dir>/<name>.pth`` = ``torch.save`` of ``{'model': state_dict, ['optimizer': ...], ['scheduler': ...],
  **args}``;
* ``<save_dir>/checkpoint.pkl`` = pickled list of the checkpoint file names, oldest first; at most
  ``max_checkpoints`` are kept, the oldest file is deleted when the list overflows;
* ``load(f=None)`` prefers the newest entry of ``checkpoint.pkl`` over ``f``, restores model / optimizer / scheduler
  and merges everything else into ``args``; with nothing to load it prints ``No checkpoint found.`` and returns ``{}``.

Differences: ``save_dir`` is created with its parents; a ``module.`` prefix mismatch between the checkpoint and the
model (``torch.nn.DataParallel`` on one side only, ``lib/modeling/build.py:11-12``) is repaired instead of raising;
tensors are loaded with ``weights_only`` semantics when the file allows it.
"""
import os
import pickle

import torch

INDEX = 'checkpoint.pkl'


def _match_prefix(state, model):
    """add / strip the DataParallel ``module.`` prefix so that ``state`` fits ``model``"""
    want = next(iter(model.state_dict()), '')
    have = next(iter(state), '')
    if want.startswith('module.') and not have.startswith('module.'):
        return {'module.' + k: v for k, v in state.items()}
    if have.startswith('module.') and not want.startswith('module.'):
        return {k[len('module.'):]: v for k, v in state.items()}
    return state


class Checkpointer:
    def __init__(self, model, optimizer=None, scheduler=None, args=None, max_checkpoints=10, save_dir='',
                 trust_pickle=False):
        """``trust_pickle``: allow ``torch.load(weights_only=False)`` for checkpoint files whose extra ``args`` hold
        arbitrary pickled objects (the reference's loader always does, lib/utils/checkpoint.py:64).  Unpickling runs
        code from the file, so it is OFF unless the caller vouches for the file."""
        self.trust_pickle = bool(trust_pickle)
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.args = {} if args is None else args
        self.max_checkpoints = max_checkpoints
        self.save_dir = save_dir
        if save_dir:
            os.makedirs(save_dir, exist_ok=True)

    # ------------------------------------------------------------------ index file
    def _index_path(self):
        return os.path.join(self.save_dir, INDEX)

    def _read_index(self):
        if not self.has_checkpoint():
            return []
        with open(self._index_path(), 'rb') as f:
            return list(pickle.load(f))

    def has_checkpoint(self):
        return os.path.exists(self._index_path())

    def get_checkpoint_file(self):
        return os.path.join(self.save_dir, self._read_index()[-1])

    def update_checkpoint(self, last_filename):
        names = self._read_index()
        names.append(os.path.basename(last_filename))
        if len(names) > self.max_checkpoints:
            oldest = names.pop(0)
            path = os.path.join(self.save_dir, oldest)
            if oldest not in names and os.path.exists(path):      # the same name may have been saved again
                os.remove(path)
        with open(self._index_path(), 'wb') as f:
            pickle.dump(names, f)

    # ------------------------------------------------------------------ save / load
    def save(self, name):
        data = {'model': self.model.state_dict()}
        if self.optimizer is not None:
            data['optimizer'] = self.optimizer.state_dict()
        if self.scheduler is not None:
            data['scheduler'] = self.scheduler.state_dict()
        data.update(self.args)
        path = os.path.join(self.save_dir, '{}.pth'.format(name))
        torch.save(data, path)
        self.update_checkpoint(path)

    def _load_file(self, f):
        try:
            return torch.load(f, map_location=torch.device('cpu'), weights_only=True)
        except pickle.UnpicklingError as e:
            if not self.trust_pickle:
                raise pickle.UnpicklingError(
                    '%s holds objects that torch.load(weights_only=True) refuses (%s). Pass trust_pickle=True to '
                    'Checkpointer only for files from a trusted source: full unpickling executes code.' % (f, e))
            import warnings
            warnings.warn('loading %s with full unpickling (trust_pickle=True)' % f)
            return torch.load(f, map_location=torch.device('cpu'), weights_only=False)

    def load(self, f=None):
        if self.has_checkpoint():
            f = self.get_checkpoint_file()
        if not f:
            print('No checkpoint found.')
            return {}
        print('Loading checkpoint from {}'.format(f))
        ckpt = self._load_file(f)
        self.model.load_state_dict(_match_prefix(ckpt.pop('model'), self.model))
        if 'optimizer' in ckpt and self.optimizer:
            print('Loading optimizer from {}'.format(f))
            self.optimizer.load_state_dict(ckpt.pop('optimizer'))
        if 'scheduler' in ckpt and self.scheduler:
            print('Loading scheduler from {}'.format(f))
            self.scheduler.load_state_dict(ckpt.pop('scheduler'))
        self.args.update(ckpt)
