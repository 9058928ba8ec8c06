"""``MultiDSprites`` with the reference's contract (lib/data/dsprite.py:11-32): ``root/images/<i>.png`` and
``root/masks/<i>.npy``; ``__getitem__`` returns ``(image float32 [C,H,W] in [0,1], masks float64 [N,H,W])`` -- the
reference converts the mask with ``astype(np.float)``, i.e. float64, and does not call ``.float()`` on it.

Difference: the reference hard-codes ``__len__ = 60000``; here it is the number of images present, capped at that.
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from . import transforms as T


class MultiDSprites(Dataset):
    max_len = 60000                              # dsprite.py:31

    def __init__(self, root, mode=None):
        self.root = root
        d = os.path.join(root, 'images')
        n = sum(1 for f in os.scandir(d) if f.name.endswith('.png')) if os.path.isdir(d) else 0
        self.n = min(n, self.max_len)

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        img = T.imread(os.path.join(self.root, 'images/{}.png'.format(index)))
        mask = np.load(os.path.join(self.root, 'masks/{}.npy'.format(index)))
        return T.to_tensor(img), torch.from_numpy(mask.astype(np.float64))
