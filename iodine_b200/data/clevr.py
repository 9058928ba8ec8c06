"""``CLEVR`` dataset with the reference's contract (lib/data/clevr.py:10-80): ``root/images/*`` sorted by path,
optional ``root/masks/<same file name>`` colour-coded segmentation; ``__getitem__`` returns
``(image float32 [3,128,128] in [0,1], masks float32 [N,128,128] of 0/1 or None)`` after a 192 centre crop and a
resize to 128 (bilinear for the image, nearest for every mask).

Differences: images are decoded with PIL instead of skimage (same bytes for PNG/JPEG), ``np.float`` (removed from
numpy) is not used, and ``sep`` is vectorised -- its masks come out in ascending packed-RGB order instead of the
reference's ``set`` iteration order, which no consumer depends on (ARI is permutation invariant).
"""
import os

import numpy as np
import torch
from PIL import Image, ImageFile
from torch.utils.data import Dataset

from . import transforms as T

ImageFile.LOAD_TRUNCATED_IMAGES = True          # as the reference (clevr.py:8)
BACKGROUND = (64, 64, 64)                        # clevr.py:72


class CLEVR(Dataset):
    crop, size = 192, 128                        # clevr.py:27-28

    def __init__(self, root, mode=None):
        self.root = root
        assert os.path.exists(root), 'Path {} does not exist'.format(root)
        self.img_paths = sorted(f.path for f in os.scandir(os.path.join(root, 'images')))

    def __len__(self):
        return len(self.img_paths)

    def _image(self, arr):
        img = Image.fromarray(np.ascontiguousarray(arr[:, :, :3]))
        return T.to_tensor(T.resize_shorter(T.center_crop(img, self.crop), self.size, Image.BILINEAR))

    def _mask(self, m):
        img = Image.fromarray(m.astype(np.uint8))                         # mode 'L', values 0/1
        return np.asarray(T.resize_shorter(T.center_crop(img, self.crop), self.size, Image.NEAREST))

    def __getitem__(self, index):
        img_path = self.img_paths[index]
        img = self._image(T.imread(img_path))
        mask = None
        mask_path = os.path.join(self.root, 'masks', os.path.split(img_path)[-1])
        if os.path.exists(mask_path):
            seps = self.sep(T.imread(mask_path))
            mask = torch.from_numpy(np.stack([self._mask(m) for m in seps], axis=0).astype(np.float32))
        return img, mask

    @staticmethod
    def sep(img):
        """colour-coded ``(H, W, >=3)`` segmentation -> list of ``(H, W)`` bool masks, one per colour other than the
        background grey (clevr.py:56-80).  Raises ``KeyError`` when the background colour is absent, as the reference's
        ``set.remove`` does."""
        rgb = np.asarray(img)[:, :, :3].astype(np.uint32)
        key = (rgb[:, :, 0] << 16) | (rgb[:, :, 1] << 8) | rgb[:, :, 2]
        colours = np.unique(key)
        bg = (BACKGROUND[0] << 16) | (BACKGROUND[1] << 8) | BACKGROUND[2]
        if bg not in colours:
            raise KeyError(BACKGROUND)
        return [key == c for c in colours if c != bg]
