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

try:
    import pandas as _pd
except ImportError:                                # optional: only speeds up the colour census of a mask image
    _pd = None

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

    def _masks(self, seg):
        """all object masks of a colour-coded segmentation after the mask transform (crop, NEAREST resize).
        Nearest-neighbour resampling commutes with the per-pixel colour test, so the packed-colour image is cropped
        and resized ONCE (as a 32-bit integer image, same sampling grid as the reference's per-mask 'L' images) and
        compared against every colour afterwards -- identical output, one resize instead of one per object."""
        key, colours = self._colours(seg)
        # (an image smaller than the crop is padded: with -1, which is no colour, where the reference pads each mask with 0)
        small = T.resize_shorter(T.center_crop(Image.fromarray(key.astype(np.int32)), self.crop, fill=-1), self.size,
                                 Image.NEAREST)
        small = np.asarray(small)
        return np.stack([(small == c) for c in colours], axis=0).astype(np.float32) if len(colours) else \
            np.zeros((0,) + small.shape, np.float32)

    def __getitem__(self, index):
        img_path = self.img_paths[index]
        img = self._image(T.imread(img_path))
        mask = None
        mask_path = os.path.join(self.root, 'masks', os.path.split(img_path)[-1])
        if os.path.exists(mask_path):
            mask = torch.from_numpy(self._masks(T.imread(mask_path)))
        return img, mask

    @staticmethod
    def _colours(img):
        """packed 24-bit colour per pixel and the sorted object colours (everything but the background grey);
        ``KeyError`` when the background colour is absent, as the reference's ``set.remove`` (clevr.py:72)."""
        rgb = np.asarray(img)[:, :, :3].astype(np.uint32)
        key = (rgb[:, :, 0] << 16) | (rgb[:, :, 1] << 8) | rgb[:, :, 2]
        if _pd is not None:                          # hash-based unique: O(n) instead of a sort of H*W keys
            colours = np.sort(_pd.unique(key.ravel()))
        else:
            colours = np.unique(key)
        bg = (BACKGROUND[0] << 16) | (BACKGROUND[1] << 8) | BACKGROUND[2]
        if bg not in colours:
            raise KeyError(BACKGROUND)
        return key, [c for c in colours if c != bg]

    @staticmethod
    def sep(img):
        """colour-coded ``(H, W, >=3)`` segmentation -> list of ``(H, W)`` bool masks, one per colour other than the
        background grey (clevr.py:56-80).  Raises ``KeyError`` when the background colour is absent, as the reference's
        ``set.remove`` does."""
        key, colours = CLEVR._colours(img)
        return [key == c for c in colours]
